"""Dev tool: per-launch kernel times of one local_thickness call (CUDA events via the ABI)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import porespy_b200 as psb
from porespy_b200 import _lib, _host

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
im = psb.generators.blobs([size] * 3, porosity=0.6, blobiness=2, seed=0, rng="philox", as_numpy=False)
torch.cuda.empty_cache()
ctx = _lib.context(0)
if os.environ.get("BIT_TMAX"):
    ctx.set_bit_tmax(int(os.environ["BIT_TMAX"]))
if os.environ.get("FOOT"):
    ctx.set_foot(int(os.environ["FOOT"]))
if os.environ.get("YFLAGS"):
    ctx.set_yflags(int(os.environ["YFLAGS"]))
if os.environ.get("ZWIDE"):
    ctx.set_zwide(int(os.environ["ZWIDE"]))
if os.environ.get("EDT_H"):
    ctx.set_edt_h(int(os.environ["EDT_H"]))
if os.environ.get("YDIRECT"):
    ctx.set_ydirect(int(os.environ["YDIRECT"]))
if os.environ.get("BITQUAD"):
    ctx.set_bitquad(int(os.environ["BITQUAD"]))
if os.environ.get("XBITS"):
    ctx.set_xbits(int(os.environ["XBITS"]))
if os.environ.get("YCOARSE"):
    ctx.set_ycoarse(int(os.environ["YCOARSE"]))
for _ in range(2):
    psb.filters.local_thickness(im, sizes=25)
torch.cuda.synchronize()
ctx.set_profile(True); ctx.profile_read()
psb.filters.local_thickness(im, sizes=25)
recs = ctx.profile_records()
from porespy_b200 import _device as d
d2 = d.edt_sq(ctx, im, im.shape); mx = d.max_u32(ctx, d2)
T, R = _host.effective_thresholds(_host.reference_sizes(25, mx), mx)
print("max d2", mx, "T", list(T))
k = 0
for name, ms in recs:
    tag = ""
    if name in ("lt_xy", "lt_point"):
        tag = f"k={k} T={T[k]} W={int(np.sqrt(T[k]-1))}"; 
    if name in ("lt_z", "lt_point"):
        k += 1
    print(f"{name:14s} {ms:8.3f} ms {tag}")
