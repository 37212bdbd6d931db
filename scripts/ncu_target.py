"""ncu target: one local_thickness(sizes=25) on a device-generated blobs volume (no warm-up,
no timing -- numbers printed under a profiler are never bench values)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import porespy_b200 as psb

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = sys.argv[2] if len(sys.argv) > 2 else "lt"
im = psb.generators.blobs([size] * 3, porosity=bench.POROSITY, blobiness=bench.BLOBINESS, seed=0, rng="philox", as_numpy=False)
torch.cuda.empty_cache()
torch.cuda.synchronize()
if mode == "lt":
    out = psb.filters.local_thickness(im, sizes=bench.SIZES)
elif mode == "poro":
    out = psb.filters.porosimetry(im, sizes=bench.SIZES)
elif mode == "poro50":
    # BASELINE config 2: access-limited drainage from the z = 0 face, 50 radii
    inl = torch.zeros((size,) * 3, dtype=torch.uint8, device=im.device)
    inl[0] = 1
    out = psb.filters.porosimetry(im, sizes=50, inlets=inl, mode="dt")
elif mode == "edt":
    out = psb.edt(im)
torch.cuda.synchronize()
print("done", tuple(out.shape))
