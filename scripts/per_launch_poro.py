"""Dev tool: per-kernel time totals of one porosimetry call (CUDA events via the ABI).
usage: per_launch_poro.py [size] [sizes] [faces|zface]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import numpy as np, torch
import porespy_b200 as psb
from porespy_b200 import _lib

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sizes = int(sys.argv[2]) if len(sys.argv) > 2 else 25
mode = sys.argv[3] if len(sys.argv) > 3 else "faces"
im = psb.generators.blobs([size] * 3, porosity=0.6, blobiness=2, seed=0, rng="philox", as_numpy=False)
inlets = None
if mode == "zface":
    inlets = torch.zeros_like(im, dtype=torch.bool)
    inlets[0] = True
torch.cuda.empty_cache()
ctx = _lib.context(0)
once = len(sys.argv) > 4 and sys.argv[4] == "once"      # under ncu: a single call
for _ in range(0 if once else 2):
    out = psb.porosimetry_index(im, sizes=sizes, inlets=inlets)
    del out
torch.cuda.synchronize()
ctx.set_profile(True); ctx.profile_read()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); out = psb.porosimetry_index(im, sizes=sizes, inlets=inlets); b.record(); torch.cuda.synchronize()
recs = ctx.profile_records()
tot = collections.OrderedDict()
for name, ms in recs:
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += ms
print(f"call {a.elapsed_time(b):.2f} ms (profiling on), {len(recs)} launches")
for name, (c, ms) in tot.items():
    print(f"{name:16s} x{c:4d} {ms:9.3f} ms")
if os.environ.get("VERBOSE"):
    for name, ms in recs:
        if name.startswith("uf_"):
            print(f"  {name:16s} {ms:8.3f}")
