"""Dev tool: per-kernel time totals of one ShardedVolume.porosimetry call on a single rank (the step-level ABI)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import torch
import porespy_b200 as psb
from porespy_b200 import _lib
from porespy_b200.sharded import ShardedVolume

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
im = psb.generators.blobs([size] * 3, porosity=0.6, blobiness=2, seed=0, rng="philox", as_numpy=False)
ctx = _lib.context(0)
sv = ShardedVolume((size,) * 3, ctx=ctx)
if os.environ.get("UF_RECORDS"):
    sv.backend.uf_records = bool(int(os.environ["UF_RECORDS"]))
for _ in range(2):
    out = sv.porosimetry(im, sizes=25, as_index=True)
    del out
torch.cuda.synchronize()
ctx.set_profile(True); ctx.profile_read()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); out = sv.porosimetry(im, sizes=25, as_index=True); b.record(); torch.cuda.synchronize()
recs = ctx.profile_records()
tot = collections.OrderedDict()
for name, ms in recs:
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += ms
print(f"call {a.elapsed_time(b):.2f} ms (profiling on), {len(recs)} launches, kernels {sum(m for _, m in recs):.2f} ms")
for name, (c, ms) in tot.items():
    print(f"{name:16s} x{c:4d} {ms:9.3f} ms")
if os.environ.get("VERBOSE"):
    for name, ms in recs:
        if name.startswith("uf_"):
            print(f"  {name:16s} {ms:8.3f}")
