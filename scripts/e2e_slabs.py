"""Dev tool: numpy -> numpy local_thickness(1024^3, 25) wall time for several slab counts of the pipelined epilogue."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import porespy_b200 as psb
from porespy_b200 import filters as F

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
im = psb.generators.blobs([size] * 3, porosity=0.6, blobiness=2, seed=0, rng="philox")
ref = None
for slabs in (0, 2, 4, 6, 8):
    F.SLAB_PIPELINE.update(enabled=slabs > 0, slabs=max(slabs, 2))
    ts = []
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = psb.filters.local_thickness(im, sizes=25)
        ts.append((time.perf_counter() - t0) * 1e3)
        if ref is None:
            ref = out.copy()
        same = bool(np.array_equal(out, ref))
        del out
    print(json.dumps({"slabs": slabs, "ms": [round(t, 1) for t in ts], "equal_to_one_piece": same}), flush=True)
