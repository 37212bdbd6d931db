"""Dev tool: wall time of simulations.drainage (numpy in, Results out) with one flood for all pressure steps vs one per step."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import porespy_b200 as psb
from porespy_b200 import simulations, _lib

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
im = psb.generators.blobs([size] * 3, porosity=0.6, blobiness=size / 512, seed=0, rng="philox")
ctx = _lib.context(0)
ref = None
for one in (True, False):
    simulations.ONE_FLOOD = one
    ts = []
    for rep in range(3):
        ctx.set_profile(rep == 2); ctx.profile_read()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = psb.simulations.drainage(im=im, voxel_size=1e-5, bins=25)
        ts.append((time.perf_counter() - t0) * 1e3)
    prof = ctx.profile_read(); ctx.set_profile(False)
    if ref is None: ref = r.im_pc.copy()
    print(json.dumps({"edge": size, "one_flood": one, "wall_ms": [round(t, 1) for t in ts],
                      "kernel_ms": {k: round(m, 2) for k, (m, c) in sorted(prof.items())},
                      "same_im_pc": bool(np.array_equal(r.im_pc, ref, equal_nan=True))}), flush=True)
