"""Times the flood users at 1024^3 (numpy in -> numpy out through the public API, CUDA-event time of the
flood itself from the kernel profile) and checks them against size-independent properties."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import porespy_b200 as psb
from porespy_b200 import _lib


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ctx = _lib.context(0)
    im = psb.generators.blobs([S] * 3, porosity=0.6, blobiness=2, seed=0, rng="philox")
    f = psb.filters
    inl, outl = np.zeros_like(im), np.zeros_like(im)
    inl[0], outl[-1] = True, True
    cases = (("find_disconnected_voxels conn=6", lambda: f.find_disconnected_voxels(im, conn=6)),
             ("find_disconnected_voxels conn=26", lambda: f.find_disconnected_voxels(im)),
             ("trim_nonpercolating_paths z-faces", lambda: f.trim_nonpercolating_paths(im, inl, outl)))
    res = {}
    for name, fn in cases:
        out = fn()
        ctx.set_profile(True)
        ctx.profile_read()
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        prof = ctx.profile_read()
        ctx.set_profile(False)
        res[name] = out
        print(json.dumps({"case": name, "edge": S, "e2e_ms": round(dt * 1e3, 1), "result_voxels": int(out.sum()),
                          "kernel_ms": {k: round(m, 2) for k, (m, c) in sorted(prof.items())}}), flush=True)
    h6, h26, perc = (res[c[0]] for c in cases)
    # properties: holes lie in the foreground; 26-connected holes are a subset of the 6-connected ones;
    # percolating paths contain no hole and touch both faces
    ok = {"holes_in_foreground": bool((h6 <= im).all() and (h26 <= im).all()),
          "h26_subset_h6": bool((h26 <= h6).all()),
          "percolating_disjoint_from_holes": bool(not (perc & h6).any()),
          "percolating_touches_both_faces": bool(perc[0].any() and perc[-1].any()),
          "idempotent": bool(np.array_equal(f.trim_nonpercolating_paths(perc, inl, outl), perc))}
    print(json.dumps({"properties": ok}), flush=True)
    assert all(ok.values())


if __name__ == "__main__":
    main()
