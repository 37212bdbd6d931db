"""A/B of one context option on the 1024^3 workloads: kernel-family times (CUDA events) of
local_thickness(sizes=25) and of local_thickness(100 linear radii) for each value of the option.
    python scripts/ab_option.py <option> <v0> <v1> ..."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import porespy_b200 as psb
from porespy_b200 import _lib


def main():
    opt, vals = sys.argv[1], [int(v) for v in sys.argv[2:]]
    ctx = _lib.context(0)
    im = bench.device_blobs((1024,) * 3, 0.6, 2, 0, torch.device("cuda", 0))
    sizes100 = np.linspace(1, 46.0, 100)
    for v in vals:
        _lib.check(ctx.lib.psb200_set_option(ctx.handle, opt.encode(), v))
        for name, fn in (("lt25", lambda: psb.filters.local_thickness(im, sizes=25)),
                         ("lt100", lambda: psb.filters.local_thickness(im, sizes=sizes100))):
            for _ in range(2):
                out = fn()
                del out
            torch.cuda.synchronize()
            ctx.set_profile(True)
            ctx.profile_read()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            prof = ctx.profile_read()
            ctx.set_profile(False)
            del out
            print(json.dumps({"option": opt, "value": v, "case": name, "ms": round(e0.elapsed_time(e1), 2),
                              "kernels": {k: round(m, 2) for k, (m, c) in sorted(prof.items())}}), flush=True)


if __name__ == "__main__":
    main()
