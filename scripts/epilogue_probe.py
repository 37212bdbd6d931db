"""Sweep of the host-result epilogue (psb200_expand_idx_f64_to_host) on the GPU box: share of the
volume widened by host threads vs on the device, and thread count.  Prints one JSON line per point.
    python scripts/epilogue_probe.py [edge]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from porespy_b200 import _device as dev
from porespy_b200 import _lib


def main():
    edge = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n = edge ** 3
    ctx = _lib.context(0)
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    idx = torch.randint(0, 24, (n,), generator=g, device="cuda", dtype=torch.uint8)
    lut = np.concatenate([[0.0], np.logspace(np.log10(46.0), 0, 23)])
    ncpu = os.cpu_count()
    print(json.dumps({"edge": edge, "os_cpu_count": ncpu, "affinity": len(os.sched_getaffinity(0))}), flush=True)
    ref = None
    for threads in (0, 8, 32):
        for pm in (0, 300, 500, 600, 700, 800, 1000):
            if pm == 0 and threads != 0:
                continue
            ts = []
            for rep in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = dev.expand_idx_to_host(ctx, idx, lut, (n,), cpu_permille=pm, nthreads=threads)
                ts.append(time.perf_counter() - t0)
                if rep == 0 and threads == 0 and pm in (0, 600):
                    chk = out[::4097].copy()
                    if ref is None:
                        ref = lut[idx[::4097].cpu().numpy()]
                    assert np.array_equal(chk, ref), f"mismatch at permille {pm}"
                del out
            print(json.dumps({"permille": pm, "threads": threads, "ms": [round(t * 1e3, 1) for t in ts]}), flush=True)


if __name__ == "__main__":
    main()
