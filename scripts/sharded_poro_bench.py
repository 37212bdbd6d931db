"""Sharded access-limited porosimetry at full per-GPU volume (BASELINE config 4, second half): every rank
holds a [S, S, S] slab of a [N*S, S, S] blobs-like volume; times ShardedVolume.porosimetry(sizes=25, default
face inlets) on the device (CUDA events, max over ranks) and prints one JSON line on rank 0.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_poro_bench.py [S]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from porespy_b200 import _lib
from porespy_b200.sharded import ShardedVolume


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ctx = _lib.context(local)
    shape = (world * S, S, S)
    job = ShardedVolume(shape, ctx)
    im = job.blobs(porosity=0.6, blobiness=2.0 * float(np.mean(shape)) / S, seed=0).view(*job.local_shape)
    torch.cuda.empty_cache()
    times = []
    for it in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if it == 2:
            ctx.set_profile(True)
            ctx.profile_read()
        e0.record()
        out = job.porosimetry(im, sizes=25)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        times.append(ms)
        invaded = float((out != 0).sum().item())
        del out
    prof = ctx.profile_read()
    ctx.set_profile(False)
    if rank == 0:
        print(json.dumps({"workload": f"ShardedVolume.porosimetry(blobs-like {list(shape)}, sizes=25, inlets=faces)",
                          "n_gpus": world, "ms": [round(t, 1) for t in times],
                          "voxels_per_s": float(np.prod(shape)) / (times[-1] * 1e-3),
                          "flood_sweeps_per_radius": job.flood_sweeps, "invaded_voxels_rank0": invaded,
                          "kernel_ms_rank0": {k: round(m, 2) for k, (m, c) in sorted(prof.items())}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
