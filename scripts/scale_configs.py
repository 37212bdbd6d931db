"""BASELINE configurations 3 and 4 on N GPUs of one box, one JSON line per measurement on rank 0:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/scale_configs.py [--size S]
                                    [--what check,lt,poro,cfg3] [--stated-blobiness] [--sha-out FILE]
    python scripts/scale_configs.py --what lt --global-shape 2048,2048,2048 --sha-out FILE     (one GPU, same image)

  check : tests/sharded_gpu_check.py cases (small volumes, every path against the CPU oracle, incl. the generator)
  lt    : local_thickness(sizes=25) of ONE global blobs image (Philox noise per global voxel: the same image for
          every N), device-resident, CUDA events, max over ranks; SHA-256 of every 128-plane chunk of the radius
          INDEX map, so that runs with different N (incl. one GPU) can be compared offline
  poro  : porosimetry(sizes=25, default face inlets), the same way
  cfg3  : local_thickness(sizes=linspace(1, max dt, 100)) (config 3: float64 radii, ~60 of them on the byte pipeline)
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from porespy_b200 import _host, _lib
from porespy_b200.sharded import ShardedVolume

CHUNK = 128


def sha_chunks(idx, lshape, z0):
    """{global first plane: sha256} for every CHUNK-plane chunk of this slab's index map."""
    out = {}
    plane = lshape[1] * lshape[2]
    h = idx.cpu().numpy()
    for z in range(0, lshape[0], CHUNK):
        out[str(z0 + z)] = hashlib.sha256(h[z * plane:(z + CHUNK) * plane].tobytes()).hexdigest()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--what", default="check,lt,poro")
    ap.add_argument("--stated-blobiness", action="store_true")
    ap.add_argument("--global-shape", default="")
    ap.add_argument("--sha-out", default="")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ctx = _lib.context(local)
    what = args.what.split(",")
    say = (lambda d: print(json.dumps(d), flush=True)) if rank == 0 else (lambda d: None)

    if "check" in what:
        from tests import sharded_gpu_check as chk
        chk.run_cases(ctx, rank, world, chk.CASES, verbose=False)
        chk.check_sharded_blobs(ctx, rank, world)
        say({"check": "tests/sharded_gpu_check.py cases + sharded blobs vs one-GPU blobs", "n_gpus": world, "ok": True})

    shape = tuple(int(v) for v in args.global_shape.split(",")) if args.global_shape else bench.global_shape(args.size, world)
    blobiness = 2.0 if args.stated_blobiness else 2.0 * float(np.mean(shape)) / args.size
    job = ShardedVolume(shape, ctx)
    lshape, z0 = job.local_shape, job.zstarts[job.rank]
    if any(w in what for w in ("lt", "poro", "cfg3")):
        im = job.blobs(porosity=0.6, blobiness=blobiness, seed=0).view(*lshape)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    def timed(fn, reps):
        times, res = [], None
        for _ in range(reps):
            del res
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=device, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            times.append(ms)
        return times, res

    def gather_sha(idx):
        mine = sha_chunks(idx, lshape, z0)
        if world == 1:
            return mine
        allsha = [None] * world
        dist.all_gather_object(allsha, mine)
        merged = {}
        for d in allsha:
            merged.update(d)
        return merged

    if world == 1:
        # one GPU: the public single-GPU API (not the sharded driver's step-level path)
        import porespy_b200 as psb

        class One:
            last_max_d2 = None

            @staticmethod
            def local_thickness(im_, sizes=25, as_index=True):
                m = psb.local_thickness_index(im_, sizes=sizes)
                return m.idx, m.values

            @staticmethod
            def porosimetry(im_, sizes=25, as_index=True):
                m = psb.porosimetry_index(im_, sizes=sizes)
                return m.idx, m.values
        run = One
    else:
        run = job
    shas = {}
    nvox = float(np.prod(shape))
    base = {"n_gpus": world, "shape": list(shape), "blobiness": blobiness}
    if "lt" in what:
        times, (idx, lut) = timed(lambda: run.local_thickness(im, sizes=25, as_index=True), args.reps)
        shas["lt"] = gather_sha(idx)
        say(dict(base, workload="local_thickness(sizes=25), index form", ms=[round(t, 2) for t in times],
                 voxels_per_s=nvox / (min(times[1:] or times) * 1e-3), edt_path=getattr(job, "edt_path", "one GPU"),
                 max_d2=getattr(run, "last_max_d2", None), radii=len(lut) - 1))
        del idx
    if "poro" in what:
        times, (idx, lut) = timed(lambda: run.porosimetry(im, sizes=25, as_index=True), args.reps)
        shas["poro"] = gather_sha(idx)
        say(dict(base, workload="porosimetry(sizes=25, inlets=faces), index form", ms=[round(t, 2) for t in times],
                 voxels_per_s=nvox / (min(times[1:] or times) * 1e-3), flood_sweeps_per_radius=getattr(job, "flood_sweeps", None)))
        del idx
    if "cfg3" in what:
        if world > 1:
            job.local_thickness(im, sizes=2, as_index=True)
            max_d2 = job.last_max_d2
        else:
            from porespy_b200 import _device as pdev
            _, max_d2 = pdev.edt_run(ctx, im.reshape(-1), shape, want_max=True)
        dmax = float(np.sqrt(np.float32(max_d2)))
        sizes = np.linspace(1, dmax, 100)
        times, (idx, lut) = timed(lambda: run.local_thickness(im, sizes=sizes, as_index=True), max(2, args.reps - 1))
        shas["cfg3"] = gather_sha(idx)
        say(dict(base, workload="local_thickness(sizes=linspace(1, max dt, 100)), index form", ms=[round(t, 2) for t in times],
                 voxels_per_s=nvox / (min(times[1:] or times) * 1e-3), effective_radii=len(lut) - 1, max_d2=int(max_d2)))
        del idx
    if args.sha_out and rank == 0:
        with open(args.sha_out, "w") as f:
            json.dump(dict(base, chunk_planes=CHUNK, sha=shas), f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
