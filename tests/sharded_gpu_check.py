"""Run under torchrun on N GPUs (N = WORLD_SIZE >= 1):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/sharded_gpu_check.py
Every rank checks its z-slab of the sharded EDT / local_thickness (NCCL all-to-all + halo
exchange, CUDA kernels through the C ABI) bit-exactly against the CPU oracle on the whole volume."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpu as oc                      # checker only
from porespy_b200 import _lib
from porespy_b200.sharded import ShardedVolume


CASES = [((96, 88, 128), 12, 200), ((97, 70, 160), 10, 0), ((64, 64, 96), [9, 6.5, 4, 2.5, 1.5, 1], 20)]


def run_cases(ctx, rank, world, cases, verbose=True):
    """Raises AssertionError on the first mismatch."""
    for shape, sizes, bit_tmax in cases:
        im = oc.blobs(list(shape), porosity=0.6, blobiness=1.5, seed=5)
        job = ShardedVolume(shape, ctx)
        job.backend.bit_tmax = bit_tmax
        sl = job.local_slice()
        want = oc.edt_sq(im)
        # all-to-all transposes (halo 0), the input-halo fast path (16 planes: deeper than every distance here)
        # and its fallback (2 planes: the exactness bound fails)
        for halo in (0, 2, 16):
            job.edt_halo = halo
            d2, mx = job.edt_sq(job.backend.to_u8(im[sl]))
            path = "halo" if world > 1 and 0 < halo <= min(job.zcounts) and int(want.max()) < (halo + 1) ** 2 else "all-to-all"
            assert job.edt_path == path, (halo, job.edt_path, path)
            assert mx == int(want.max()), (mx, int(want.max()))
            got = d2.cpu().numpy().view(np.uint32).reshape(want[sl].shape)
            assert np.array_equal(got, want[sl]), f"rank {rank}: sharded edt differs {shape} halo={halo}"
        assert np.array_equal(job.edt(im[sl]).cpu().numpy(), oc.edt(im)[sl]), "edt float"
        lt = job.local_thickness(im[sl], sizes=sizes).cpu().numpy()
        ref = oc.local_thickness(im, sizes=sizes, mode="dt")
        if not np.array_equal(lt, ref[sl]):
            bad = np.argwhere(lt != ref[sl])
            raise AssertionError(f"rank {rank}: local_thickness {shape}: {len(bad)} voxels differ, first {bad[:5].tolist()}")
        # access-limited: slab-local union-find + face-flag exchange (psb200_uf_*)
        for inl in (None, "z0", "x0"):
            mask = None
            if inl is not None:
                mask = np.zeros(shape, dtype=bool)
                if inl == "z0":
                    mask[0] = True
                else:
                    mask[:, :, 0] = True
            mip = job.porosimetry(im[sl], sizes=sizes, inlets=None if mask is None else mask[sl]).cpu().numpy()
            ref = oc.porosimetry(im, sizes=sizes, inlets=mask, mode="dt")
            if not np.array_equal(mip, ref[sl]):
                bad = np.argwhere(mip != ref[sl])
                raise AssertionError(f"rank {rank}: porosimetry {shape} inlets={inl}: {len(bad)} voxels differ, "
                                     f"first {bad[:5].tolist()}")
        if rank == 0 and verbose:
            print(f"sharded x{world} ok: {shape} sizes={sizes} bit_tmax={bit_tmax}", flush=True)


def check_sharded_blobs(ctx, rank, world):
    """The sharded generator yields the slabs of the SAME image the one-GPU generator draws from Philox noise
    (the z filter reads noise planes beyond the slab; statistics are fixed-order per-plane sums)."""
    import porespy_b200 as psb
    for shape, blob in (((120, 64, 96), 2), ((90, 50, 64), [1, 2, 3])):
        whole = psb.generators.blobs(list(shape), porosity=0.6, blobiness=blob, seed=11, rng="philox")
        job = ShardedVolume(shape, ctx)
        mine = job.blobs(porosity=0.6, blobiness=blob, seed=11).cpu().numpy().reshape(job.local_shape).astype(bool)
        assert np.array_equal(mine, whole[job.local_slice()]), f"rank {rank}: sharded blobs differ {shape}"
    if rank == 0:
        print(f"sharded x{world} blobs ok", flush=True)


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.context(local)
    run_cases(ctx, rank, world, CASES)
    check_sharded_blobs(ctx, rank, world)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_GPU_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
