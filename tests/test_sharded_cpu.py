"""world_size-2/3 gloo runs of the z-slab sharded driver (porespy_b200/sharded.py) on CPU with the
numpy step backend: the sharded result must equal the oracle's result on the whole volume."""
import os
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cpu as oc


def _free_port():
    """A rendezvous file, not a TCP port: a probed-then-released port can be taken by somebody else
    before the workers bind it (seen as a rare failure of this test)."""
    import tempfile
    fd, path = tempfile.mkstemp(prefix="psb200_gloo_")
    os.close(fd)
    os.unlink(path)                    # FileStore creates it
    return path


def _worker(rank, world, port, shape, sizes, bit_tmax, errq):
    try:
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
        dist.init_process_group("gloo", init_method=f"file://{port}", rank=rank, world_size=world)
        from porespy_b200.sharded import ShardedVolume
        from tests.cpu_backend import CpuBackend
        im = oc.blobs(list(shape), porosity=0.6, blobiness=1.5, seed=3)
        job = ShardedVolume(shape, backend=CpuBackend(bit_tmax=bit_tmax))
        sl = job.local_slice()
        # EDT through both all-to-all transposes (edt_halo = 0), through the input-halo fast path (halo deeper than
        # the largest distance) and through its fallback (halo too shallow: the bound fails, all-to-all runs)
        want = oc.edt_sq(im)
        deep = min(job.zcounts)
        for halo in (0, 1, deep):
            path = "halo" if halo > 0 and int(want.max()) < (halo + 1) ** 2 else "all-to-all"
            job.edt_halo = halo
            d2, mx = job.edt_sq(job.backend.to_u8(im[sl]))
            assert job.edt_path == path, (halo, job.edt_path, path)
            assert mx == int(want.max()), (mx, int(want.max()))
            assert np.array_equal(d2.numpy().view(np.uint32).reshape(want[sl].shape), want[sl]), f"edt slab halo={halo}"
        assert np.array_equal(job.edt(im[sl]).numpy(), oc.edt(im)[sl]), "edt float"
        # radius loop with halo exchange (byte and bit pipelines)
        lt = job.local_thickness(im[sl], sizes=sizes).numpy()
        ref = oc.local_thickness(im, sizes=sizes, mode="dt")
        assert lt.dtype == np.float64
        if not np.array_equal(lt, ref[sl]):
            bad = np.argwhere(lt != ref[sl])
            raise AssertionError(f"rank {rank}: {len(bad)} voxels differ, first {bad[:5].tolist()}")
        # access-limited porosimetry: slab-local flooding + face-flag exchange (default face inlets,
        # then a single-face mask: the path from the inlet face crosses every slab boundary)
        for inl in (None, "z0", "x0"):
            mask = None
            if inl is not None:
                mask = np.zeros(shape, dtype=bool)
                if inl == "z0":
                    mask[0] = True
                else:
                    mask[:, :, 0] = True
            mip = job.porosimetry(im[sl], sizes=sizes, inlets=None if mask is None else mask[sl]).numpy()
            ref = oc.porosimetry(im, sizes=sizes, inlets=mask, mode="dt")
            if not np.array_equal(mip, ref[sl]):
                bad = np.argwhere(mip != ref[sl])
                raise AssertionError(f"rank {rank}: porosimetry inlets={inl}: {len(bad)} voxels differ, "
                                     f"first {bad[:5].tolist()}")
            assert max(job.flood_sweeps) >= 1
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


def _run_world(world, shape, sizes, bit_tmax):
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, sizes, bit_tmax, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            msgs.append("worker timed out")
    if any(p.exitcode != 0 for p in procs) and not msgs:
        msgs.append(f"worker exit codes {[p.exitcode for p in procs]}")
    return msgs


@pytest.mark.parametrize("world,shape,sizes,bit_tmax", [
    (2, (24, 20, 32), 8, 200),          # bit pipeline for every radius
    (2, (25, 21, 32), 8, 0),            # byte pipeline only, uneven slabs and pencils
    (3, (31, 26, 64), [5, 3.2, 2, 1], 6),   # mixed: large radii bytes, small radii bits; 3 ranks
])
def test_sharded_local_thickness_gloo(world, shape, sizes, bit_tmax):
    msgs = _run_world(world, shape, sizes, bit_tmax)
    if msgs and not any("differ" in m or "AssertionError" in m for m in msgs):
        # rendezvous trouble (the probed port was taken before the workers bound it): one more try;
        # a wrong RESULT is never retried
        msgs = _run_world(world, shape, sizes, bit_tmax)
    assert not msgs, "\n".join(msgs)


def test_partition_helpers():
    from porespy_b200.sharded import ceil_split_counts, split_counts
    assert split_counts(10, 3) == [4, 3, 3]
    assert split_counts(8, 8) == [1] * 8
    assert ceil_split_counts(10, 4) == (3, [3, 3, 3, 1])
    assert ceil_split_counts(1024, 8) == (128, [128] * 8)
