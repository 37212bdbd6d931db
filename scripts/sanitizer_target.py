"""compute-sanitizer target: every kernel family of the hot path on small volumes, results
checked against the oracle.  Run as
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitizer_target.py
(the sanitizer slows kernels down 10-100x: sizes are tens of voxels per edge)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import porespy_b200 as psb
from oracle import cpu as oc                        # checker only
from porespy_b200 import _lib

ctx = _lib.context(0)
for shape, sizes in (((40, 36, 64), 10), ((24, 300, 32), 8), ((37, 29, 41), 7), ((64, 96), 9)):
    im = oc.blobs(list(shape), porosity=0.6, blobiness=1.0, seed=3)
    assert np.array_equal(psb.edt(im), oc.edt(im)), "edt"
    for tmax in (200, 0):
        ctx.set_bit_tmax(tmax)
        assert np.array_equal(psb.filters.local_thickness(im, sizes=sizes), oc.local_thickness(im, sizes=sizes, mode="dt")), "lt"
        assert np.array_equal(psb.filters.porosimetry(im, sizes=sizes), oc.porosimetry(im, sizes=sizes, mode="dt")), "poro"
    ctx.set_bit_tmax(200)
    inl = np.zeros(shape, bool)
    inl[0] = True
    for strel in (None, oc._cross(im.ndim)):
        assert np.array_equal(psb.filters.trim_disconnected_blobs(im, inl, strel=strel),
                              oc.trim_disconnected_blobs(im, inl, strel=strel)), "flood"
# flood on rows of several 128-voxel segments / partial segments / odd lengths, both implementations, and the
# slab-coupled forms (link records + join times, and job lists + per-radius marking) with the shard helpers
import torch
from porespy_b200 import _host
from porespy_b200.sharded import CudaBackend, split_counts
rng = np.random.default_rng(2)
for records in (True, False):
    ctx.set_uf_records(records)
    for shape, p in (((6, 9, 261), 0.33), ((5, 7, 130), 0.3), ((40, 259), 0.6), ((3, 1, 400), 0.8)):
        im = rng.random(shape) < p
        inl = np.zeros_like(im)
        inl[..., 0] = True
        for strel in (None, oc._cross(len(shape))):
            assert np.array_equal(psb.filters.trim_disconnected_blobs(im, inl, strel=strel),
                                  oc.trim_disconnected_blobs(im, inl, strel=strel)), "flood shapes"
    im = oc.blobs([12, 20, 300], porosity=0.55, blobiness=1.0, seed=5)
    assert np.array_equal(psb.filters.porosimetry(im, sizes=8), oc.porosimetry(im, sizes=8, mode="dt")), "poro wide rows"
    be = CudaBackend(ctx)
    be.uf_records = records
    shape = (30, 24, 200)
    im = oc.blobs(list(shape), porosity=0.55, blobiness=1.5, seed=11)
    d2 = oc.edt_sq(im)
    T, R = _host.effective_thresholds(np.array([5.0, 3.0, 2.0, 1.0]), int(d2.max()))
    counts = split_counts(shape[0], 2)
    sts = []
    for s0, c in zip((0, counts[0]), counts):
        d2s = torch.from_numpy(d2[s0:s0 + c].astype(np.uint32).view(np.int32).copy()).cuda().reshape(-1)
        sts.append(be.uf_begin(be.classify(d2s, T), None, (c, shape[1], shape[2]), s0, shape[0]))
    for k in range(len(T)):
        for st in sts:
            be.uf_activate(st, k - 1, k)
        while True:
            faces = [(be.uf_face(st, k, 0), be.uf_face(st, k, st.shape[0] - 1)) for st in sts]
            be.uf_inject(sts[0], k, sts[0].shape[0] - 1, faces[1][0])
            be.uf_inject(sts[1], k, 0, faces[0][1])
            if not max(be.uf_changed(st) for st in sts):
                break
        for st in sts:
            be.uf_settle(st, k)
    for st in sts:
        be.uf_resolve(st)
    for k, Tk in enumerate(T):
        want = oc.trim_disconnected_blobs(d2 >= Tk, oc.border_faces(shape), strel=oc._cross(3))
        got = np.concatenate([(st.rcls.cpu().numpy().reshape(st.shape) <= k) for st in sts], axis=0)
        assert np.array_equal(got, want), "slab flood"
ctx.set_uf_records(True)
be = CudaBackend(ctx)
src = torch.from_numpy((rng.random(4096 * 5 + 8) < 0.4).astype(np.uint8)).cuda()
back = torch.empty_like(src)
be.mask_unpack(be.mask_pack(src), back)
assert bool((back == src).all())
reach = torch.from_numpy(rng.integers(0, 9, (10, 12, 20)).astype(np.uint8)).cuda().reshape(-1)
be.halo_cone(reach, (10, 12, 20), 6, 0), be.halo_cone(reach, (10, 12, 20), 6, 1)
b = psb.generators.blobs([32, 40, 64], porosity=0.6, blobiness=1, seed=1)
assert abs(b.mean() - 0.6) < 0.1
print("SANITIZER_TARGET_OK", ctx.launch_count(), "launches")
