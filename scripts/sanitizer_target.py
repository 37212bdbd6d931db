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
b = psb.generators.blobs([32, 40, 64], porosity=0.6, blobiness=1, seed=1)
assert abs(b.mean() - 0.6) < 0.1
print("SANITIZER_TARGET_OK", ctx.launch_count(), "launches")
