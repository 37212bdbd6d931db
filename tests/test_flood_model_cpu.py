"""The flood's record rules (tests/model_flood.py, a sequential model of csrc/flood_kernels.cuh) against plain
labelling, on the CPU: random and smooth class maps, inlets inside and outside the sets, both connectivities, both
inlet modes, segment lengths that cut rows at various places."""
import numpy as np
import pytest
import scipy.ndimage as spim

from tests.model_flood import reached_class, reached_class_by_labelling


@pytest.mark.parametrize("inlets_in_set", [False, True])
@pytest.mark.parametrize("conn", [6, 26])
def test_record_rules_match_labelling(conn, inlets_in_set):
    rng = np.random.default_rng(100 * conn + int(inlets_in_set))
    records = voxels = 0
    for trial in range(36):
        shape = (int(rng.integers(1, 5)), int(rng.integers(1, 7)), int(rng.integers(1, 40)))
        nk = int(rng.integers(1, 6))
        if trial % 3 == 0:
            cls = rng.integers(0, nk, shape)
        else:
            sm = spim.gaussian_filter(rng.random(shape), 1.0)
            cls = ((sm - sm.min()) / (np.ptp(sm) + 1e-9) * nk).astype(int).clip(0, nk - 1)
        r = rng.random(shape)
        cls = np.where(r < 0.3, 255, np.where(r < 0.4, 254, cls))
        inlet = rng.random(shape) < (0.05 if trial % 2 else 0.0)
        if trial % 4 == 0:
            inlet[:, :, 0] = True
        if trial % 7 == 0:
            inlet[0] = True
        got, nrec = reached_class(cls, inlet, conn, seg=int(rng.choice([4, 16, 128])), inlets_in_set=inlets_in_set)
        want = reached_class_by_labelling(cls, inlet, conn, inlets_in_set=inlets_in_set)
        assert np.array_equal(got, want), (trial, shape, np.argwhere(got != want)[:5])
        records += nrec
        voxels += int((cls < 254).sum())
    assert records < (3 if conn == 6 else 13) * voxels          # far fewer than one job per voxel and direction
