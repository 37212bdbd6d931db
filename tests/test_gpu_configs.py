"""GPU suite, BASELINE.json configurations at their STATED sizes: the CUDA path through the
public API against the CPU oracle on the same input, bit-exact.  The oracle side is the C restatement of the
reference loop (oracle.cpu.porosimetry_c, pinned against the numpy restatement in tests/test_oracle.py); it
needs a few minutes of host time at 1024^3, so PSB200_BIG_TESTS=0 skips the 1024^3 cases.

Inputs come from the device blobs generator with numpy's seeded noise stream, i.e. they are
`ps.generators.blobs(shape, porosity=0.6, blobiness=2, seed=0)` (tests/test_gpu_parity.py checks the generator
against the host restatement)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cpu as oc                       # noqa: E402  (checker only)
from tests.test_gpu_parity import assert_same      # noqa: E402

BIG = os.environ.get("PSB200_BIG_TESTS", "1") != "0"


@pytest.fixture(scope="module")
def psb():
    import torch
    assert torch.cuda.is_available(), "GPU suite needs a CUDA device"
    import porespy_b200 as psb
    return psb


@pytest.fixture(scope="module")
def blobs512(psb):
    return psb.generators.blobs([512, 512, 512], porosity=0.6, blobiness=2, seed=0)


@pytest.fixture(scope="module")
def blobs1024(psb):
    if not BIG:
        pytest.skip("PSB200_BIG_TESTS=0")
    return psb.generators.blobs([1024, 1024, 1024], porosity=0.6, blobiness=2, seed=0)


def same_big(got, want, what):
    """array_equal on multi-GB arrays without a full-size temporary per comparison operator."""
    assert got.shape == want.shape and got.dtype == want.dtype, what
    step = max(1, (1 << 27) // max(1, int(np.prod(got.shape[1:]))))
    for z in range(0, got.shape[0], step):
        if not np.array_equal(got[z:z + step], want[z:z + step]):
            assert_same(got[z:z + step], want[z:z + step], f"{what} planes {z}..{z + step}")


def test_config1_edt_512(psb, blobs512):
    """config 1: standalone exact EDT of a 512^3 blobs volume -- squared distances bit-exact, float32 equal."""
    from porespy_b200.edt import edt_sq_u32
    want = oc.edt_sq(blobs512)
    same_big(edt_sq_u32(blobs512), want, "d2 512^3")
    dt = psb.edt(blobs512)
    assert dt.dtype == np.float32
    same_big(dt, oc.edt(blobs512, parallel=0), "edt f32 512^3")


def test_config3_linspace100_512(psb, blobs512):
    """config 3's radii (sizes = linspace(1, max dt, 100): float64 compare path, ~60 radii on the byte
    pipeline) at 512^3."""
    dmax = float(oc.edt(blobs512, parallel=0).max())
    sizes = np.linspace(1, dmax, 100)
    got = psb.filters.local_thickness(blobs512, sizes=sizes)
    want = oc.local_thickness_c(blobs512, sizes=sizes)
    same_big(got, want, "local_thickness linspace(100) 512^3")


def test_config0_style_lt25_1024(psb, blobs1024):
    """the benchmarked configuration itself: local_thickness(blobs(1024^3, 0.6, 2), sizes=25)."""
    got = psb.filters.local_thickness(blobs1024, sizes=25)
    want = oc.local_thickness_c(blobs1024, sizes=25)
    same_big(got, want, "local_thickness(25) 1024^3")


def test_config2_porosimetry50_zface_1024(psb, blobs1024):
    """config 2: porosimetry(sizes=50, inlets = the z=0 face, access_limited, mode='dt') at 1024^3."""
    inlets = np.zeros(blobs1024.shape, dtype=bool)
    inlets[0, :, :] = True
    got = psb.filters.porosimetry(blobs1024, sizes=50, inlets=inlets, access_limited=True, mode="dt")
    want = oc.porosimetry_c(blobs1024, sizes=50, inlets=inlets, access_limited=True)
    del inlets
    same_big(got, want, "porosimetry(50, z-face) 1024^3")
