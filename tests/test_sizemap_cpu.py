"""CPU suite: host logic of porespy_b200.sizemap (the reference's post-processing functions evaluated on one
representative voxel per (value, mask) combination) against the reference-generated goldens and the plain
numpy restatements, with a numpy stand-in for the histogram / expansion kernels."""
import numpy as np
import pytest

from oracle import cpu as oc
from porespy_b200 import sizemap as sm
from tests.cpu_sizemap import CpuIndexMap as M


def test_goldens(golden):
    g, b = golden.sizemap, golden.blobs100
    im, lt, mip = b.mask("im"), b.rmap("lt_dt_25"), b.rmap("poro_inlet0_dt_12")
    assert np.array_equal(sm.size_to_satn(M(lt)), g.rmap("lt_satn_dr"))
    assert np.array_equal(sm.size_to_satn(M(lt), mode="imbibition"), g.rmap("lt_satn_im"))
    assert np.array_equal(sm.size_to_satn(M(lt), bins=12), g.rmap("lt_satn_bins12"))
    assert np.array_equal(sm.size_to_satn(M(mip), im=im), g.rmap("mip_satn_im_mask"))
    assert np.array_equal(sm.size_to_seq(M(lt)), g.rmap("lt_seq_dr"))
    assert np.array_equal(sm.size_to_seq(M(lt), mode="imbibition"), g.rmap("lt_seq_im"))
    assert np.array_equal(sm.size_to_seq(M(mip), im=im), g.rmap("mip_seq_mask"))
    assert np.array_equal(sm.size_to_seq(M(lt), bins=10), g.rmap("lt_seq_bins10"))
    seq = g.rmap("lt_seq_dr").astype(np.int64)
    assert np.array_equal(sm.seq_to_satn(M(seq)), g.rmap("seq_satn_dr"))
    assert np.array_equal(sm.seq_to_satn(M(seq), mode="imbibition"), g.rmap("seq_satn_im"))
    assert np.array_equal(sm.seq_to_satn(M(g.rmap("mseq").astype(np.int64)), im=im), g.rmap("mseq_satn_mask"))
    for name, kw in (("psd_default", {}), ("psd_lin20", dict(bins=20, log=False)), ("psd_vox", dict(bins=7, voxel_size=2.5))):
        r = sm.pore_size_distribution(M(lt), **kw)
        for f in ("pdf", "cdf", "satn", "bin_centers", "bin_edges", "bin_widths"):
            assert np.array_equal(getattr(r, f), g.raw(f"{name}__{f}")), (name, f)
    r = sm.pc_curve(im, sizes=M(lt), voxel_size=1e-5)
    assert np.array_equal(r.pc, g.raw("pc_lt__pc")) and np.array_equal(r.snwp, g.raw("pc_lt__snwp"))
    r = sm.pc_curve(None, sizes=M(mip))
    assert np.array_equal(r.pc, g.raw("pc_mip__pc")) and np.array_equal(r.snwp, g.raw("pc_mip__snwp"))
    small = g.raw("small")
    assert np.array_equal(sm.size_to_satn(M(small)), g.raw("small_satn"))
    assert np.array_equal(sm.size_to_seq(M(small)), g.raw("small_seq"))
    assert np.array_equal(sm.size_to_seq(M(small), mode="imbibition"), g.raw("small_seq_im"))
    assert np.array_equal(sm.seq_to_satn(M(g.raw("small_seq"))), g.raw("small_seq_satn"))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_maps_vs_numpy_restatement(seed):
    rng = np.random.default_rng(seed)
    shape = (17, 23, 11)
    radii = np.concatenate([[0.0], np.sort(rng.uniform(1, 30, 9))])
    size = radii[rng.integers(0, len(radii), shape)]
    size[rng.random(shape) < 0.05] = -1
    im = rng.random(shape) < 0.7
    for kw in (dict(), dict(mode="imbibition"), dict(bins=7), dict(im=im), dict(im=im, mode="imbibition")):
        assert np.array_equal(sm.size_to_satn(M(size), **kw), oc.size_to_satn(size, **kw)), kw
        assert np.array_equal(sm.size_to_seq(M(size), **kw), oc.size_to_seq(size, **kw)), kw
    seq = oc.size_to_seq(size)
    for kw in (dict(), dict(mode="imbibition"), dict(im=im)):
        assert np.array_equal(sm.seq_to_satn(M(seq), **kw), oc.seq_to_satn(seq, **kw)), kw
    with pytest.raises(NotImplementedError):
        sm.pc_curve(im, pc=size)
