"""GPU suite: radius-map post-processing on the index form (porespy_b200.sizemap) through
the device histogram / expansion / index-build kernels, against the reference-generated goldens
(tests/golden/make_golden_sizemap.py) and the numpy restatements in oracle/cpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cpu as oc                       # noqa: E402  (checker only)
from tests.test_gpu_parity import assert_same      # noqa: E402


@pytest.fixture(scope="module")
def psb():
    import torch
    assert torch.cuda.is_available()
    import porespy_b200 as psb
    return psb


def test_index_map_roundtrip(psb, golden):
    b = golden.blobs100
    im, lt = b.mask("im"), b.rmap("lt_dt_25")
    m = psb.local_thickness_index(im, sizes=25, mode="dt")
    assert m.idx.numel() == im.size and m.idx_bytes == 1
    assert_same(m.to_numpy(), lt, "index map of local_thickness")
    inl = np.zeros_like(im)
    inl[0] = True
    p = psb.porosimetry_index(im, sizes=12, inlets=inl, mode="dt")
    assert_same(p.to_numpy(), b.rmap("poro_inlet0_dt_12"), "index map of porosimetry")
    # arbitrary arrays: float64 with -1 / -0.0, int64 sequences beyond 256 values, float32
    rng = np.random.default_rng(0)
    a = rng.choice(np.array([-1.0, -0.0, 0.0, 1.5, 2.25, 1e-300, 7e12]), (30, 41, 17))
    back = psb.IndexMap.from_array(a).to_numpy()
    assert np.array_equal(back, a)
    s = rng.integers(-1, 3000, (64, 50, 31))
    ms = psb.IndexMap.from_array(s)
    assert ms.idx_bytes == 2 and np.array_equal(ms.to_numpy(), s) and np.array_equal(ms.values, np.unique(s))
    c = ms.counts()
    assert np.array_equal(c, np.unique(s, return_counts=True)[1])
    f = rng.choice(np.array([0, 1.25, 3.5], dtype=np.float32), (20, 20))
    mf = psb.IndexMap.from_array(f)
    assert mf.to_numpy().dtype == np.float32 and np.array_equal(mf.to_numpy(), f)
    with pytest.raises(ValueError):
        psb.IndexMap.from_array(rng.random((300, 300)))


def test_goldens(psb, golden):
    f, me = psb.filters, psb.metrics
    g, b = golden.sizemap, golden.blobs100
    im, lt, mip = b.mask("im"), b.rmap("lt_dt_25"), b.rmap("poro_inlet0_dt_12")
    ltm = psb.local_thickness_index(im, sizes=25)                       # straight from the radius loop
    for src in (lt, ltm):
        assert_same(f.size_to_satn(src), g.rmap("lt_satn_dr"), "satn dr")
        assert_same(f.size_to_satn(src, mode="imbibition"), g.rmap("lt_satn_im"), "satn im")
        assert_same(f.size_to_satn(src, bins=12), g.rmap("lt_satn_bins12"), "satn bins")
        assert_same(f.size_to_seq(src), g.rmap("lt_seq_dr").astype(np.int64), "seq dr")
        assert_same(f.size_to_seq(src, mode="imbibition"), g.rmap("lt_seq_im").astype(np.int64), "seq im")
        assert_same(f.size_to_seq(src, bins=10), g.rmap("lt_seq_bins10").astype(np.int64), "seq bins")
        for name, kw in (("psd_default", {}), ("psd_lin20", dict(bins=20, log=False)), ("psd_vox", dict(bins=7, voxel_size=2.5))):
            r = me.pore_size_distribution(src, **kw)
            for fld in ("pdf", "cdf", "satn", "bin_centers", "bin_edges", "bin_widths"):
                assert np.array_equal(getattr(r, fld), g.raw(f"{name}__{fld}")), (name, fld)
        r = me.pc_curve(im, sizes=src, voxel_size=1e-5)
        assert np.array_equal(r.pc, g.raw("pc_lt__pc")) and np.array_equal(r.snwp, g.raw("pc_lt__snwp"))
    assert_same(f.size_to_satn(mip, im=im), g.rmap("mip_satn_im_mask"), "satn mask")
    assert_same(f.size_to_seq(mip, im=im), g.rmap("mip_seq_mask").astype(np.int64), "seq mask")
    seq = g.rmap("lt_seq_dr").astype(np.int64)
    assert_same(f.seq_to_satn(seq), g.rmap("seq_satn_dr"), "seq->satn dr")
    assert_same(f.seq_to_satn(seq, mode="imbibition"), g.rmap("seq_satn_im"), "seq->satn im")
    assert_same(f.seq_to_satn(g.rmap("mseq").astype(np.int64), im=im), g.rmap("mseq_satn_mask"), "seq->satn mask")
    r = me.pc_curve(None, sizes=mip)
    assert np.array_equal(r.pc, g.raw("pc_mip__pc")) and np.array_equal(r.snwp, g.raw("pc_mip__snwp"))
    small = g.raw("small")
    assert_same(f.size_to_satn(small), g.raw("small_satn"), "small satn")
    assert_same(f.size_to_seq(small), g.raw("small_seq"), "small seq")


@pytest.mark.parametrize("seed", [0, 1])
def test_random_maps_vs_numpy_restatement(psb, seed):
    import torch
    f = psb.filters
    rng = np.random.default_rng(seed)
    shape = (37, 23, 64)
    radii = np.concatenate([[0.0], np.sort(rng.uniform(1, 30, 300))])       # > 256 values: two-byte indices
    size = radii[rng.integers(0, len(radii), shape)]
    size[rng.random(shape) < 0.05] = -1
    im = rng.random(shape) < 0.7
    for kw in (dict(), dict(mode="imbibition"), dict(bins=7), dict(im=im), dict(im=im, mode="imbibition")):
        assert_same(f.size_to_satn(size, **kw), oc.size_to_satn(size, **kw), f"satn {kw}")
        assert_same(f.size_to_seq(size, **kw), oc.size_to_seq(size, **kw), f"seq {kw}")
    seq = oc.size_to_seq(size)
    for kw in (dict(), dict(mode="imbibition"), dict(im=im)):
        assert_same(f.seq_to_satn(seq, **kw), oc.seq_to_satn(seq, **kw), f"seq->satn {kw}")
    # device tensors in, binary-mask validation
    assert_same(f.size_to_satn(torch.from_numpy(size).cuda(), im=torch.from_numpy(im).cuda()), oc.size_to_satn(size, im=im), "tensors")
    with pytest.raises(NotImplementedError):
        f.size_to_satn(size, im=im.astype(int) * 2)


def test_psd_of_bench_like_volume_without_float_map(psb):
    """The point of the index form: local thickness -> pore size distribution without the 8 B/voxel map."""
    im = psb.generators.blobs([160, 160, 160], porosity=0.6, blobiness=1.5, seed=2)
    m = psb.local_thickness_index(im, sizes=25)
    r = psb.metrics.pore_size_distribution(m, bins=10)
    want = oc.pore_size_distribution(oc.local_thickness_c(im, sizes=25), bins=10)
    for fld in ("pdf", "cdf", "satn", "bin_centers", "bin_edges", "bin_widths"):
        assert np.array_equal(getattr(r, fld), want[fld]), fld
