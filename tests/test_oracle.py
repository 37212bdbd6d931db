"""CPU suite: pins the oracle (oracle/cpu.py + oracle/edt_oracle.c) against the golden
vectors produced by the reference's own source (tests/golden/make_golden.py) and against
the golden numbers asserted by the reference's tests."""
import numpy as np
import pytest

from oracle import cpu as oc
from tests.golden_io import sha


def _scipy_d2(mask):
    import scipy.ndimage as spim
    idx = spim.distance_transform_edt(mask, return_distances=False, return_indices=True)
    d2 = np.zeros(mask.shape, dtype=np.int64)
    grids = np.meshgrid(*[np.arange(n) for n in mask.shape], indexing="ij", sparse=True)
    for ax in range(mask.ndim):
        d2 += (idx[ax].astype(np.int64) - grids[ax]) ** 2
    return d2.astype(np.uint32)


@pytest.mark.parametrize("shape,p", [((37, 29, 41), 0.5), ((64, 64, 64), 0.97), ((1, 50, 60), 0.8),
                                     ((5, 1, 7), 0.6), ((17, 251), 0.9), ((129,), 0.95)])
def test_edt_oracle_vs_scipy(shape, p):
    rng = np.random.default_rng(hash(shape) & 0xFFFF)
    im = rng.random(shape) < p
    im.flat[rng.integers(im.size)] = False          # at least one background voxel
    assert np.array_equal(oc.edt_sq(im), _scipy_d2(im))


def test_edt_oracle_degenerate():
    assert np.all(oc.edt_sq(np.ones((4, 5, 6), bool)) == oc.INF_U32)
    assert np.all(np.isinf(oc.edt(np.ones((4, 5, 6), bool))))
    assert np.all(oc.edt_sq(np.zeros((4, 5, 6), bool)) == 0)
    one = np.ones((9, 9, 9), bool)
    one[4, 4, 4] = False
    z, y, x = np.mgrid[-4:5, -4:5, -4:5]
    assert np.array_equal(oc.edt_sq(one), (x * x + y * y + z * z).astype(np.uint32))


def test_edt_oracle_golden(golden):
    g = golden.blobs100
    im = g.mask("im")
    assert im.sum() / im.size == 0.499829                 # TF:17
    d2 = oc.edt_sq(im)
    assert sha(d2) == str(g.raw("d2_sha"))
    assert int(d2.max()) == int(g.raw("d2_max"))
    assert np.array_equal(d2[:, :, 50], g.raw("d2_slice50"))
    assert np.array_equal(oc.edt_sq(im, nthreads=1), d2)


def test_blobs_restatement_reproduces_reference_image(golden):
    np.random.seed(0)
    im = oc.blobs(shape=[100, 100, 100], blobiness=2)
    assert np.array_equal(im, golden.blobs100.mask("im"))
    im = oc.blobs([200, 200], porosity=0.55, blobiness=2, seed=0)
    assert np.array_equal(im, golden.trim.mask("im2d"))


def test_strels_golden(golden):
    g = golden.strels
    assert list(g.raw("sums")) == [25, 93, 29, 123]        # test_tools.py:309-316
    assert np.array_equal(oc.ps_round(3, 2), g.mask("disk3"))
    assert np.array_equal(oc.ps_round(3, 3), g.mask("ball3"))
    assert np.array_equal(oc.ps_round(3, 2, smooth=False), g.mask("disk3_rough"))
    assert np.array_equal(oc.ps_round(3, 3, smooth=False), g.mask("ball3_rough"))
    assert np.array_equal(oc.ps_round(np.float32(4.2426405), 3), g.mask("ball_4p2426"))


def test_porosimetry_num_points(golden):
    im = golden.blobs100.mask("im")
    mip = oc.porosimetry(im=im, sizes=10)
    ans = np.array([0.00000000, 1.00000000, 1.37871571, 1.61887041, 1.90085700, 2.23196205,
                    2.62074139, 3.07724114, 3.61325732])    # TF:39-41
    assert np.allclose(np.unique(mip), ans)
    assert np.array_equal(mip, golden.blobs100.rmap("poro_hybrid_sizes10"))
    assert np.array_equal(oc.porosimetry(im=im, sizes=10, mode="dt"), mip)


def test_porosimetry_modes_and_sizes(golden):
    g = golden.blobs100
    im = g.mask("im")
    sizes = np.arange(25, 1, -1)
    assert np.array_equal(oc.porosimetry(im, sizes=sizes, mode="dt"), g.rmap("poro_dt_arange_3d"))
    im2d = im[:, :, 50]
    assert np.array_equal(oc.porosimetry(im2d, sizes=sizes, mode="dt"), g.rmap("poro_dt_arange_2d"))
    assert np.array_equal(oc.porosimetry(im2d, sizes=sizes, mode="hybrid"), g.rmap("poro_dt_arange_2d"))
    s = np.logspace(0.01, 0.6, 5)
    mip = oc.porosimetry(im=im, sizes=s, mode="dt")
    assert np.allclose(np.unique(mip)[1:], s)                # TF:53-56
    assert np.array_equal(mip, g.rmap("poro_logsizes"))


def test_local_thickness_golden(golden):
    g = golden.blobs100
    im = g.mask("im")
    lt = oc.local_thickness(im, mode="dt")
    np.testing.assert_almost_equal(lt.max(), oc.edt(im).max(), decimal=6)   # TF:266-272
    assert np.array_equal(lt, g.rmap("lt_dt_25"))
    assert np.array_equal(oc.local_thickness(im[:, :, 50]), g.rmap("lt_2d_25"))
    assert np.array_equal(oc.local_thickness(im, sizes=[6, 4.5, 3, 2, 1], mode="dt"),
                          g.rmap("lt_list_sizes"))
    assert np.array_equal(oc.porosimetry(im[:, :, 50], sizes=9, access_limited=False),
                          g.rmap("poro_2d_noaccess"))


def test_porosimetry_single_face_inlet(golden):
    g = golden.blobs100
    im = g.mask("im")
    inlets = np.zeros_like(im)
    inlets[0, ...] = True
    assert np.array_equal(oc.porosimetry(im, sizes=12, inlets=inlets, mode="dt"),
                          g.rmap("poro_inlet0_dt_12"))


def test_trim_disconnected_blobs_golden(golden):
    g = golden.trim
    im, inl = g.mask("im2d"), g.mask("inlets2d")
    assert np.array_equal(oc.trim_disconnected_blobs(im, inl), g.mask("out8"))
    assert np.array_equal(oc.trim_disconnected_blobs(im, inl, strel=oc._cross(2)), g.mask("out4"))
    im, inl = g.mask("im3d"), g.mask("inlets3d")
    assert np.array_equal(oc.trim_disconnected_blobs(im, inl), g.mask("out26"))
    assert np.array_equal(oc.trim_disconnected_blobs(im, inl, strel=oc._cross(3)), g.mask("out6"))
    with pytest.raises(Exception, match="inlets not valid"):
        oc.trim_disconnected_blobs(im, np.zeros((3, 3, 3)))


def test_misc2d_golden(golden):
    g = golden.misc2d
    lt = oc.local_thickness(g.mask("rsa"), sizes=[20, 10])
    assert np.all(np.unique(lt) == [0, 10, 20])               # TF:274-279
    assert np.array_equal(lt, g.rmap("lt_rsa"))
    drn = g.mask("drn")
    lt = oc.local_thickness(drn)
    assert np.array_equal(lt, g.rmap("lt_drn"))
    assert (lt > 25).sum() / drn.sum() == 0.34427115020497745  # test_drainage.py:17-18,49


def test_numpy_integer_scalar_is_one_radius(golden):
    """sizes=np.int64(n) is NOT n log-spaced radii (F:1131-1134)."""
    im = golden.blobs100.mask("im")[:40, :40, :40]
    lt = oc.local_thickness(im, sizes=np.int64(2), mode="dt")
    assert set(np.unique(lt)) <= {0.0, 2.0}


def test_flood_users_golden(golden):
    """find_disconnected_voxels / fill_blind_pores / trim_floating_solid / trim_nonpercolating_paths
    restatements against what the reference's own source returned (tests/golden/flood_users.npz,
    incl. the reference's golden counts TF:107-121 and the surface=True label-0 quirk)."""
    g = golden.flood_users
    im = g.mask("im")
    sl = im[:, :, 0]
    assert [int(oc.find_disconnected_voxels(a, conn=c).sum()) for a, c in ((sl, None), (sl, 4), (im, None), (im, 6))] \
        == [477, 652, 55, 202]
    assert np.array_equal(oc.find_disconnected_voxels(sl), g.mask("h2d8"))
    assert np.array_equal(oc.find_disconnected_voxels(sl, conn=4), g.mask("h2d4"))
    assert np.array_equal(oc.find_disconnected_voxels(im), g.mask("h26"))
    assert np.array_equal(oc.find_disconnected_voxels(im, conn=6), g.mask("h6"))
    assert np.array_equal(oc.find_disconnected_voxels(im, surface=True), g.mask("h26_surface"))
    assert np.array_equal(oc.find_disconnected_voxels(im, conn=6, surface=True), g.mask("h6_surface"))
    assert np.array_equal(oc.find_disconnected_voxels(sl, conn=4, surface=True), g.mask("h2d4_surface"))
    assert np.array_equal(oc.find_disconnected_voxels(g.mask("cap"), conn=6, surface=True), g.mask("cap_surface"))
    assert np.array_equal(oc.fill_blind_pores(im), g.mask("fill_blind"))
    assert np.array_equal(oc.fill_blind_pores(im, conn=6, surface=True), g.mask("fill_blind6s"))
    assert np.array_equal(oc.trim_floating_solid(im), g.mask("trim_solid"))
    assert np.array_equal(oc.trim_floating_solid(im, conn=6), g.mask("trim_solid6"))
    with pytest.raises(Exception, match="conn is not valid"):
        oc.find_disconnected_voxels(im, conn=5)
    for name, key, axes in (("np2d_im", "np2d_ax", (0, 1)), ("np3d_im", "np3d_ax", (0, 1, 2))):
        b = g.mask(name)
        for ax in axes:
            inl, outl = np.zeros_like(b), np.zeros_like(b)
            inl[(slice(None),) * ax + (0,)] = True
            outl[(slice(None),) * ax + (-1,)] = True
            assert np.array_equal(oc.trim_nonpercolating_paths(b, inl, outl), g.mask(f"{key}{ax}"))
    b = g.mask("np2d_none_im")
    inl, outl = np.zeros_like(b), np.zeros_like(b)
    inl[:, 0], outl[:, -1] = True, True
    assert oc.trim_nonpercolating_paths(b, inl, outl).sum() == 0                  # TF:149-160


def _trapped_cases(g):
    seq2 = g.raw("seq2d").astype(np.int64)
    outl = np.zeros(seq2.shape, bool)
    outl[-1, :] = True
    seq3 = g.raw("seq3d").astype(np.int64)
    outl3 = np.zeros(seq3.shape, bool)
    outl3[-1] = True
    return [(dict(seq=seq2), "t2d_faces_25"), (dict(seq=seq2, outlets=outl), "t2d_outlet_25"),
            (dict(seq=seq2, outlets=outl, bins=None), "t2d_outlet_all"), (dict(seq=seq2, outlets=outl, bins=7), "t2d_outlet_7"),
            (dict(seq=seq3, outlets=outl3), "t3d_outlet_25"), (dict(seq=seq3, bins=None), "t3d_faces_all")], seq2, outl


def test_find_trapped_regions_golden(golden):
    """find_trapped_regions restatement against the reference's own output (tests/golden/trapped.npz)."""
    g = golden.trapped
    cases, seq2, outl = _trapped_cases(g)
    for kw, key in cases:
        assert np.array_equal(oc.find_trapped_regions(**kw), g.mask(key)), key
    assert np.array_equal(oc.find_trapped_regions(seq2, outlets=outl, bins=None, return_mask=False), g.raw("s2d_outlet_seq"))


@pytest.mark.parametrize("shape,seed", [((60, 50, 70), 1), ((40, 80), 2), ((33, 31, 64), 3)])
def test_c_loop_matches_numpy_restatement(shape, seed):
    """oracle_porosimetry_dt (the radius loop in C, used at 1024^3) == the numpy restatement of F:1124-1212."""
    im = oc.blobs(list(shape), porosity=0.6, blobiness=1.5, seed=seed)
    for sizes in (12, np.linspace(1, 8, 17), [5, 3, 2], np.int64(3)):
        assert np.array_equal(oc.porosimetry_c(im, sizes=sizes), oc.porosimetry(im, sizes=sizes, mode="dt"))
        assert np.array_equal(oc.local_thickness_c(im, sizes=sizes), oc.local_thickness(im, sizes=sizes, mode="dt"))
    inl = np.zeros(shape, bool)
    inl[0] = True
    assert np.array_equal(oc.porosimetry_c(im, sizes=10, inlets=inl), oc.porosimetry(im, sizes=10, inlets=inl, mode="dt"))
    pt = np.zeros(shape, bool)
    pt[tuple(s // 2 for s in shape)] = True
    assert np.array_equal(oc.porosimetry_c(im, sizes=7, inlets=pt), oc.porosimetry(im, sizes=7, inlets=pt, mode="dt"))
    where = (np.array([0, 1]), np.array([2, 3]), np.array([4, 5]))      # F:1252-1255 (axis-0 fancy index quirk)
    assert np.array_equal(oc.porosimetry_c(im, sizes=5, inlets=where), oc.porosimetry(im, sizes=5, inlets=where, mode="dt"))


def test_c_loop_golden(golden):
    g = golden.blobs100
    im = g.mask("im")
    assert np.array_equal(oc.local_thickness_c(im, sizes=25), g.rmap("lt_dt_25"))
    assert np.array_equal(oc.porosimetry_c(im, sizes=np.arange(25, 1, -1)), g.rmap("poro_dt_arange_3d"))
    inlets = np.zeros_like(im)
    inlets[0, ...] = True
    assert np.array_equal(oc.porosimetry_c(im, sizes=12, inlets=inlets), g.rmap("poro_inlet0_dt_12"))


def test_sizemap_restatements_match_reference_goldens(golden):
    """oracle.cpu size_to_seq / size_to_satn / seq_to_satn / pore_size_distribution / pc_curve_sizes against the
    reference's own outputs (tests/golden/make_golden_sizemap.py)."""
    g, b = golden.sizemap, golden.blobs100
    im, lt, mip = b.mask("im"), b.rmap("lt_dt_25"), b.rmap("poro_inlet0_dt_12")
    assert np.array_equal(oc.size_to_satn(lt), g.rmap("lt_satn_dr"))
    assert np.array_equal(oc.size_to_satn(lt, mode="imbibition"), g.rmap("lt_satn_im"))
    assert np.array_equal(oc.size_to_satn(lt, bins=12), g.rmap("lt_satn_bins12"))
    assert np.array_equal(oc.size_to_satn(mip, im=im), g.rmap("mip_satn_im_mask"))
    assert np.array_equal(oc.size_to_seq(lt), g.rmap("lt_seq_dr"))
    assert np.array_equal(oc.size_to_seq(lt, mode="imbibition"), g.rmap("lt_seq_im"))
    assert np.array_equal(oc.size_to_seq(mip, im=im), g.rmap("mip_seq_mask"))
    assert np.array_equal(oc.size_to_seq(lt, bins=10), g.rmap("lt_seq_bins10"))
    seq = g.rmap("lt_seq_dr")
    assert np.array_equal(oc.seq_to_satn(seq), g.rmap("seq_satn_dr"))
    assert np.array_equal(oc.seq_to_satn(seq, mode="imbibition"), g.rmap("seq_satn_im"))
    assert np.array_equal(oc.seq_to_satn(g.rmap("mseq"), im=im), g.rmap("mseq_satn_mask"))
    for name, kw in (("psd_default", {}), ("psd_lin20", dict(bins=20, log=False))):
        r = oc.pore_size_distribution(lt, **kw)
        for f in ("pdf", "cdf", "satn", "bin_centers", "bin_edges", "bin_widths"):
            assert np.array_equal(r[f], g.raw(f"{name}__{f}")), (name, f)
    x, y = oc.pc_curve_sizes(im, lt, voxel_size=1e-5)
    assert np.array_equal(x, g.raw("pc_lt__pc")) and np.array_equal(y, g.raw("pc_lt__snwp"))
    small = g.raw("small")
    assert np.array_equal(oc.size_to_satn(small), g.raw("small_satn"))
    assert np.array_equal(oc.size_to_seq(small), g.raw("small_seq"))
    assert np.array_equal(oc.size_to_seq(small, mode="imbibition"), g.raw("small_seq_im"))
    assert np.array_equal(oc.seq_to_satn(g.raw("small_seq")), g.raw("small_seq_satn"))


def _drainage_same(got, g, name):
    assert np.array_equal(got["im_pc"], g.rmap(name + "_im_pc")), name + " im_pc"
    assert np.array_equal(got["im_satn"], g.rmap(name + "_im_satn")), name + " im_satn"
    if got["im_trapped"] is not None:
        assert np.array_equal(got["im_trapped"], g.mask(name + "_im_trapped")), name + " trapped"
    assert np.array_equal(np.asarray(got["pc"]), g.raw(name + "_pc")), name + " pc"
    assert np.array_equal(np.asarray(got["snwp"]), g.raw(name + "_snwp")), name + " snwp"


def test_drainage_restatement_matches_reference_goldens(golden):
    """oracle.cpu.drainage against the reference's own simulations.drainage (tests/golden/make_golden_drainage.py)."""
    g = golden.drainage
    im, inl, out, res = g.mask("a_im"), g.mask("a_inlets"), g.mask("a_outlets"), g.mask("a_residual")
    vs = 1e-4
    _drainage_same(oc.drainage(im, vs, inlets=inl, g=0), g, "a1")
    _drainage_same(oc.drainage(im, vs, inlets=inl, outlets=out, residual=res, g=0), g, "a4")
    _drainage_same(oc.drainage(im, vs, inlets=inl, outlets=out), g, "a5")
    _drainage_same(oc.drainage(im, vs, inlets=inl, bins=[300.0, 900.0, 2000.0, 1500.0, 8000.0], delta_rho=-997, g=9.81,
                               sigma=0.05, theta=140), g, "a6")
    _drainage_same(oc.drainage(im, np.float64(vs), inlets=inl, bins=12), g, "a7")
    im3, out3 = g.mask("b_im"), g.mask("b_outlets")
    _drainage_same(oc.drainage(im3, 1e-5), g, "b1")
    _drainage_same(oc.drainage(im3, 1e-5, pc=g.raw("b_pc_user"), bins=10), g, "b3")
