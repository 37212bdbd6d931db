"""GPU suite (`-m gpu`): the CUDA path, called through the public API / C ABI, must be
bit-exact against the CPU oracle and the reference-generated golden vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cpu as oc                       # noqa: E402  (checker only)
from tests.golden_io import sha                    # noqa: E402


@pytest.fixture(scope="module")
def psb():
    import torch
    assert torch.cuda.is_available(), "GPU suite needs a CUDA device"
    import porespy_b200 as psb
    return psb


def assert_same(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    if np.array_equal(got, want):
        return
    bad = np.argwhere(got != want)
    lines = [f"{what}: {len(bad)} of {got.size} voxels differ"]
    for c in bad[:12]:
        lines.append(f"  at {tuple(int(i) for i in c)}: got {got[tuple(c)]!r} want {want[tuple(c)]!r}")
    lines.append(f"  got values {np.unique(got)[:20]}  want values {np.unique(want)[:20]}")
    raise AssertionError("\n".join(lines))


def set_algo(psb, name):
    """fast: default (bit-parallel kernels for T <= 200, byte pipeline above); nobit: byte pipeline
    only; allbit: bit-parallel kernels up to T = 400; generic: full lower-envelope EDT per radius."""
    from porespy_b200 import _lib
    ctx = _lib.context()
    ctx.set_algo(_lib.ALGO_GENERIC if name == "generic" else _lib.ALGO_FAST)
    ctx.set_bit_tmax({"nobit": 0, "allbit": 400, "allbit1": 400}.get(name, 200))
    ctx.set_bit4(name != "allbit1")          # allbit1: one- / two-words-per-lane kernels only


def rand_image(shape, p, seed):
    rng = np.random.default_rng(seed)
    return rng.random(shape) < p


# ---------------------------------------------------------------------------------- EDT
@pytest.fixture(params=["edt16", "edt32"])
def edt_mode(request):
    """Both forms of the EDT y/z passes: 16-bit kernel with the gated uint32 fallback, uint32 only."""
    from porespy_b200 import _lib
    ctx = _lib.context()
    ctx.set_edt16(request.param == "edt16")
    yield request.param
    ctx.set_edt16(True)


EDT_CASES = [((37, 29, 41), 0.5), ((64, 64, 64), 0.97), ((13, 37, 101), 0.9), ((1, 50, 60), 0.8),
             ((50, 1, 60), 0.8), ((50, 60, 1), 0.8), ((5, 1, 7), 0.6), ((17, 251), 0.9), ((129,), 0.95),
             ((3, 300, 5), 0.99), ((127, 31, 33), 0.995), ((2, 2, 2), 0.5), ((1,), 1.0), ((70, 130), 0.999)]


@pytest.mark.parametrize("shape,p", EDT_CASES)
def test_edt_random(psb, edt_mode, shape, p):
    im = rand_image(shape, p, seed=len(shape) * 1000 + shape[0])
    from porespy_b200.edt import edt_sq_u32
    assert_same(edt_sq_u32(im), oc.edt_sq(im), f"edt_sq {shape}")
    dt = psb.edt(im)
    assert dt.dtype == np.float32
    assert_same(dt, oc.edt(im), f"edt {shape}")


def test_edt_adversarial(psb, edt_mode):
    from porespy_b200.edt import edt_sq_u32
    one_bg = np.ones((33, 40, 47), bool)
    one_bg[16, 20, 23] = False
    assert_same(edt_sq_u32(one_bg), oc.edt_sq(one_bg), "single background voxel")
    one_fg = np.zeros((33, 40, 47), bool)
    one_fg[16, 20, 23] = True
    assert_same(edt_sq_u32(one_fg), oc.edt_sq(one_fg), "single foreground voxel")
    z, y, x = np.indices((24, 24, 24))
    checker = ((z + y + x) % 2).astype(bool)
    assert_same(edt_sq_u32(checker), oc.edt_sq(checker), "checkerboard")
    empty = np.zeros((9, 10, 11), bool)
    assert np.all(psb.edt(empty) == 0)
    full = np.ones((9, 10, 11), bool)
    assert np.all(edt_sq_u32(full) == 0xFFFFFFFF)
    assert np.all(np.isinf(psb.edt(full)))              # black_border=False, note N8
    corner = np.ones((40, 41, 300), bool)
    corner[0, 0, 0] = False                             # long distances, 299^2+40^2+39^2
    assert_same(edt_sq_u32(corner), oc.edt_sq(corner), "far corner")
    # distances that straddle the 16-bit cap (32767) of the two-voxels-per-instruction passes:
    # y pass overflows / z pass overflows / neither
    for shp, bg in (((3, 400, 260), (1, 0, 0)), ((260, 5, 190), (0, 2, 0)), ((200, 150, 8), (199, 149, 7))):
        far = np.ones(shp, bool)
        far[bg] = False
        assert_same(edt_sq_u32(far), oc.edt_sq(far), f"cap straddle {shp}")
        assert_same(psb.edt(far), oc.edt(far), f"cap straddle f32 {shp}")
    tall = rand_image((300, 40, 64), 0.999, 11)           # z tiles with halo (n > 224), sparse background
    assert_same(edt_sq_u32(tall), oc.edt_sq(tall), "tall sparse")
    wide = rand_image((2, 700, 130), 0.9995, 12)          # y tiles with halo, scans beyond the staged halo
    assert_same(edt_sq_u32(wide), oc.edt_sq(wide), "wide sparse")
    # kwargs PoreSpy passes (F:1186-1189, _snows.py:600-607)
    assert_same(psb.edt(data=one_bg, parallel=0), oc.edt(one_bg), "kwargs")
    assert_same(psb.edtsq(one_bg), oc.edt_sq(one_bg).astype(np.float32), "edtsq")
    with pytest.raises(NotImplementedError):
        psb.edt(one_bg, black_border=True)


def test_edt_golden_blobs100(psb, edt_mode, golden):
    from porespy_b200.edt import edt_sq_u32
    g = golden.blobs100
    d2 = edt_sq_u32(g.mask("im"))
    assert sha(d2) == str(g.raw("d2_sha"))
    assert_same(d2[:, :, 50], g.raw("d2_slice50"), "d2 slice")


def test_edt_config1_shape_sample(psb, edt_mode):
    """BASELINE config 1 at a size the oracle finishes in seconds (256^3 of the 512^3 recipe)."""
    from porespy_b200.edt import edt_sq_u32
    im = oc.blobs([256, 256, 256], porosity=0.6, blobiness=2, seed=0)
    assert_same(edt_sq_u32(im), oc.edt_sq(im), "blobs 256^3")


# ------------------------------------------------------------------- local thickness
LT_CASES = [((40, 36, 44), 12, 0.6), ((48, 52), 9, 0.6), ((30, 30, 30), [5, 3.5, 2, 1.2], 0.6),
            ((33, 31, 29), np.arange(8, 1, -1), 0.5), ((61, 67, 131), 25, 0.7), ((1, 90, 140), 10, 0.6),
            ((150, 3, 40), 6, 0.8), ((64, 200), 25, 0.75), ((257,), 5, 0.9),
            # nx % 16 == 0: the streaming three-kernel path
            ((45, 70, 64), 25, 0.6), ((130, 140, 160), 25, 0.65), ((300, 256), 25, 0.7),
            ((300, 1, 32), 8, 0.8), ((20, 300, 16), 12, 0.7), ((256,), 6, 0.9), ((9, 520, 144), 30, 0.8),
            # nx % 32 == 0: bit-parallel kernels (incl. rows longer than 1024 voxels: 30-word segments)
            ((40, 50, 96), 20, 0.6), ((3, 40, 1120), 14, 0.7), ((5, 4, 2048), [9, 4, 2.5, 1], 0.8),
            ((70, 2080), 16, 0.7),
            # even word counts above 32: two words per lane (64-word rows, 60-word segments with halo lanes)
            ((6, 24, 2112), [9, 4, 2.5, 1], 0.8), ((2, 30, 4160), 14, 0.7), ((40, 2048), 12, 0.7),
            # rows of exactly 1024 / 2048 / 4096 voxels: four words per lane (interior and border rows)
            ((40, 50, 1024), 20, 0.6), ((6, 21, 1024), [9, 4, 2.5, 1], 0.8), ((3, 12, 4096), 10, 0.7),
            ((37, 1024), 14, 0.7)]


@pytest.mark.parametrize("algo", ["fast", "nobit", "allbit", "allbit1", "generic"])
@pytest.mark.parametrize("shape,sizes,por", LT_CASES)
def test_local_thickness_vs_oracle(psb, algo, shape, sizes, por):
    set_algo(psb, algo)
    try:
        if len(shape) == 1:
            im = rand_image(shape, por, 5)
            im[::17] = False
        else:
            im = oc.blobs(list(shape), porosity=por, blobiness=1.5, seed=7)
        got = psb.filters.local_thickness(im, sizes=sizes, mode="dt")
        want = oc.local_thickness(im, sizes=sizes, mode="dt")
        assert got.dtype == np.float64
        assert_same(got, want, f"local_thickness {algo} {shape}")
    finally:
        set_algo(psb, "fast")


@pytest.mark.parametrize("algo", ["fast", "generic"])
def test_local_thickness_goldens(psb, golden, algo):
    set_algo(psb, algo)
    try:
        g = golden.blobs100
        im = g.mask("im")
        lt = psb.filters.local_thickness(im, mode="dt")
        assert_same(lt, g.rmap("lt_dt_25"), "lt_dt_25")
        np.testing.assert_almost_equal(lt.max(), psb.edt(im).max(), decimal=6)        # TF:266-272
        assert_same(psb.filters.local_thickness(im, mode="hybrid"), g.rmap("lt_dt_25"), "hybrid")
        assert_same(psb.filters.local_thickness(im[:, :, 50]), g.rmap("lt_2d_25"), "lt_2d_25")
        assert_same(psb.filters.local_thickness(im, sizes=[6, 4.5, 3, 2, 1]), g.rmap("lt_list_sizes"),
                    "lt_list_sizes")
        assert_same(psb.filters.porosimetry(im[:, :, 50], sizes=9, access_limited=False),
                    g.rmap("poro_2d_noaccess"), "poro_2d_noaccess")
        m = golden.misc2d
        lt = psb.filters.local_thickness(m.mask("rsa"), sizes=[20, 10])
        assert np.all(np.unique(lt) == [0, 10, 20])                                   # TF:274-279
        assert_same(lt, m.rmap("lt_rsa"), "lt_rsa")
        drn = m.mask("drn")
        lt = psb.filters.local_thickness(drn)
        assert_same(lt, m.rmap("lt_drn"), "lt_drn")
        assert (lt > 25).sum() / drn.sum() == 0.34427115020497745                    # test_drainage.py:49
    finally:
        set_algo(psb, "fast")


def test_input_not_mutated_and_quirks(psb, golden):
    im = golden.blobs100.mask("im")[:50, :50, :50].copy()
    keep = im.copy()
    a = psb.filters.local_thickness(im, sizes=np.int64(2))      # numpy int scalar: ONE radius (App. B1)
    assert np.array_equal(im, keep) and a is not im
    assert set(np.unique(a)) <= {0.0, 2.0}
    assert_same(a, oc.local_thickness(im, sizes=np.int64(2), mode="dt"), "np.int64 sizes")
    b = psb.filters.local_thickness(im[None, :, :, None, :1].repeat(1, axis=0), sizes=5)   # squeeze
    assert b.shape == (50, 50)
    with pytest.raises(Exception, match="Unrecognized mode"):
        psb.filters.local_thickness(im, mode="nope")
    assert np.all(psb.filters.local_thickness(np.zeros((8, 9, 10), bool)) == 0)      # all background
    full = psb.filters.local_thickness(np.ones((6, 7, 8), bool), sizes=[3, 2])
    assert np.all(full == 3.0)                                                       # no background (N8)
    neg = psb.filters.local_thickness(im, sizes=[2.5, 0, -1])
    assert_same(neg, oc.local_thickness(im, sizes=[2.5, 0, -1], mode="dt"), "non-positive radii")


def test_many_thresholds_grouping(psb):
    """more than 253 distinct thresholds -> several device calls chained through idx KEEP marks"""
    im = oc.blobs([70, 60, 50], porosity=0.75, blobiness=0.6, seed=3)
    dmax = float(oc.edt(im).max())
    sizes = np.linspace(1.0, dmax, 900)
    got = psb.filters.local_thickness(im, sizes=sizes)
    want = oc.local_thickness(im, sizes=sizes, mode="dt")
    assert_same(got, want, "900 radii")


# ------------------------------------------------------------------------ porosimetry
@pytest.mark.parametrize("algo", ["fast", "generic"])
def test_porosimetry_goldens(psb, golden, algo):
    set_algo(psb, algo)
    try:
        g = golden.blobs100
        im = g.mask("im")
        mip = psb.filters.porosimetry(im=im, sizes=10)
        ans = np.array([0.00000000, 1.00000000, 1.37871571, 1.61887041, 1.90085700, 2.23196205,
                        2.62074139, 3.07724114, 3.61325732])                          # TF:39-41
        assert np.allclose(np.unique(mip), ans)
        assert_same(mip, g.rmap("poro_hybrid_sizes10"), "poro sizes=10")
        sizes = np.arange(25, 1, -1)
        for mode in ("hybrid", "dt", "mio"):                                         # TF:44-51
            assert_same(psb.filters.porosimetry(im, sizes=sizes, mode=mode), g.rmap("poro_dt_arange_3d"),
                        f"3d {mode}")
        assert_same(psb.filters.porosimetry(im[:, :, 50], sizes=sizes, mode="dt"),
                    g.rmap("poro_dt_arange_2d"), "2d")                               # TF:27-34
        s = np.logspace(0.01, 0.6, 5)
        mip = psb.filters.porosimetry(im=im, sizes=s)
        assert np.allclose(np.unique(mip)[1:], s)                                    # TF:53-56
        assert_same(mip, g.rmap("poro_logsizes"), "logsizes")
        inlets = np.zeros_like(im)
        inlets[0, ...] = True
        assert_same(psb.filters.porosimetry(im, sizes=12, inlets=inlets, mode="dt"),
                    g.rmap("poro_inlet0_dt_12"), "inlet0")
    finally:
        set_algo(psb, "fast")


@pytest.mark.parametrize("shape,por", [((44, 40, 36), 0.55), ((90, 110), 0.6), ((31, 64, 129), 0.5),
                                       ((44, 40, 48), 0.55), ((90, 112), 0.6), ((131, 64, 128), 0.5),
                                       ((36, 44, 64), 0.6), ((90, 96), 0.6)])
@pytest.mark.parametrize("algo", ["fast", "nobit"])
def test_porosimetry_vs_oracle(psb, algo, shape, por):
    set_algo(psb, algo)
    try:
        _porosimetry_vs_oracle(psb, shape, por)
    finally:
        set_algo(psb, "fast")


def _porosimetry_vs_oracle(psb, shape, por):
    im = oc.blobs(list(shape), porosity=por, blobiness=1.5, seed=11)
    assert_same(psb.filters.porosimetry(im, sizes=10), oc.porosimetry(im, sizes=10, mode="dt"), "faces")
    inlets = np.zeros(shape, dtype=int)
    inlets[..., 0] = 1
    assert_same(psb.filters.porosimetry(im, sizes=8, inlets=inlets),
                oc.porosimetry(im, sizes=8, inlets=inlets, mode="dt"), "x-face inlet")
    inlets = np.zeros(shape, dtype=bool)
    inlets[tuple(s // 2 for s in shape)] = True                                     # a single (maybe solid) voxel
    assert_same(psb.filters.porosimetry(im, sizes=6, inlets=inlets),
                oc.porosimetry(im, sizes=6, inlets=inlets, mode="dt"), "point inlet")
    with pytest.raises(Exception, match="inlets not valid"):
        psb.filters.porosimetry(im, inlets=np.zeros(shape))


@pytest.mark.parametrize("shape,por", [((20, 24, 300), 0.55), ((40, 517), 0.6), ((9, 33, 131), 0.5)])
@pytest.mark.parametrize("records", [True, False])
def test_porosimetry_flood_paths(psb, shape, por, records):
    """Access-limited loop on the row-rooted forest + link records (default) and on the per-voxel job lists:
    rows of several 128-voxel segments, partial segments, rows that are not a multiple of 4 voxels."""
    from porespy_b200 import _lib
    ctx = _lib.context()
    ctx.set_uf_records(records)
    try:
        _porosimetry_vs_oracle(psb, shape, por)
    finally:
        ctx.set_uf_records(True)


@pytest.mark.parametrize("records", [True, False])
def test_flood_random_shapes(psb, records):
    """psb200_flood, 6/26 (4/8) connectivity, near the percolation threshold where components are large and
    tortuous; shapes with partial segments and odd row lengths; inlets inside the solid as well."""
    from porespy_b200 import _lib
    ctx = _lib.context()
    ctx.set_uf_records(records)
    try:
        rng = np.random.default_rng(5)
        for shape, p in (((12, 14, 130), 0.33), ((7, 9, 261), 0.3), ((60, 259), 0.6), ((3, 5, 7), 0.5),
                         ((1, 1, 400), 0.9), ((16, 1, 129), 0.7)):
            im = rng.random(shape) < p
            inl = np.zeros_like(im)
            inl[..., 0] = True
            inl |= rng.random(shape) < 0.002
            for strel in (None, oc._cross(len(shape))):
                assert_same(psb.filters.trim_disconnected_blobs(im, inl, strel=strel),
                            oc.trim_disconnected_blobs(im, inl, strel=strel), f"{shape} {strel is None}")
    finally:
        ctx.set_uf_records(True)


def test_trim_disconnected_blobs(psb, golden):
    g = golden.trim
    im, inl = g.mask("im2d"), g.mask("inlets2d")
    assert_same(psb.filters.trim_disconnected_blobs(im, inl), g.mask("out8"), "8-conn")   # TF:201-210
    assert_same(psb.filters.trim_disconnected_blobs(im, inl, strel=oc._cross(2)), g.mask("out4"), "4-conn")
    im, inl = g.mask("im3d"), g.mask("inlets3d")
    assert_same(psb.filters.trim_disconnected_blobs(im, inl), g.mask("out26"), "26-conn")
    assert_same(psb.filters.trim_disconnected_blobs(im, inl, strel=oc._cross(3)), g.mask("out6"), "6-conn")
    rng = np.random.default_rng(0)
    im = rng.random((40, 50, 60)) < 0.35                                             # near percolation
    inl = np.zeros_like(im)
    inl[:, :, -1] = True
    for strel in (None, oc._cross(3)):
        assert_same(psb.filters.trim_disconnected_blobs(im, inl, strel=strel),
                    oc.trim_disconnected_blobs(im, inl, strel=strel), "random")


def test_config0_golden(psb, golden):
    """BASELINE config 0: 200^3 blobs(0.6, 2), sizes=25 -- reference output pinned by SHA-256."""
    g = golden.config0
    im = oc.blobs([200, 200, 200], porosity=0.6, blobiness=2, seed=0)
    assert sha(im) == str(g.raw("im_sha"))
    from porespy_b200.edt import edt_sq_u32
    assert sha(edt_sq_u32(im)) == str(g.raw("d2_sha"))
    assert sha(psb.edt(im)) == str(g.raw("dt_sha"))
    lt = psb.filters.local_thickness(im, sizes=25)
    v, c = np.unique(lt, return_counts=True)
    assert np.array_equal(v, g.raw("lt_values")) and np.array_equal(c, g.raw("lt_counts"))
    assert sha(lt) == str(g.raw("lt_sha"))
    assert sha(psb.filters.porosimetry(im, sizes=25)) == str(g.raw("poro_faces_sha"))
    inl = np.zeros_like(im)
    inl[0, ...] = True
    assert sha(psb.filters.porosimetry(im, sizes=25, inlets=inl)) == str(g.raw("poro_inlet0_sha"))


def test_sharded_driver_single_rank(psb):
    """The z-slab driver with one rank runs every step-level C-ABI entry point it uses
    (edt_xy / edt_z / lt_classify / lt_xy / lt_z / lt_pack / lt_bitball / lt_wmask)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "sharded_gpu_check.py")], cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_GPU_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_sharded_two_gpus(psb):
    """NCCL all-to-all + halo exchange on 2 GPUs (skipped on a 1-GPU box)."""
    import subprocess, sys, os, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(root, "tests", "sharded_gpu_check.py")], cwd=root,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SHARDED_GPU_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("records", [True, False])
@pytest.mark.parametrize("nslabs", [2, 3])
def test_flood_face_exchange_one_gpu(psb, nslabs, records):
    """psb200_uf_begin / activate / face / inject / mark (job lists) and psb200_uf_*_records (link records + join
    times, one resolve pass): the slab-local union-finds of a volume cut into z-slabs, coupled only through the
    face flags (what ShardedVolume._flood_exchange sends between ranks), must reproduce
    trim_disconnected_blobs on the whole volume for every radius."""
    import torch
    from porespy_b200 import _lib, _host
    from porespy_b200.sharded import CudaBackend, split_counts
    be = CudaBackend(_lib.context())
    be.uf_records = records           # link records (default) / per-voxel job lists
    shape = (60, 48, 200) if records else (60, 48, 64)     # 200: a full and a partial 128-voxel segment per row
    nz, ny, nx = shape
    im = oc.blobs(list(shape), porosity=0.55, blobiness=1.5, seed=11)
    d2 = oc.edt_sq(im)
    radii = np.array([6.0, 4.5, 3.0, 2.0, 1.0])
    T, R = _host.effective_thresholds(radii, int(d2.max()))
    counts = split_counts(nz, nslabs)
    starts = [sum(counts[:i]) for i in range(nslabs)]
    for inl_kind in ("faces", "z0"):
        mask = None
        if inl_kind == "z0":
            mask = np.zeros(shape, dtype=bool)
            mask[0] = True
        inlets_full = mask if mask is not None else oc.border_faces(shape)
        sts = []
        for s0, c in zip(starts, counts):
            d2s = torch.from_numpy(d2[s0:s0 + c].astype(np.uint32).view(np.int32).copy()).cuda().reshape(-1)
            cls = be.classify(d2s, T)
            inl = None if mask is None else be.to_u8(mask[s0:s0 + c])
            sts.append(be.uf_begin(cls, inl, (c, ny, nx), s0, nz))
        for k, Tk in enumerate(T):
            for st in sts:
                be.uf_activate(st, k - 1, k)
            sweeps = 0
            while True:
                sweeps += 1
                faces = [(be.uf_face(st, k, 0), be.uf_face(st, k, st.shape[0] - 1)) for st in sts]
                for i, st in enumerate(sts):
                    if i > 0:
                        be.uf_inject(st, k, 0, faces[i - 1][1])
                    if i < nslabs - 1:
                        be.uf_inject(st, k, st.shape[0] - 1, faces[i + 1][0])
                if not max(be.uf_changed(st) for st in sts):
                    break
                assert sweeps < 50
            for st in sts:
                be.uf_settle(st, k)
        for st in sts:
            be.uf_resolve(st)
        for k, Tk in enumerate(T):
            seeds = d2 >= Tk
            want = oc.trim_disconnected_blobs(seeds, inlets_full, strel=oc._cross(3))
            got = np.concatenate([(st.rcls.cpu().numpy().reshape(st.shape) <= k) for st in sts], axis=0)
            assert_same(got, want, f"reached seeds, inlets={inl_kind}, k={k}, T={Tk}")


def test_slab_pipelined_numpy_result(psb):
    """numpy -> numpy local_thickness with the radius loop run slab by slab and the host epilogue of each slab
    overlapping the next slab's kernels (filters.SLAB_PIPELINE) equals the one-piece path and the oracle."""
    from porespy_b200 import filters as F
    im = oc.blobs([120, 72, 96], porosity=0.6, blobiness=1.5, seed=4)
    saved = dict(F.SLAB_PIPELINE)
    try:
        F.SLAB_PIPELINE.update(enabled=False)
        whole = psb.filters.local_thickness(im, sizes=12)
        for slabs in (2, 3, 4):
            F.SLAB_PIPELINE.update(enabled=True, min_voxels=1, slabs=slabs)
            assert F._slab_plan((120, 72, 96), [50, 20, 5]) is not None
            got = psb.filters.local_thickness(im, sizes=12)
            assert got.dtype == np.float64 and got.shape == im.shape
            assert_same(got, whole, f"{slabs} slabs vs one piece")
        assert_same(whole, oc.local_thickness_c(im, sizes=12), "vs oracle")
    finally:
        F.SLAB_PIPELINE.clear()
        F.SLAB_PIPELINE.update(saved)


def test_shard_halo_helpers(psb):
    """psb200_mask_pack_u8 / psb200_mask_unpack_u8 (EDT halo planes as bits) and psb200_lt_halo_cone (one plane of cone
    values per face instead of W planes of reach bytes) against numpy."""
    import torch
    from porespy_b200 import _lib
    from porespy_b200.sharded import CudaBackend
    from tests.cpu_backend import CpuBackend
    be, cb = CudaBackend(_lib.context()), CpuBackend()
    rng = np.random.default_rng(3)
    for n in (64, 8 * 1000 + 8, 4096 * 33):
        src = (rng.random(n) < 0.4).astype(np.uint8) * rng.integers(1, 255, n).astype(np.uint8)
        bits = be.mask_pack(torch.from_numpy(src).cuda())
        assert np.array_equal(bits.cpu().numpy(), np.packbits(src != 0, bitorder="little"))
        back = torch.empty(n, dtype=torch.uint8, device="cuda")
        be.mask_unpack(bits, back)
        assert np.array_equal(back.cpu().numpy(), (src != 0).astype(np.uint8))
    for shape in ((9, 12, 16), (5, 7, 9), (40, 8, 12)):
        reach = rng.integers(0, 12, shape).astype(np.uint8) * (rng.random(shape) < 0.3)
        t = torch.from_numpy(reach.reshape(-1).copy())
        for depth in (1, 4, 60):
            for side in (0, 1):
                got = be.halo_cone(t.cuda(), shape, depth, side).cpu().numpy()
                assert np.array_equal(got, cb.halo_cone(t, shape, depth, side).numpy()), (shape, depth, side)


@pytest.mark.parametrize("permille,threads", [(0, 0), (500, 3), (1000, 0), (730, 16)])
def test_host_epilogue_split(psb, permille, threads):
    """psb200_expand_idx_f64_to_host: any split between host-thread widening of index bytes and
    device-side widening gives the same float64 map (odd length: partial chunks, unaligned tail)."""
    import torch
    from porespy_b200 import _device as dev
    from porespy_b200 import _lib
    ctx = _lib.context()
    n = 3 * (1 << 24) + 12345
    g = torch.Generator(device="cuda")
    g.manual_seed(permille)
    idx = torch.randint(0, 40, (n,), generator=g, device="cuda", dtype=torch.uint8)
    lut = np.concatenate([[0.0], np.sort(np.random.default_rng(1).uniform(1, 50, 39))[::-1]])
    out = dev.expand_idx_to_host(ctx, idx, lut, (n,), cpu_permille=permille, nthreads=threads, chunk=1 << 22)
    assert out.dtype == np.float64 and out.shape == (n,)
    assert np.array_equal(out, lut[idx.cpu().numpy()])


def test_local_thickness_large_numpy_result(psb):
    """Public API on a volume whose float64 result goes through the host-result epilogue (>= 2^28 bytes)."""
    _large_numpy_result(psb)


def _large_numpy_result(psb):
    im = oc.blobs([320, 320, 352], porosity=0.6, blobiness=2, seed=3)
    a = psb.filters.local_thickness(im, sizes=12)
    b = psb.filters.local_thickness(im, sizes=12)            # second call reuses the recycled pinned buffer
    want = oc.local_thickness(im, sizes=12, mode="dt")
    assert_same(a, want, "large numpy result")
    assert_same(b, want, "large numpy result, second call")


@pytest.mark.parametrize("dtype", [np.bool_, np.uint8])
def test_upload_mask_bit_packed(psb, dtype):
    """psb200_upload_mask_u8: a host volume uploaded as bits arrives as (byte != 0) bytes, for sizes
    that are not multiples of the chunk, of 16 or of 8, and for uint8 values other than 0 / 1."""
    from porespy_b200 import _device as dev
    from porespy_b200 import _lib
    ctx = _lib.context()
    rng = np.random.default_rng(5)
    for n in (3 * (1 << 23) + 8 * 1001 + 5, (1 << 23), 77):
        a = rng.integers(0, 4, n).astype(np.uint8)
        a[rng.random(n) < 0.5] = 0
        if dtype == np.bool_:
            a = a != 0
        got = dev.upload_mask(ctx, a.view(np.uint8)).cpu().numpy()
        assert np.array_equal(got, (a != 0).astype(np.uint8)), n
    big = oc.blobs([160, 640, 704], porosity=0.6, blobiness=2, seed=1)      # 72 MB: takes the packed path
    assert_same(psb.edt(big), oc.edt(big), "edt through the bit-packed upload")


def test_flood_users(psb, golden):
    """The other users of the flood kernel against the reference-generated goldens
    and the oracle: find_disconnected_voxels (incl. surface=True and its label-0 quirk), fill_blind_pores,
    trim_floating_solid, trim_nonpercolating_paths."""
    f = psb.filters
    g = golden.flood_users
    im = g.mask("im")
    sl = im[:, :, 0]
    for got, key in ((f.find_disconnected_voxels(sl), "h2d8"), (f.find_disconnected_voxels(sl, conn=4), "h2d4"),
                     (f.find_disconnected_voxels(im), "h26"), (f.find_disconnected_voxels(im, conn=6), "h6"),
                     (f.find_disconnected_voxels(im, surface=True), "h26_surface"),
                     (f.find_disconnected_voxels(im, conn=6, surface=True), "h6_surface"),
                     (f.find_disconnected_voxels(sl, conn=4, surface=True), "h2d4_surface"),
                     (f.find_disconnected_voxels(g.mask("cap"), conn=6, surface=True), "cap_surface"),
                     (f.fill_blind_pores(im), "fill_blind"), (f.fill_blind_pores(im, conn=6, surface=True), "fill_blind6s"),
                     (f.trim_floating_solid(im), "trim_solid"), (f.trim_floating_solid(im, conn=6), "trim_solid6")):
        assert_same(got, g.mask(key), key)
    assert int(f.find_disconnected_voxels(im).sum()) == 55 and int(f.find_disconnected_voxels(im, conn=6).sum()) == 202
    with pytest.raises(Exception, match="conn is not valid"):
        f.find_disconnected_voxels(im, conn=5)
    for name, key, axes in (("np2d_im", "np2d_ax", (0, 1)), ("np3d_im", "np3d_ax", (0, 1, 2))):
        b = g.mask(name)
        for ax in axes:
            inl, outl = np.zeros_like(b), np.zeros_like(b)
            inl[(slice(None),) * ax + (0,)] = True
            outl[(slice(None),) * ax + (-1,)] = True
            assert_same(f.trim_nonpercolating_paths(im=b, inlets=inl, outlets=outl), g.mask(f"{key}{ax}"), f"{key}{ax}")
    b3 = g.mask("np3d_im")
    inl, outl = np.zeros_like(b3), np.zeros_like(b3)
    inl[0], outl[-1] = True, True
    assert_same(f.trim_nonpercolating_paths(b3, inl, outl, strel=np.ones((3, 3, 3))), g.mask("np3d_ax0_cube"), "cube strel")
    b = g.mask("np2d_none_im")
    inl, outl = np.zeros_like(b), np.zeros_like(b)
    inl[:, 0], outl[:, -1] = True, True
    assert f.trim_nonpercolating_paths(b, inl, outl).sum() == 0
    # random media against the oracle, odd shapes
    for shape, p, conn in (((37, 41, 29), 0.45, 6), ((37, 41, 29), 0.3, None), ((90, 77), 0.55, 4), ((64, 50), 0.5, 8)):
        r = rand_image(shape, p, 7)
        for surface in (False, True):
            assert_same(f.find_disconnected_voxels(r, conn=conn, surface=surface),
                        oc.find_disconnected_voxels(r, conn=conn, surface=surface), f"random {shape} {conn} {surface}")


@pytest.mark.parametrize("one_flood", [True, False])
def test_find_trapped_regions(psb, golden, one_flood):
    """find_trapped_regions (one flood with join times for all bins / one device flood per bin) against the
    reference-generated goldens, and against the numpy restatement on a larger random sequence."""
    from tests.test_oracle import _trapped_cases
    from porespy_b200 import filters as F
    saved = F.ONE_FLOOD_TRAPPED
    F.ONE_FLOOD_TRAPPED = one_flood
    try:
        g = golden.trapped
        cases, seq2, outl = _trapped_cases(g)
        for kw, key in cases:
            got = psb.filters.find_trapped_regions(**kw)
            assert got.dtype == np.bool_
            assert_same(got, g.mask(key), key)
        s = psb.filters.find_trapped_regions(seq2, outlets=outl, bins=None, return_mask=False)
        assert np.array_equal(s, g.raw("s2d_outlet_seq"))
        assert np.array_equal(seq2, g.raw("seq2d"))                 # input not modified
        rng = np.random.default_rng(9)
        for shape in ((40, 36, 140), (90, 131)):
            im = oc.blobs(list(shape), porosity=0.6, blobiness=1.5, seed=3)
            seq = (rng.integers(1, 40, shape) * im).astype(np.int64)
            outl = np.zeros(shape, dtype=bool)
            outl[-1] = True
            for bins in (25, 7, None):
                assert_same(psb.filters.find_trapped_regions(seq, outlets=outl, bins=bins),
                            oc.find_trapped_regions(seq, outlets=outl, bins=bins), f"{shape} bins={bins}")
    finally:
        F.ONE_FLOOD_TRAPPED = saved


# -------------------------------------------------------------------- round-2 additions
def test_back_to_back_uploads_keep_both_masks(psb):
    """Two different >= 64 MiB host volumes uploaded through the shared page-locked staging buffer one
    right after the other (trim_disconnected_blobs: fg then inlets) must both arrive intact."""
    import torch
    from porespy_b200 import _device as dev
    from porespy_b200 import _lib
    ctx = _lib.context()
    n = (1 << 26) + 4099
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2, n, dtype=np.uint8)
    b = rng.integers(0, 2, n, dtype=np.uint8)
    for _ in range(3):
        ta = dev.to_device_u8(a, ctx)
        tb = dev.to_device_u8(b, ctx)
        tc = dev.to_device_u8(a ^ b, ctx)
        torch.cuda.synchronize()
        assert np.array_equal(ta.cpu().numpy(), a)
        assert np.array_equal(tb.cpu().numpy(), b)
        assert np.array_equal(tc.cpu().numpy(), a ^ b)


def test_divs_tensor_inlets_and_signed_images(psb):
    import torch
    im = oc.blobs([48, 40, 64], porosity=0.6, blobiness=1.5, seed=21)
    want = oc.local_thickness(im, sizes=9, mode="dt")
    assert_same(psb.filters.local_thickness(im, sizes=9, divs=2), want, "divs=2")           # F:947-952 signature
    assert_same(psb.filters.porosimetry(im, sizes=9, divs=[2, 1, 2], access_limited=False), want, "divs list")
    # negative voxels are background (F:1126 `im > 0`), for numpy and for device tensors
    signed = im.astype(np.int16)
    signed[~im] = -1
    assert_same(psb.filters.local_thickness(signed, sizes=9), want, "signed image")
    got = psb.filters.local_thickness(torch.from_numpy(signed.astype(np.float32)).cuda(), sizes=9)
    assert_same(got.cpu().numpy(), want, "signed float tensor image")
    # inlets as a device tensor
    inl = np.zeros(im.shape, dtype=bool)
    inl[:, 0, :] = True
    wantp = oc.porosimetry(im, sizes=9, inlets=inl, mode="dt")
    assert_same(psb.filters.porosimetry(im, sizes=9, inlets=torch.from_numpy(inl).cuda()), wantp, "tensor inlets")
    with pytest.raises(Exception, match="inlets not valid"):
        psb.filters.porosimetry(im, inlets=torch.zeros(im.shape, dtype=torch.uint8).cuda())


def test_patch_install_rebinds_every_module(psb):
    """install() after the (stand-in) PoreSpy modules did `from edt import edt`: every porespy.* module that
    holds the original function is rebound, uninstall() restores them."""
    import sys, types
    from porespy_b200 import patch

    def original_edt(data, **kw):
        raise AssertionError("the original edt must not be called after install()")

    fake_edt = types.ModuleType("edt")
    fake_edt.edt = original_edt
    mods = {"edt": fake_edt}
    for name in ("porespy", "porespy.filters", "porespy.filters._funcs", "porespy.filters._snows",
                 "porespy.tools", "porespy.tools._funcs", "porespy.simulations", "porespy.simulations._drainage"):
        mods[name] = types.ModuleType(name)
    mods["porespy"].filters = mods["porespy.filters"]
    mods["porespy.filters"]._funcs = mods["porespy.filters._funcs"]
    for name in ("porespy.filters._funcs", "porespy.filters._snows", "porespy.tools._funcs",
                 "porespy.simulations._drainage"):
        mods[name].edt = original_edt                                  # `from edt import edt`
    mods["porespy.filters._funcs"].porosimetry = lambda *a, **k: None
    mods["porespy.filters"].porosimetry = mods["porespy.filters._funcs"].porosimetry

    def ps_ball(r):                                                    # T:1149-1155, through the module's `edt`
        probe = np.ones([2 * int(np.ceil(r)) + 1] * 3, dtype=bool)
        probe[(int(np.ceil(r)),) * 3] = False
        return mods["porespy.tools._funcs"].edt(probe) < r
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        patch.install()
        assert sys.modules["edt"].edt is psb.edt
        for name in ("porespy.filters._funcs", "porespy.filters._snows", "porespy.tools._funcs",
                     "porespy.simulations._drainage"):
            assert mods[name].edt is psb.edt, name
        assert mods["porespy.filters"].porosimetry is psb.filters.porosimetry
        assert int(ps_ball(3).sum()) == 93                             # test_tools.py:309-316
        patch.uninstall()
        assert sys.modules["edt"] is fake_edt
        assert mods["porespy.filters._snows"].edt is original_edt
    finally:
        patch.uninstall()
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# ---------------------------------------------------------------- blobs generator (8(f) rank 4b)
@pytest.mark.parametrize("shape,por,blob", [((64, 64, 64), 0.6, 1), ((50, 70, 90), 0.5, 2), ((120, 130), 0.55, 1),
                                            ((33, 40, 200), 0.7, [1, 2, 3]), ((100, 100, 100), 0.499, 2)])
def test_blobs_generator_vs_host(psb, shape, por, blob):
    """numpy-seeded noise: the device generator reproduces the host blobs() (scipy gaussian_filter + numpy
    norm_to_uniform); only voxels within rounding of the threshold may differ."""
    want = oc.blobs(list(shape), porosity=por, blobiness=blob, seed=4)
    got = psb.generators.blobs(list(shape), porosity=por, blobiness=blob, seed=4)
    assert got.dtype == np.bool_ and got.shape == tuple(shape)
    assert (got != want).sum() <= max(1, got.size // 200000), f"{(got != want).sum()} voxels differ"
    field = psb.generators.blobs(list(shape), porosity=None, blobiness=blob, seed=4)
    wantf = oc.blobs(list(shape), porosity=None, blobiness=blob, seed=4)
    assert field.dtype == np.float64
    np.testing.assert_allclose(field, wantf, rtol=0, atol=1e-12)


def test_blobs_generator_philox(psb):
    a = psb.generators.blobs([96, 80, 128], porosity=0.6, blobiness=2, seed=7, rng="philox")
    b = psb.generators.blobs([96, 80, 128], porosity=0.6, blobiness=2, seed=7, rng="philox")
    c = psb.generators.blobs([96, 80, 128], porosity=0.6, blobiness=2, seed=8, rng="philox")
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean() - 0.6) < 0.02
