"""CPU suite: host-side logic of the product (thresholds, radii, inlets), the numpy model of the
bounded pipeline against the oracle, and the C-ABI surface (symbols only, no compute)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import cpu as oc
from porespy_b200 import _host as host
from porespy_b200 import _lib
from tests import model_fast as mf


def test_threshold_examples():
    # r = 4.2426405 -> T = 18, r = 3.9947 -> T = 16
    assert host.threshold_of(np.float32(4.2426405)) == 18
    assert host.threshold_of(np.float32(3.9947)) == 16
    assert host.threshold_of(np.int64(20)) == 400
    assert host.threshold_of(3.0) == 9
    assert host.threshold_of(1.0) == 1
    assert host.threshold_of(0.0) == 0 and host.threshold_of(-2.5) == 0
    assert host.threshold_of(np.nan) is None and host.threshold_of(np.inf) is None


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_threshold_matches_float_comparison(dtype):
    rng = np.random.default_rng(3)
    d2 = np.arange(0, 5000, dtype=np.uint32)
    dt = np.sqrt(d2.astype(np.float32))
    radii = np.concatenate([rng.uniform(0.5, 70, 200), np.sqrt(np.arange(1, 60))]).astype(dtype)
    for r in radii:
        T = host.threshold_of(r)
        assert np.array_equal(dt >= r, d2 >= T), r
        assert np.array_equal(dt < r, d2 < T), r


def test_reference_sizes_dtype_and_values():
    r = host.reference_sizes(25, 86)
    assert r.dtype == np.float32 and len(r) == 25 and r[-1] == 1.0     # NEP 50 trap (note N1)
    T, R = host.effective_thresholds(r, 86)
    assert list(T) == [86, 72, 60, 50, 41, 34, 29, 24, 20, 17, 14, 12, 10, 8, 7, 6, 5, 4, 3, 2, 1]
    assert R[0] == float(r[0]) and R[-1] == 1.0
    r = host.reference_sizes([20, 10], 1000)
    assert r.dtype.kind == "i" and list(r) == [20, 10]
    r = host.reference_sizes(np.int64(7), 1000)           # numpy integer scalar: ONE radius
    assert list(r) == [7]
    T, R = host.effective_thresholds(host.reference_sizes([3, 2, 0, -1, 50], 30), 30)
    assert list(T) == [9, 4] and list(R) == [3.0, 2.0]


def test_all_background_yields_no_thresholds():
    r = host.reference_sizes(10, 0)
    T, R = host.effective_thresholds(r, 0)
    assert len(T) == 0


def test_normalise_inlets():
    m = np.zeros((4, 5), bool)
    m[0] = True
    assert np.array_equal(host.normalise_inlets(m.astype(int), (4, 5)), m)
    with pytest.raises(Exception, match="inlets not valid"):
        host.normalise_inlets(np.zeros((4, 5)), (4, 5))
    with pytest.raises(Exception, match="inlets not valid"):
        host.normalise_inlets(m, (5, 4))
    # tuple inlets follow the reference's own idiom literally (np.copy of the tuple becomes a
    # 2-D integer array that indexes axis 0, F:1252-1255) -- quirk preserved, not "fixed"
    where = (np.array([0, 0, 1]), np.array([1, 2, 3]))
    ref = np.zeros((4, 5), bool)
    ref[np.copy(where)] = True
    assert np.array_equal(host.normalise_inlets(where, (4, 5)), ref)


def _idx_to_map(idx, R):
    return np.concatenate([[0.0], R])[idx]


@pytest.mark.parametrize("shape,sizes", [((40, 36, 44), 12), ((48, 52), 9), ((30, 30, 30), [5, 3.5, 2, 1.2]),
                                         ((33, 31, 29), np.arange(8, 1, -1))])
def test_model_matches_oracle_local_thickness(shape, sizes):
    im = oc.blobs(list(shape), porosity=0.6, blobiness=1.5, seed=7)
    d2 = oc.edt_sq(im)
    radii = host.reference_sizes(sizes, int(d2.max()))
    T, R = host.effective_thresholds(radii, int(d2.max()))
    got = _idx_to_map(mf.lt_idx(d2, T).reshape(shape), R)
    want = oc.local_thickness(im, sizes=sizes, mode="dt")
    assert np.array_equal(got, want)


def test_model_matches_oracle_porosimetry():
    shape = (36, 40, 32)
    im = oc.blobs(list(shape), porosity=0.55, blobiness=1.5, seed=11)
    d2 = oc.edt_sq(im)
    inlets = np.zeros(shape, bool)
    inlets[0] = True
    radii = host.reference_sizes(10, int(d2.max()))
    T, R = host.effective_thresholds(radii, int(d2.max()))

    def trim(k, seeds):
        return oc.trim_disconnected_blobs(seeds, inlets, strel=oc._cross(3))
    got = _idx_to_map(mf.lt_idx(d2, T, seeds_filter=trim).reshape(shape), R)
    want = oc.porosimetry(im, sizes=10, inlets=inlets, mode="dt")
    assert np.array_equal(got, want)


def test_model_golden(golden):
    g = golden.blobs100
    im = g.mask("im")[:60, :60, :60]
    d2 = oc.edt_sq(im)
    radii = host.reference_sizes(25, int(d2.max()))
    T, R = host.effective_thresholds(radii, int(d2.max()))
    got = _idx_to_map(mf.lt_idx(d2, T).reshape(im.shape), R)
    assert np.array_equal(got, oc.local_thickness(im, mode="dt"))


# ------------------------------------------------------------------------------ C ABI
def _declared_symbols():
    text = open(_lib.HEADER).read()
    return sorted(set(re.findall(r"\b(psb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    lib.psb200_version.restype = ctypes.c_int
    assert lib.psb200_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import porespy_b200 as psb
    with pytest.raises(Exception):
        psb.filters.local_thickness(np.ones((8, 8, 8), bool))
    with pytest.raises(Exception):
        psb.edt(np.ones((8, 8), bool))


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert "oracle" not in src.replace("no CPU fallback", "").lower() or fn == "_lib.py", fn


def test_edt_rejects_what_it_does_not_implement():
    """Options / inputs of the upstream `edt` that PoreSpy never uses must not silently differ."""
    import porespy_b200 as psb
    im = np.ones((4, 5), bool)
    for kw in (dict(anisotropy=(1.0, 2.0)), dict(black_border=True), dict(voxel_graph=np.zeros((4, 5), np.uint8))):
        with pytest.raises(NotImplementedError):
            psb.edt(im, **kw)
    with pytest.raises(NotImplementedError):
        psb.edt(np.array([[0, 1, 2], [3, 3, 0]]))          # multi-label image
    with pytest.raises(ValueError):
        psb.edt(np.ones((2, 2, 2, 2), bool))


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints one JSON line with the
    contract's keys; it needs no GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-edge", "48"], capture_output=True, text=True, cwd=root, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "local_thickness_voxels_per_s"
    assert line["unit"] == "voxels/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_patch_install_logic_without_gpu():
    """install()/uninstall() rebinding (sys.modules['edt'], every porespy.* module that bound the original `edt`,
    the filters) on stand-in modules; no kernel is called."""
    import importlib
    import sys
    import types
    import porespy_b200
    from porespy_b200 import patch
    edt_mod = importlib.import_module("porespy_b200.edt")

    def orig(data, **kw):
        return "orig"

    fake = types.ModuleType("edt")
    fake.edt = orig
    mods = {n: types.ModuleType(n) for n in ("porespy", "porespy.filters", "porespy.filters._funcs",
                                             "porespy.tools._funcs", "porespy.networks._getnet")}
    mods["porespy"].filters = mods["porespy.filters"]
    for n in ("porespy.filters._funcs", "porespy.tools._funcs", "porespy.networks._getnet"):
        mods[n].edt = orig
    mods["porespy.filters._funcs"].porosimetry = lambda: 0
    mods["porespy.filters"].porosimetry = mods["porespy.filters._funcs"].porosimetry
    mods["edt"] = fake
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        patch.install()
        assert sys.modules["edt"].edt is edt_mod.edt
        for n in ("porespy.filters._funcs", "porespy.tools._funcs", "porespy.networks._getnet"):
            assert mods[n].edt is edt_mod.edt
        assert mods["porespy.filters"].porosimetry is porespy_b200.filters.porosimetry
        patch.uninstall()
        assert sys.modules["edt"] is fake
        assert mods["porespy.networks._getnet"].edt is orig
    finally:
        patch.uninstall()
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_slab_plan_of_the_pipelined_numpy_result():
    """filters._slab_plan: the slabs tile [0, nz) in order, every slab carries the largest reach as halo (clipped at
    the volume), and the plan is declined when halos would outweigh the slabs or the volume is small."""
    from porespy_b200 import filters as F
    from porespy_b200 import _host
    saved = dict(F.SLAB_PIPELINE)
    try:
        F.SLAB_PIPELINE.update(enabled=True, min_voxels=1 << 28, slabs=6)
        T = [1700, 900, 300, 50, 2]                      # thresholds descend: reach of the first one is the halo
        W = _host.isqrt(T[0] - 1)
        plan = F._slab_plan((1024, 1024, 1024), T)
        assert plan is not None and len(plan) == 6
        assert plan[0][0] == 0 and plan[-1][1] == 1024
        for (z0, z1, e0, e1), nxt in zip(plan, plan[1:] + [None]):
            assert z0 < z1 and e0 == max(0, z0 - W) and e1 == min(1024, z1 + W)
            if nxt is not None:
                assert nxt[0] == z1
        assert plan[0][1] - plan[0][0] < plan[1][1] - plan[1][0]            # thin first slab: the epilogue starts early
        assert F._slab_plan((256, 256, 256), T) is None                    # below min_voxels
        assert F._slab_plan((1024, 1024, 1024), []) is None
        assert F._slab_plan((1024, 1024, 1024), list(range(400, 0, -1))) is None      # > 253 thresholds: grouped path
        F.SLAB_PIPELINE.update(min_voxels=1)
        assert F._slab_plan((60, 64, 64), [1700, 5]) is None               # 2 W + 1 planes of halo per 10-plane slab
        F.SLAB_PIPELINE.update(enabled=False)
        assert F._slab_plan((1024, 1024, 1024), T) is None
    finally:
        F.SLAB_PIPELINE.clear()
        F.SLAB_PIPELINE.update(saved)
