"""GPU suite: `porespy_b200.simulations.drainage` against the outputs of the reference's own
`ps.simulations.drainage` (tests/golden/make_golden_drainage.py; cases after test/integration/test_drainage.py) and
against the numpy restatement in oracle/cpu.py on other inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import cpu as oc                       # noqa: E402  (checker only)
from tests.test_gpu_parity import assert_same      # noqa: E402


@pytest.fixture(scope="module", params=["one_flood", "flood_per_step"])
def psb(request):
    """Both forms of the pressure loop: one flood with join times for all (ascending) steps, and one flood per step."""
    import torch
    assert torch.cuda.is_available()
    import porespy_b200 as psb
    from porespy_b200 import filters, simulations
    saved = simulations.ONE_FLOOD, filters.ONE_FLOOD_TRAPPED
    simulations.ONE_FLOOD = filters.ONE_FLOOD_TRAPPED = request.param == "one_flood"
    yield psb
    simulations.ONE_FLOOD, filters.ONE_FLOOD_TRAPPED = saved


def same(r, want, name):
    get = (lambda k: want.rmap(name + "_" + k)) if hasattr(want, "rmap") else (lambda k: want[k])
    assert_same(r.im_pc, get("im_pc"), name + " im_pc")
    assert_same(r.im_satn, get("im_satn"), name + " im_satn")
    if hasattr(want, "rmap"):
        has_tr = (name + "_im_trapped__bits") in want.keys()
        tr = want.mask(name + "_im_trapped") if has_tr else None
        pc, snwp = want.raw(name + "_pc"), want.raw(name + "_snwp")
    else:
        tr, pc, snwp = want["im_trapped"], np.asarray(want["pc"]), np.asarray(want["snwp"])
    if tr is None:
        assert r.im_trapped is None
    else:
        assert_same(r.im_trapped, tr, name + " im_trapped")
    assert np.array_equal(np.asarray(r.pc), pc), name + " pc"
    assert np.array_equal(np.asarray(r.snwp), snwp), name + " snwp"


def test_goldens_2d(psb, golden):
    d = psb.simulations.drainage
    g = golden.drainage
    im, inl, out, res = g.mask("a_im"), g.mask("a_inlets"), g.mask("a_outlets"), g.mask("a_residual")
    vs = 1e-4
    r1 = d(im=im, voxel_size=vs, inlets=inl, g=0)
    same(r1, g, "a1")
    assert r1.snwp[0] == 0 and r1.snwp[-1] == 1                         # test_drainage.py:46,52
    same(d(im=im, voxel_size=vs, inlets=inl, outlets=out, g=0), g, "a2")
    same(d(im=im, voxel_size=vs, inlets=inl, residual=res, g=0), g, "a3")
    same(d(im=im, voxel_size=vs, inlets=inl, outlets=out, residual=res, g=0), g, "a4")
    same(d(im=im, voxel_size=vs, inlets=inl, outlets=out), g, "a5")
    same(d(im=im, voxel_size=vs, inlets=inl, bins=[300.0, 900.0, 2000.0, 1500.0, 8000.0], delta_rho=-997, g=9.81,
           sigma=0.05, theta=140), g, "a6")
    same(d(im=im, voxel_size=np.float64(vs), inlets=inl, bins=12), g, "a7")


def test_goldens_3d(psb, golden):
    d = psb.simulations.drainage
    g = golden.drainage
    im3, out3 = g.mask("b_im"), g.mask("b_outlets")
    same(d(im=im3, voxel_size=1e-5), g, "b1")
    same(d(im=im3, voxel_size=1e-5, outlets=out3, bins=15, g=0), g, "b2")
    pc_user = g.raw("b_pc_user")
    keep = pc_user.copy()
    same(d(im=im3, voxel_size=1e-5, pc=pc_user, bins=10), g, "b3")
    assert np.array_equal(pc_user, keep)


@pytest.mark.parametrize("shape,seed", [((96, 80, 64), 1), ((150, 130), 2)])
def test_vs_numpy_restatement(psb, shape, seed):
    im = oc.blobs(list(shape), porosity=0.65, blobiness=1.5, seed=seed)
    out = np.zeros_like(im)
    out[-1] = True
    for kw in (dict(), dict(outlets=out, bins=12), dict(g=0, bins=9, sigma=0.03)):
        same(psb.simulations.drainage(im=im, voxel_size=2e-5, **kw), oc.drainage(im, 2e-5, **kw), f"{shape} {sorted(kw)}")


@pytest.mark.parametrize("conn", [6, 26])
def test_flood_classes_vs_labelling(conn):
    """psb200_flood_classes: first step at which a voxel of nested sets is connected to the inlets, against one
    scipy labelling per step (inlet voxels are nodes from step 0 on, F:1265)."""
    import scipy.ndimage as spim
    import torch
    from porespy_b200 import _device as dev
    from porespy_b200 import _lib
    ctx = _lib.context()
    rng = np.random.default_rng(7)
    for shape, nsteps in (((24, 30, 140), 6), ((1, 60, 131), 4), ((9, 11, 13), 9)):
        sm = spim.gaussian_filter(rng.random(shape), 2.0)
        cls = np.clip(((sm - sm.min()) / np.ptp(sm) * (nsteps + 2)).astype(int), 0, nsteps + 1)
        cls = np.where(cls >= nsteps, 254, cls)
        cls = np.where(rng.random(shape) < 0.25, 255, cls).astype(np.uint8)
        inl = np.zeros(shape, dtype=bool)
        inl[..., 0] = True
        inl |= rng.random(shape) < 0.001
        c = conn if shape[0] > 1 else (4 if conn == 6 else 8)
        got = dev.flood_classes(ctx, torch.from_numpy(cls.reshape(-1)).cuda(), torch.from_numpy(inl.reshape(-1).view(np.uint8)).cuda(),
                                nsteps, c, shape).cpu().numpy().reshape(shape)
        st = spim.generate_binary_structure(3, 1 if conn == 6 else 3)
        want = np.where(cls == 255, 255, 254).astype(np.uint8)
        for k in range(nsteps - 1, -1, -1):
            nodes = (cls <= k) | inl
            lab = spim.label(nodes, structure=st)[0]
            keep = np.unique(lab[inl])
            want[np.isin(lab, keep[keep > 0]) & (cls <= k)] = k
        assert np.array_equal(got, want), (shape, conn, np.argwhere(got != want)[:5])
