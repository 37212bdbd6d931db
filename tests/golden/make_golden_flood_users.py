"""Generates tests/golden/flood_users.npz by running the REFERENCE's own source (dev container only):

    python tests/golden/make_golden_flood_users.py

`find_disconnected_voxels`, `fill_blind_pores`, `trim_floating_solid`
(`src/porespy/filters/_funcs.py:352-503`) and `trim_nonpercolating_paths` (`:506-555`), imported
unmodified through `oracle/ref_shim.py` (skimage's `clear_border` and the 3x3(x3) strels are
restated there in numpy).  The cases mirror test/unit/test_filters.py:107-121, 123-199, 212-222.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
from make_golden import pack, save  # noqa: E402  (its module-level import of the reference is reused)

ps = ref_shim.import_reference()


def main():
    np.random.seed(0)
    im = ps.generators.blobs(shape=[100, 100, 100], blobiness=2)
    assert ps.metrics.porosity(im) == 0.499829                      # TF:17
    f = ps.filters
    h2d8, h2d4 = f.find_disconnected_voxels(im[:, :, 0]), f.find_disconnected_voxels(im[:, :, 0], conn=4)
    h26, h6 = f.find_disconnected_voxels(im), f.find_disconnected_voxels(im, conn=6)
    assert (h2d8.sum(), h2d4.sum(), h26.sum(), h6.sum()) == (477, 652, 55, 202)      # TF:107-121
    out = dict(im=pack(im), h2d8=pack(h2d8), h2d4=pack(h2d4), h26=pack(h26), h6=pack(h6),
               h26_surface=pack(f.find_disconnected_voxels(im, surface=True)),
               h6_surface=pack(f.find_disconnected_voxels(im, conn=6, surface=True)),
               h2d4_surface=pack(f.find_disconnected_voxels(im[:, :, 0], conn=4, surface=True)),
               fill_blind=pack(f.fill_blind_pores(im)), fill_blind6s=pack(f.fill_blind_pores(im, conn=6, surface=True)),
               trim_solid=pack(f.trim_floating_solid(im)), trim_solid6=pack(f.trim_floating_solid(im, conn=6)))
    # a face without any background voxel: label 0 drops out of `keep` (F:413-420) and the solid counts as holes
    cap = im.copy()
    cap[0] = True
    out["cap"] = pack(cap)
    out["cap_surface"] = pack(f.find_disconnected_voxels(cap, conn=6, surface=True))
    # trim_nonpercolating_paths, TF:123-199
    np.random.seed(0)
    b2 = ps.generators.blobs([200, 200], porosity=0.55, blobiness=2)
    for ax in (0, 1):
        inl, outl = np.zeros_like(b2), np.zeros_like(b2)
        inl[(slice(None),) * ax + (0,)] = 1
        outl[(slice(None),) * ax + (-1,)] = 1
        out[f"np2d_ax{ax}"] = pack(f.trim_nonpercolating_paths(im=b2, inlets=inl, outlets=outl))
    out["np2d_im"] = pack(b2)
    np.random.seed(0)
    b25 = ps.generators.blobs([200, 200], porosity=0.25, blobiness=2)
    inl, outl = np.zeros_like(b25), np.zeros_like(b25)
    inl[:, 0] = 1
    outl[:, -1] = 1
    none = f.trim_nonpercolating_paths(im=b25, inlets=inl, outlets=outl)
    assert none.sum() == 0                                          # TF:149-160
    out["np2d_none_im"] = pack(b25)
    np.random.seed(0)
    b3 = ps.generators.blobs([100, 100, 100], porosity=0.55, blobiness=2)
    out["np3d_im"] = pack(b3)
    for ax in (0, 1, 2):
        inl, outl = np.zeros_like(b3), np.zeros_like(b3)
        inl[(slice(None),) * ax + (0,)] = 1
        outl[(slice(None),) * ax + (-1,)] = 1
        out[f"np3d_ax{ax}"] = pack(f.trim_nonpercolating_paths(im=b3, inlets=inl, outlets=outl))
    from skimage.morphology import cube
    inl, outl = np.zeros_like(b3), np.zeros_like(b3)
    inl[0], outl[-1] = 1, 1
    out["np3d_ax0_cube"] = pack(f.trim_nonpercolating_paths(im=b3, inlets=inl, outlets=outl, strel=cube(3)))
    save("flood_users", **out)


if __name__ == "__main__":
    main()
