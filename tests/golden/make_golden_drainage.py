"""Golden vectors for `ps.simulations.drainage`, produced by the REFERENCE's own source
(/root/reference/src/porespy/simulations/_drainage.py) through oracle/ref_shim.py -- dev container only:

    python tests/golden/make_golden_drainage.py

Cases follow test/integration/test_drainage.py:7-56 (2-D blobs, with / without trapping and residual), plus
gravity, a 3-D image with default inlets, explicit pressure bins and a user-supplied pc map.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_shim  # noqa: E402

ps = ref_shim.import_reference()
flat = {}


def pack(name, mask):
    mask = np.asarray(mask, dtype=bool)
    flat[name + "__shape"] = np.array(mask.shape, dtype=np.int64)
    flat[name + "__bits"] = np.packbits(mask.ravel())


def enc(name, m):
    vals, inv = np.unique(m, return_inverse=True)
    assert len(vals) < 256
    flat[name + "__values"] = vals.astype(np.float64)
    flat[name + "__idx"] = inv.reshape(m.shape).astype(np.uint8)


def store(name, r):
    enc(name + "_im_pc", r.im_pc)
    enc(name + "_im_satn", r.im_satn)
    if r.im_trapped is not None:
        pack(name + "_im_trapped", r.im_trapped)
    flat[name + "_pc"] = np.asarray(r.pc, dtype=np.float64)
    flat[name + "_snwp"] = np.asarray(r.snwp, dtype=np.float64)


np.random.seed(6)
im = ps.generators.blobs(shape=[200, 200], porosity=0.7, blobiness=1.5)
inlets = np.zeros_like(im)
inlets[0, :] = True
outlets = np.zeros_like(im)
outlets[-1, :] = True
im = ps.filters.trim_nonpercolating_paths(im=im, inlets=inlets, outlets=outlets)
lt = ps.filters.local_thickness(im)
residual = lt > 14
pack("a_im", im), pack("a_inlets", inlets), pack("a_outlets", outlets), pack("a_residual", residual)
vs = 1e-4
store("a1", ps.simulations.drainage(im=im, voxel_size=vs, inlets=inlets, g=0))
store("a2", ps.simulations.drainage(im=im, voxel_size=vs, inlets=inlets, outlets=outlets, g=0))
store("a3", ps.simulations.drainage(im=im, voxel_size=vs, inlets=inlets, residual=residual, g=0))
store("a4", ps.simulations.drainage(im=im, voxel_size=vs, inlets=inlets, outlets=outlets, residual=residual, g=0))
store("a5", ps.simulations.drainage(im=im, voxel_size=vs, inlets=inlets, outlets=outlets))            # gravity, defaults
store("a6", ps.simulations.drainage(im=im, voxel_size=vs, inlets=inlets, bins=[300.0, 900.0, 2000.0, 1500.0, 8000.0],
                                    delta_rho=-997, g=9.81, sigma=0.05, theta=140))
store("a7", ps.simulations.drainage(im=im, voxel_size=np.float64(vs), inlets=inlets, bins=12))          # numpy scalar: float64 products

np.random.seed(3)
im3 = ps.generators.blobs(shape=[48, 40, 56], porosity=0.65, blobiness=1)
out3 = np.zeros_like(im3)
out3[-1] = True
pack("b_im", im3), pack("b_outlets", out3)
store("b1", ps.simulations.drainage(im=im3, voxel_size=1e-5))                                         # default inlets
store("b2", ps.simulations.drainage(im=im3, voxel_size=1e-5, outlets=out3, bins=15, g=0))
from edt import edt as ref_edt
pc_user = 2 * 0.072 / (ref_edt(im3) * 1e-5).astype(np.float64)
pc_user[~im3] = 0
flat["b_pc_user"] = pc_user.astype(np.float64)
store("b3", ps.simulations.drainage(im=im3, voxel_size=1e-5, pc=pc_user.copy(), bins=10))

path = os.path.join(HERE, "drainage.npz")
np.savez_compressed(path, **flat)
print(f"drainage: {os.path.getsize(path) / 1024:.0f} KiB, {len(flat)} arrays")
