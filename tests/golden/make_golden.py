"""Generates tests/golden/*.npz by running the REFERENCE's own source.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference (`/root/reference/src/porespy`) is imported unmodified through
`oracle/ref_shim.py` (scipy-backed exact `edt.edt`, trivial skimage strels).  Every output
below is what `ps.filters.porosimetry / local_thickness / trim_disconnected_blobs`
(`src/porespy/filters/_funcs.py:947-1270`) returned for the stored input.  The cases
mirror the reference's own tests of the path (test/unit/test_filters.py:27-56, 201-210,
266-288; test/unit/test_tools.py:309-316; test/integration/test_drainage.py:7-18,47-56;
examples/filters/reference/porosimetry.ipynb cell 11).

Storage: boolean images are bit-packed; float64 radius maps are stored losslessly as
(`values` = sorted unique float64, `idx` = uint8 index of every voxel into `values`).
Volumes too large to commit (200^3) are stored as SHA-256 of the raw float64 bytes plus
the value histogram.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import ref_shim  # noqa: E402

ps = ref_shim.import_reference()


def pack(mask):
    mask = np.asarray(mask, dtype=bool)
    return dict(shape=np.array(mask.shape, dtype=np.int64), bits=np.packbits(mask.ravel()))


def enc_map(m):
    vals, inv = np.unique(m, return_inverse=True)
    assert len(vals) < 256
    return dict(values=vals.astype(np.float64), idx=inv.reshape(m.shape).astype(np.uint8))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name, **arrays):
    flat = {}
    for k, v in arrays.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[f"{k}__{kk}"] = vv
        else:
            flat[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    t0 = time.time()
    # ------------------------------------------------------------------ FilterTest image
    np.random.seed(0)
    im = ps.generators.blobs(shape=[100, 100, 100], blobiness=2)
    assert ps.metrics.porosity(im) == 0.499829          # TF:17
    from edt import edt
    dt = edt(im)
    d2 = np.rint(dt.astype(np.float64) ** 2).astype(np.uint32)
    assert np.array_equal(np.sqrt(d2.astype(np.float32)), dt)
    im2d = im[:, :, 50]
    sizes_int = np.arange(25, 1, -1)
    sizes_log = np.logspace(0.01, 0.6, 5)
    inlet0 = np.zeros_like(im, dtype=bool)
    inlet0[0, ...] = True
    out = dict(
        im=pack(im),
        d2_sha=sha(d2), d2_max=int(d2.max()), d2_sum=int(d2.astype(np.int64).sum()),
        d2_slice50=d2[:, :, 50],
        poro_hybrid_sizes10=enc_map(ps.filters.porosimetry(im=im, sizes=10)),       # TF:36-42
        poro_dt_arange_3d=enc_map(ps.filters.porosimetry(im, sizes=sizes_int, mode="dt")),  # TF:44-51
        poro_dt_arange_2d=enc_map(ps.filters.porosimetry(im2d, sizes=sizes_int, mode="dt")),  # TF:27-34
        poro_logsizes=enc_map(ps.filters.porosimetry(im=im, sizes=sizes_log)),      # TF:53-56
        lt_dt_25=enc_map(ps.filters.local_thickness(im, mode="dt")),                # TF:266-272
        lt_2d_25=enc_map(ps.filters.local_thickness(im2d)),                         # TF:281-288
        poro_inlet0_dt_12=enc_map(ps.filters.porosimetry(im, sizes=12, inlets=inlet0,
                                                         mode="dt")),               # ipynb cell 11
        poro_2d_noaccess=enc_map(ps.filters.porosimetry(im2d, sizes=9, access_limited=False)),
        lt_list_sizes=enc_map(ps.filters.local_thickness(im, sizes=[6, 4.5, 3, 2, 1], mode="dt")),
    )
    hyb = ps.filters.porosimetry(im, sizes=sizes_int, mode="hybrid")
    assert np.array_equal(hyb, ps.filters.porosimetry(im, sizes=sizes_int, mode="dt"))
    assert np.isclose(out["lt_dt_25"]["values"].max(), dt.max(), atol=1e-6)
    save("blobs100", **out)

    # ------------------------------------------------- trim_disconnected_blobs (TF:201-210)
    np.random.seed(0)
    b2 = ps.generators.blobs([200, 200], porosity=0.55, blobiness=2)
    inl = np.zeros_like(b2)
    inl[0, ...] = 1
    h8 = ps.filters.trim_disconnected_blobs(im=b2, inlets=inl)
    from skimage.morphology import disk, ball
    h4 = ps.filters.trim_disconnected_blobs(im=b2, inlets=inl, strel=disk(1))
    np.random.seed(1)
    b3 = ps.generators.blobs([60, 50, 40], porosity=0.45, blobiness=1.5)
    inl3 = np.zeros_like(b3)
    inl3[:, 0, :] = 1
    save("trim", im2d=pack(b2), inlets2d=pack(inl), out8=pack(h8), out4=pack(h4),
         im3d=pack(b3), inlets3d=pack(inl3),
         out26=pack(ps.filters.trim_disconnected_blobs(im=b3, inlets=inl3)),
         out6=pack(ps.filters.trim_disconnected_blobs(im=b3, inlets=inl3, strel=ball(1))))

    # ---------------------------------------------------- strels (test_tools.py:309-316)
    save("strels",
         disk3=pack(ps.tools.ps_disk(3)), ball3=pack(ps.tools.ps_ball(3)),
         disk3_rough=pack(ps.tools.ps_disk(3, smooth=False)),
         ball3_rough=pack(ps.tools.ps_ball(3, smooth=False)),
         ball_4p2426=pack(ps.tools.ps_ball(np.float32(4.2426405))),
         sums=np.array([ps.tools.ps_disk(3).sum(), ps.tools.ps_ball(3).sum(),
                        ps.tools.ps_disk(3, smooth=False).sum(),
                        ps.tools.ps_ball(3, smooth=False).sum()]))

    # ------------------------------------- known sizes (TF:274-279) and drainage residual
    np.random.seed(0)
    rsa = np.zeros(shape=[300, 300])
    rsa = ps.generators.random_spheres(im=rsa, r=20)
    rsa = ps.generators.random_spheres(im=rsa, r=10)
    lt_rsa = ps.filters.local_thickness(rsa, sizes=[20, 10])
    assert np.all(np.unique(lt_rsa) == [0, 10, 20])
    np.random.seed(6)
    drn = ps.generators.blobs(shape=[500, 500], porosity=0.7, blobiness=1.5)
    inlets = np.zeros_like(drn)
    inlets[0, :] = True
    outlets = np.zeros_like(drn)
    outlets[-1, :] = True
    drn = ps.filters.trim_nonpercolating_paths(im=drn, inlets=inlets, outlets=outlets)
    lt_drn = ps.filters.local_thickness(drn)
    ratio = (lt_drn > 25).sum() / drn.sum()
    assert ratio == 0.34427115020497745, ratio          # test_drainage.py:17-18,49
    save("misc2d", rsa=pack(rsa > 0), lt_rsa=enc_map(lt_rsa),
         drn=pack(drn), lt_drn=enc_map(lt_drn), drn_ratio=np.float64(ratio))

    # ----------------------------------------------- BASELINE config 0 (200^3, sizes=25)
    np.random.seed(0)
    big = ps.generators.blobs(shape=[200, 200, 200], porosity=0.6, blobiness=2)
    lt = ps.filters.local_thickness(big, sizes=25)
    mip = ps.filters.porosimetry(big, sizes=25, inlets=None)
    inl0 = np.zeros_like(big)
    inl0[0, ...] = True
    mip0 = ps.filters.porosimetry(big, sizes=25, inlets=inl0)
    dbig = edt(big)
    d2big = np.rint(dbig.astype(np.float64) ** 2).astype(np.uint32)
    v, c = np.unique(lt, return_counts=True)
    v1, c1 = np.unique(mip, return_counts=True)
    v2, c2 = np.unique(mip0, return_counts=True)
    save("config0", im_sha=sha(big), porosity=np.float64(big.mean()),
         d2_sha=sha(d2big), d2_max=int(d2big.max()), dt_sha=sha(dbig),
         lt_sha=sha(lt), lt_values=v, lt_counts=c,
         poro_faces_sha=sha(mip), poro_faces_values=v1, poro_faces_counts=c1,
         poro_inlet0_sha=sha(mip0), poro_inlet0_values=v2, poro_inlet0_counts=c2)
    print(f"done in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
