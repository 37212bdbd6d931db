"""Golden vectors for the radius-map post-processing functions, produced by the
REFERENCE's own source through oracle/ref_shim.py (dev container only):

    python tests/golden/make_golden_sizemap.py

`size_to_seq`, `size_to_satn`, `seq_to_satn` (filters/_size_seq_satn.py:16-221), `pore_size_distribution`
(metrics/_funcs.py:558-632) and the `sizes` branch of `pc_curve` (metrics/_funcs.py:1073-1090), on the
local-thickness / porosimetry maps of the 100^3 blobs image of the reference's tests (test_filters.py:13-17) and
on the hand-made image of test/unit/test_filters_size_seq_satn.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_shim  # noqa: E402

ps = ref_shim.import_reference()
from tests.golden_io import Golden  # noqa: E402

g = Golden().blobs100
im = g.mask("im")
lt = g.rmap("lt_dt_25")
mip = g.rmap("poro_inlet0_dt_12")                 # zeros inside the pore space: uninvaded voxels stay 0
out = {}
out["lt_satn_dr"] = ps.filters.size_to_satn(lt)
out["lt_satn_im"] = ps.filters.size_to_satn(lt, mode="imbibition")
out["lt_satn_bins12"] = ps.filters.size_to_satn(lt, bins=12)
out["mip_satn_im_mask"] = ps.filters.size_to_satn(mip, im=im)
out["lt_seq_dr"] = ps.filters.size_to_seq(lt)
out["lt_seq_im"] = ps.filters.size_to_seq(lt, mode="imbibition")
out["mip_seq_mask"] = ps.filters.size_to_seq(mip, im=im)
out["lt_seq_bins10"] = ps.filters.size_to_seq(lt, bins=10)
seq = out["lt_seq_dr"]
out["seq_satn_dr"] = ps.filters.seq_to_satn(seq)
out["seq_satn_im"] = ps.filters.seq_to_satn(seq, mode="imbibition")
mseq = out["mip_seq_mask"].copy()
mseq[(mip == 0) & im] = -1                         # uninvaded
out["mseq"] = mseq
out["mseq_satn_mask"] = ps.filters.seq_to_satn(mseq, im=im)
for name, kw in (("psd_default", {}), ("psd_lin20", dict(bins=20, log=False)), ("psd_vox", dict(bins=7, voxel_size=2.5))):
    r = ps.metrics.pore_size_distribution(lt, **kw)
    for f in ("pdf", "cdf", "satn", "bin_centers", "bin_edges", "bin_widths"):
        out[f"{name}__{f}"] = np.asarray(getattr(r, f))
r = ps.metrics.pc_curve(sizes=lt, im=im, sigma=0.072, theta=180, voxel_size=1e-5)
out["pc_lt__pc"], out["pc_lt__snwp"] = np.asarray(r.pc), np.asarray(r.snwp)
r = ps.metrics.pc_curve(im=None, sizes=mip)
out["pc_mip__pc"], out["pc_mip__snwp"] = np.asarray(r.pc), np.asarray(r.snwp)

# the reference's own unit-test image (test/unit/test_filters_size_seq_satn.py:13-30 style): small hand-made maps
small = np.array([[0, 0, 0, 0, 0, 0], [0, 3, 3, 2, 2, -1], [0, 3, 1, 1, 2, -1], [0, 0, 1.5, 1.5, 0, 0]])
out["small"] = small
out["small_satn"] = ps.filters.size_to_satn(small)
out["small_seq"] = ps.filters.size_to_seq(small)
out["small_seq_im"] = ps.filters.size_to_seq(small, mode="imbibition")
out["small_seq_satn"] = ps.filters.seq_to_satn(out["small_seq"])

flat = {}
for k, v in out.items():
    v = np.asarray(v)
    if v.ndim == 3:                                # volumes: (values, idx) like make_golden.enc_map
        vals, inv = np.unique(v, return_inverse=True)
        flat[k + "__values"] = vals
        flat[k + "__idx"] = inv.reshape(v.shape).astype(np.uint8 if len(vals) <= 256 else np.uint16)
    else:
        flat[k] = v
path = os.path.join(HERE, "sizemap.npz")
np.savez_compressed(path, **flat)
print(f"sizemap: {os.path.getsize(path) / 1024:.0f} KiB, {len(flat)} arrays")
