"""Generates tests/golden/trapped.npz by running the REFERENCE's own source (dev container only):

    python tests/golden/make_golden_trapped.py

`find_trapped_regions` (`src/porespy/filters/_funcs.py:73-147`) on invasion sequences made by the
reference itself (`porosimetry` -> `size_to_seq`), imported unmodified through `oracle/ref_shim.py`.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
from make_golden import pack, save  # noqa: E402

ps = ref_shim.import_reference()


def main():
    f = ps.filters
    out = {}
    np.random.seed(0)
    im2 = ps.generators.blobs([200, 200], porosity=0.6, blobiness=1.5)
    inl = np.zeros_like(im2)
    inl[0, :] = True
    seq2 = f.size_to_seq(f.porosimetry(im2, sizes=15, inlets=inl))
    outl = np.zeros_like(im2)
    outl[-1, :] = True
    out["seq2d"] = seq2.astype(np.int32)
    out["t2d_faces_25"] = pack(f.find_trapped_regions(seq2))
    out["t2d_outlet_25"] = pack(f.find_trapped_regions(seq2, outlets=outl))
    out["t2d_outlet_all"] = pack(f.find_trapped_regions(seq2, outlets=outl, bins=None))
    out["t2d_outlet_7"] = pack(f.find_trapped_regions(seq2, outlets=outl, bins=7))
    out["s2d_outlet_seq"] = f.find_trapped_regions(seq2, outlets=outl, bins=None, return_mask=False).astype(np.int32)
    np.random.seed(2)
    im3 = ps.generators.blobs([60, 50, 40], porosity=0.55, blobiness=1.2)
    inl3 = np.zeros_like(im3)
    inl3[0] = True
    outl3 = np.zeros_like(im3)
    outl3[-1] = True
    seq3 = f.size_to_seq(f.porosimetry(im3, sizes=12, inlets=inl3))
    out["seq3d"] = seq3.astype(np.int32)
    out["t3d_outlet_25"] = pack(f.find_trapped_regions(seq3, outlets=outl3))
    out["t3d_faces_all"] = pack(f.find_trapped_regions(seq3, bins=None))
    print({k: int(np.unpackbits(v["bits"]).sum()) for k, v in out.items() if isinstance(v, dict)})
    save("trapped", **out)


if __name__ == "__main__":
    main()
