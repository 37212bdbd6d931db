"""Radius / sequence maps in INDEX form and the reference's post-processing functions evaluated on it
(SURVEY 8(f) rank 3):

* `size_to_seq`, `size_to_satn`, `seq_to_satn`  /root/reference/src/porespy/filters/_size_seq_satn.py:16-221
* `pore_size_distribution`                      /root/reference/src/porespy/metrics/_funcs.py:558-632
* `pc_curve` (the `sizes` branch)               /root/reference/src/porespy/metrics/_funcs.py:1073-1090

`porosimetry` / `local_thickness` write one of <= 254 radii into every voxel; on the device that map IS one
index byte per voxel plus a table of radii (`IndexMap`).  Each function above only needs (a) set operations on
the distinct values, (b) voxel counts per (value, pore-mask) combination and (c) a pointwise value -> value
map.  (b) is one histogram kernel over the index map, (c) one table-expansion kernel (csrc/sizemap_kernels.cuh);
(a) and the arithmetic in between are the reference's OWN numpy statements, executed on one representative
voxel per (value, mask) combination with the counts as weights.  The 8 B/voxel float64 radius map is never
produced unless the caller asks for a float64 result.

Arbitrary arrays (a float64 map from an earlier call, an integer sequence map) are converted on the device
(`IndexMap.from_array`: distinct-value hash table + index-of pass); up to 65536 distinct values.
"""
import ctypes

import numpy as np

from . import _device as dev
from . import _host as host
from . import _lib

__all__ = ["IndexMap", "local_thickness_index", "porosimetry_index", "size_to_seq", "size_to_satn", "seq_to_satn",
           "pore_size_distribution", "pc_curve", "Results"]

_EMPTY = np.uint64(0x8000000000000000)      # -0.0 / INT64_MIN: never a key (floats are canonicalised with + 0.0)
_TABLE_CAP = 1 << 18
MAX_VALUES = 65536


class Results(dict):
    """Attribute-style result container (the reference returns `porespy.tools.Results`, a dataclass-like bag
    of named arrays)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class IndexMap:
    """`values[idx]` is the map.  idx: flat device tensor, uint8 (<= 256 values) or int16 holding uint16;
    values: 1-D numpy array (any order, any 8-byte-representable dtype); shape: the image shape."""

    def __init__(self, ctx, idx, values, shape):
        self.ctx, self.idx, self.shape = ctx, idx, tuple(int(s) for s in shape)
        self.values = np.asarray(values)
        self.idx_bytes = idx.element_size()
        assert self.idx_bytes in (1, 2) and len(self.values) <= (256 if self.idx_bytes == 1 else MAX_VALUES)

    @property
    def size(self):
        return int(np.prod(self.shape)) if len(self.shape) else 1

    # ---------------------------------------------------------------------------- construction
    @classmethod
    def from_array(cls, arr, ctx=None):
        """Any numpy array / torch tensor with <= 65536 distinct values -> index form (device)."""
        if isinstance(arr, IndexMap):
            return arr
        torch = dev._torch()
        ctx = ctx if ctx is not None else _lib.context()
        device = f"cuda:{ctx.device}"
        if isinstance(arr, torch.Tensor):
            np_dtype = np.dtype(str(arr.dtype).replace("torch.", "")) if arr.dtype != torch.bool else np.dtype(bool)
            t = arr.to(device)
            if t.dtype.is_floating_point:
                kind, x = 0, (t.to(torch.float64) + 0.0).reshape(-1)           # + 0.0: -0.0 and 0.0 are one value
            else:
                kind, x = 1, t.to(torch.int64).reshape(-1)
            shape = tuple(arr.shape)
        else:
            a = np.asarray(arr)
            np_dtype, shape = a.dtype, a.shape
            if a.dtype.kind == "f":
                kind, h = 0, np.ascontiguousarray(a, dtype=np.float64).reshape(-1) + 0.0
            elif a.dtype.kind in "iub":
                kind, h = 1, np.ascontiguousarray(a).astype(np.int64).reshape(-1)
            else:
                raise TypeError(f"IndexMap.from_array: unsupported dtype {a.dtype}")
            x = torch.from_numpy(h).to(device)
        n = x.numel()
        table = torch.empty(_TABLE_CAP, dtype=torch.int64, device=device)
        ovf = torch.zeros(1, dtype=torch.int32, device=device)
        _lib.check(ctx.lib.psb200_distinct64(ctx.handle, dev.ptr(x), n, dev.ptr(table), _TABLE_CAP, dev.ptr(ovf),
                                             dev.stream_ptr()))
        tab = table.cpu().numpy().view(np.uint64)
        keys = tab[tab != _EMPTY]
        if int(ovf.item()) or len(keys) > MAX_VALUES:
            raise ValueError("IndexMap.from_array: more than 65536 distinct values")
        if len(keys) == 0:
            keys = np.zeros(1, dtype=np.uint64)
        keys = np.sort(keys.view(np.float64 if kind == 0 else np.int64))
        K = len(keys)
        idx = torch.empty(n, dtype=torch.uint8 if K <= 256 else torch.int16, device=device)
        kd = torch.from_numpy(keys.view(np.int64).copy()).to(device)
        _lib.check(ctx.lib.psb200_index_of64(ctx.handle, dev.ptr(x), n, dev.ptr(kd), K, kind, dev.ptr(idx),
                                             idx.element_size(), dev.stream_ptr()))
        return cls(ctx, idx, keys.astype(np_dtype), shape)

    # ------------------------------------------------------------------------------- device ops
    def _mask(self, im):
        """Binary pore mask `im` of the image shape -> flat uint8 device tensor (non-binary images are rejected:
        the reference mixes `im == 0`, `im > 0` and `im == 1`, which only agree on binary images)."""
        torch = dev._torch()
        if isinstance(im, torch.Tensor):
            if tuple(im.shape) != self.shape:
                raise ValueError("im must have the shape of the map")
            if im.dtype != torch.bool and bool(((im != 0) & (im != 1)).any().item()):
                raise NotImplementedError("im must be a binary image")
            return dev.to_device_u8(im, self.ctx).reshape(-1)
        im = np.asarray(im)
        if im.shape != self.shape:
            raise ValueError("im must have the shape of the map")
        if im.dtype != np.bool_ and im.size and (im.min() < 0 or im.max() > 1 or np.any(im != im.astype(bool))):
            raise NotImplementedError("im must be a binary image")
        return dev.to_device_u8(im, self.ctx).reshape(-1)

    def counts(self, mask=None):
        """Voxels per value: int64 [K]; with a mask [2][K] (row 0: mask == 0, row 1: mask != 0)."""
        torch = dev._torch()
        K = len(self.values)
        out = torch.empty(K * (2 if mask is not None else 1), dtype=torch.int64, device=self.idx.device)
        _lib.check(self.ctx.lib.psb200_hist_idx(self.ctx.handle, dev.ptr(self.idx), self.idx_bytes, dev.ptr(mask),
                                                self.idx.numel(), K, dev.ptr(out), dev.stream_ptr()))
        c = out.cpu().numpy()
        return c.reshape(2, K) if mask is not None else c

    def expand(self, lut, mask=None, as_numpy=True):
        """out[v] = lut[(mask[v] ? K : 0) + idx[v]] as an array of lut.dtype (8-byte dtypes) and the map's shape."""
        torch = dev._torch()
        lut = np.ascontiguousarray(lut)
        if lut.dtype.itemsize != 8:
            lut = lut.astype(np.float64 if lut.dtype.kind == "f" else np.int64)
        K = len(self.values)
        assert len(lut) == K * (2 if mask is not None else 1)
        lut_d = torch.from_numpy(lut.view(np.int64).copy()).to(self.idx.device)
        out = torch.empty(self.idx.numel(), dtype=torch.int64, device=self.idx.device)
        _lib.check(self.ctx.lib.psb200_expand_lut8(self.ctx.handle, dev.ptr(self.idx), self.idx_bytes, dev.ptr(mask),
                                                   dev.ptr(lut_d), dev.ptr(out), out.numel(), K, dev.stream_ptr()))
        if not as_numpy:
            return (out.view(torch.float64) if lut.dtype.kind == "f" else out).view(*self.shape)
        return dev.to_host(out).view(lut.dtype).reshape(self.shape)

    def expand_mask(self, lut_bool, mask=None):
        """Device uint8 mask: out[v] = lut_bool[(mask[v] ? K : 0) + idx[v]] (flat tensor)."""
        torch = dev._torch()
        K = len(self.values)
        lut = np.ascontiguousarray(lut_bool, dtype=np.uint8)
        assert len(lut) == K * (2 if mask is not None else 1)
        lut_d = torch.from_numpy(lut).to(self.idx.device)
        out = torch.empty(self.idx.numel(), dtype=torch.uint8, device=self.idx.device)
        _lib.check(self.ctx.lib.psb200_expand_lut1(self.ctx.handle, dev.ptr(self.idx), self.idx_bytes, dev.ptr(mask),
                                                   dev.ptr(lut_d), dev.ptr(out), out.numel(), K, dev.stream_ptr()))
        return out

    def to_numpy(self):
        """The map itself (float64 for radius maps), e.g. to hand it to code that wants the reference's array."""
        vals = self.values
        if vals.dtype.itemsize != 8:
            return self.expand(vals.astype(np.float64 if vals.dtype.kind == "f" else np.int64)).astype(vals.dtype)
        return self.expand(vals)

    # -------------------------------------------------------- representative voxels (one per combination)
    def representatives(self, im=None):
        """(s, m, w, present, mask): value, pore flag and voxel count of every (value, mask) combination that
        occurs; `present` maps them back into the 2K (or K) table; `mask` is the device mask (or None)."""
        K = len(self.values)
        if im is None:
            c = self.counts()
            present = np.flatnonzero(c > 0)
            return self.values[present], None, c[present].astype(np.int64), present, None
        mask = self._mask(im)
        c = self.counts(mask).reshape(-1)
        present = np.flatnonzero(c > 0)
        return self.values[present % K], present >= K, c[present].astype(np.int64), present, mask

    def scatter(self, out_rep, present, mask):
        """Representative results -> per-voxel array."""
        K = len(self.values)
        lut = np.zeros(K * (2 if mask is not None else 1), dtype=out_rep.dtype)
        lut[present] = out_rep
        return self.expand(lut, mask)


# ------------------------------------------------------------------------------ producing index maps
def porosimetry_index(im, sizes=25, inlets=None, access_limited=True, mode="hybrid", divs=1):
    """`filters.porosimetry` (F:1032-1212) returning the map in index form: no float64 volume is written."""
    from . import filters
    return filters.porosimetry(im, sizes=sizes, inlets=inlets, access_limited=access_limited, mode=mode, divs=divs,
                               _as_index=True)


def local_thickness_index(im, sizes=25, mode="hybrid", divs=1):
    """`filters.local_thickness` (F:947-1029) in index form."""
    return porosimetry_index(im, sizes=sizes, access_limited=False, mode=mode, divs=divs)


# --------------------------------------------------------------- the reference functions on the index form
def size_to_satn(size, im=None, bins=None, mode="drainage"):
    r"""Invasion sizes -> non-wetting phase saturation (filters/_size_seq_satn.py:86-149).  `size`: IndexMap,
    numpy array or device tensor; returns the numpy float64 array the reference returns."""
    sm = IndexMap.from_array(size)
    s, m, w, present, mask = sm.representatives(im)
    if bins is None:
        bins = np.unique(s[s > 0])                                   # :135
    elif isinstance(bins, int):
        bins = np.linspace(0, s.max(), bins)                         # :137
    if m is None:
        m = ~(s == 0)                                                # :139
    void_vol = w[m].sum()                                            # :140  im.sum()
    satn = -np.ones_like(s, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore"):
        if mode.startswith("im"):
            for r in bins:
                hits = (s <= r) * (s > 0)
                temp = w[hits].sum() / void_vol                      # hits.sum()/void_vol
                satn[hits * (satn == -1)] = temp
        elif mode.startswith("dr"):
            for r in bins[-1::-1]:
                hits = (s >= r) * (s > 0)
                temp = w[hits].sum() / void_vol
                satn[hits * (satn == -1)] = temp
    satn *= (m > 0)                                                  # :148
    return sm.scatter(satn, present, mask)


def size_to_seq(size, im=None, bins=None, mode="drainage"):
    r"""Invasion sizes -> invasion sequence (filters/_size_seq_satn.py:16-83); int64 like the reference."""
    sm = IndexMap.from_array(size)
    s, m, w, present, mask = sm.representatives(im)
    solid = (s == 0) if m is None else (m == 0)                      # :62-65
    uninvaded = s == -1
    if bins is None:
        bins = np.unique(s)                                          # np.unique(size)
    elif isinstance(bins, int):
        bins = np.linspace(0, s.max(), bins)
    vals = np.digitize(s, bins=bins, right=True)
    if mode.startswith("im"):
        vals[solid] = 0
        vals[uninvaded] = -1
        vals = host.make_contiguous_symmetric(vals)
    if mode.startswith("dr"):
        vals = host.make_contiguous_symmetric(vals)
        vals = vals.max() + 1 - vals
        vals[solid] = 0
        vals[uninvaded] = -1
    return sm.scatter(np.asarray(vals, dtype=np.int64), present, mask)


def _seq_to_satn_rep(s, m, w, size, mode="drainage"):
    """Body of seq_to_satn (filters/_size_seq_satn.py:196-221) on representative voxels: s = sequence value,
    m = pore flag (None: not given), w = voxel count of the combination, size = voxels of the image."""
    q = np.copy(s).astype(int)
    solid_mask = (q == 0) if m is None else (m == 0)
    uninvaded_mask = q == -1
    q[q <= 0] = 0
    if mode.startswith("im"):
        q = q.max() - q + 1
        q[solid_mask] = 0
        q[uninvaded_mask] = 0
    q = np.unique(q, return_inverse=True)[1].reshape(-1)             # rankdata(seq, method='dense') - 1
    b = np.zeros(int(q.max()) + 1 if q.size else 1, dtype=np.int64)
    np.add.at(b, q, w)                                               # np.bincount(seq) over all voxels
    if (w[solid_mask].sum() > 0) or (w[uninvaded_mask].sum() > 0):
        b[0] = 0
    c = np.cumsum(b)
    with np.errstate(divide="ignore", invalid="ignore"):
        satn = c[q] / (size - w[solid_mask].sum())
    satn[solid_mask] = 0
    satn[uninvaded_mask] = -1
    return satn


def seq_to_satn(seq, im=None, mode="drainage"):
    r"""Invasion sequence -> saturation (filters/_size_seq_satn.py:152-221).  `rankdata(seq, 'dense') - 1` is
    evaluated as the dense rank `np.unique(..., return_inverse=True)` (integer, as the reference's bincount needs;
    SciPy >= 1.18 returns floats there and breaks the reference itself)."""
    sm = IndexMap.from_array(seq)
    s, m, w, present, mask = sm.representatives(im)
    return sm.scatter(_seq_to_satn_rep(s, m, w, sm.size, mode), present, mask)


def _pc_to_satn_rep(s, m, w, size, mode="drainage"):
    """pc_to_satn (filters/_size_seq_satn.py:338-342) on representative voxels (m: pore flag, mandatory)."""
    a = np.digitize(s, bins=np.unique(s))
    a[~m] = 0
    a[np.where(s == np.inf)] = -1
    return _seq_to_satn_rep(a, m, w, size, mode)


def _satn_to_seq_rep(satn, m, mode="drainage"):
    """satn_to_seq (filters/_size_seq_satn.py:384-399) on representative voxels (m: pore flag)."""
    uninvaded = satn == -1
    values = np.unique(satn)
    seq = np.digitize(satn, bins=values)
    seq[satn == -1] = -1
    seq[~m] = 0
    seq = host.make_contiguous_symmetric(seq)
    if mode.startswith("im"):
        seq = (seq.max() + 1) - seq
        seq[~m] = 0
    seq[uninvaded] = -1
    return seq


def _pc_curve_pc_rep(s, m, w):
    """pc_curve's `pc` branch (metrics/_funcs.py:1091-1108) on representative voxels -> (Ps, snwp)."""
    Ps = np.unique(s[m])
    if Ps[-1] == np.inf:
        Ps[-1] = Ps[-2] * 2
    if Ps[0] == -np.inf:
        Ps[0] = Ps[1] - np.abs(Ps[1] / 2)
    else:
        Ps = np.hstack((Ps[0] - np.abs(Ps[0] / 2), Ps))
    y = []
    Vp = w[m].sum(dtype=np.int64)
    temp, tw = s[m], w[m]
    for p in Ps:
        y.append(tw[temp <= p].sum(dtype=np.int64) / Vp)
    return Ps, y


def _parse_histogram(h, voxel_size=1, density=True):
    """metrics/_funcs.py:861-884."""
    delta_x = h[1]
    P = h[0]
    bin_widths = delta_x[1:] - delta_x[:-1]
    temp = P * (bin_widths)
    C = np.cumsum(temp[-1::-1])[-1::-1]
    S = P * (bin_widths)
    if not density:
        P /= np.max(P)
        temp_sum = np.sum(P * bin_widths)
        C /= temp_sum
        S /= temp_sum
    hist = Results()
    hist.pdf = P
    hist.cdf = C
    hist.relfreq = S
    hist.bin_centers = ((delta_x[1:] + delta_x[:-1]) / 2) * voxel_size
    hist.bin_edges = delta_x * voxel_size
    hist.bin_widths = (bin_widths) * voxel_size
    return hist


def pore_size_distribution(im, bins=10, log=True, voxel_size=1):
    r"""Pore-size distribution of a `porosimetry` / `local_thickness` map (metrics/_funcs.py:558-632): the
    histogram of the voxel values is the histogram of the distinct radii weighted by their voxel counts."""
    sm = IndexMap.from_array(im)
    s, _, w, _, _ = sm.representatives()
    keep = s > 0
    vals = s[keep] * voxel_size                                      # :619-620
    if log:
        vals = np.log10(vals)
    h = _parse_histogram(np.histogram(vals, bins=bins, weights=w[keep], density=True))
    cld = Results()
    cld[f"{log * 'Log' + 'R'}"] = h.bin_centers
    cld.pdf = h.pdf
    cld.cdf = h.cdf
    cld.satn = h.relfreq
    cld.bin_centers = h.bin_centers
    cld.bin_edges = h.bin_edges
    cld.bin_widths = h.bin_widths
    return cld


def pc_curve(im, sizes=None, pc=None, seq=None, sigma=0.072, theta=180, voxel_size=1):
    r"""Capillary pressure curve from an invasion-size map (metrics/_funcs.py:1073-1090, the `sizes` branch; the
    `seq` / `pc` branches are outside this path)."""
    if seq is not None or pc is not None or sizes is None:
        raise NotImplementedError("porespy_b200.pc_curve implements the `sizes` branch only")
    sm = IndexMap.from_array(sizes)
    s, m, w, _, _ = sm.representatives(im)
    if m is None:
        m = ~(s == 0)                                                # :1075
    sz = np.unique(s)[:0:-1]
    sz = np.hstack((sz[0] * 2, sz))
    x, y = [], []
    total = w[m == 1].sum(dtype=np.int64)                            # im.sum()
    for n in sz:
        r = n * voxel_size
        p = -2 * sigma * np.cos(np.deg2rad(theta)) / r
        x.append(p)
        snwp = w[(s >= n) * (m == 1)].sum(dtype=np.int64) / total
        y.append(snwp)
    res = Results()
    res.pc = x
    res.snwp = y
    return res
