"""Host-side arithmetic of the hot path: radii -> integer thresholds, exactly as numpy would
compare them in the reference (SURVEY 8(a) notes N1-N4, N6).  Pure numpy, no device code."""
import numpy as np

INF_U32 = 0xFFFFFFFF
_N_LIMIT = 1 << 32


def shape3(shape):
    """(nz, ny, nx) view of a squeezed 1-D / 2-D / 3-D shape."""
    shape = tuple(int(s) for s in shape)
    if len(shape) == 3:
        return shape
    if len(shape) == 2:
        return (1,) + shape
    if len(shape) == 1:
        return (1, 1) + shape
    raise ValueError(f"only 1-D, 2-D and 3-D images are supported, got shape {shape}")


def isqrt(v):
    """floor(sqrt(v)) for a non-negative integer (reach of a ball: W = isqrt(T - 1))."""
    import math
    return math.isqrt(int(v))


def dt_max_f32(max_d2):
    """np.amax(edt(im)) as the float32 scalar the reference sees (F:1132)."""
    if max_d2 == INF_U32:
        return np.float32(np.inf)
    return np.sqrt(np.float32(max_d2))


def reference_sizes(sizes, max_d2):
    """The radii array the reference loops over (F:1131-1134), same dtype, same order."""
    if isinstance(sizes, int):
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            return np.logspace(start=np.log10(dt_max_f32(max_d2)), stop=0, num=sizes)
    return np.unique(sizes)[-1::-1]


def threshold_of(r):
    """T(r) = min{n in N : np.sqrt(np.float32(n)) >= r}, compared with numpy's own promotion
    of a float32 array against the scalar `r` (float32 / float64 / int64 ...).  Returns None if
    no representable n satisfies it (r = nan, r = +inf, r beyond the supported range).

    seeds (dt >= r)  <=>  d2 >= T(r);   fill (edt(~seeds) < r)  <=>  d2' < T(r)   (note N3).
    """
    def ok(n):
        with np.errstate(invalid="ignore"):
            return bool((np.sqrt(np.array([n], dtype=np.float32)) >= r)[0])
    if ok(0):
        return 0
    hi = 1
    while not ok(hi):
        hi *= 2
        if hi >= _N_LIMIT:
            return None
    lo = hi // 2          # not ok(lo), ok(hi)
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if ok(mid):
            hi = mid
        else:
            lo = mid
    return hi


def effective_thresholds(radii, max_d2, has_background=True):
    """Descending radii -> the (threshold, radius) pairs that can change the result.

    Dropped: radii with T == 0 (r <= 0: fill is empty, N6), radii with no seeds (T > max d2,
    N4), and a radius whose T equals the previous kept one (same seeds, same fill: every voxel
    it could write is already written).  Returns (T uint32 array, radii float64 array)."""
    Ts, Rs = [], []
    last = None
    for r in radii:
        T = threshold_of(r)
        if T is None or T == 0:
            continue
        if max_d2 != INF_U32 and T > max_d2:
            continue
        if T == last:
            continue
        if last is not None and T > last:
            raise ValueError("radii must be visited in descending order")
        Ts.append(T)
        Rs.append(float(r))
        last = T
    return np.array(Ts, dtype=np.uint32), np.array(Rs, dtype=np.float64)


def normalise_inlets(inlets, shape):
    """Validation and conversion of `inlets` as trim_disconnected_blobs does it (F:1252-1259)."""
    if isinstance(inlets, tuple):
        where = np.copy(inlets)
        mask = np.zeros(shape, dtype=bool)
        mask[where] = True
        return mask
    inlets = np.asarray(inlets)
    if (inlets.shape == tuple(shape)) and (inlets.max() == 1):
        return inlets.astype(bool)
    raise Exception("inlets not valid, refer to docstring for info")


def border_faces(shape):
    """`get_border(shape, mode='faces')` (generators/_borders.py:93-100): every voxel with an index of 0
    or n - 1 along some axis."""
    out = np.zeros(tuple(shape), dtype=bool)
    for ax in range(out.ndim):
        sl = [slice(None)] * out.ndim
        for side in (0, -1):
            sl[ax] = side
            out[tuple(sl)] = True
    return out


def make_contiguous_symmetric(im):
    """`make_contiguous(im, mode='symmetric')` (tools/_funcs.py:842-847): positive values become their rank
    1..n among the positive values, negative ones -1..-m by magnitude, zeros stay."""
    im = np.array(im)

    def rank_positive(a):
        vals = np.unique(a)
        vals = vals[vals > 0]
        fw = np.zeros(int(a.max()) + 1 if a.size else 1, dtype=a.dtype)
        fw[vals] = np.arange(1, len(vals) + 1)
        return fw[a]

    return rank_positive(im * (im >= 0)) - rank_positive(-im * (im < 0))
