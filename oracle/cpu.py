"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the EDT -> local_thickness / porosimetry path.

Nothing under `porespy_b200/` imports this module.  Allowed callers: `tests/`,
`__graft_entry__.smoke()`, and `bench.py`'s `cpu_baseline` / `--impl reference` legs.

It restates, on the CPU and in *float semantics* (float32 distance map compared against
the radius objects exactly the way numpy does it in the reference), the algorithm of

* `porosimetry`              /root/reference/src/porespy/filters/_funcs.py:1124-1148, 1177-1212
* `local_thickness`          /root/reference/src/porespy/filters/_funcs.py:1027-1029
* `trim_disconnected_blobs`  /root/reference/src/porespy/filters/_funcs.py:1252-1270
* `get_border(mode='faces')` /root/reference/src/porespy/generators/_borders.py:93-100
* `ps_round/ps_ball/ps_disk` /root/reference/src/porespy/tools/_funcs.py:1149-1156
* `fftmorphology` dilation   /root/reference/src/porespy/filters/_fftmorphology.py:75-93
* `blobs` / `norm_to_uniform`/root/reference/src/porespy/generators/_imgen.py:1023-1051,
                             /root/reference/src/porespy/tools/_funcs.py:963-969
* `edt.edt` (third-party, unpinned, not vendored) -> `oracle/edt_oracle.c`.

Parity pinning: `tests/test_oracle.py` checks these restatements against the golden
vectors in `tests/golden/`, which `tests/golden/make_golden.py` produced by running the
reference's own source (via `oracle/ref_shim.py`), and against the golden numbers the
reference's tests assert (test/unit/test_filters.py:36-42, 53-56, 266-279;
test/unit/test_tools.py:309-316).

The product (CUDA) path works on integer squared distances and integer thresholds; this
oracle deliberately does NOT -- it compares float32 distances with the radii like the
reference, so the product's threshold logic is checked rather than mirrored.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
INF_U32 = 0xFFFFFFFF


def build(force=False):
    """Compile oracle/edt_oracle.c -> oracle/liboracle.so (gcc + OpenMP)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "edt_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.oracle_edt_sq.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        lib.oracle_edt_sq.restype = ctypes.c_int
        lib.oracle_sqrt_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_int]
        lib.oracle_sqrt_f32.restype = ctypes.c_int
        lib.oracle_num_threads.restype = ctypes.c_int
        lib.oracle_porosimetry_dt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                              ctypes.c_int]
        lib.oracle_porosimetry_dt.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def num_threads():
    return int(_lib().oracle_num_threads())


def _as3d(shape):
    if len(shape) == 1:
        return (1, 1, shape[0])
    if len(shape) == 2:
        return (1, shape[0], shape[1])
    if len(shape) == 3:
        return tuple(shape)
    raise ValueError("oracle supports 1-D, 2-D and 3-D arrays")


def edt_sq(data, nthreads=0):
    """Exact integer squared EDT (uint32; INF_U32 where no background exists)."""
    a = np.ascontiguousarray(np.asarray(data) != 0, dtype=np.uint8)
    out = np.empty(a.shape, dtype=np.uint32)
    if a.size == 0:
        return out
    nz, ny, nx = _as3d(a.shape)
    rc = _lib().oracle_edt_sq(a.ctypes.data, out.ctypes.data, nz, ny, nx, int(nthreads))
    assert rc == 0
    return out


def edt(data, anisotropy=None, black_border=False, order="K", parallel=1,
        voxel_graph=None):
    """`edt.edt` as PoreSpy calls it: float32(sqrt(d2)); `parallel<=0` = all cores."""
    if anisotropy is not None or black_border or voxel_graph is not None:
        raise NotImplementedError("oracle restates only the options PoreSpy uses")
    d2 = edt_sq(data, nthreads=0 if parallel <= 0 else parallel)
    out = np.empty(d2.shape, dtype=np.float32)
    if d2.size:
        _lib().oracle_sqrt_f32(d2.ctypes.data, out.ctypes.data, d2.size, 0)
    return out


# ----------------------------------------------------------------------------- inputs
def norm_to_uniform(im, scale=None):
    """T:963-969 -- normal -> uniform greyscale via the error function."""
    from scipy.special import erfc
    lo, hi = (im.min(), im.max()) if scale is None else scale
    im = (im - np.mean(im)) / np.std(im)
    im = 1 / 2 * erfc(-im / np.sqrt(2))
    im = (im - im.min()) / (im.max() - im.min())
    return im * (hi - lo) + lo


def blobs(shape, porosity=0.5, blobiness=1, seed=None):
    """_imgen.py:1023-1051 -- noise, gaussian blur, uniformise, threshold (divs=1)."""
    import scipy.ndimage as spim
    if seed is not None:
        np.random.seed(seed)
    if isinstance(shape, int):
        shape = [shape] * 3
    if len(shape) == 1:
        shape = [shape[0]] * 3
    shape = np.array(shape)
    if isinstance(blobiness, int):
        blobiness = [blobiness] * len(shape)
    sigma = np.mean(shape) / (40 * np.array(blobiness))
    field = spim.gaussian_filter(np.random.random(shape), sigma=sigma)
    field = norm_to_uniform(field, scale=[0, 1])
    return field < porosity if porosity else field


def border_faces(shape, thickness=1):
    """_borders.py:93-100 -- True on every face voxel (2-D / 3-D; other ndim: all True)."""
    t = thickness
    out = np.ones(shape, dtype=bool)
    if len(shape) == 2:
        out[t:-t, t:-t] = False
    elif len(shape) == 3:
        out[t:-t, t:-t, t:-t] = False
    return out


def _cross(ndim):
    import scipy.ndimage as spim
    return spim.generate_binary_structure(ndim, 1)   # == skimage ball(1) / disk(1)


def _full(ndim):
    return np.ones((3,) * ndim, dtype=bool)          # == skimage cube(3) / square(3)


def ps_round(r, ndim, smooth=True):
    """T:1149-1156 -- EDT-defined ball/disk; strict `<` when smooth."""
    half = int(np.ceil(r))
    probe = np.ones([2 * half + 1] * ndim, dtype=bool)
    probe[(half,) * ndim] = False
    d = edt(probe)
    return d < r if smooth else d <= r


# ------------------------------------------------------------------------ the hot path
def trim_disconnected_blobs(im, inlets, strel=None):
    """F:1252-1270 -- keep foreground whose `strel`-component of (inlets | im) holds an inlet."""
    import scipy.ndimage as spim
    if isinstance(inlets, tuple):
        where = np.copy(inlets)
        inlets = np.zeros_like(im, dtype=bool)
        inlets[where] = True
    elif (inlets.shape == im.shape) and (inlets.max() == 1):
        inlets = inlets.astype(bool)
    else:
        raise Exception("inlets not valid, refer to docstring for info")
    if strel is None:
        strel = _full(im.ndim)
    lab = spim.label(inlets + (im > 0), structure=strel)[0]
    wanted = np.unique(lab[inlets])
    wanted = wanted[wanted > 0]
    return np.isin(lab, wanted) * im


def _conn_strel(ndim, conn):
    """F:393-406 -- conn 4/6: cross, None/8/26: full cube; anything else is the reference's exception."""
    small, big = (4, 8) if ndim == 2 else (6, 26)
    if conn == small:
        return _cross(ndim)
    if conn in (None, big):
        return _full(ndim)
    raise Exception("Received conn is not valid")


def find_disconnected_voxels(im, conn=None, surface=False):
    """F:391-421 -- label with the conn strel; holes = labels that do not touch the border
    (skimage clear_border), or with surface=True the labels that do not touch EVERY face.  Label 0
    (the background) goes through the same set arithmetic as in the reference: it is a hole when
    some face has no background voxel at all."""
    import scipy.ndimage as spim
    im = np.asarray(im)
    if im.ndim not in (2, 3):
        raise Exception("Received conn is not valid")      # (the reference falls through to an unbound strel)
    labels = spim.label(im, structure=_conn_strel(im.ndim, conn))[0]
    if not surface:
        touching = set()
        for ax in range(labels.ndim):
            for side in (0, -1):
                touching.update(np.unique(np.take(labels, side, axis=ax)).tolist())
        touching.discard(0)
        return (labels > 0) & ~np.isin(labels, list(touching))
    keep = set(np.unique(labels).tolist())
    for ax in range(labels.ndim):
        keep.intersection_update(np.unique(np.take(labels, 0, axis=ax)).tolist())
        keep.intersection_update(np.unique(np.take(labels, -1, axis=ax)).tolist())
    return np.isin(labels, list(keep), invert=True)


def fill_blind_pores(im, conn=None, surface=False):
    """F:459-462."""
    im = np.copy(im)
    im[find_disconnected_voxels(im, conn=conn, surface=surface)] = False
    return im


def trim_floating_solid(im, conn=None, surface=False):
    """F:500-503."""
    im = np.copy(im)
    im[find_disconnected_voxels(~im, conn=conn, surface=surface)] = True
    return im


def trim_nonpercolating_paths(im, inlets, outlets, strel=None):
    """F:550-555 -- components (scipy's default cross connectivity unless `strel`) that hold an inlet
    voxel AND an outlet voxel."""
    import scipy.ndimage as spim
    labels = spim.label(im, structure=strel)[0]
    IN = np.unique(labels * inlets)
    OUT = np.unique(labels * outlets)
    hits = np.array(list(set(IN.tolist()).intersection(set(OUT.tolist()))))
    return np.isin(labels, hits[hits > 0]) if hits.size else np.zeros(labels.shape, dtype=bool)


def make_contiguous_symmetric(im):
    """T:842-847 (`make_contiguous(mode='symmetric')`): positive values are ranked 1..n, negative values
    -1..-m by magnitude, zeros stay."""
    im = np.array(im)

    def relabel(a):                      # skimage relabel_sequential: sorted unique positive values -> 1..n
        vals = np.unique(a)
        vals = vals[vals > 0]
        fw = np.zeros(int(a.max()) + 1 if a.size else 1, dtype=a.dtype)
        fw[vals] = np.arange(1, len(vals) + 1)
        return fw[a]

    return relabel(im * (im >= 0)) - relabel(-im * (im < 0))


def find_trapped_regions(seq, outlets=None, bins=25, return_mask=True):
    """F:115-147 -- for every bin value i (descending): the voxels with seq >= i whose component (scipy's
    default cross connectivity) holds no outlet voxel are trapped."""
    import scipy.ndimage as spim
    seq = np.copy(seq)
    if outlets is None:
        outlets = border_faces(seq.shape)
    trapped = np.zeros_like(outlets)
    if bins is None:
        bins = np.unique(seq)[-1::-1]
        bins = bins[bins > 0]
    elif isinstance(bins, int):
        bins = np.linspace(seq.max(), 1, bins)
    for i in bins:
        temp = seq >= i
        labels = spim.label(temp)[0]
        keep = np.setdiff1d(np.unique(labels[outlets]), np.array([0]))
        trapped += temp * np.isin(labels, keep, invert=True)
    if return_mask:
        return trapped
    seq[trapped] = -1
    return make_contiguous_symmetric(seq)


def _dilate_fft(mask, strel):
    """_fftmorphology.py:75-93 -- zero-pad by 1, fftconvolve 'same' > 0.1, crop."""
    from scipy.signal import fftconvolve
    padded = np.pad(mask, pad_width=1, mode="constant", constant_values=0)
    hit = fftconvolve(padded, strel, mode="same") > 0.1
    return hit[(slice(1, -1),) * mask.ndim]


def porosimetry(im, sizes=25, inlets=None, access_limited=True, mode="hybrid", divs=1,
                nthreads=0):
    """F:1124-1148 + loop bodies F:1177-1192 ('dt') and F:1193-1209 ('hybrid').

    `divs` is accepted and ignored (results are chunk-invariant, F:1517-1520);
    mode 'mio' is outside the path.
    """
    if mode not in ("dt", "hybrid"):
        if mode == "mio":
            raise NotImplementedError("mode 'mio' is outside the oracle's scope")
        raise Exception("Unrecognized mode " + mode)
    im = np.squeeze(im)
    par = nthreads if nthreads > 0 else 0
    dt = edt(im > 0, parallel=par)
    if inlets is None:
        inlets = border_faces(im.shape)
    if isinstance(sizes, int):
        sizes = np.logspace(start=np.log10(np.amax(dt)), stop=0, num=sizes)
    else:
        sizes = np.unique(sizes)[-1::-1]
    conn = _cross(im.ndim)
    out = np.zeros(np.shape(im))
    for r in sizes:
        seeds = dt >= r
        if access_limited:
            seeds = trim_disconnected_blobs(seeds, inlets, strel=conn)
        if not np.any(seeds):
            continue
        if mode == "dt":
            fill = edt(~seeds, parallel=par) < r
        else:
            fill = _dilate_fft(seeds, ps_round(r, im.ndim))
        out[(out == 0) * fill] = r
    return out


def local_thickness(im, sizes=25, mode="hybrid", divs=1, nthreads=0):
    """F:1027-1029 -- porosimetry without access limitation."""
    return porosimetry(im, sizes=sizes, access_limited=False, mode=mode, divs=divs,
                       nthreads=nthreads)


# ------------------------------------------------------- the same loop in one C call (big volumes)
def porosimetry_c(im, sizes=25, inlets=None, access_limited=True, nthreads=0):
    """`porosimetry(mode='dt')` with the radius loop in C (oracle_porosimetry_dt, edt_oracle.c): the same
    statements F:1124-1192 without numpy temporaries, for volumes where the restatement above would take
    many minutes.  The prologue (squeeze, radii) is the numpy code of F:1124-1134; tests/test_oracle.py
    pins the C loop against `porosimetry` above.  `nthreads=0`: every hardware thread (independent of
    OMP_NUM_THREADS, which torchrun sets to 1)."""
    im = np.squeeze(im)
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    fg = np.ascontiguousarray(im > 0, dtype=np.uint8)
    if isinstance(sizes, int):
        dtmax = np.amax(edt(fg, parallel=nthreads))
        sizes = np.logspace(start=np.log10(dtmax), stop=0, num=sizes)
    else:
        sizes = np.unique(sizes)[-1::-1]
    radii = np.ascontiguousarray(sizes, dtype=np.float64)        # float32 / int64 -> float64 is exact
    inl = None
    if access_limited:
        if inlets is None:
            inlets = border_faces(im.shape)
        if isinstance(inlets, tuple):
            where = np.copy(inlets)
            inlets = np.zeros_like(im, dtype=bool)
            inlets[where] = True
        elif not ((inlets.shape == im.shape) and (inlets.max() == 1)):
            raise Exception("inlets not valid, refer to docstring for info")
        inl = np.ascontiguousarray(inlets, dtype=bool).view(np.uint8)
    out = np.empty(im.shape, dtype=np.float64)
    nz, ny, nx = _as3d(im.shape)
    rc = _lib().oracle_porosimetry_dt(fg.ctypes.data, radii.ctypes.data, len(radii),
                                      inl.ctypes.data if inl is not None else None, out.ctypes.data,
                                      nz, ny, nx, int(nthreads))
    assert rc == 0, rc
    return out


def local_thickness_c(im, sizes=25, nthreads=0):
    return porosimetry_c(im, sizes=sizes, access_limited=False, nthreads=nthreads)


# ------------------------------------------- radius-map post-processing, plain numpy
def size_to_seq(size, im=None, bins=None, mode="drainage"):
    """filters/_size_seq_satn.py:62-83."""
    solid = (size == 0) if im is None else (im == 0)
    uninvaded = size == -1
    if bins is None:
        bins = np.unique(size)
    elif isinstance(bins, int):
        bins = np.linspace(0, size.max(), bins)
    vals = np.digitize(size, bins=bins, right=True)
    if mode.startswith("im"):
        vals[solid] = 0
        vals[uninvaded] = -1
        vals = make_contiguous_symmetric(vals)
    if mode.startswith("dr"):
        vals = make_contiguous_symmetric(vals)
        vals = vals.max() + 1 - vals
        vals[solid] = 0
        vals[uninvaded] = -1
    return vals


def size_to_satn(size, im=None, bins=None, mode="drainage"):
    """filters/_size_seq_satn.py:134-149."""
    if bins is None:
        bins = np.unique(size[size > 0])
    elif isinstance(bins, int):
        bins = np.linspace(0, size.max(), bins)
    if im is None:
        im = ~(size == 0)
    void_vol = im.sum()
    satn = -np.ones_like(size, dtype=float)
    if mode.startswith("im"):
        for r in bins:
            hits = (size <= r) * (size > 0)
            satn[hits * (satn == -1)] = hits.sum() / void_vol
    elif mode.startswith("dr"):
        for r in bins[-1::-1]:
            hits = (size >= r) * (size > 0)
            satn[hits * (satn == -1)] = hits.sum() / void_vol
    satn *= (im > 0)
    return satn


def seq_to_satn(seq, im=None, mode="drainage"):
    """filters/_size_seq_satn.py:196-221 (rankdata 'dense' - 1 as the integer dense rank, see oracle/ref_shim.py)."""
    seq = np.copy(seq).astype(int)
    solid_mask = (seq == 0) if im is None else (im == 0)
    uninvaded_mask = seq == -1
    seq[seq <= 0] = 0
    if mode.startswith("im"):
        seq = seq.max() - seq + 1
        seq[solid_mask] = 0
        seq[uninvaded_mask] = 0
    seq = np.unique(seq, return_inverse=True)[1].reshape(-1)
    b = np.bincount(seq)
    if (solid_mask.sum(dtype=np.int64) > 0) or (uninvaded_mask.sum(dtype=np.int64) > 0):
        b[0] = 0
    c = np.cumsum(b)
    seq = np.reshape(seq, solid_mask.shape)
    satn = c[seq] / (seq.size - solid_mask.sum(dtype=np.int64))
    satn[solid_mask] = 0
    satn[uninvaded_mask] = -1
    return satn


def pore_size_distribution(im, bins=10, log=True, voxel_size=1):
    """metrics/_funcs.py:619-632 + _parse_histogram :861-884 -> dict of arrays."""
    im = im.flatten()
    vals = im[im > 0] * voxel_size
    if log:
        vals = np.log10(vals)
    P, edges = np.histogram(vals, bins=bins, density=True)
    widths = edges[1:] - edges[:-1]
    return dict(pdf=P, cdf=np.cumsum((P * widths)[-1::-1])[-1::-1], satn=P * widths,
                bin_centers=((edges[1:] + edges[:-1]) / 2) * 1, bin_edges=edges * 1, bin_widths=widths * 1)


def pc_curve_sizes(im, sizes, sigma=0.072, theta=180, voxel_size=1):
    """metrics/_funcs.py:1073-1090."""
    if im is None:
        im = ~(sizes == 0)
    sz = np.unique(sizes)[:0:-1]
    sz = np.hstack((sz[0] * 2, sz))
    x, y = [], []
    for n in sz:
        r = n * voxel_size
        x.append(-2 * sigma * np.cos(np.deg2rad(theta)) / r)
        y.append(((sizes >= n) * (im == 1)).sum(dtype=np.int64) / im.sum(dtype=np.int64))
    return np.asarray(x), np.asarray(y)


# ----------------------------------------------------- simulations.drainage, plain numpy
def pc_to_satn(pc, im, mode="drainage"):
    """filters/_size_seq_satn.py:338-342."""
    a = np.digitize(pc, bins=np.unique(pc))
    a[~im] = 0
    a[np.where(pc == np.inf)] = -1
    return seq_to_satn(seq=a, im=im, mode=mode)


def satn_to_seq(satn, im=None, mode="drainage"):
    """filters/_size_seq_satn.py:384-399."""
    if im is None:
        im = satn > 0
    uninvaded = satn == -1
    values = np.unique(satn)
    seq = np.digitize(satn, bins=values)
    seq[satn == -1] = -1
    seq[~im] = 0
    seq = make_contiguous_symmetric(seq)
    if mode.startswith("im"):
        seq = (seq.max() + 1) - seq
        seq[~im] = 0
    seq[uninvaded] = -1
    return seq


def pc_curve_pc(im, pc):
    """metrics/_funcs.py:1091-1108 (the `pc` branch of pc_curve) -> (Ps, snwp)."""
    Ps = np.unique(pc[im])
    if Ps[-1] == np.inf:
        Ps[-1] = Ps[-2] * 2
    if Ps[0] == -np.inf:
        Ps[0] = Ps[1] - np.abs(Ps[1] / 2)
    else:
        Ps = np.hstack((Ps[0] - np.abs(Ps[0] / 2), Ps))
    Vp = im.sum(dtype=np.int64)
    temp = pc[im]
    return Ps, [(temp <= p).sum(dtype=np.int64) / Vp for p in Ps]


def _insert_spheres(inv, new, radii_map, v):
    """tools/_sphere_insertions.py:327-385 for all new points at once: the union of the balls
    {o : sqrt(|o|^2) <= r - 0.001} around the points of radius r, written where inv == 0."""
    import scipy.ndimage as spim
    for r in np.unique(radii_map[new]):
        r = int(r)
        if r < 1:
            continue
        ax = np.arange(-r, r + 1)
        grids = np.meshgrid(*([ax] * inv.ndim), indexing="ij")
        ball = np.sqrt(sum(gg.astype(float) ** 2 for gg in grids)) <= r - 0.001
        cover = spim.binary_dilation(new & (radii_map == r), structure=ball)
        inv[cover & (inv == 0)] = v
    return inv


def drainage(im, voxel_size, pc=None, inlets=None, outlets=None, residual=None, bins=25, delta_rho=1000, g=9.81,
             sigma=0.072, theta=180):
    """simulations/_drainage.py:104-185 -> dict(im_pc, im_satn, im_trapped, pc, snwp)."""
    im = np.array(im, dtype=bool)
    dt = edt(im)
    if pc is None:
        with np.errstate(divide="ignore", invalid="ignore"):
            pc = -(im.ndim - 1) * sigma * np.cos(np.deg2rad(theta)) / (dt * voxel_size)
    else:
        pc = np.array(pc, dtype=np.float64)
    pc[~im] = 0
    h = np.ones_like(im, dtype=bool)
    h[0, ...] = False
    h = (edt(h) + 1) * voxel_size
    rgh = delta_rho * g * h
    fn = pc + rgh
    if inlets is None:
        inlets = np.zeros_like(im)
        inlets[0, ...] = True
    if isinstance(bins, int):
        vmax = fn[fn < np.inf].max()
        vmin = fn[im][fn[im] > -np.inf].min()
        Ps = np.linspace(vmin, vmax * 1.1, bins)
    else:
        Ps = bins
    inv = np.zeros_like(im, dtype=float)
    seeds = np.zeros_like(im, dtype=bool)
    radii_map = dt.astype(int)
    mask = None
    if (residual is not None) and (outlets is not None):
        mask = im * (~residual)
        mask = trim_disconnected_blobs(mask, inlets=inlets)
    for p in Ps:
        temp = (fn <= p) * im
        if residual is not None:
            temp = temp + residual
        new_seeds = trim_disconnected_blobs(temp, inlets=inlets)
        if mask is not None:
            new_seeds = new_seeds * mask
        temp = new_seeds * (~seeds)
        seeds += new_seeds
        inv = _insert_spheres(inv, temp, radii_map, p)
    inv[(inv == 0) * im] = np.inf
    if residual is not None:
        inv[residual] = -np.inf
    trapped = None
    satn = pc_to_satn(pc=inv, im=im)
    if outlets is not None:
        seq = satn_to_seq(satn=satn, im=im)
        trapped = find_trapped_regions(seq=seq, outlets=outlets)
        trapped[seq == -1] = True
        inv[trapped] = np.inf
        if residual is not None:
            inv[residual] = -np.inf
        satn = pc_to_satn(pc=inv, im=im)
    Pc, snwp = pc_curve_pc(im, inv)
    return dict(im_pc=inv, im_satn=satn, im_trapped=trapped, pc=Pc, snwp=snwp)
