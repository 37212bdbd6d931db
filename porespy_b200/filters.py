"""Drop-in replacements for the reference's hot-path filters, same signatures and results:

* `porosimetry`             /root/reference/src/porespy/filters/_funcs.py:1032-1212
* `local_thickness`         /root/reference/src/porespy/filters/_funcs.py:947-1029
* `trim_disconnected_blobs` /root/reference/src/porespy/filters/_funcs.py:1215-1270
* the other users of the same flood: `find_disconnected_voxels` :352-421,
  `fill_blind_pores` :424-462, `trim_floating_solid` :465-503, `trim_nonpercolating_paths` :506-555,
  `find_trapped_regions` :73-147

Host code here only does what numpy does in the reference prologue (squeeze, radii, inlet
validation); all voxel work runs in libpsb200.so on the GPU.  There is no CPU fallback.
"""
import logging

import numpy as np

from . import _device as dev
from . import _host as host
from . import _lib

logger = logging.getLogger(__name__)

__all__ = ["porosimetry", "local_thickness", "trim_disconnected_blobs", "find_disconnected_voxels",
           "fill_blind_pores", "trim_floating_solid", "trim_nonpercolating_paths", "find_trapped_regions",
           "size_to_seq", "size_to_satn", "seq_to_satn"]


def _result_for_no_background(shape, radii):
    """Image without any background voxel: edt == +inf everywhere (black_border=False), so every
    radius r with `inf >= r` seeds the whole image and `edt(all False) == 0 < r` fills it: the
    first positive, non-NaN radius wins everywhere (parity unpinned upstream)."""
    out = np.zeros(shape)
    for r in radii:
        if np.inf >= r and 0 < r:      # both False for NaN
            out[...] = r
            break
    return out


# numpy results of large 3-D volumes without access limitation: the radius loop runs z-slab by z-slab (halo = the
# largest reach, so every slab is exact on its own -- DESIGN.md lemma i) and the host epilogue of a finished slab
# (index bytes over PCIe, widened to float64 by host threads: 8 B/voxel of host-memory writes, the longest leg of
# the numpy -> numpy call) overlaps the kernels of the next one.
SLAB_PIPELINE = {"min_voxels": 1 << 28, "slabs": 6, "enabled": True}


def _slab_plan(shape3_, T):
    """[(z0, z1, e0, e1)] -- own planes and extended planes of every slab -- or None when slabs do not pay."""
    nz, ny, nx = shape3_
    cfg = SLAB_PIPELINE
    if not cfg["enabled"] or nz * ny * nx < cfg["min_voxels"] or len(T) == 0 or len(T) > _lib.MAX_THRESHOLDS:
        return None
    W = host.isqrt(int(T[0]) - 1)                       # thresholds descend: the first radius reaches farthest
    nslabs = int(cfg["slabs"])
    if nslabs < 2 or nz // nslabs < 2 * W + 1:          # halo work above the slab's own
        return None
    # a thin first slab: the host epilogue (the longest leg) starts as early as possible; the rest evenly
    first = max(2 * W + 1, nz // (4 * nslabs))
    cuts = [0] + [first + ((nz - first) * i) // (nslabs - 1) for i in range(nslabs)]
    return [(z0, z1, max(0, z0 - W), min(nz, z1 + W)) for z0, z1 in zip(cuts[:-1], cuts[1:])]


def _run_loop_slabs(ctx, d2, shape, T, R, plan):
    from concurrent.futures import ThreadPoolExecutor
    torch = dev._torch()
    nz, ny, nx = host.shape3(shape)
    plane = ny * nx
    n = nz * plane
    lut = np.concatenate([[0.0], R])
    out = torch.empty(n, dtype=torch.float64, pin_memory=True)
    side = torch.cuda.Stream(device=d2.device)
    ws = torch.empty(2 * (1 << 22) * 8 + 256, dtype=torch.uint8, device=d2.device)       # epilogue's own device chunks
    comp = torch.cuda.current_stream()
    with ThreadPoolExecutor(max_workers=1) as pool:      # one epilogue at a time: each uses every host thread
        jobs = []
        for z0, z1, e0, e1 in plan:
            idx = torch.empty((e1 - e0) * plane, dtype=torch.uint8, device=d2.device)
            dev.local_thickness_idx(ctx, d2[e0 * plane:e1 * plane], T, idx, None, _lib.INLETS_NONE, 3, (e1 - e0, ny, nx), 0)
            ready = torch.cuda.Event()
            ready.record(comp)
            jobs.append(pool.submit(dev.expand_idx_slice_to_host, ctx, idx, (z0 - e0) * plane, (z1 - z0) * plane, lut,
                                    out, z0 * plane, ready, side, ws))
        for j in jobs:
            j.result()
    return out.numpy().reshape(shape)


def _run_loop(ctx, d2, shape, ndim, T, R, inlets_u8, inlet_mode, as_numpy=True, as_index=False):
    """Device loop over the effective thresholds (in groups of <= 253) -> float64 radius map, or with
    `as_index` the map in index form (sizemap.IndexMap: index byte per voxel + radius table)."""
    torch = dev._torch()
    n = int(np.prod(shape))
    idx = torch.empty(n, dtype=torch.uint8, device=d2.device)
    G = _lib.MAX_THRESHOLDS
    ngroups = max(1, -(-len(T) // G))
    if as_index:
        from .sizemap import IndexMap
        if ngroups == 1:
            dev.local_thickness_idx(ctx, d2, T, idx, inlets_u8, inlet_mode, ndim, shape, 0)
            return IndexMap(ctx, idx, np.concatenate([[0.0], R]), shape)
        return IndexMap.from_array(_run_loop(ctx, d2, shape, ndim, T, R, inlets_u8, inlet_mode, as_numpy=False), ctx)
    if ngroups == 1 and as_numpy and inlet_mode == _lib.INLETS_NONE and ndim == 3:
        plan = _slab_plan(host.shape3(shape), T)
        if plan is not None:
            del idx
            return _run_loop_slabs(ctx, d2, shape, T, R, plan)
    if ngroups == 1 and as_numpy:
        # common case: the float64 map (F:1178) is only materialised on the host
        dev.local_thickness_idx(ctx, d2, T, idx, inlets_u8, inlet_mode, ndim, shape, 0)
        return dev.expand_idx_to_host(ctx, idx, np.concatenate([[0.0], R]), shape)
    out = torch.empty(n, dtype=torch.float64, device=d2.device)
    for g in range(ngroups):
        Tg, Rg = T[g * G:(g + 1) * G], R[g * G:(g + 1) * G]
        flags = 0
        if g > 0:
            dev.mark_written(ctx, out, idx)
            flags = _lib.FLAG_IDX_PREINIT
        dev.local_thickness_idx(ctx, d2, Tg, idx, inlets_u8, inlet_mode, ndim, shape, flags)
        lut = np.concatenate([[0.0], Rg])
        dev.expand_idx(ctx, idx, lut, out, merge=(g > 0))
    del idx
    out = out.view(*shape) if len(shape) else out
    if not as_numpy:
        return out
    return dev.to_host(out)


def porosimetry(im, sizes: int = 25, inlets=None, access_limited: bool = True,
                mode: str = "hybrid", divs=1, _as_index=False):
    r"""Porosimetry simulation (sphere insertion from the inlets); see the reference docstring
    (F:1040-1122) for the meaning of every argument -- they are unchanged.

    `mode`: 'hybrid', 'dt' and 'mio' give identical results in the reference (its own tests
    assert it, test/unit/test_filters.py:27-51); all three run the same exact integer
    pipeline here.  `divs` is accepted for compatibility and ignored (the result of the
    reference's chunked path is chunk-invariant by construction, F:1517-1520).
    """
    torch = dev._torch()
    as_numpy = not isinstance(im, torch.Tensor)
    if mode not in ("hybrid", "dt", "mio"):
        raise Exception("Unrecognized mode " + mode)
    if as_numpy:
        im = np.squeeze(np.asarray(im))
    else:
        im = torch.squeeze(im)
    shape = tuple(int(s) for s in im.shape)
    ndim = len(shape)
    if ndim == 0 or int(np.prod(shape)) == 0:
        return np.zeros(shape)
    ctx = _lib.context()
    as_index = _as_index
    im_u8 = dev.to_device_u8(im, ctx, positive=True)          # F:1126 edt(im > 0)
    d2, max_d2 = dev.edt_run(ctx, im_u8, shape, want_max=True)    # max fused into the last pass
    del im_u8
    radii = host.reference_sizes(sizes, max_d2)

    inlet_mode, inlets_u8 = _lib.INLETS_NONE, None
    if access_limited:
        if inlets is None:
            inlet_mode = _lib.INLETS_FACES
        else:
            if isinstance(inlets, torch.Tensor):
                # device-resident mask: the same validation as F:1256-1259 without a host round trip
                if tuple(inlets.shape) != shape or int(inlets.max().item()) != 1:
                    raise Exception("inlets not valid, refer to docstring for info")
                mask = inlets
            else:
                mask = host.normalise_inlets(inlets, shape)
            inlet_mode, inlets_u8 = _lib.INLETS_MASK, dev.to_device_u8(mask, ctx)

    if max_d2 == host.INF_U32:
        res = _result_for_no_background(shape, radii)
        if as_index:
            from .sizemap import IndexMap
            return IndexMap.from_array(res, ctx)
        return res if as_numpy else torch.from_numpy(res).to(d2.device)
    T, R = host.effective_thresholds(radii, max_d2)
    return _run_loop(ctx, d2, shape, ndim, T, R, inlets_u8, inlet_mode, as_numpy=as_numpy, as_index=as_index)


def local_thickness(im, sizes: int = 25, mode: str = "hybrid", divs: int = 1):
    r"""Radius of the largest sphere that covers each voxel and fits in the foreground.
    Identical to `porosimetry(..., access_limited=False)` (F:1027-1029)."""
    return porosimetry(im=im, sizes=sizes, access_limited=False, mode=mode, divs=divs)


def trim_disconnected_blobs(im, inlets, strel=None):
    r"""Removes foreground voxels not connected to the inlets (F:1215-1270).

    `strel` is the connectivity neighbourhood.  The reference passes it to
    `scipy.ndimage.label`; here the 3x3(x3) structuring elements PoreSpy itself uses are
    recognised: the cross (`ball(1)`/`disk(1)`: 6-/4-connectivity) and the full cube
    (`cube(3)`/`square(3)`: 26-/8-connectivity, the default).  Other shapes are rejected.
    """
    im = np.asarray(im)
    if im.ndim not in (2, 3):
        raise ValueError("trim_disconnected_blobs supports 2-D and 3-D images")
    mask = host.normalise_inlets(inlets, im.shape)
    full = 26 if im.ndim == 3 else 8
    cross = 6 if im.ndim == 3 else 4
    if strel is None:
        conn = full
    else:
        s = np.asarray(strel) != 0
        if s.shape != (3,) * im.ndim:
            raise NotImplementedError("only 3x3(x3) connectivity structuring elements are supported")
        grid = np.indices(s.shape) - 1
        if np.array_equal(s, np.abs(grid).sum(axis=0) <= 1):
            conn = cross
        elif s.all():
            conn = full
        else:
            raise NotImplementedError("strel must be the cross (ball(1)/disk(1)) or the full cube")
    if im.size == 0:
        return np.zeros(im.shape, dtype=im.dtype)
    ctx = _lib.context()
    fg = dev.to_device_u8(im > 0, ctx)
    inl = dev.to_device_u8(mask, ctx)
    shape3 = host.shape3(im.shape)
    keep = dev.to_host(dev.flood(ctx, fg, inl, conn, shape3).view(*im.shape)).astype(bool)
    return keep * im


# ------------------------------------------------------------------ other users of the flood
def _conn_of(ndim, conn):
    """F:393-406: 4 / 6 = faces only, None / 8 / 26 = full neighbourhood."""
    small, big = (4, 8) if ndim == 2 else (6, 26)
    if conn == small:
        return small
    if conn in (None, big):
        return big
    raise Exception("Received conn is not valid")


def _strel_conn(ndim, strel, default):
    """Connectivity of a 3x3(x3) structuring element as `scipy.ndimage.label` would use it."""
    full = 26 if ndim == 3 else 8
    cross = 6 if ndim == 3 else 4
    if strel is None:
        return default
    s = np.asarray(strel) != 0
    if s.shape != (3,) * ndim:
        raise NotImplementedError("only 3x3(x3) connectivity structuring elements are supported")
    grid = np.indices(s.shape) - 1
    if np.array_equal(s, np.abs(grid).sum(axis=0) <= 1):
        return cross
    if s.all():
        return full
    raise NotImplementedError("strel must be the cross (ball(1)/disk(1)) or the full cube")


def _face_views(t):
    """The 2 * ndim faces of a device tensor as (index tuple) list."""
    out = []
    for ax in range(t.dim()):
        for side in (0, -1):
            out.append((slice(None),) * ax + (side,))
    return out


def _mask_to_device(a, ctx, shape):
    """Host array -> 0/1 uint8 device tensor of `shape` without a host-side copy for bool input (the bytes of
    a numpy bool array already are 0/1; anything else is normalised on the device)."""
    torch = dev._torch()
    a = np.asarray(a)
    t = dev.to_device_u8(a, ctx).view(*shape)
    if a.dtype != np.bool_:
        t = (t != 0).to(torch.uint8)
    return t


def _mask_to_host(t):
    """0/1 device tensor -> numpy bool (the downloaded bytes are reinterpreted, not converted)."""
    torch = dev._torch()
    return dev.to_host(t.to(torch.uint8)).view(np.bool_)


def _reached_from(ctx, fg, seeds, conn):
    """Foreground voxels connected (within the foreground) to the non-zero voxels of `seeds` (a subset
    of the foreground): psb200_flood with the seeds as inlets."""
    return dev.flood(ctx, fg.reshape(-1), seeds.reshape(-1), conn, host.shape3(tuple(fg.shape))).view(fg.shape)


def find_disconnected_voxels(im, conn=None, surface=False):
    r"""Voxels of `im` that are not connected to the image border (F:352-421); with `surface=True`
    the voxels of every region that does not touch ALL faces.  Same arguments, same quirks: a region
    is a `scipy.ndimage.label` component under the `conn` neighbourhood, and with `surface=True`
    the background is reported as well when some face holds no background voxel (its label 0 then
    drops out of the reference's `keep` set)."""
    torch = dev._torch()
    im = np.asarray(im)
    if im.ndim not in (2, 3):
        raise Exception("Received conn is not valid")
    c = _conn_of(im.ndim, conn)
    if im.size == 0:
        return np.zeros(im.shape, dtype=bool)
    ctx = _lib.context()
    fg = _mask_to_device(im, ctx, im.shape)
    if not surface:
        seeds = torch.zeros_like(fg)
        for f in _face_views(fg):
            seeds[f] = fg[f]
        reached = _reached_from(ctx, fg, seeds, c)
        holes = (fg != 0) & (reached == 0)
        return _mask_to_host(holes)
    keep = None
    bg_on_every_face = True
    for f in _face_views(fg):
        seeds = torch.zeros_like(fg)
        seeds[f] = fg[f]
        r = _reached_from(ctx, fg, seeds, c) != 0
        keep = r if keep is None else (keep & r)
        bg_on_every_face &= bool((fg[f] == 0).any().item())
    holes = (fg != 0) & ~keep
    if not bg_on_every_face:
        holes |= fg == 0
    return _mask_to_host(holes)


def fill_blind_pores(im, conn=None, surface=False):
    r"""Fills the pores that `find_disconnected_voxels` reports (F:424-462)."""
    im = np.copy(im)
    im[find_disconnected_voxels(im, conn=conn, surface=surface)] = False
    return im


def trim_floating_solid(im, conn=None, surface=False):
    r"""Removes the solid that `find_disconnected_voxels(~im)` reports (F:465-503)."""
    im = np.copy(im)
    im[find_disconnected_voxels(~im, conn=conn, surface=surface)] = True
    return im


def trim_nonpercolating_paths(im, inlets, outlets, strel=None):
    r"""Keeps the regions of `im` that hold an inlet voxel and an outlet voxel (F:506-555).  `strel`
    is the connectivity `scipy.ndimage.label` gets: None = its default, the cross (4 / 6 neighbours);
    the cross and the full 3x3(x3) cube are recognised."""
    torch = dev._torch()
    im = np.asarray(im)
    if im.ndim not in (2, 3):
        raise ValueError("trim_nonpercolating_paths supports 2-D and 3-D images")
    c = _strel_conn(im.ndim, strel, 6 if im.ndim == 3 else 4)
    if im.size == 0:
        return np.zeros(im.shape, dtype=bool)
    ctx = _lib.context()
    fg = _mask_to_device(im, ctx, im.shape)
    hit = None
    for mask in (inlets, outlets):
        m = _mask_to_device(mask, ctx, im.shape)
        r = _reached_from(ctx, fg, m * fg, c) != 0
        hit = r if hit is None else (hit & r)
    return _mask_to_host(hit)


ONE_FLOOD_TRAPPED = True      # find_trapped_regions: one flood with join times for all bins (False: one flood per bin)


def _trapped_mask(ctx, shape, temp_of, bins, out_t):
    """F:131-137 on the device: for every bin value i the voxels of `temp = seq >= i` whose component (cross
    neighbourhood, scipy's default) holds no outlet voxel.  temp_of(i) -> flat uint8 device mask."""
    torch = dev._torch()
    conn = 6 if len(shape) == 3 else 4
    out_flat = out_t.reshape(-1)
    bins = list(bins)
    if ONE_FLOOD_TRAPPED and 0 < len(bins) <= 253 and all(b1 <= b0 for b0, b1 in zip(bins[:-1], bins[1:])):
        # descending bins: the sets `seq >= i` are nested, so ONE flood with join times gives the first bin at which
        # every voxel is connected to an outlet voxel of its set (psb200_flood_classes, outlets counted from their own
        # bin on); connectivity only grows with the set, so a voxel is trapped for SOME bin iff it is not connected at
        # the first bin that holds it
        n = int(np.prod(shape))
        cls = torch.full((n,), 254, dtype=torch.uint8, device=out_t.device)
        for k in range(len(bins) - 1, -1, -1):                       # smallest k wins: write from the last bin down
            _lib.check(ctx.lib.psb200_set_where_u8(ctx.handle, dev.ptr(cls), dev.ptr(temp_of(bins[k])), k, n, dev.stream_ptr()))
        rcls = dev.flood_classes(ctx, cls, out_flat, len(bins), conn, host.shape3(shape), inlets_in_set=True)
        return ((cls != 254) & (rcls != cls)).view(*shape)
    trapped = torch.zeros(int(np.prod(shape)), dtype=torch.bool, device=out_t.device)
    for i in bins:
        temp = temp_of(i)
        reached = dev.flood(ctx, temp, temp * out_flat, conn, host.shape3(shape))
        trapped |= (temp != 0) & (reached == 0)
    return trapped.view(*shape)


def find_trapped_regions(seq, outlets=None, bins: int = 25, return_mask: bool = True):
    r"""Trapped regions of an invasion sequence (F:73-147): for every bin value i, descending, the
    voxels with `seq >= i` that are not connected (cross neighbourhood, scipy's default) to an outlet
    voxel of the same set.  One flood per bin on the device, as in the reference's loop; the bins
    (`None`: every sequence value, int: `np.linspace(seq.max(), 1, bins)`) are the reference's own
    numpy expressions.  `return_mask=False` relabels on the host like `make_contiguous('symmetric')`."""
    seq = np.copy(seq)
    if seq.ndim not in (2, 3):
        raise ValueError("find_trapped_regions supports 2-D and 3-D images")
    if outlets is None:
        outlets = host.border_faces(seq.shape)
    outlets = np.asarray(outlets)
    if outlets.dtype != np.bool_:
        raise NotImplementedError("outlets must be a boolean mask")
    if bins is None:
        bins = np.unique(seq)[-1::-1]
        bins = bins[bins > 0]
    elif isinstance(bins, int):
        bins = np.linspace(seq.max(), 1, bins)
    ctx = _lib.context()
    out_t = _mask_to_device(outlets, ctx, seq.shape)
    # `seq >= i` is numpy's own comparison (F:133: any dtype numpy accepts, its promotion rules); only
    # the resulting mask goes to the device
    trapped = _trapped_mask(ctx, seq.shape, lambda i: _mask_to_device(seq >= i, ctx, seq.shape).reshape(-1), bins, out_t)
    trapped = _mask_to_host(trapped)
    if return_mask:
        return trapped
    seq[trapped] = -1
    return host.make_contiguous_symmetric(seq)


# radius-map post-processing (filters/_size_seq_satn.py:16-221), evaluated on the index form of the map
from .sizemap import seq_to_satn, size_to_satn, size_to_seq  # noqa: E402,F401
