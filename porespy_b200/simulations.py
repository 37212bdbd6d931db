"""`ps.simulations.drainage`: image-based drainage with gravity by sphere insertion,
/root/reference/src/porespy/simulations/_drainage.py:18-188, on the device.

Per applied pressure the reference thresholds fn = pc + rho g h, keeps what is connected to the inlets
(`trim_disconnected_blobs`, default full neighbourhood), and paints a sphere of radius int(dt) around every
newly invaded voxel (`_insert_disks_at_points`, tools/_sphere_insertions.py:327-385).  Here: the EDT, the
threshold (fn evaluated per voxel in the reference's own float32 / float64 mix), the flood, and a power-diagram
min-plus pass per axis for the variable-radius spheres all run in libpsb200.so (csrc/drainage_kernels.cuh); the
invasion map lives on the device as one code byte per voxel (code -> applied pressure), and the epilogue
(`pc_to_satn`, `satn_to_seq`, `find_trapped_regions`, `pc_curve`) works on that index form (sizemap.py).
"""
import ctypes

import numpy as np

from . import _device as dev
from . import _host as host
from . import _lib
from . import sizemap as sm
from .sizemap import IndexMap, Results

__all__ = ["drainage"]

MAX_PRESSURES = 250
ONE_FLOOD = True          # ascending pressures: one flood with join times for all steps (False: one flood per step)


def _fn_args(ndim, voxel_size, delta_rho, g, sigma, theta, inner):
    """Scalars of fn = pc + rgh and the precisions numpy gives the three products (F:107, F:113-114) for the
    scalar types the caller passed (python scalars adopt float32 from `dt`, numpy float64 scalars promote)."""
    probe = np.zeros(1, dtype=np.float32)
    den = probe * voxel_size
    h = (probe + 1) * voxel_size
    rgh = delta_rho * g * h
    flags = (1 if den.dtype == np.float64 else 0) | (2 if h.dtype == np.float64 else 0) | (4 if rgh.dtype == np.float64 else 0)
    c0 = float(-(ndim - 1) * sigma * np.cos(np.deg2rad(theta)))
    return dict(inner=int(inner), c0=c0, voxel_size=float(voxel_size), rho_g=float(delta_rho * g), flags=flags)


def drainage(im, voxel_size, pc=None, inlets=None, outlets=None, residual=None, bins=25, delta_rho=1000, g=9.81,
             sigma=0.072, theta=180):
    r"""Simulate drainage using image-based sphere insertion, optionally including gravity.  Arguments and the
    returned `Results` (im_pc, im_satn, im_trapped, pc, snwp) as in the reference (F:18-99).  A caller-supplied
    `pc` map is taken as float64 and is not modified (the reference zeroes it outside `im` in place)."""
    torch = dev._torch()
    im = np.array(im, dtype=bool)
    if im.ndim not in (2, 3):
        raise ValueError("drainage supports 2-D and 3-D images")
    shape, n = im.shape, im.size
    shape3 = host.shape3(shape)
    ctx = _lib.context()
    lib, h_ = ctx.lib, ctx.handle
    device = f"cuda:{ctx.device}"
    im_u8 = dev.to_device_u8(im, ctx).reshape(-1)
    dt, max_d2 = dev.edt_run(ctx, im_u8, shape, as_f32=True, want_max=True)         # F:105  dt = edt(im)
    if max_d2 == host.INF_U32:
        raise NotImplementedError("drainage: the image has no solid voxel (infinite distances)")
    fa = _fn_args(im.ndim, voxel_size, delta_rho, g, sigma, theta, n // shape[0])
    pc_d = None
    if pc is not None:
        pc_d = torch.from_numpy(np.ascontiguousarray(pc, dtype=np.float64).reshape(-1)).to(device)
    fnargs = (n, fa["inner"], fa["c0"], fa["voxel_size"], fa["rho_g"], fa["flags"])

    def to_mask(a):
        a = np.asarray(a)
        if a.shape != shape:
            raise ValueError("inlets / outlets / residual must have the shape of im")
        return dev.to_device_u8(a.astype(bool), ctx).reshape(-1)

    if inlets is None:                                                              # F:117-119
        inlets = np.zeros_like(im)
        inlets[0, ...] = True
    inl_d = to_mask(inlets)
    res_d = to_mask(residual) if residual is not None else None
    conn = 26 if im.ndim == 3 else 8                         # trim_disconnected_blobs' default strel (F:1260-1264)

    if isinstance(bins, int):                                                       # F:121-124
        nb = 148 * 8
        part = torch.empty(2 * nb, dtype=torch.float64, device=device)
        _lib.check(lib.psb200_drain_stats(h_, dev.ptr(dt), dev.ptr(im_u8), dev.ptr(pc_d), *fnargs, dev.ptr(part), nb,
                                          dev.stream_ptr()))
        ph = part.cpu().numpy()
        vmax, vmin = np.float64(ph[0::2].max()), np.float64(ph[1::2].min())
        Ps = np.linspace(vmin, vmax * 1.1, bins)
    else:
        Ps = bins
    Ps = list(Ps)
    if len(Ps) > MAX_PRESSURES:
        raise NotImplementedError(f"drainage supports up to {MAX_PRESSURES} pressure steps")

    def flood(mask_d):
        return dev.flood(ctx, mask_d, inl_d, conn, shape3)

    mask_d = None
    if (residual is not None) and (outlets is not None):                            # F:129-132
        mask_d = flood(im_u8 * (1 - res_d))
    inv = torch.zeros(n, dtype=torch.uint8, device=device)
    rad = torch.empty(n, dtype=torch.int16, device=device)
    stats = torch.zeros(4, dtype=torch.int64, device=device)                        # [count (u64), max radius (int)]
    pws = ctx.lib.psb200_drain_paint_workspace_bytes(h_, *shape3)
    # Ascending pressures (the default bins, and any sorted user list): the sets (fn <= p) * im [+ residual] are
    # nested, so the first step at which every voxel is invaded comes from ONE flood with join times
    # (psb200_flood_classes) instead of one flood per step; otherwise the reference's loop step by step.
    pf = np.asarray([float(p) for p in Ps], dtype=np.float64)
    nested = ONE_FLOOD and len(pf) <= 253 and not np.isnan(pf).any() and bool(np.all(np.diff(pf) >= 0))
    rcls = None
    if nested:
        cls = torch.empty(n, dtype=torch.uint8, device=device)
        _lib.check(lib.psb200_drain_classify(h_, dev.ptr(dt), dev.ptr(im_u8), dev.ptr(pc_d), dev.ptr(res_d), *fnargs,
                                             pf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(pf), dev.ptr(cls),
                                             dev.stream_ptr()))
        rcls = dev.flood_classes(ctx, cls, inl_d, len(pf), conn, shape3)             # inlets: nodes from step 0 on
        del cls
    else:
        seeds = torch.zeros(n, dtype=torch.uint8, device=device)
        temp = torch.empty(n, dtype=torch.uint8, device=device)
    for k, p in enumerate(Ps):                                                      # F:133-154
        if nested:
            _lib.check(lib.psb200_drain_newly_rcls(h_, dev.ptr(rcls), k, dev.ptr(mask_d), dev.ptr(dt), dev.ptr(rad), n,
                                                   ctypes.c_void_p(stats.data_ptr()), ctypes.c_void_p(stats.data_ptr() + 8),
                                                   dev.stream_ptr()))
        else:
            _lib.check(lib.psb200_drain_threshold(h_, dev.ptr(dt), dev.ptr(im_u8), dev.ptr(pc_d), dev.ptr(res_d), *fnargs,
                                                  float(p), dev.ptr(temp), dev.stream_ptr()))
            reached = flood(temp)
            _lib.check(lib.psb200_drain_newly(h_, dev.ptr(reached), dev.ptr(mask_d), dev.ptr(seeds), dev.ptr(dt), dev.ptr(rad), n,
                                              ctypes.c_void_p(stats.data_ptr()), ctypes.c_void_p(stats.data_ptr() + 8),
                                              dev.stream_ptr()))
        st = stats.cpu().numpy()
        count, rmax = int(st[0]), int(st[1:2].view(np.int32)[0])
        if count == 0 or rmax == 0 or float(p) == 0.0:       # (a value of 0 leaves the reference's array "unwritten")
            continue
        ws = ctx.workspace(pws)
        _lib.check(lib.psb200_drain_paint(h_, dev.ptr(rad), rmax, dev.ptr(inv), k + 1, *shape3, dev.ptr(ws), ws.numel(),
                                          dev.stream_ptr()))
    del rad, rcls

    # ---- epilogue on the code map: values[code] is the invasion pressure
    nP = len(Ps)
    K_INF, K_NINF = nP + 1, nP + 2
    values = np.array([0.0] + [float(p) for p in Ps] + [np.inf, -np.inf], dtype=np.float64)
    zero_lut = np.zeros(256, dtype=np.uint8)
    zero_lut[:len(values)] = values == 0
    zl = torch.from_numpy(zero_lut).to(device)
    _lib.check(lib.psb200_set_zero_codes_u8(h_, dev.ptr(inv), dev.ptr(im_u8), dev.ptr(zl), K_INF, n, dev.stream_ptr()))   # F:157
    if res_d is not None:
        _lib.check(lib.psb200_set_where_u8(h_, dev.ptr(inv), dev.ptr(res_d), K_NINF, n, dev.stream_ptr()))              # F:160-161
    inv_map = IndexMap(ctx, inv, values, shape)

    def satn_of(m_):
        s, m, w, present, mask = m_.representatives(im_u8.view(*shape))
        return sm._pc_to_satn_rep(s, m, w, n), (s, m, w, present, mask)

    satn_rep, rep = satn_of(inv_map)                                                # F:166
    trapped = None
    if outlets is not None:                                                         # F:168-175
        s, m, w, present, mask = rep
        seq_rep = sm._satn_to_seq_rep(satn_rep, m)
        K = len(values)
        out_d = to_mask(outlets)
        seq_max = seq_rep.max()
        tbins = np.linspace(seq_max, 1, 25)                                         # find_trapped_regions(bins=25)

        def temp_of(i):
            lut = np.zeros(2 * K, dtype=np.uint8)
            lut[present] = seq_rep >= i
            return inv_map.expand_mask(lut, mask)

        from .filters import _trapped_mask
        tr = _trapped_mask(ctx, shape, temp_of, tbins, out_d.view(*shape)).reshape(-1)
        lut = np.zeros(2 * K, dtype=np.uint8)
        lut[present] = seq_rep == -1
        tr = (tr | (inv_map.expand_mask(lut, mask) != 0)).to(torch.uint8)            # trapped[seq == -1] = True
        _lib.check(lib.psb200_set_where_u8(h_, dev.ptr(inv), dev.ptr(tr), K_INF, n, dev.stream_ptr()))    # inv[trapped] = inf
        if res_d is not None:
            _lib.check(lib.psb200_set_where_u8(h_, dev.ptr(inv), dev.ptr(res_d), K_NINF, n, dev.stream_ptr()))
        trapped = dev.to_host(tr).view(np.bool_).reshape(shape)
        satn_rep, rep = satn_of(inv_map)
    s, m, w, present, mask = rep
    results = Results()
    results.im_satn = inv_map.scatter(satn_rep, present, mask)
    results.im_pc = inv_map.to_numpy()
    results.im_trapped = trapped
    results.pc, results.snwp = sm._pc_curve_pc_rep(s, m, w)                         # F:181-183
    return results
