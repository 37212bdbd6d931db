"""Thin device-side helpers: torch tensors are used only as device allocations and for the
host<->device copies; every computation is a call into libpsb200.so."""
import ctypes
import os

import numpy as np

from . import _lib
from ._host import shape3


def _torch():
    import torch
    return torch


def stream_ptr():
    return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


PIN_MIN_BYTES = 1 << 20      # below this a pageable copy is as fast as a pinned one

def pinned_empty(shape, dtype=np.uint8):
    """numpy array backed by page-locked host memory (torch's caching host allocator owns it and
    recycles the block when the array is garbage-collected).  Volumes handed to the public API in
    such arrays are uploaded by one DMA at PCIe speed instead of through the driver's staging
    copies of pageable memory."""
    torch = _torch()
    tdt = {np.dtype(np.bool_): torch.bool, np.dtype(np.uint8): torch.uint8, np.dtype(np.float32): torch.float32,
           np.dtype(np.float64): torch.float64, np.dtype(np.uint32): torch.int32,
           np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
    t = torch.empty(tuple(int(s) for s in shape), dtype=tdt, pin_memory=True)
    a = t.numpy()
    return a.view(np.uint32) if np.dtype(dtype) == np.dtype(np.uint32) else a


def to_pinned(arr):
    """Copy of a numpy array in page-locked memory (see pinned_empty)."""
    arr = np.asarray(arr)
    out = pinned_empty(arr.shape, arr.dtype)
    np.copyto(out, arr)
    return out


def to_host(t):
    """Device tensor -> numpy.  Large results land in page-locked memory (one DMA at PCIe speed;
    a pageable destination costs several times more), small ones take the plain path."""
    torch = _torch()
    if t.numel() * t.element_size() < PIN_MIN_BYTES:
        return t.cpu().numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


# share (per mille) of the volume whose radius index crosses PCIe as bytes and is widened to float64
# by the library's host threads (psb200_expand_idx_f64_to_host); the rest is widened on the device.
# Tuned on the B200 boxes of this pool (scripts/epilogue_probe.py); 0 = all-device path.
HOST_WIDEN_PERMILLE = int(os.environ.get("PSB200_HOST_WIDEN_PERMILLE", "1000"))
HOST_WIDEN_THREADS = int(os.environ.get("PSB200_HOST_WIDEN_THREADS", "0"))      # 0: all hardware threads


def expand_idx_to_host(ctx, idx, lut, shape, chunk=1 << 26, cpu_permille=None, nthreads=None):
    """Radius-index map (uint8, device) -> float64 numpy in page-locked memory, without ever holding
    the 8 B/voxel map in HBM (psb200_expand_idx_f64_to_host): part of the volume leaves the device
    as index bytes and is widened by host threads of the library, the rest is widened on the
    device chunk by chunk while the previous chunk is on its way over PCIe."""
    torch = _torch()
    n = idx.numel()
    if n * 8 < PIN_MIN_BYTES:
        out = torch.empty(n, dtype=torch.float64, device=idx.device)
        expand_idx(ctx, idx, lut, out)
        return out.cpu().numpy().reshape(shape)
    cpu_permille = HOST_WIDEN_PERMILLE if cpu_permille is None else int(cpu_permille)
    nthreads = HOST_WIDEN_THREADS if nthreads is None else int(nthreads)
    host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    stage_n = (n * cpu_permille) // 1000
    stage = torch.empty(max(stage_n, 1), dtype=torch.uint8, pin_memory=True)
    chunk = min(chunk, n)
    ws = ctx.workspace(2 * chunk * 8 + 256)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    _lib.check(ctx.lib.psb200_expand_idx_f64_to_host(
        ctx.handle, ptr(idx), lut.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(lut),
        ctypes.c_void_p(host.data_ptr()), n, ctypes.c_void_p(stage.data_ptr()), stage.numel(),
        ptr(ws), ws.numel(), cpu_permille, nthreads, 0, stream_ptr()))
    del stage
    return host.numpy().reshape(shape)


def expand_idx_slice_to_host(ctx, idx, off, cnt, lut, host, host_off, ready_event, side_stream, ws, nthreads=None):
    """One slab of the host-result epilogue, callable from a worker thread: widen idx[off : off + cnt] (uint8,
    device) into host[host_off : host_off + cnt] (float64, page-locked) once `ready_event` has fired.  The call
    returns when the slab is complete in host memory; it only touches `side_stream`, the library's copy streams
    and host threads, so the caller's compute stream keeps running the next slab meanwhile."""
    torch = _torch()
    torch.cuda.set_device(ctx.device)
    side_stream.wait_event(ready_event)
    nthreads = HOST_WIDEN_THREADS if nthreads is None else int(nthreads)
    stage = torch.empty(max(cnt, 1), dtype=torch.uint8, pin_memory=True)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    _lib.check(ctx.lib.psb200_expand_idx_f64_to_host(
        ctx.handle, ctypes.c_void_p(idx.data_ptr() + off), lut.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(lut),
        ctypes.c_void_p(host.data_ptr() + 8 * host_off), cnt, ctypes.c_void_p(stage.data_ptr()), stage.numel(),
        ptr(ws), ws.numel(), HOST_WIDEN_PERMILLE, nthreads, 0, ctypes.c_void_p(side_stream.cuda_stream)))
    del stage


def to_device_u8(arr, ctx, positive=False):
    """Host array or torch tensor -> contiguous uint8 device tensor whose non-zero bytes mark the
    foreground (the kernels test `byte != 0`, so bool / uint8 data is passed through as is).
    `positive`: the foreground of a signed / float image is `im > 0` (F:1126 `edt(im > 0)`, F:1265), not
    `im != 0` (masks, and the `edt` module itself, where every non-zero label is an object)."""
    torch = _torch()
    dev = f"cuda:{ctx.device}"
    if isinstance(arr, torch.Tensor):
        t = arr.to(dev)
        if t.dtype == torch.bool:
            return t.contiguous().view(torch.uint8)
        if t.dtype == torch.uint8:
            return t.contiguous()
        return ((t > 0) if positive else (t != 0)).to(torch.uint8).contiguous()
    a = np.asarray(arr)
    if a.dtype == np.bool_ or a.dtype == np.uint8:
        a = np.ascontiguousarray(a).view(np.uint8)
    else:
        a = np.ascontiguousarray((a > 0) if positive else (a != 0)).view(np.uint8)
    if a.nbytes >= UPLOAD_PACK_MIN_BYTES:
        return upload_mask(ctx, a)
    return torch.from_numpy(a).to(dev, non_blocking=False)      # a single DMA when `a` is page-locked


UPLOAD_PACK_MIN_BYTES = int(os.environ.get("PSB200_UPLOAD_PACK_MIN_BYTES", str(1 << 26)))
_upload_stage = {}


def upload_mask(ctx, a):
    """Large host volumes cross PCIe as one bit per voxel (psb200_upload_mask_u8): host threads of the
    library pack `byte != 0` chunk by chunk while earlier chunks are in flight, a kernel spreads the
    bits to 0/1 bytes.  The page-locked staging buffer is kept per device (grow-only); the call returns
    once the last copy has left it, so back-to-back uploads (fg, inlets, outlets) may share it."""
    torch = _torch()
    n = a.size
    nb = (n + 7) // 8
    stage = _upload_stage.get(ctx.device)
    if stage is None or stage.numel() < nb:
        stage = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
        _upload_stage[ctx.device] = stage
    out = torch.empty(a.shape, dtype=torch.uint8, device=f"cuda:{ctx.device}")
    ws = ctx.workspace(nb + 256)
    _lib.check(ctx.lib.psb200_upload_mask_u8(
        ctx.handle, ctypes.c_void_p(a.ctypes.data), n, ptr(out), ctypes.c_void_p(stage.data_ptr()), stage.numel(),
        ptr(ws), ws.numel(), 0, stream_ptr()))
    return out


def edt_run(ctx, im_u8, shape, as_f32=False, want_max=False, zmax=None):
    """uint8 device volume -> (uint32 squared distances | float32 distances, max d2 or None).
    One call of psb200_edt_u8: sqrt and max are fused into the last pass.  `zmax = (z0, z1)`: the maximum
    over those planes only (psb200_edt_u8_zmax)."""
    torch = _torch()
    nz, ny, nx = shape3(shape)
    out = torch.empty(nz * ny * nx, dtype=torch.float32 if as_f32 else torch.int32, device=im_u8.device)
    mx = torch.empty(1, dtype=torch.int32, device=im_u8.device) if want_max else None
    nbytes = ctx.lib.psb200_edt_workspace_bytes(ctx.handle, nz, ny, nx)
    ws = ctx.workspace(nbytes)
    if zmax is None:
        _lib.check(ctx.lib.psb200_edt_u8(ctx.handle, ptr(im_u8), ptr(out), 1 if as_f32 else 0, ptr(mx),
                                         nz, ny, nx, ptr(ws), ws.numel(), stream_ptr()))
    else:
        _lib.check(ctx.lib.psb200_edt_u8_zmax(ctx.handle, ptr(im_u8), ptr(out), 1 if as_f32 else 0, ptr(mx),
                                              nz, ny, nx, int(zmax[0]), int(zmax[1]), ptr(ws), ws.numel(), stream_ptr()))
    if want_max:
        return out, int(mx.cpu().numpy().view(np.uint32)[0])
    return out, None


def edt_sq(ctx, im_u8, shape):
    """uint8 device volume -> uint32 squared distances (device tensor)."""
    return edt_run(ctx, im_u8, shape)[0]


def max_u32(ctx, d2):
    torch = _torch()
    out = torch.zeros(1, dtype=torch.int32, device=d2.device)
    _lib.check(ctx.lib.psb200_max_u32(ctx.handle, ptr(d2), d2.numel(), ptr(out), stream_ptr()))
    return int(out.cpu().numpy().view(np.uint32)[0])


def sqrt_f32(ctx, d2):
    torch = _torch()
    out = torch.empty(d2.numel(), dtype=torch.float32, device=d2.device)
    _lib.check(ctx.lib.psb200_sqrt_f32(ctx.handle, ptr(d2), ptr(out), d2.numel(), stream_ptr()))
    return out


def local_thickness_idx(ctx, d2, T, idx, inlets_u8, inlet_mode, ndim, shape, flags=0):
    nz, ny, nx = shape3(shape)
    nbytes = ctx.lib.psb200_local_thickness_workspace_bytes(ctx.handle, nz, ny, nx, inlet_mode)
    ws = ctx.workspace(nbytes)
    T = np.ascontiguousarray(T, dtype=np.uint32)
    _lib.check(ctx.lib.psb200_local_thickness_idx(
        ctx.handle, ptr(d2), T.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(T), ptr(idx),
        ptr(inlets_u8), inlet_mode, ndim, nz, ny, nx, flags, ptr(ws), ws.numel(), stream_ptr()))


def expand_idx(ctx, idx, lut, out, merge=False):
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    _lib.check(ctx.lib.psb200_expand_idx_f64(
        ctx.handle, ptr(idx), lut.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(lut),
        ptr(out), idx.numel(), _lib.FLAG_EXPAND_MERGE if merge else 0, stream_ptr()))


def mark_written(ctx, out, idx):
    _lib.check(ctx.lib.psb200_mark_written(ctx.handle, ptr(out), ptr(idx), idx.numel(), stream_ptr()))


def flood_classes(ctx, cls_u8, inlets_u8, nsteps, conn, shape, inlets_in_set=False):
    """Nested-set flood (psb200_flood_classes): first step at which every voxel is a node connected to the inlets.
    `inlets_in_set`: an inlet voxel only counts from the step at which it joins the set."""
    torch = _torch()
    nz, ny, nx = shape3(shape)
    out = torch.empty(nz * ny * nx, dtype=torch.uint8, device=cls_u8.device)
    ws = ctx.workspace(ctx.lib.psb200_flood_workspace_bytes(ctx.handle, nz, ny, nx))
    _lib.check(ctx.lib.psb200_flood_classes(ctx.handle, ptr(cls_u8), ptr(inlets_u8), 1 if inlets_in_set else 0, ptr(out), int(nsteps), conn,
                                            nz, ny, nx, ptr(ws), ws.numel(), stream_ptr()))
    return out


def flood(ctx, mask_u8, inlets_u8, conn, shape):
    torch = _torch()
    nz, ny, nx = shape3(shape)
    out = torch.empty(nz * ny * nx, dtype=torch.uint8, device=mask_u8.device)
    nbytes = ctx.lib.psb200_flood_workspace_bytes(ctx.handle, nz, ny, nx)
    ws = ctx.workspace(nbytes)
    _lib.check(ctx.lib.psb200_flood(ctx.handle, ptr(mask_u8), ptr(inlets_u8), ptr(out), conn,
                                    nz, ny, nx, ptr(ws), ws.numel(), stream_ptr()))
    return out
