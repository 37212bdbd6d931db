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
# Opt-in (PSB200_HOST_PREZERO=1): float64 results from 2^28 bytes on are zeroed by background host
# threads during the GPU phase and the epilogue skips all-zero lines.  Measured on this pool's B200
# boxes it LOSES (e2e 135 -> 154 ms at 1024^3): the host memory system is the bottleneck of the
# epilogue, and zeroing adds 8.6 GB of stores to it for the 3 GB it saves later.
PREZERO_MIN_BYTES = (1 << 28) if os.environ.get("PSB200_HOST_PREZERO", "0") == "1" else (1 << 62)


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


class HostResult:
    """The float64 result array of one call, in page-locked host memory, being zeroed by background
    host threads of the library (psb200_host_zero_begin) while the GPU computes.  `finish()` must
    run exactly once (it joins the threads); `array()` hands the numpy view out."""

    def __init__(self, ctx, shape, nthreads=None):
        torch = _torch()
        self.ctx, self.shape = ctx, tuple(shape)
        self.n = int(np.prod(self.shape)) if len(self.shape) else 1
        self.host = torch.empty(self.n, dtype=torch.float64, pin_memory=True)
        self.job = ctypes.c_void_p()
        if nthreads is None:
            # leave two hardware threads to the launching thread and the driver
            nthreads = max(1, (os.cpu_count() or 1) - 2)
        _lib.check(ctx.lib.psb200_host_zero_begin(ctypes.c_void_p(self.host.data_ptr()), self.n, int(nthreads),
                                                  ctypes.byref(self.job)))

    def finish(self):
        if self.job is not None and self.job.value:
            _lib.check(self.ctx.lib.psb200_host_zero_wait(self.job))
        self.job = None

    def array(self):
        self.finish()
        return self.host.numpy().reshape(self.shape)

    def __del__(self):
        try:
            self.finish()
        except Exception:
            pass


def expand_idx_to_host(ctx, idx, lut, shape, chunk=1 << 26, cpu_permille=None, nthreads=None, result=None):
    """Radius-index map (uint8, device) -> float64 numpy in page-locked memory, without ever holding
    the 8 B/voxel map in HBM (psb200_expand_idx_f64_to_host): part of the volume leaves the device
    as index bytes and is widened by host threads of the library, the rest is widened on the
    device chunk by chunk while the previous chunk is on its way over PCIe.  `result`: a
    HostResult started earlier in the call (its buffer is zero by now, so all-zero lines are skipped)."""
    torch = _torch()
    n = idx.numel()
    if n * 8 < PIN_MIN_BYTES:
        if result is not None:
            result.finish()
        out = torch.empty(n, dtype=torch.float64, device=idx.device)
        expand_idx(ctx, idx, lut, out)
        return out.cpu().numpy().reshape(shape)
    cpu_permille = HOST_WIDEN_PERMILLE if cpu_permille is None else int(cpu_permille)
    nthreads = HOST_WIDEN_THREADS if nthreads is None else int(nthreads)
    flags = 0
    if result is not None:
        result.finish()
        host, flags = result.host, _lib.FLAG_HOST_PREZEROED
    else:
        host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    stage_n = (n * cpu_permille) // 1000
    stage = torch.empty(max(stage_n, 1), dtype=torch.uint8, pin_memory=True)
    chunk = min(chunk, n)
    ws = ctx.workspace(2 * chunk * 8 + 256)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    _lib.check(ctx.lib.psb200_expand_idx_f64_to_host(
        ctx.handle, ptr(idx), lut.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(lut),
        ctypes.c_void_p(host.data_ptr()), n, ctypes.c_void_p(stage.data_ptr()), stage.numel(),
        ptr(ws), ws.numel(), cpu_permille, nthreads, flags, stream_ptr()))
    del stage
    return host.numpy().reshape(shape)


def to_device_u8(arr, ctx):
    """Host array or torch tensor -> contiguous uint8 device tensor whose non-zero bytes mark the
    foreground (the kernels test `byte != 0`, so bool / uint8 data is passed through as is)."""
    torch = _torch()
    dev = f"cuda:{ctx.device}"
    if isinstance(arr, torch.Tensor):
        t = arr.to(dev)
        if t.dtype == torch.bool:
            return t.contiguous().view(torch.uint8)
        if t.dtype == torch.uint8:
            return t.contiguous()
        return (t != 0).to(torch.uint8).contiguous()
    a = np.asarray(arr)
    if a.dtype == np.bool_ or a.dtype == np.uint8:
        a = np.ascontiguousarray(a).view(np.uint8)
    else:
        a = np.ascontiguousarray(a != 0).view(np.uint8)
    if a.nbytes >= UPLOAD_PACK_MIN_BYTES:
        return upload_mask(ctx, a)
    return torch.from_numpy(a).to(dev, non_blocking=False)      # a single DMA when `a` is page-locked


UPLOAD_PACK_MIN_BYTES = int(os.environ.get("PSB200_UPLOAD_PACK_MIN_BYTES", str(1 << 26)))
_upload_stage = {}


def upload_mask(ctx, a):
    """Large host volumes cross PCIe as one bit per voxel (psb200_upload_mask_u8): host threads of the
    library pack `byte != 0` chunk by chunk while earlier chunks are in flight, a kernel spreads the
    bits to 0/1 bytes.  The page-locked staging buffer is kept per device (grow-only): the copies
    read it asynchronously, and every public call that uploads a numpy volume ends with a
    synchronising download, so it is idle again before the next upload."""
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


def edt_run(ctx, im_u8, shape, as_f32=False, want_max=False):
    """uint8 device volume -> (uint32 squared distances | float32 distances, max d2 or None).
    One call of psb200_edt_u8: sqrt and max are fused into the last pass."""
    torch = _torch()
    nz, ny, nx = shape3(shape)
    out = torch.empty(nz * ny * nx, dtype=torch.float32 if as_f32 else torch.int32, device=im_u8.device)
    mx = torch.empty(1, dtype=torch.int32, device=im_u8.device) if want_max else None
    nbytes = ctx.lib.psb200_edt_workspace_bytes(ctx.handle, nz, ny, nx)
    ws = ctx.workspace(nbytes)
    _lib.check(ctx.lib.psb200_edt_u8(ctx.handle, ptr(im_u8), ptr(out), 1 if as_f32 else 0, ptr(mx),
                                     nz, ny, nx, ptr(ws), ws.numel(), stream_ptr()))
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


def flood(ctx, mask_u8, inlets_u8, conn, shape):
    torch = _torch()
    nz, ny, nx = shape3(shape)
    out = torch.empty(nz * ny * nx, dtype=torch.uint8, device=mask_u8.device)
    nbytes = ctx.lib.psb200_flood_workspace_bytes(ctx.handle, nz, ny, nx)
    ws = ctx.workspace(nbytes)
    _lib.check(ctx.lib.psb200_flood(ctx.handle, ptr(mask_u8), ptr(inlets_u8), ptr(out), conn,
                                    nz, ny, nx, ptr(ws), ws.numel(), stream_ptr()))
    return out
