"""Device-side `ps.generators.blobs` (reference: /root/reference/src/porespy/generators/_imgen.py:960-1051,
`norm_to_uniform` /root/reference/src/porespy/tools/_funcs.py:935-969) -- the input generator of every
benchmark configuration.  Same signature, same return value.

Host code evaluates the reference's own numpy expressions for sigma and for scipy's gaussian kernel
(`scipy.ndimage._filters._gaussian_kernel1d`, truncate=4.0); noise, the three 1-D correlations, the
statistics and norm_to_uniform + threshold run in libpsb200.so (csrc/blobs_kernels.cuh), float64 in scipy's
arithmetic order.  `rng='numpy'` (default) draws the noise from numpy's global MT19937 stream on the host
exactly like the reference (`np.random.seed(seed); np.random.random(shape)`), so the image is the
reference's image for that seed (up to voxels whose uniformised value lies within an ulp of `porosity`:
mean/std are fixed-order sums, not numpy's pairwise ones).  `rng='philox'` draws the noise on the device as a
function of the global voxel index, which z-slab shards use to generate their own part of one global image
without communication (`sharded.ShardedVolume.blobs`).
"""
import ctypes

import numpy as np

from . import _device as dev
from . import _lib

__all__ = ["blobs"]


def _prologue(shape, blobiness, divs):
    """F(_imgen.py):1023-1039 -- shape / blobiness normalisation and sigma."""
    if isinstance(shape, int):
        shape = [shape] * 3
    if len(shape) == 1:
        shape = [shape[0]] * 3
    shape = np.array(shape)
    if isinstance(blobiness, int):
        blobiness = [blobiness] * len(shape)
    blobiness = np.array(blobiness)
    sigma = np.mean(shape) / (40 * blobiness)
    sigma = np.broadcast_to(np.asarray(sigma, dtype=np.float64), (len(shape),))
    return tuple(int(s) for s in shape), sigma


def gaussian_kernel(sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter1d's kernel (order 0): radius = int(truncate * sigma + 0.5), weights
    exp(-0.5 / sigma^2 * x^2) normalised by their sum.  Returns (radius, w[0..radius]) from the outermost tap
    to the centre (the kernel is symmetric)."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    sigma2 = sd * sd
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    w = phi[::-1]                                       # correlate1d gets the reversed kernel
    return radius, np.ascontiguousarray(w[:radius + 1], dtype=np.float64)


def _stat(ctx, x, nplanes, plane, mean, mode):
    """Per-plane partials (device, fixed order) -> numpy [nplanes][chunks]."""
    torch = dev._torch()
    ch = ctx.lib.psb200_stats_chunks()
    part = torch.empty(nplanes * ch, dtype=torch.float64, device=x.device)
    _lib.check(ctx.lib.psb200_stats_f64(ctx.handle, dev.ptr(x), nplanes, plane, float(mean), mode, dev.ptr(part),
                                        dev.stream_ptr()))
    return part.cpu().numpy().reshape(nplanes, ch)


def combine_sum(parts):
    """Sum of per-plane partials in plane order (sequential float64: the same on every rank and for every
    sharding)."""
    total = 0.0
    for v in np.asarray(parts, dtype=np.float64).reshape(-1):
        total += float(v)
    return total


def gauss_axis(ctx, src, dst, axis, sigma, shape3, z_out0=0, z_in0=0, nz_in=None, nz_glob=None):
    """One 1-D correlation of gaussian_filter on the device (psb200_gauss_axis_f64)."""
    nz, ny, nx = shape3
    radius, w = gaussian_kernel(sigma)
    ws = ctx.workspace(ctx.lib.psb200_gauss_workspace_bytes(ctx.handle, radius))
    _lib.check(ctx.lib.psb200_gauss_axis_f64(
        ctx.handle, dev.ptr(src), dev.ptr(dst), axis, w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), radius,
        nz, ny, nx, int(z_out0), int(z_in0), int(nz if nz_in is None else nz_in),
        int(nz if nz_glob is None else nz_glob), dev.ptr(ws), ws.numel(), dev.stream_ptr()))


def field_stats(ctx, f, nplanes, plane):
    """(sum partials, min, max) of a filtered slab; the squared deviations need the global mean first."""
    s = _stat(ctx, f, nplanes, plane, 0.0, 0)
    lo = float(_stat(ctx, f, nplanes, plane, 0.0, 2).min())
    hi = float(_stat(ctx, f, nplanes, plane, 0.0, 3).max())
    return s, lo, hi


def finish(ctx, f, n, mean, sd, fmin, fmax, porosity):
    """norm_to_uniform + threshold (or the uniform field when porosity is None / 0)."""
    torch = dev._torch()
    if porosity:
        out = torch.empty(n, dtype=torch.uint8, device=f.device)
        _lib.check(ctx.lib.psb200_blobs_finish(ctx.handle, dev.ptr(f), n, mean, sd, fmin, fmax, float(porosity),
                                               dev.ptr(out), None, dev.stream_ptr()))
        return out
    out = torch.empty(n, dtype=torch.float64, device=f.device)
    _lib.check(ctx.lib.psb200_blobs_finish(ctx.handle, dev.ptr(f), n, mean, sd, fmin, fmax, 0.0, None, dev.ptr(out),
                                           dev.stream_ptr()))
    return out


def philox_noise(ctx, n, seed, first=0):
    torch = dev._torch()
    out = torch.empty(int(n), dtype=torch.float64, device=f"cuda:{ctx.device}")
    _lib.check(ctx.lib.psb200_noise_philox_f64(ctx.handle, dev.ptr(out), int(n), int(seed) & (2 ** 64 - 1), int(first),
                                               dev.stream_ptr()))
    return out


def blobs(shape, porosity: float = 0.5, blobiness: int = 1, divs: int = 1, seed=None, rng="numpy",
          as_numpy=True):
    r"""Generates an image containing amorphous blobs; arguments as in the reference (`divs` is accepted and
    ignored: it only selects the reference's chunked CPU filter, whose result equals the unchunked one).

    rng : 'numpy' (host MT19937 stream like the reference) or 'philox' (device noise, `seed` keys it;
          seed=None draws a key from numpy's global stream).
    as_numpy : False returns the device tensor (uint8 0/1, or float64 when `porosity` is None / 0).
    """
    torch = dev._torch()
    shape, sigma = _prologue(shape, blobiness, divs)
    if len(shape) not in (2, 3):
        raise ValueError("blobs supports 2-D and 3-D shapes")
    ctx = _lib.context()
    n = int(np.prod(shape))
    shape3 = (1,) + shape if len(shape) == 2 else shape
    sig3 = ([None] + list(sigma)) if len(shape) == 2 else list(sigma)
    if rng == "numpy":
        if seed is not None:
            np.random.seed(seed)
        a = torch.from_numpy(np.random.random(shape).reshape(-1)).to(f"cuda:{ctx.device}")
    elif rng == "philox":
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        a = philox_noise(ctx, n, seed)
    else:
        raise ValueError("rng must be 'numpy' or 'philox'")
    b = torch.empty_like(a)
    for axis in range(3):                                # gaussian_filter: axis 0 first
        if sig3[axis] is None:                           # 2-D image: no z axis
            continue
        if float(sig3[axis]) <= 1e-15:                   # scipy skips axes with sigma ~ 0
            continue
        gauss_axis(ctx, a, b, axis, sig3[axis], shape3)
        a, b = b, a
    del b
    nplanes, plane = shape3[0], shape3[1] * shape3[2]
    s, lo, hi = field_stats(ctx, a, nplanes, plane)
    mean = combine_sum(s) / n
    sd = float(np.sqrt(combine_sum(_stat(ctx, a, nplanes, plane, mean, 1)) / n))
    out = finish(ctx, a, n, mean, sd, lo, hi, porosity)
    out = out.view(*shape)
    if not as_numpy:
        return out
    h = dev.to_host(out)
    return h.view(np.bool_) if porosity else h
