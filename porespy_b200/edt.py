"""`edt`-compatible module: the third-party package PoreSpy imports as `from edt import edt`
(call sites /root/reference/src/porespy/filters/_funcs.py:5,1126,1186-1191;
tools/_funcs.py:6,1153; filters/_snows.py:600-607; beta/_gdd.py:9,126).

Only the behaviour PoreSpy relies on is implemented: binary input, isotropic,
black_border=False.  Other options raise instead of silently differing.
"""
import numpy as np

from . import _device as dev
from . import _host as host
from . import _lib

__all__ = ["edt", "edtsq"]


def _check(anisotropy, black_border, voxel_graph, data):
    if anisotropy is not None and any(float(a) != 1.0 for a in np.atleast_1d(anisotropy)):
        raise NotImplementedError("porespy_b200.edt: anisotropy is not supported")
    if black_border:
        raise NotImplementedError("porespy_b200.edt: black_border=True is not supported")
    if voxel_graph is not None:
        raise NotImplementedError("porespy_b200.edt: voxel_graph is not supported")


def _run(data, squared):
    torch = dev._torch()
    as_numpy = not isinstance(data, torch.Tensor)
    if as_numpy:
        data = np.asarray(data)
        if data.dtype != np.bool_ and data.size and data.max() > 1:
            # the upstream package treats every label as its own object (distances to label boundaries);
            # PoreSpy only passes binary images, anything else must not silently turn binary
            raise NotImplementedError("porespy_b200.edt: multi-label images are not supported (binary input only)")
    shape = tuple(int(s) for s in data.shape)
    if len(shape) == 0 or int(np.prod(shape)) == 0:
        return np.zeros(shape, dtype=np.float32)
    if len(shape) > 3:
        raise ValueError("edt supports 1-D, 2-D and 3-D arrays")
    ctx = _lib.context()
    im_u8 = dev.to_device_u8(data, ctx)
    if squared:
        # squared distances are exact integers < 2^24 for every supported volume of edge
        # <= 2048; returned as float32 like edt.edtsq
        d2 = dev.edt_sq(ctx, im_u8, shape)
        out = d2.view(torch.int32).to(torch.float32)
        out = torch.where(d2 == -1, torch.full_like(out, float("inf")), out)
    else:
        out = dev.edt_run(ctx, im_u8, shape, as_f32=True)[0]      # sqrt fused into the last pass
    out = out.view(*shape)
    return dev.to_host(out) if as_numpy else out


def edt(data, anisotropy=None, black_border=False, order="K", parallel=1, voxel_graph=None):
    """Exact Euclidean distance of every non-zero voxel to the nearest zero voxel, float32.
    `order` and `parallel` are accepted and ignored (the GPU path has no thread count)."""
    _check(anisotropy, black_border, voxel_graph, data)
    return _run(data, squared=False)


def edtsq(data, anisotropy=None, black_border=False, order="K", parallel=1, voxel_graph=None):
    """Squared distances (float32, exact integers)."""
    _check(anisotropy, black_border, voxel_graph, data)
    return _run(data, squared=True)


def edt_sq_u32(data):
    """Exact squared distances as uint32 numpy (0xFFFFFFFF where the image has no background)."""
    data = np.asarray(data)
    shape = tuple(int(s) for s in data.shape)
    ctx = _lib.context()
    d2 = dev.edt_sq(ctx, dev.to_device_u8(data, ctx), shape)
    return dev.to_host(d2).view(np.uint32).reshape(shape)
