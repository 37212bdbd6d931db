"""TEST INFRASTRUCTURE ONLY -- never imported by the product (`porespy_b200/`).

Import shim that lets the *unmodified* reference source under `/root/reference/src`
run in the dev container, where its third-party native dependencies (`edt`,
`skimage`, `dask`, `openpnm`, `matplotlib`, `pywt`, ...) are not installed.

It is used for exactly two things, both in the dev container only (the GPU box has
no `/root/reference`):

* `tests/golden/make_golden.py` -- generates the committed golden vectors by calling
  the reference's own `ps.filters.porosimetry / local_thickness /
  trim_disconnected_blobs`  (`/root/reference/src/porespy/filters/_funcs.py:947-1270`).
* cross-checking the restatements in `oracle/` against the real reference source.

What is real and what is stubbed
--------------------------------
* `edt.edt` -- the PyPI package `edt` (seung-lab/euclidean-distance-transform-3d,
  unpinned in `/root/reference/pyproject.toml:29`) is absent.  Here it is supplied by
  scipy's exact EDT: `distance_transform_edt(return_indices=True)` gives the nearest
  background voxel, from which the exact *integer* squared distance is recomputed and
  `np.sqrt(float32(d2))` returned -- the same `float32(sqrt(d2))` contract as the
  wheel (black_border=False: the image border is not background).
* `skimage.morphology.{ball,disk,square,cube}` -- trivial restatements.
* `dask.delayed/compute` -- serial; `skimage.segmentation.relabel_sequential`, `clear_border` -- numpy.
* `scipy.stats.rankdata(method='dense')` inside `filters/_size_seq_satn.py` -- integer return type of SciPy < 1.18
  restored (the installed SciPy 1.18 returns float64 and `np.bincount` rejects it).
* everything else: inert MagicMock attributes (never on the hot path).
"""
import importlib.abc
import importlib.machinery
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_SRC = "/root/reference/src"

_STUB_ROOTS = (
    "edt", "skimage", "dask", "openpnm", "matplotlib", "pywt", "pyevtk", "stl",
    "trimesh", "imageio", "tifffile", "pyimagej", "imagej", "scyjava", "cupy",
    "cupyx", "nanomesh", "loguru", "pypardiso", "transforms3d", "networkx", "h5py",
    "docrep", "chemicals", "thermo", "sympy", "jsonschema", "flatdict", "traits",
    "pyfastnoisesimd", "setuptools_scm", "mpl_toolkits",
)


def _edt_sq_scipy(data):
    """Exact integer squared EDT (int64) of the non-zero voxels of `data`."""
    import scipy.ndimage as spim
    data = np.asarray(data) != 0
    if data.size == 0:
        return np.zeros(data.shape, dtype=np.int64)
    if not np.any(~data):
        # no background anywhere: black_border=False => distance is infinite.
        # (parity unpinned; oracle and product both return +inf.)
        return np.full(data.shape, -1, dtype=np.int64)
    idx = spim.distance_transform_edt(data, return_distances=False, return_indices=True)
    d2 = np.zeros(data.shape, dtype=np.int64)
    grids = np.meshgrid(*[np.arange(n) for n in data.shape], indexing="ij", sparse=True)
    for ax in range(data.ndim):
        diff = idx[ax].astype(np.int64) - grids[ax]
        d2 += diff * diff
    return d2


def edt_shim(data, anisotropy=None, black_border=False, order="K", parallel=1,
             voxel_graph=None):
    """Signature of `edt.edt` as PoreSpy uses it (F:1126, F:1186-1191, T:1153)."""
    assert anisotropy is None and not black_border and voxel_graph is None
    d2 = _edt_sq_scipy(data)
    out = np.sqrt(d2.astype(np.float32))
    out[d2 < 0] = np.inf
    return out


def _ball(radius, dtype=np.uint8):
    n = 2 * radius + 1
    z, y, x = np.mgrid[-radius:radius + 1, -radius:radius + 1, -radius:radius + 1]
    return np.array(x * x + y * y + z * z <= radius * radius, dtype=dtype).reshape(n, n, n)


def _disk(radius, dtype=np.uint8):
    y, x = np.mgrid[-radius:radius + 1, -radius:radius + 1]
    return np.array(x * x + y * y <= radius * radius, dtype=dtype)


def _square(width, dtype=np.uint8):
    return np.ones((width, width), dtype=dtype)


def _cube(width, dtype=np.uint8):
    return np.ones((width, width, width), dtype=dtype)


def _relabel_sequential(label_field, offset=1):
    vals = np.unique(label_field)
    vals = vals[vals > 0]
    fw = np.zeros(int(label_field.max()) + 1, dtype=label_field.dtype)
    fw[vals] = np.arange(offset, offset + len(vals))
    return fw[label_field], fw, None


class _Delayed:
    def __init__(self, func):
        self.func = func

    def __call__(self, *a, **k):
        return _Lazy(self.func, a, k)


class _Lazy:
    def __init__(self, f, a, k):
        self.f, self.a, self.k = f, a, k

    def compute(self, **kw):
        return self.f(*self.a, **self.k)


def _dask_compute(*items, **kw):
    return tuple(i.compute() if isinstance(i, _Lazy) else i for i in items)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name == "__version__":
            return "3.0.0"
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        m = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        name = module.__name__
        if name == "edt":
            module.edt = edt_shim
            module.edtsq = lambda data, **k: _edt_sq_scipy(data).astype(np.float32)
        elif name == "skimage.morphology":
            module.ball, module.disk = _ball, _disk
            module.square, module.cube = _square, _cube
        elif name == "skimage.segmentation":
            module.relabel_sequential = _relabel_sequential
            module.clear_border = _clear_border
        elif name == "dask":
            module.delayed = _Delayed
            module.compute = _dask_compute


def _clear_border(labels, **kwargs):
    """skimage.segmentation.clear_border for a label image: labels that touch any face become 0
    (default buffer_size=0, bgval=0 -- the only form PoreSpy uses, filters/_funcs.py:411)."""
    labels = np.array(labels, copy=True)
    touching = set()
    for ax in range(labels.ndim):
        for side in (0, -1):
            touching.update(np.unique(np.take(labels, side, axis=ax)).tolist())
    touching.discard(0)
    if touching:
        labels[np.isin(labels, list(touching))] = 0
    return labels


_installed = False


def install():
    """Install the stub finder and put the reference source on sys.path."""
    global _installed
    if _installed:
        return
    import os
    if not os.path.isdir(REFERENCE_SRC):
        raise RuntimeError(f"{REFERENCE_SRC} not present (only exists in the dev container)")
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REFERENCE_SRC)
    _installed = True


def import_reference():
    """Return the reference `porespy` package, imported from /root/reference/src."""
    install()
    import logging
    import porespy as ps  # noqa: the real reference source
    ps.settings.tqdm["disable"] = True
    # SciPy >= 1.18 returns float64 from rankdata(..., 'dense'); the reference (written against older SciPy,
    # where dense ranks were integers) feeds the result to np.bincount (filters/_size_seq_satn.py:211-212).
    # Restore the integer return type of the dependency the reference was written for.
    import scipy.stats as _st
    import porespy.filters._size_seq_satn as _sss
    _sss.rankdata = lambda a, method="average", **k: (
        _st.rankdata(a, method=method, **k).astype(np.int64) if method == "dense" else _st.rankdata(a, method=method, **k))
    logging.getLogger().handlers.clear()
    return ps
