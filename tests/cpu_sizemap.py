"""numpy stand-in for the device side of porespy_b200.sizemap.IndexMap (histogram / table expansion / index build),
so that the host logic of the post-processing functions is covered by the CPU suite."""
import numpy as np

from porespy_b200.sizemap import IndexMap


class CpuIndexMap(IndexMap):
    def __init__(self, arr):
        arr = np.asarray(arr)
        a = arr + 0.0 if arr.dtype.kind == "f" else arr
        self.values, inv = np.unique(a, return_inverse=True)
        self.values = self.values.astype(arr.dtype)
        self.idx_np = inv.reshape(-1)
        self.shape = arr.shape
        self.ctx = None

    def _mask(self, im):
        im = np.asarray(im)
        assert im.shape == self.shape
        if im.dtype != np.bool_ and im.size and (im.min() < 0 or im.max() > 1):
            raise NotImplementedError("im must be a binary image")
        return (im != 0).reshape(-1)

    def counts(self, mask=None):
        K = len(self.values)
        if mask is None:
            return np.bincount(self.idx_np, minlength=K).astype(np.int64)
        return np.bincount(self.idx_np + K * mask.astype(np.int64), minlength=2 * K).astype(np.int64).reshape(2, K)

    def expand(self, lut, mask=None, as_numpy=True):
        K = len(self.values)
        sel = self.idx_np if mask is None else self.idx_np + K * mask.astype(np.int64)
        return np.asarray(lut)[sel].reshape(self.shape)
