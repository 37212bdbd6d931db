"""Readers for the fixtures written by tests/golden/make_golden.py."""
import hashlib
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class Case:
    def __init__(self, path):
        self._z = np.load(path)

    def keys(self):
        return list(self._z.keys())

    def raw(self, name):
        return self._z[name]

    def mask(self, name):
        shape = tuple(int(s) for s in self._z[f"{name}__shape"])
        n = int(np.prod(shape))
        return np.unpackbits(self._z[f"{name}__bits"])[:n].astype(bool).reshape(shape)

    def rmap(self, name):
        """float64 radius map stored as (values, idx)."""
        return self._z[f"{name}__values"][self._z[f"{name}__idx"]]


class Golden:
    def __getattr__(self, name):
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        if not os.path.exists(path):
            raise AttributeError(name)
        c = Case(path)
        setattr(self, name, c)
        return c
