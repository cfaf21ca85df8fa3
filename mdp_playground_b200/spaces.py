"""Minimal space descriptors for VectorRLToyEnv.

gymnasium is not a dependency of this package; these carry the attributes
callers read off `env.observation_space` / `env.action_space` (n, shape,
dtype, low/high) plus a seeded `sample()` for convenience.  When gymnasium is
importable the user can wrap them trivially; the step path never uses them.
"""
import numpy as np

from .config import np_random


class DiscreteSpace:
    def __init__(self, n, seed=None, dtype=np.int64):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(dtype)
        self._rng, _ = np_random(seed)

    def sample(self, size=None):
        return self._rng.integers(self.n, size=size)

    def contains(self, x):
        return bool(np.all((np.asarray(x) >= 0) & (np.asarray(x) < self.n)))

    def __repr__(self):
        return f"DiscreteSpace({self.n})"


class BoxSpace:
    def __init__(self, low, high, shape, dtype=np.float32, seed=None):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low).astype(self.dtype)
        self.high = np.full(self.shape, high).astype(self.dtype)
        self._rng, _ = np_random(seed)

    def sample(self, size=None):
        shp = self.shape if size is None else (size,) + self.shape
        if np.all(np.isfinite(self.low)) and np.all(np.isfinite(self.high)):
            return self._rng.uniform(self.low, self.high, size=shp).astype(self.dtype)
        return self._rng.normal(size=shp).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape[-len(self.shape):] == self.shape
                    and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"BoxSpace({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
