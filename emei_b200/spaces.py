"""Minimal gym-0.26-compatible space shims (gym is not a dependency of this package).

Only what emei's API surface uses: ``shape``, ``dtype``, ``contains``, ``sample``, ``seed``
(emei/envs/classic_control/base_control.py:65-66, cartpole.py:45-46, emei/core.py:144-147).
``sample_batch`` is additive: one action per environment of a batched env.
"""
import numpy as np


class Space:
    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self._rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(None)))

    def seed(self, seed=None):
        self._rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        return [seed]


class Discrete(Space):
    def __init__(self, n: int):
        super().__init__((), np.int64)
        self.n = int(n)

    def contains(self, x) -> bool:
        if isinstance(x, (int, np.integer)):
            return 0 <= int(x) < self.n
        x = np.asarray(x)
        return bool(np.issubdtype(x.dtype, np.integer) and x.shape == () and 0 <= int(x) < self.n)

    def sample(self):
        return int(self._rng.integers(self.n))

    def sample_batch(self, batch_size: int):
        return self._rng.integers(self.n, size=batch_size)

    def __repr__(self):
        return f"Discrete({self.n})"


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        low = np.asarray(low, dtype=dtype)
        high = np.asarray(high, dtype=dtype)
        if shape is None:
            shape = low.shape
        super().__init__(shape, dtype)
        self.low = np.broadcast_to(low, self.shape).copy()
        self.high = np.broadcast_to(high, self.shape).copy()

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(
            np.can_cast(x.dtype, self.dtype) and x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)
        )

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def sample_batch(self, batch_size: int):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi, size=(batch_size,) + self.shape).astype(self.dtype)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
