"""Minimal gym.spaces.Box stand-in (gym/gymnasium are optional; used when neither is installed)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        low = np.asarray(low, dtype=np.float64)
        high = np.asarray(high, dtype=np.float64)
        if shape is not None:
            low = np.broadcast_to(low, shape)
            high = np.broadcast_to(high, shape)
        self.low = low.astype(self.dtype)
        self.high = high.astype(self.dtype)
        self.shape = self.low.shape
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


def make_box(low, high, dtype=np.float32):
    """A Box from gymnasium / gym when importable (so SB-style callers get the real type), else the stand-in."""
    for mod in ("gymnasium", "gym"):
        try:
            spaces = __import__(mod + ".spaces", fromlist=["Box"])
            return spaces.Box(low=np.asarray(low, dtype=dtype), high=np.asarray(high, dtype=dtype), dtype=dtype)
        except Exception:
            continue
    return Box(low, high, dtype=dtype)
