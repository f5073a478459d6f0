"""Minimal gymnasium.spaces (fallback, see gymnasium/__init__.py)."""
from __future__ import annotations

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(int(s) for s in shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = np.random.default_rng(seed)

    @property
    def shape(self):
        return self._shape

    def seed(self, seed=None):
        self._np_random = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        raise NotImplementedError

    def contains(self, x) -> bool:
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)

    def _key(self):
        return (self._shape, self.dtype)

    def __eq__(self, other):
        return type(self) is type(other) and self._key() == other._key()

    def __hash__(self):
        return hash((type(self).__name__, str(self._key())))


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        super().__init__((), np.int64, seed)
        self.n, self.start = int(n), int(start)

    def sample(self):
        return int(self.start + self._np_random.integers(self.n))

    def contains(self, x):
        try:
            return self.start <= int(x) < self.start + self.n
        except (TypeError, ValueError):
            return False

    def _key(self):
        return (self.n, self.start)

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self):
        return (self._np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.nvec.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

    def _key(self):
        return tuple(self.nvec.reshape(-1).tolist())

    def __repr__(self):
        return f"MultiDiscrete({self.nvec.tolist()})"


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        super().__init__(shape, dtype, seed)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self._shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self._shape).copy()

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return (lo + (hi - lo) * self._np_random.random(self._shape)).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self._shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def _key(self):
        return (self._shape, self.dtype, self.low.tobytes(), self.high.tobytes())

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self._shape}, {self.dtype})"


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        super().__init__(None, None, seed)
        self.spaces = tuple(spaces)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def _key(self):
        return tuple((type(s).__name__, str(s._key())) for s in self.spaces)

    def __repr__(self):
        return f"Tuple({self.spaces})"


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kwargs):
        super().__init__(None, None, seed)
        self.spaces = dict(spaces or {}, **kwargs)

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def contains(self, x):
        return isinstance(x, dict) and x.keys() == self.spaces.keys() and all(self.spaces[k].contains(v) for k, v in x.items())

    def keys(self):
        return self.spaces.keys()

    def values(self):
        return self.spaces.values()

    def items(self):
        return self.spaces.items()

    def __getitem__(self, k):
        return self.spaces[k]

    def __setitem__(self, k, v):
        self.spaces[k] = v

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __contains__(self, k):
        return k in self.spaces if isinstance(k, str) else self.contains(k)

    def _key(self):
        return tuple((k, type(s).__name__, str(s._key())) for k, s in self.spaces.items())

    def __repr__(self):
        return f"Dict({self.spaces})"
