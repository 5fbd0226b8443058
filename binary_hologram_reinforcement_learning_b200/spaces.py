"""gymnasium when importable, else minimal stand-ins with the same surface.

The reference env subclasses ``gym.Env`` and declares ``spaces.Dict/Box/Discrete``
(env.py:18-19,42-52).  gymnasium is not installed in the build image, so the
env must import without it; when gymnasium is present the real classes are
used and Stable-Baselines3 sees a genuine ``gym.Env``.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as gym
    from gymnasium import spaces
    Env = gym.Env
    Box, Dict, Discrete, MultiDiscrete = spaces.Box, spaces.Dict, spaces.Discrete, spaces.MultiDiscrete
    HAVE_GYMNASIUM = True
except Exception:
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape=None, dtype=None, seed=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            super().__init__(shape, dtype, seed)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return self._rng.integers(self.low, self.high + 1, dtype=self.dtype)
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class Discrete(_Space):
        def __init__(self, n, seed=None, start=0):
            super().__init__((), np.int64, seed)
            self.n, self.start = int(n), int(start)

        def sample(self):
            return int(self._rng.integers(self.start, self.start + self.n))

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n

    class MultiDiscrete(_Space):
        def __init__(self, nvec, dtype=np.int64, seed=None):
            self.nvec = np.asarray(nvec, dtype=dtype)
            super().__init__(self.nvec.shape, dtype, seed)

        def sample(self):
            return (self._rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

    class Dict(_Space):
        def __init__(self, spaces=None, seed=None):
            super().__init__(None, None, seed)
            self.spaces = dict(spaces or {})

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

    class Env:
        metadata = {}
        observation_space = None
        action_space = None

        def reset(self, seed=None, options=None):
            raise NotImplementedError

        def step(self, action):
            raise NotImplementedError

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self
