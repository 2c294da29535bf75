"""gymnasium.spaces when gymnasium is installed, otherwise a minimal stand-in (Box / Dict) with the same attributes."""
import numpy as np

try:  # pragma: no cover - depends on the environment
    from gymnasium import spaces as _sp
    Box, Dict, Space = _sp.Box, _sp.Dict, _sp.Space
    HAVE_GYMNASIUM = True
except Exception:
    HAVE_GYMNASIUM = False

    class Space:
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    class Dict(Space):
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}
