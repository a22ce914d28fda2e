"""gymnasium when it is installed, else a minimal stand-in with the same surface (Env, spaces.Box/Discrete,
register/make, seeding).  gymnasium is not in this image and cannot be installed (no network); the stand-in
seeds exactly like gymnasium: ``Generator(PCG64(SeedSequence(seed)))``."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as gym
    from gymnasium import spaces
    from gymnasium.utils import seeding

    HAVE_GYMNASIUM = True
    Env = gym.Env
    register = gym.register
    make = gym.make
    np_random = seeding.np_random
    Wrapper = gym.Wrapper
except ImportError:
    HAVE_GYMNASIUM = False

    def np_random(seed=None):
        seq = np.random.SeedSequence(seed)
        return np.random.Generator(np.random.PCG64(seq)), seq.entropy

    class _Space:
        def __init__(self, shape=None, dtype=None):
            self.shape, self.dtype = shape, dtype
            self._rng = None

        def seed(self, seed=None):
            self._rng, s = np_random(seed)
            return [s]

        @property
        def np_random(self):
            if self._rng is None:
                self._rng, _ = np_random(None)
            return self._rng

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            super().__init__(tuple(shape) if shape is not None else np.shape(low), dtype)
            self.low = np.full(self.shape, low, dtype=dtype)
            self.high = np.full(self.shape, high, dtype=dtype)

        def sample(self):
            return self.np_random.uniform(-1, 1, size=self.shape).astype(self.dtype)

        def contains(self, x):
            return np.shape(x) == self.shape

        def __repr__(self):
            return f"Box({self.shape}, {np.dtype(self.dtype).name})"

    class Discrete(_Space):
        def __init__(self, n, start=0):
            super().__init__((), np.int64)
            self.n, self.start = int(n), int(start)

        def sample(self):
            return int(self.start + self.np_random.integers(self.n))

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class Tuple(_Space):
        def __init__(self, spaces_):
            super().__init__(None, None)
            self.spaces = tuple(spaces_)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

        def contains(self, x):
            return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

        def __len__(self):
            return len(self.spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def __repr__(self):
            return "Tuple(" + ", ".join(repr(s) for s in self.spaces) + ")"

    class _Spaces:
        Box = Box
        Discrete = Discrete
        Tuple = Tuple
        Space = _Space

    spaces = _Spaces()

    class Env:
        metadata = {"render_modes": []}
        render_mode = None
        spec = None
        _np_random = None

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random, _ = np_random(None)
            return self._np_random

        @np_random.setter
        def np_random(self, value):
            self._np_random = value

        @property
        def unwrapped(self):
            return self

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random, _ = np_random(seed)

        def close(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def reset(self, *, seed=None, options=None):
            return self.env.reset(seed=seed, options=options)

        def step(self, action):
            return self.env.step(action)

        def close(self):
            return self.env.close()

    _registry = {}

    def register(id, entry_point=None, **kwargs):
        _registry[id] = (entry_point, kwargs)

    def make(id, **kwargs):
        import importlib

        entry, extra = _registry[id]
        if isinstance(entry, str):
            mod, cls = entry.split(":")
            entry = getattr(importlib.import_module(mod), cls)
        env = entry(**dict(extra.get("kwargs", {}), **kwargs))
        for wrapper in extra.get("additional_wrappers", ()):  # classes here; WrapperSpec objects under gymnasium
            env = wrapper(env)
        return env
