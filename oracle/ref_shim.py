"""TEST INFRASTRUCTURE ONLY -- import shim that lets the UNMODIFIED reference run here.

The reference (``/root/reference``) imports ``gymnasium``, ``pygame`` and
``matplotlib`` at module level on the hot path's import chain
(``ttrl_env/__init__.py:4``, ``envs/common/abstract.py:7-11``,
``envs/common/graphics.py:7``, ``vehicle/dynamics.py:5``); none of the three is
installed in this image and there is no network.  This module installs inert
stand-ins in ``sys.modules`` (only the names the reference touches) and puts
the reference on ``sys.path``.  It is used by the golden-vector generator
(``tests/golden/make_golden.py``) and by the oracle pinning tests; it never
runs on the GPU box (``/root/reference`` does not exist there) and the product
package never imports it.

``Env.reset(seed)`` builds ``np.random.Generator(np.random.PCG64(SeedSequence(seed)))``
exactly like ``gymnasium.utils.seeding.np_random`` so seeded reference episodes
are reproducible.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("TTRL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ttrl_env"))


class _Inert(types.ModuleType):
    """Module whose attributes are inert callables/classes (never reached on the hot path)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[:1].isupper():  # e.g. pygame.Surface is subclassed at import time
            sub = type(name, (), {"__init__": lambda self, *a, **k: None})
            setattr(self, name, sub)
            return sub
        sub = _Inert(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


def _np_random(seed=None):
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    rng = np.random.Generator(np.random.PCG64(seed_seq))
    return rng, np_seed


def _make_gymnasium():
    gym = types.ModuleType("gymnasium")

    class Space:
        def __init__(self, shape=None, dtype=None, seed=None):
            self.shape = shape
            self.dtype = dtype
            self._np_random = None

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random, _ = _np_random(None)
            return self._np_random

        def seed(self, seed=None):
            self._np_random, s = _np_random(seed)
            return [s]

        def sample(self):
            raise NotImplementedError

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            if shape is None:
                shape = np.shape(low)
            super().__init__(tuple(shape), dtype, seed)
            self.low = np.full(self.shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low, dtype=dtype)
            self.high = np.full(self.shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high, dtype=dtype)

        def sample(self):
            return self.np_random.uniform(-1, 1, size=self.shape).astype(self.dtype)

    class Discrete(Space):
        def __init__(self, n, seed=None, start=0):
            super().__init__((), np.int64, seed)
            self.n = int(n)
            self.start = start

        def sample(self):
            return int(self.start + self.np_random.integers(self.n))

    class Tuple(Space):
        def __init__(self, spaces_, seed=None):
            super().__init__(None, None, seed)
            self.spaces = tuple(spaces_)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

    class Dict(Space):
        def __init__(self, spaces_=None, seed=None, **kw):
            super().__init__(None, None, seed)
            self.spaces = dict(spaces_ or {}, **kw)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

    spaces = types.ModuleType("gymnasium.spaces")
    for c in (Space, Box, Discrete, Tuple, Dict):
        setattr(spaces, c.__name__, c)

    class Env:
        metadata = {"render_modes": []}
        render_mode = None
        spec = None
        _np_random = None

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random, _ = _np_random(None)
            return self._np_random

        @np_random.setter
        def np_random(self, value):
            self._np_random = value

        @property
        def unwrapped(self):
            return self

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random, _ = _np_random(seed)

        def step(self, action):
            raise NotImplementedError

        def close(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env

        @classmethod
        def wrapper_spec(cls, **kwargs):
            return (cls.__name__, kwargs)

        def __getattr__(self, name):
            return getattr(self.env, name)

    class RecordConstructorArgs:
        def __init__(self, **kwargs):
            pass

    gym.Env = Env
    gym.Wrapper = Wrapper
    gym.spaces = spaces
    gym.Space = Space

    utils = types.ModuleType("gymnasium.utils")
    utils.RecordConstructorArgs = RecordConstructorArgs
    seeding = types.ModuleType("gymnasium.utils.seeding")
    seeding.np_random = _np_random
    utils.seeding = seeding
    gym.utils = utils

    wrappers = types.ModuleType("gymnasium.wrappers")

    class RecordVideo(Wrapper):
        pass

    class RecordEpisodeStatistics(Wrapper):
        pass

    wrappers.RecordVideo = RecordVideo
    wrappers.RecordEpisodeStatistics = RecordEpisodeStatistics
    gym.wrappers = wrappers

    registry = {}
    envs = types.ModuleType("gymnasium.envs")
    registration = types.ModuleType("gymnasium.envs.registration")

    def register(id, entry_point=None, **kwargs):
        registry[id] = (entry_point, kwargs)

    registration.register = register
    registration.registry = registry
    envs.registration = registration
    gym.envs = envs
    gym.register = register

    core = types.ModuleType("gymnasium.core")
    core.Env = Env
    core.Wrapper = Wrapper
    gym.core = core

    logger = types.ModuleType("gymnasium.logger")
    logger.warn = lambda *a, **k: None
    logger.info = lambda *a, **k: None
    gym.logger = logger
    error = types.ModuleType("gymnasium.error")

    class Error(Exception):
        pass

    error.Error = Error
    gym.error = error

    def make(id, **kwargs):
        import importlib

        entry, _ = registry[id]
        mod, cls = entry.split(":")
        return getattr(importlib.import_module(mod), cls)(**kwargs)

    gym.make = make
    mods = {
        "gymnasium": gym,
        "gymnasium.spaces": spaces,
        "gymnasium.utils": utils,
        "gymnasium.utils.seeding": seeding,
        "gymnasium.wrappers": wrappers,
        "gymnasium.envs": envs,
        "gymnasium.envs.registration": registration,
        "gymnasium.core": core,
        "gymnasium.logger": logger,
        "gymnasium.error": error,
    }
    return mods


_installed = False


def install() -> None:
    """Install the stubs and make ``import ttrl_env`` / ``import ttrl_agent`` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if "gymnasium" not in sys.modules:
        try:
            import gymnasium  # noqa: F401  (use the real one when present)
        except ImportError:
            sys.modules.update(_make_gymnasium())
    for name in ("pygame", "pygame.gfxdraw", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "seaborn", "tensorboardX"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = _Inert(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
