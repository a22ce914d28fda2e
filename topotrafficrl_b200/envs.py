"""Single-env, gymnasium-shaped front ends over the batched simulator, with the reference's ids, class names,
config keys, return tuple and RNG stream (``ttrl_env/__init__.py:22-56``, ``envs/common/abstract.py:188-250``,
``envs/intersection_env.py:17-139``).  ``scripts/example.py``-style code runs unchanged apart from the import.

The env's ``np_random`` is a numpy ``Generator(PCG64)`` seeded like gymnasium; spawn draws and the
"shuffled" observation permutation are taken from it in the reference's order (observe -> shuffle, then
clear -> spawn), so a seeded episode reproduces the reference episode.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import abi, scenes
from ._gym import Env, register, spaces
from .reset import draw_spawn, reset_intersection
from .sim import Sim
from .vector_env import _SimResetBackend


class AbstractEnv(Env):
    """Common part: config handling, spaces, the step/reset tuple."""

    metadata = {"render_modes": ["human", "rgb_array"]}
    PERCEPTION_DISTANCE = 5.0 * 40.0
    SCENE = "intersection"

    def __init__(self, config: Optional[dict] = None, render_mode: Optional[str] = None, device: int = 0) -> None:
        super().__init__()
        self.config = self.default_config()
        self.configure(config)
        self.render_mode = render_mode
        self.device_index = device
        self.sim: Optional[Sim] = None
        self.time = self.steps = 0
        self.done = False
        self._build()
        self.reset()

    @classmethod
    def default_config(cls) -> dict:
        return scenes.merged_config(scenes.BASE_CONFIG, None)

    def configure(self, config: Optional[dict]) -> None:
        if config:
            self.config.update(config)  # shallow, like the reference (abstract.py:111-113)

    # ---- to be provided by the scene ---------------------------------------------------------------
    def _build(self) -> None:
        raise NotImplementedError

    def _reset(self) -> None:
        raise NotImplementedError

    def render(self):
        raise NotImplementedError("rendering (pygame) is outside the B200 hot path (SURVEY.md section 2 row 20)")

    def close(self) -> None:
        self.done = True
        if self.sim is not None:
            self.sim.close()
            self.sim = None


class IntersectionEnv(AbstractEnv):
    ACTIONS = {0: "SLOWER", 1: "IDLE", 2: "FASTER"}
    ACTIONS_INDEXES = {v: k for k, v in ACTIONS.items()}

    @classmethod
    def default_config(cls) -> dict:
        return scenes.merged_config(scenes.INTERSECTION_CONFIG, None)

    def _build(self) -> None:
        if self.sim is not None:
            self.sim.close()
        self.net = scenes.make_intersection_network()
        self.table = self.net.to_table(scenes.intersection_exit_predicate)
        self.cfg = scenes.build_config(self.table, self.config, "intersection")
        self.sim = Sim(self.cfg, self.table, 1, 32, self.device_index, scenes.intersection_spawn_routes(self.net, self.table))
        self.sim.set_autoreset(False)
        shape = ((self.cfg.n_features, self.cfg.grid_w, self.cfg.grid_h) if self.cfg.obs_type == abi.OBS_GRID
                 else (self.cfg.obs_vehicles, self.cfg.n_features))
        self._obs_shape = shape
        self.observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=shape, dtype=np.float32)
        self.action_space = spaces.Discrete(5 if self.cfg.action_mode == abi.ACT_ALL else 3)
        self._built_for = repr(self.config)

    def _shuffle(self, obs: np.ndarray) -> np.ndarray:
        if self.cfg.obs_type == abi.OBS_KINEMATICS and self.cfg.order == abi.ORDER_SHUFFLED:
            self.np_random.shuffle(obs[1:])  # observation.py:272-273
        return obs

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        super().reset(seed=seed, options=options)
        if options and "config" in options:
            self.configure(options["config"])
        if repr(self.config) != self._built_for:
            self._build()
        self.time = self.steps = 0
        self.done = False
        reset_intersection(_SimResetBackend(self.sim), [self.np_random], self.net, self.table, self.config, self.cfg)
        obs = self._observe()
        return self._shuffle(obs), self._info(None)

    def _observe(self) -> np.ndarray:
        import torch

        buf = torch.zeros(self.sim.obs_size, dtype=torch.float32, device=f"cuda:{self.device_index}")
        self.sim.observe_ptr(buf.data_ptr(), int(torch.cuda.current_stream().cuda_stream))
        return buf.cpu().numpy().reshape(self._obs_shape)

    def _info(self, action) -> dict:
        st = self.sim.get_state()
        ego = int(st.env_i[abi.EI_EGO, 0])
        return {"speed": float(st.veh_d[abi.D_SPEED, 0, ego]),
                "crashed": bool(st.veh_i[abi.I_FLAGS, 0, ego] & abi.FL_CRASHED), "action": action}

    def step(self, action: int):
        if self.sim is None:
            raise NotImplementedError("The road and vehicle must be initialized in the environment implementation")
        # RNG order of the reference: observe() shuffles first, then _spawn_vehicle draws (intersection_env.py:135-139)
        shuffled = self.cfg.obs_type == abi.OBS_KINEMATICS and self.cfg.order == abi.ORDER_SHUFFLED
        perm = None
        if shuffled:
            perm = np.arange(self.cfg.obs_vehicles - 1)
            self.np_random.shuffle(perm)
        draws = (abi.SpawnDraw * 1)()
        saved = draw_spawn(self.np_random, float(self.config["spawn_probability"]), draws[0])
        self.sim.inject_spawn(draws)
        obs, reward, term, trunc = self.sim.step_host(np.array([int(action)], np.int32))
        if saved is not None and not self.sim.spawn_accepted()[0]:
            self.np_random.bit_generator.state = saved
        obs = obs.reshape(self._obs_shape)
        if perm is not None:
            obs[1:] = obs[1:][perm]
        self.time += 1 / self.config["policy_frequency"]
        self.steps += int(self.config["simulation_frequency"] // self.config["policy_frequency"])
        info = self._info(action)
        info["agents_rewards"] = (float(reward[0]),)
        info["agents_terminated"] = (bool(term[0]),)
        return obs, float(reward[0]), bool(term[0]), bool(trunc[0]), info


def _register_ttrl_envs() -> None:
    register(id="intersection-v0", entry_point="topotrafficrl_b200.envs:IntersectionEnv")


_register_ttrl_envs()
