"""Single-env, gymnasium-shaped front ends over the batched simulator, with the reference's ids, class names,
config keys, return tuple and RNG stream (``ttrl_env/__init__.py:22-56``, ``envs/common/abstract.py:188-250``,
``envs/intersection_env.py:17-139``, ``envs/roundabout_env.py``, ``envs/u_turn_env.py``).  ``scripts/example.py``-style
code runs unchanged apart from the import.

The env's ``np_random`` is a numpy ``Generator(PCG64)`` seeded like gymnasium; reset draws, spawn draws and the
"shuffled" observation permutations are taken from it in the reference's order (observe -> shuffle per agent, then
clear -> spawn), so a seeded episode reproduces the reference episode.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import abi, scenes
from ._gym import HAVE_GYMNASIUM, Env, Wrapper, register, spaces
from .reset import draw_spawn, reset_intersection, reset_roundabout, reset_uturn
from .sim import Sim
from .vector_env import _SimResetBackend


class AbstractEnv(Env):
    """Common part: config handling, spaces, the step/reset tuple (abstract.py:25-250)."""

    metadata = {"render_modes": ["human", "rgb_array"]}
    PERCEPTION_DISTANCE = 5.0 * 40.0
    SCENE = "intersection"
    VCAP = 32
    EGO_LANES = 1  # lanes of the ego's road at reset: default ``features_range`` of Kinematics (observation.py:213-225)

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

    # ---- provided by the scene -------------------------------------------------------------------------
    def _make_network(self):
        raise NotImplementedError

    def _exit_predicate(self):
        return None

    def _spawn_routes(self):
        return None

    def _reset_state(self) -> None:
        """Place the vehicles of a new episode on the device, drawing from ``self.np_random`` like the reference's ``_reset``."""
        raise NotImplementedError

    # ---- generic -----------------------------------------------------------------------------------------
    def _build(self) -> None:
        if self.sim is not None:
            self.sim.close()
        self.net = self._make_network()
        self.table = self.net.to_table(self._exit_predicate())
        self.cfg = scenes.build_config(self.table, self.config, self.SCENE, ego_lanes_count=self.EGO_LANES)
        self.sim = Sim(self.cfg, self.table, 1, self.VCAP, self.device_index, self._spawn_routes())
        self.sim.set_autoreset(False)
        self.sim.host_info(copy=False)  # info["speed"], ["crashed"], ["rewards"] of every step_host call
        self._info_dev = None
        self.num_agents = K = self.sim.num_agents
        shape = scenes.obs_shape(self.cfg)
        self._obs_shape = shape
        box = spaces.Box(low=-np.inf, high=np.inf, shape=shape, dtype=np.float32)
        act = spaces.Discrete(5 if self.cfg.action_mode == abi.ACT_ALL else 3)
        multi = self.config["observation"]["type"] == "MultiAgentObservation"
        self._multi = multi
        self.observation_space = spaces.Tuple([box] * K) if multi else box       # observation.py:597-600
        self.action_space = spaces.Tuple([act] * K) if self.config["action"]["type"] == "MultiAgentAction" else act  # action.py:311-314
        self._built_for = repr(self.config)

    def _shuffled(self) -> bool:
        return self.cfg.obs_type == abi.OBS_KINEMATICS and self.cfg.order == abi.ORDER_SHUFFLED

    def _format_obs(self, flat: np.ndarray, perms=None):
        obs = flat.reshape((self.num_agents,) + self._obs_shape)
        if perms is not None:
            for k, perm in enumerate(perms):
                obs[k, 1:] = obs[k, 1:][perm]
        return tuple(np.array(o) for o in obs) if self._multi else obs[0]

    def _draw_perms(self):
        """``np_random.shuffle(obs[1:])`` per agent observation (observation.py:272-273): the permutation depends on the
        row count only, so it is drawn on an index vector and applied to the device's rows."""
        if not self._shuffled():
            return None
        perms = []
        for _ in range(self.num_agents):
            perm = np.arange(self.cfg.obs_vehicles - 1)
            self.np_random.shuffle(perm)
            perms.append(perm)
        return perms

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        super().reset(seed=seed, options=options)
        if options and "config" in options:
            self.configure(options["config"])
        if repr(self.config) != self._built_for:
            self._build()
        self.time = self.steps = 0
        self.done = False
        self._reset_state()
        flat = self._observe()
        return self._format_obs(flat, self._draw_perms()), self._info(None)

    def _observe(self) -> np.ndarray:
        import torch

        buf = torch.zeros(self.sim.obs_size, dtype=torch.float32, device=f"cuda:{self.device_index}")
        if self._info_dev is None:
            self._info_dev = torch.zeros((abi.NINFO, 1), dtype=torch.float64, device=f"cuda:{self.device_index}")
            self.sim.set_info_outputs_ptr(self._info_dev.data_ptr(), None)
        self.sim.observe_ptr(buf.data_ptr(), int(torch.cuda.current_stream().cuda_stream))
        return buf.cpu().numpy()

    def _info(self, action, after_step: bool = False) -> dict:
        """``AbstractEnv._info`` (abstract.py:169-186): evaluated on the device together with the reward (before the
        clear / spawn of ``IntersectionEnv.step``); at reset by the observation kernel, with no action."""
        buf = self.sim.host_info(copy=False)[0][:, 0] if after_step else self._info_dev.cpu().numpy()[:, 0]
        return {"speed": float(buf[abi.INFO_SPEED]), "crashed": bool(buf[abi.INFO_CRASHED]), "action": action,
                "rewards": {k: float(buf[abi.INFO_REWARDS + i]) for i, k in enumerate(abi.REWARD_KEYS[self.cfg.reward_type])}}

    def step(self, action):
        if self.sim is None:
            raise NotImplementedError("The road and vehicle must be initialized in the environment implementation")
        K = self.num_agents
        if self.config["action"]["type"] == "MultiAgentAction":
            assert isinstance(action, tuple)  # action.py:321
            acts = np.array([int(a) for a in action], np.int32)
        else:
            acts = np.array([int(action)], np.int32)
        # RNG order of the reference: observe() shuffles first, then _spawn_vehicle draws (intersection_env.py:135-139)
        perms = self._draw_perms()
        saved = None
        if self.cfg.spawn_enabled:
            draws = (abi.SpawnDraw * 1)()
            saved = draw_spawn(self.np_random, float(self.config["spawn_probability"]), draws[0],
                               linear=self.cfg.vehicle_model == abi.VEHICLE_LINEAR)
            self.sim.inject_spawn(draws)
        obs, reward, term, trunc = self.sim.step_host(acts.reshape(1, K))
        if saved is not None and not self.sim.spawn_accepted()[0]:
            self.np_random.bit_generator.state = saved
        self.time += 1 / self.config["policy_frequency"]
        self.steps += int(self.config["simulation_frequency"] // self.config["policy_frequency"])
        info = self._info(action, after_step=True)
        if self.SCENE == "intersection":  # IntersectionEnv._info (intersection_env.py:121-129)
            if K > 1:
                ar, at = self.sim.agent_outputs_host()
                info["agents_rewards"] = tuple(float(x) for x in ar[0])
                info["agents_terminated"] = tuple(bool(x) for x in at[0])
            else:
                info["agents_rewards"] = (float(reward[0]),)
                info["agents_terminated"] = (bool(term[0]),)
        return self._format_obs(obs, perms), float(reward[0]), bool(term[0]), bool(trunc[0]), info

    def render(self):
        """``AbstractEnv.render`` (abstract.py:275-303) for ``render_mode="rgb_array"``: a headless numpy redraw of the reference's
        picture (``render.py``); there is no window for ``"human"`` (pygame is not a dependency and the GPU box has no display)."""
        if self.render_mode is None:
            return None
        if self.render_mode != "rgb_array":
            raise NotImplementedError("render_mode 'human' needs pygame and a display: use render_mode='rgb_array'")
        from .render import render_rgb
        return render_rgb(self.net, self.sim.get_state(), self.config, 0, linear_traffic=self.cfg.vehicle_model == abi.VEHICLE_LINEAR)

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

    def _make_network(self):
        return scenes.make_intersection_network()

    def _exit_predicate(self):
        return scenes.intersection_exit_predicate

    def _spawn_routes(self):
        return scenes.intersection_spawn_routes(self.net, self.table)

    def _reset_state(self) -> None:
        reset_intersection(_SimResetBackend(self.sim), [self.np_random], self.net, self.table, self.config, self.cfg)


class MultiAgentIntersectionEnv(IntersectionEnv):
    """intersection_env.py:372-394: K controlled vehicles, tuple actions and tuple observations."""

    @classmethod
    def default_config(cls) -> dict:
        return scenes.merged_config(scenes.MULTI_AGENT_INTERSECTION_CONFIG, None)


class RoundaboutEnv(AbstractEnv):
    SCENE = "roundabout"
    VCAP = 16

    @classmethod
    def default_config(cls) -> dict:
        return scenes.merged_config(scenes.ROUNDABOUT_CONFIG, None)

    def _make_network(self):
        return scenes.make_roundabout_network()

    def _reset_state(self) -> None:
        self.sim.set_state(reset_roundabout([self.np_random], self.net, self.table, self.config, self.cfg, self.VCAP))


class UTurnEnv(AbstractEnv):
    """u_turn_env.py (default observation: ``TimeToCollision`` with a 16 s horizon)."""
    SCENE = "u-turn"
    VCAP = 16
    EGO_LANES = 2

    @classmethod
    def default_config(cls) -> dict:
        return scenes.merged_config(scenes.UTURN_CONFIG, None)

    def _make_network(self):
        return scenes.make_uturn_network()

    def _reset_state(self) -> None:
        self.sim.set_state(reset_uturn([self.np_random], self.net, self.table, self.config, self.cfg, self.VCAP))


class MultiAgentWrapper(Wrapper):
    """abstract.py:432-441: per-agent rewards and terminal flags instead of the aggregated ones."""

    def step(self, action):
        obs, _, _, truncated, info = self.env.step(action)
        return obs, info["agents_rewards"], info["agents_terminated"], truncated, info


def _register_ttrl_envs() -> None:
    """Same ids as ``ttrl_env/__init__.py:22-56`` (``intersection-v1`` = ContinuousAction: outside the hot path)."""
    register(id="intersection-v0", entry_point="topotrafficrl_b200.envs:IntersectionEnv")
    register(id="intersection-multi-agent-v0", entry_point="topotrafficrl_b200.envs:MultiAgentIntersectionEnv")
    if HAVE_GYMNASIUM:  # pragma: no cover - gymnasium is not in this image
        from gymnasium.envs.registration import WrapperSpec
        wrappers = (WrapperSpec("MultiAgentWrapper", "topotrafficrl_b200.envs:MultiAgentWrapper", None),)
    else:
        wrappers = (MultiAgentWrapper,)
    register(id="intersection-multi-agent-v1", entry_point="topotrafficrl_b200.envs:MultiAgentIntersectionEnv",
             additional_wrappers=wrappers)
    register(id="roundabout-v0", entry_point="topotrafficrl_b200.envs:RoundaboutEnv")
    register(id="u-turn-v0", entry_point="topotrafficrl_b200.envs:UTurnEnv")


_register_ttrl_envs()
