"""``TTRLVectorEnv``: E env instances advanced in lockstep on one B200 (the vector form of the reference's
gymnasium ``AbstractEnv`` API, abstract.py:188-250).  Observations, rewards and flags stay on the device as
torch tensors (PyTorch is used for device memory and streams only); ``step_host`` is the numpy/host-buffer form.

There is no CPU fallback: constructing the env without the CUDA library or without a CUDA device raises.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import abi, scenes
from ._gym import np_random as gym_np_random
from ._gym import spaces
from .reset import reset_intersection, reset_roundabout, reset_uturn
from .sim import Sim
from .state import SimState


class _SimResetBackend:
    def __init__(self, sim: Sim):
        self.sim, self.num_envs = sim, sim.num_envs

    def spawn(self, draws, longitudinal, position_deviation, speed_deviation, spawn_probability, go_straight):
        return self.sim.spawn(draws, longitudinal, position_deviation, speed_deviation, spawn_probability, go_straight)

    def substep_none(self):
        self.sim.substep_ptr(None, 0)

    def get_state(self):
        return self.sim.get_state()

    def set_state(self, st):
        self.sim.set_state(st)


class TTRLVectorEnv:
    """
    :param num_envs: number of env instances on this device
    :param scene: ``"highway"`` (synthetic multi-lane highway, BASELINE configs 2/3), ``"intersection"``
                  (``IntersectionEnv``, configs 1/4; ``config["controlled_vehicles"] = K > 1`` gives the multi-agent form:
                  actions ``[E, K]``, observations ``[E, K, ...]``, ``info["agents_rewards"]`` / ``["agents_terminated"]``
                  ``[E, K]``), ``"roundabout"`` (``RoundaboutEnv``) or ``"u-turn"`` (``UTurnEnv``)
    :param config: reference-style config dict, shallow-merged over the scene default (abstract.py:111-113)
    :param device: CUDA device index or ``"cuda:N"``
    :param seed: base seed; env e uses the stream keyed by ``first_env + e`` so shards reproduce a larger run
    :param first_env: global index of this shard's first env (multi-GPU sharding)
    :param async_reset: with ``reset_mode="device"`` and a scene whose reset has warm-up sub-steps (the intersection: 45),
                  generate every env's next episode ahead of time on a side stream (``TTRL_AUTORESET_DEVICE_ASYNC``) instead of
                  right after the step that finished it; the episodes are the same bit for bit.
    :param reset_mode: ``"device"`` (default): episodes are generated on the GPU (``ttrl_sim_reset``; Philox draws keyed
                  by (seed, global env, episode)) at ``reset()`` and, with ``autoreset``, inside the step kernel when an
                  env finishes.  ``"host"``: the first reset is driven from the host with numpy ``Generator(PCG64)``
                  streams seeded like gymnasium (the reference's own reset for the intersection scene) and finished
                  envs restart from that pool of initial states (roundabout / u-turn: ``pool_factor`` initial states
                  per env, entry p = the reference's ``reset(seed + p)``).
    """

    def __init__(self, num_envs: int, scene: str = "highway", config: Optional[dict] = None, device=0, seed: int = 0,
                 first_env: int = 0, vcap: Optional[int] = None, autoreset: bool = True, reset_mode: Optional[str] = None,
                 pool_factor: int = 4, async_reset: bool = True) -> None:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("TTRLVectorEnv needs a CUDA device (there is no CPU fallback)")
        self.torch = torch
        self.num_envs = int(num_envs)
        self.scene = scene
        self.device_index = int(str(device).split(":")[-1]) if not isinstance(device, int) else device
        self.device = torch.device("cuda", self.device_index)
        self.seed_value, self.first_env = int(seed), int(first_env)
        if scene == "highway":
            self.config = scenes.merged_config(scenes.HIGHWAY_CONFIG, config)
            self.net = scenes.make_highway_network(int(self.config["lanes_count"]), float(self.config["road_length"]),
                                                   float(self.config["speed_limit"]))
            self.table = self.net.to_table()
            self.cfg = scenes.build_config(self.table, self.config, "highway", ego_lanes_count=int(self.config["lanes_count"]))
            self.vcap = int(vcap or self.config["vehicles_count"])
            routes = None
        elif scene == "intersection":
            self.config = scenes.merged_config(scenes.INTERSECTION_CONFIG, config)
            self.net = scenes.make_intersection_network()
            self.table = self.net.to_table(scenes.intersection_exit_predicate)
            self.cfg = scenes.build_config(self.table, self.config, "intersection")
            # slot capacity: the reference has no cap; its episodes peak at 15 vehicles, a spawn is
            # rejected when all slots are taken
            self.vcap = int(vcap or 24)
            routes = scenes.intersection_spawn_routes(self.net, self.table)
        elif scene in ("roundabout", "u-turn"):
            self.config = scenes.merged_config(scenes.ROUNDABOUT_CONFIG if scene == "roundabout" else scenes.UTURN_CONFIG, config)
            self.net = scenes.make_roundabout_network() if scene == "roundabout" else scenes.make_uturn_network()
            self.table = self.net.to_table()
            self.cfg = scenes.build_config(self.table, self.config, scene, ego_lanes_count=1 if scene == "roundabout" else 2)
            self.vcap = int(vcap or 16)
            routes = None
        else:
            raise ValueError(f"unknown scene {scene!r}")
        scripted = scene in ("roundabout", "u-turn")
        if reset_mode is None:
            reset_mode = "device"
        if reset_mode not in ("device", "host"):
            raise ValueError(f"unknown reset_mode {reset_mode!r}")
        self.reset_mode = reset_mode
        self.pool_factor = max(1, int(pool_factor))
        self.sim = Sim(self.cfg, self.table, self.num_envs, self.vcap, self.device_index, routes)
        self.num_agents = self.sim.num_agents
        if scripted:
            self.sim.set_reset_params(scenes.cast_reset_params(scene, self.net, self.table, self.config))
        else:
            self.sim.set_reset_params(scenes.highway_reset_params(self.config) if scene == "highway"
                                      else scenes.intersection_reset_params(self.config))
        self.autoreset = autoreset
        # scenes whose reset runs warm-up sub-steps regenerate the next episodes on a side stream (same episodes, bit for bit)
        device_mode = abi.AUTORESET_DEVICE_ASYNC if (async_reset and scene == "intersection") else abi.AUTORESET_DEVICE
        self.sim.set_autoreset(0 if not autoreset else (device_mode if reset_mode == "device" else abi.AUTORESET_POOL))
        self.obs_shape = scenes.obs_shape(self.cfg)
        n_actions = 5 if self.cfg.action_mode == abi.ACT_ALL else 3
        self.single_observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=self.obs_shape, dtype=np.float32)
        self.single_action_space = spaces.Discrete(n_actions)
        E, K = self.num_envs, self.num_agents
        if K > 1:  # MultiAgentObservation: one observation per controlled vehicle
            self.obs_shape = (K,) + self.obs_shape
        self._agent_reward = torch.zeros((E, K), dtype=torch.float32, device=self.device)
        self._agent_term = torch.zeros((E, K), dtype=torch.uint8, device=self.device)
        self.sim.set_agent_outputs_ptr(self._agent_reward.data_ptr(), self._agent_term.data_ptr())
        self._obs = torch.zeros((E,) + self.obs_shape, dtype=torch.float32, device=self.device)
        self._reward = torch.zeros(E, dtype=torch.float32, device=self.device)
        self._term = torch.zeros(E, dtype=torch.uint8, device=self.device)
        self._trunc = torch.zeros(E, dtype=torch.uint8, device=self.device)
        # batched info dict (AbstractEnv._info abstract.py:169-186) and gymnasium's final_observation, written by the step kernel
        self._info_buf = torch.zeros((abi.NINFO, E), dtype=torch.float64, device=self.device)
        self._final_obs = torch.zeros((E,) + self.obs_shape, dtype=torch.float32, device=self.device)
        self.sim.set_info_outputs_ptr(self._info_buf.data_ptr(), self._final_obs.data_ptr())
        self._reward_keys = abi.REWARD_KEYS[self.cfg.reward_type]

    # ------------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return int(self.torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, seed: Optional[int] = None):
        if seed is not None:
            self.seed_value = int(seed)
        self.sim.seed((self.seed_value << 1) | 1, self.first_env)  # device-side Philox draws (spawns, resets)
        if self.reset_mode == "device":
            self.sim.reset_device(None, self._stream())
            self.sim.observe_ptr(self._obs.data_ptr(), self._stream())
            return self._obs, {}
        if self.scene == "highway":
            st = scenes.make_highway_state(self.num_envs, self.config, seed=self.seed_value, first_env=self.first_env, vcap=self.vcap)
            self.sim.set_state(st)
        elif self.scene in ("roundabout", "u-turn"):
            # pool entry p of this shard = the reference's reset(seed = seed + global index p); env e starts from entry e
            n_pool = self.num_envs * self.pool_factor
            rngs = [gym_np_random(self.seed_value + self.first_env * self.pool_factor + p)[0] for p in range(n_pool)]
            make = reset_roundabout if self.scene == "roundabout" else reset_uturn
            pool = make(rngs, self.net, self.table, self.config, self.cfg, self.vcap)
            st = pool.slice_envs(0, self.num_envs)
            self.sim.set_state(st)
            self.sim.set_reset_pool(pool)
            self.sim.observe_ptr(self._obs.data_ptr(), self._stream())
            return self._obs, {}
        else:
            rngs = [gym_np_random(self.seed_value + self.first_env + e)[0] for e in range(self.num_envs)]
            st = reset_intersection(_SimResetBackend(self.sim), rngs, self.net, self.table, self.config, self.cfg)
        self.sim.set_reset_pool(st)
        self.sim.observe_ptr(self._obs.data_ptr(), self._stream())
        return self._obs, {}

    def step(self, actions):
        """actions: int tensor [E] (``[E, K]`` with K controlled vehicles) on this device (int32 preferred; int64 is converted)."""
        torch = self.torch
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions), device=self.device)
        if actions.dtype != torch.int32:
            actions = actions.to(torch.int32)
        actions = actions.contiguous()
        self.sim.step_ptr(actions.data_ptr(), self._obs.data_ptr(), self._reward.data_ptr(), self._term.data_ptr(),
                          self._trunc.data_ptr(), self._stream())
        return self._obs, self._reward, self._term.bool(), self._trunc.bool(), self._info()

    def _info(self) -> dict:
        """The reference's ``info`` dict with one device tensor per key (views of buffers the next ``step`` overwrites):
        ``speed`` / ``crashed`` of ``controlled_vehicles[0]``, ``rewards`` = ``_rewards(action)``, ``agents_rewards`` /
        ``agents_terminated`` (``[E, K]``), and gymnasium's autoreset pair: ``final_observation`` ``[E, ...]`` holds the
        terminal observation of the envs flagged in ``_final_observation`` (= terminated | truncated; their row of the
        returned observation already belongs to the next episode)."""
        buf = self._info_buf
        info = {"speed": buf[abi.INFO_SPEED], "crashed": buf[abi.INFO_CRASHED] != 0,
                "rewards": {k: buf[abi.INFO_REWARDS + i] for i, k in enumerate(self._reward_keys)},
                "agents_rewards": self._agent_reward, "agents_terminated": self._agent_term.bool()}
        if self.autoreset:
            info["final_observation"] = self._final_obs
            info["_final_observation"] = (self._term | self._trunc).bool()
        return info

    def step_host(self, actions: np.ndarray, copy: bool = True, with_info: bool = False):
        """numpy in / numpy out.  ``copy=False``: zero-copy views of the page-locked staging buffers (valid until the next call).
        ``with_info``: also fetch ``speed`` / ``crashed`` / ``rewards`` / ``final_observation`` (extra D2H copies every step)."""
        if with_info:
            self.sim.host_info(copy=False)  # switches the info / final-observation outputs of the host-buffer path on
        obs, reward, term, trunc = self.sim.step_host(actions, copy=copy)
        ar, at = self.sim.agent_outputs_host(copy=copy)
        info = {"agents_rewards": ar, "agents_terminated": at.view(np.bool_)}
        if with_info:
            buf, fo = self.sim.host_info(copy=copy)
            info.update({"speed": buf[abi.INFO_SPEED], "crashed": buf[abi.INFO_CRASHED] != 0,
                         "rewards": {k: buf[abi.INFO_REWARDS + i] for i, k in enumerate(self._reward_keys)}})
            if self.autoreset:
                info["final_observation"] = fo.reshape((self.num_envs,) + self.obs_shape)
                info["_final_observation"] = (term | trunc).view(np.bool_)
        return obs.reshape((self.num_envs,) + self.obs_shape), reward, term.view(np.bool_), trunc.view(np.bool_), info

    def get_state(self) -> SimState:
        return self.sim.get_state()

    def set_state(self, st: SimState) -> None:
        self.sim.set_state(st)

    def stats(self, reset: bool = False) -> dict:
        s = self.sim.stats(reset)
        return {k: getattr(s, k) for k, _ in s._fields_}

    def close(self) -> None:
        self.sim.close()
