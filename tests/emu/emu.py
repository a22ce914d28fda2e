"""TEST-ONLY ctypes wrapper around tests/emu/libttrl_emu.so (host emulation of the CUDA phases)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from topotrafficrl_b200 import abi
from topotrafficrl_b200.state import SimState

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libttrl_emu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_LIB)
        _lib.emu_scene_create.restype = C.c_void_p
    return _lib


def _lin(st, pool=None):
    """Tell the emulator which LinearVehicle parameter blocks the next call operates on (None: class defaults)."""
    lib().emu_set_linear_params(_p(getattr(st, "lin", None)), _p(getattr(pool, "lin", None)) if pool is not None else None)


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Emulator:
    def __init__(self, cfg: abi.Config, table, spawn_routes=None):
        self.L = lib()
        self.cfg = cfg
        self.sc = C.c_void_p(self.L.emu_scene_create(C.byref(cfg), table.lanes, table.roads, _p(table.node_first), _p(table.node_roads)))
        if spawn_routes is not None:
            sl, rl, rr = (np.ascontiguousarray(a, dtype=np.int32) for a in spawn_routes)
            self.L.emu_scene_set_spawn_routes(self.sc, _p(sl), _p(rl), _p(rr))
        self.K = max(1, int(cfg.controlled_vehicles))
        from topotrafficrl_b200 import scenes
        self.obs_size = self.K * int(np.prod(scenes.obs_shape(cfg)))
        self.agent_reward = self.agent_terminated = None  # per-agent outputs of the last step()
        self.pool = None
        self.autoreset = False

    def substep(self, st: SimState, actions=None):
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32)
        _lin(st)
        self.L.emu_substep(self.sc, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d), C.c_int(st.num_envs), C.c_int(st.vcap), _p(a))

    def observe(self, st: SimState, inv_perm=None):
        obs = np.zeros((st.num_envs, self.obs_size), np.float32)
        _lin(st)
        self.L.emu_observe(self.sc, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d), C.c_int(st.num_envs), C.c_int(st.vcap),
                           _p(obs), C.c_int(self.obs_size), _p(inv_perm))
        return obs

    def step(self, st: SimState, actions, draws=None, inv_perm=None, stats=None, seed=0, first_env=0):
        E = st.num_envs
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((E, self.obs_size), np.float32)
        reward = np.zeros(E, np.float32)
        term = np.zeros(E, np.uint8)
        trunc = np.zeros(E, np.uint8)
        acc = np.zeros(E, np.int32)
        self.agent_reward = np.zeros((E, self.K), np.float32)
        self.agent_terminated = np.zeros((E, self.K), np.uint8)
        self.info = np.zeros((abi.NINFO, E), np.float64)          # ttrl_sim_set_info_outputs
        self.final_obs = np.zeros((E, self.obs_size), np.float32)
        pool = self.pool
        _lin(st, pool)
        self.L.emu_step(self.sc, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d), C.c_int(E), C.c_int(st.vcap), _p(a),
                        _p(obs), C.c_int(self.obs_size), _p(reward), _p(term), _p(trunc), draws, _p(acc), _p(inv_perm), _p(stats),
                        C.c_int(pool.num_envs if pool is not None else 0),
                        _p(pool.veh_d) if pool is not None else None, _p(pool.veh_i) if pool is not None else None,
                        _p(pool.env_i) if pool is not None else None, _p(pool.env_d) if pool is not None else None,
                        C.c_int(int(self.autoreset)), C.c_uint64(seed), C.c_int64(first_env), _p(self.agent_reward), _p(self.agent_terminated),
                        _p(self.info), _p(self.final_obs))
        return obs, reward, term, trunc, acc

    def set_reset_params(self, rp):
        self.L.emu_scene_set_reset_params(self.sc, C.byref(rp))

    def reset(self, st: SimState, seed: int, first_env: int = 0, episode: int = 0):
        _lin(st)
        self.L.emu_reset(self.sc, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d), C.c_int(st.num_envs), C.c_int(st.vcap),
                         C.c_uint64(seed), C.c_int64(first_env), C.c_int(episode))

    def reset_attempt_draw(self, seed: int, env: int, episode: int, attempt: int) -> abi.SpawnDraw:
        d = abi.SpawnDraw()
        self.L.emu_reset_attempt_draw(C.c_uint64(seed), C.c_int64(env), C.c_int(episode), C.c_int(attempt), C.byref(d))
        return d

    def reset_uniforms(self, seed: int, env: int, episode: int, idx: int):
        out = (C.c_double * 2)()
        self.L.emu_reset_uniforms(C.c_uint64(seed), C.c_int64(env), C.c_int(episode), C.c_uint32(idx), out)
        return out[0], out[1]

    def spawn(self, st, draws, longitudinal, position_deviation=1.0, speed_deviation=1.0, spawn_probability=0.6, go_straight=False):
        acc = np.zeros(st.num_envs, np.int32)
        _lin(st)
        self.L.emu_spawn(self.sc, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d), C.c_int(st.num_envs), C.c_int(st.vcap), draws,
                         C.c_double(longitudinal), C.c_double(position_deviation), C.c_double(speed_deviation),
                         C.c_double(spawn_probability), C.c_int(int(go_straight)), _p(acc))
        return acc
