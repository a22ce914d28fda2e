"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around ``libttrl_oracle.so`` (the CPU restatement in
``ttrl_oracle.c``).  Imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg; never by the product package ``topotrafficrl_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from topotrafficrl_b200 import abi
from topotrafficrl_b200.road import NetworkTable
from topotrafficrl_b200.state import SimState

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libttrl_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ttrl_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_scene_create.restype = C.c_void_p
        _lib.orc_wrap_to_pi.restype = C.c_double
        _lib.orc_wrap_to_pi.argtypes = [C.c_double]
        _lib.orc_not_zero.restype = C.c_double
        _lib.orc_not_zero.argtypes = [C.c_double]
        _lib.orc_lane_heading_at.restype = C.c_double
        _lib.orc_lane_distance_with_heading.restype = C.c_double
        _lib.orc_steering_control.restype = C.c_double
        _lib.orc_idm_acceleration.restype = C.c_double
    return _lib


def _p(a: Optional[np.ndarray], t=None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One scene (network + config) of the CPU oracle; state lives in caller-owned :class:`SimState`."""

    def __init__(self, cfg: abi.Config, table: NetworkTable, spawn_routes=None, threads: int = 1) -> None:
        self.cfg, self.table, self.threads = cfg, table, threads
        L = lib()
        self._L = L
        self._sc = C.c_void_p(L.orc_scene_create(C.byref(cfg), table.lanes, table.roads,
                                                 _p(table.node_first), _p(table.node_roads)))
        if spawn_routes is not None:
            sl, rl, rr = (np.ascontiguousarray(a, dtype=np.int32) for a in spawn_routes)
            L.orc_scene_set_spawn_routes(self._sc, _p(sl), _p(rl), _p(rr))
        L.orc_obs_size.argtypes = [C.c_void_p]
        self.obs_size = L.orc_obs_size(self._sc)
        self.K = max(1, int(cfg.controlled_vehicles))
        self.agent_reward = self.agent_terminated = None  # per-agent outputs of the last step()
        self._pool_lin = None

    def _lin(self, st) -> None:
        """LinearVehicle parameter blocks (``SimState.lin``) of the state the next call operates on and of the pool."""
        self._L.orc_scene_set_linear_params(self._sc, _p(getattr(st, "lin", None)), _p(self._pool_lin))

    def __del__(self):
        try:
            self._L.orc_scene_destroy(self._sc)
        except Exception:
            pass

    def set_reset_pool(self, pool: SimState) -> None:
        pool = pool.contiguous()
        self._pool_keep, self._pool_lin = pool, pool.lin
        self._L.orc_scene_set_reset_pool(self._sc, C.c_int(pool.num_envs), C.c_int(pool.vcap), _p(pool.veh_d), _p(pool.veh_i),
                                         _p(pool.env_i), _p(pool.env_d))

    def set_autoreset(self, on: bool) -> None:
        self._L.orc_scene_set_autoreset(self._sc, C.c_int(int(on)))

    def substep(self, st: SimState, actions: Optional[np.ndarray] = None) -> None:
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32)
        self._lin(st)
        self._L.orc_substep(self._sc, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d),
                            C.c_int(st.num_envs), C.c_int(st.vcap), _p(a), C.c_int(self.threads))

    def observe(self, st: SimState) -> np.ndarray:
        obs = np.zeros((st.num_envs, self.obs_size), np.float32)
        self._lin(st)
        self._L.orc_observe(self._sc, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d),
                            C.c_int(st.num_envs), C.c_int(st.vcap), _p(obs), C.c_int(self.threads))
        return obs

    def step(self, st: SimState, actions: Optional[np.ndarray], draws=None, stats: Optional[np.ndarray] = None):
        E = st.num_envs
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((E, self.obs_size), np.float32)
        reward = np.zeros(E, np.float32)
        term = np.zeros(E, np.uint8)
        trunc = np.zeros(E, np.uint8)
        accepted = np.zeros(E, np.int32)
        self.agent_reward = np.zeros((E, self.K), np.float32)
        self.agent_terminated = np.zeros((E, self.K), np.uint8)
        self._lin(st)
        self._L.orc_step_agents(self._sc, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d), C.c_int(E), C.c_int(st.vcap),
                                _p(a), _p(obs), _p(reward), _p(term), _p(trunc),
                                draws if draws is not None else None, _p(accepted), _p(stats), C.c_int(self.threads),
                                _p(self.agent_reward), _p(self.agent_terminated))
        return obs, reward, term, trunc, accepted

    def spawn(self, st: SimState, draws, longitudinal: float, position_deviation: float = 1.0,
              speed_deviation: float = 1.0, spawn_probability: float = 0.6, go_straight: bool = False) -> np.ndarray:
        accepted = np.zeros(st.num_envs, np.int32)
        self._lin(st)
        self._L.orc_spawn(self._sc, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d), C.c_int(st.num_envs),
                          C.c_int(st.vcap), draws, C.c_double(longitudinal), C.c_double(position_deviation),
                          C.c_double(speed_deviation), C.c_double(spawn_probability), C.c_int(int(go_straight)), _p(accepted))
        return accepted

    # ---- function-level entry points (known-answer tests against the reference) -------------------
    def lane_local(self, lane: int, x: float, y: float):
        out = (C.c_double * 2)()
        self._L.orc_lane_local(C.byref(self.table.lanes[lane]), C.c_double(x), C.c_double(y), out)
        return out[0], out[1]

    def lane_position(self, lane: int, s: float, r: float):
        out = (C.c_double * 2)()
        self._L.orc_lane_position(C.byref(self.table.lanes[lane]), C.c_double(s), C.c_double(r), out)
        return out[0], out[1]

    def lane_heading_at(self, lane: int, s: float) -> float:
        return self._L.orc_lane_heading_at(C.byref(self.table.lanes[lane]), C.c_double(s))

    def lane_distance_with_heading(self, lane: int, x: float, y: float, h: float) -> float:
        return self._L.orc_lane_distance_with_heading(C.byref(self.table.lanes[lane]), C.c_double(x), C.c_double(y), C.c_double(h))

    def closest_lane(self, x: float, y: float, h: float) -> int:
        return self._L.orc_closest_lane(self._sc, C.c_double(x), C.c_double(y), C.c_double(h))

    def steering_control(self, x, y, h, speed, target_lane) -> float:
        return self._L.orc_steering_control(self._sc, C.c_double(x), C.c_double(y), C.c_double(h), C.c_double(speed), C.c_int(target_lane))

    def speed_to_index(self, speed: float) -> int:
        return self._L.orc_speed_to_index(self._sc, C.c_double(speed))

    def idm_acceleration(self, self_delta, ego6, front6=None) -> float:
        e = (C.c_double * 6)(*ego6)
        f = (C.c_double * 6)(*front6) if front6 is not None else None
        return self._L.orc_idm_acceleration(self._sc, C.c_double(self_delta), e, f)

    def polygons_intersecting(self, va, vb, dt):
        out = (C.c_double * 4)()
        self._L.orc_polygons_intersecting((C.c_double * 4)(*va), (C.c_double * 4)(*vb), C.c_double(dt), out)
        return bool(out[0]), bool(out[1]), (out[2], out[3])

    def rotated_rectangles_intersect(self, r1, r2) -> bool:
        return bool(self._L.orc_rotated_rectangles_intersect((C.c_double * 5)(*r1), (C.c_double * 5)(*r2)))


def wrap_to_pi(x: float) -> float:
    return lib().orc_wrap_to_pi(x)


def not_zero(x: float) -> float:
    return lib().orc_not_zero(x)
