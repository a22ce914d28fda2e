"""Host-side container for the struct-of-arrays simulation state (layout of ``include/ttrl_b200.h``).

Field ``f`` of slot ``s`` of env ``e`` is ``veh_d[f, e, s]`` / ``veh_i[f, e, s]``; slots ``0..n-1`` are live, in
``Road.vehicles`` list order (the order is semantic in the reference: road.py:461-478).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi


def pack_route(route: Optional[Sequence[Tuple[int, Optional[int]]]]) -> Tuple[int, Tuple[int, ...], Tuple[int, ...]]:
    """``route`` = list of (road index, lane id | None), or None.  Returns (len, road words, lane words): entry k is
    byte ``k % 4`` of word ``k // 4`` (``abi.ROUTE_WORDS`` int32 words each)."""
    if route is None:
        return -1, (0,) * abi.ROUTE_WORDS, (0,) * abi.ROUTE_WORDS
    if len(route) > abi.ROUTE_CAP:
        raise ValueError(f"route longer than {abi.ROUTE_CAP} entries")
    rr = [0] * abi.ROUTE_WORDS
    rl = [0] * abi.ROUTE_WORDS
    for k, (road, lane) in enumerate(route):
        rr[k // 4] |= (int(road) & 0xFF) << (8 * (k % 4))
        rl[k // 4] |= (0xFF if lane is None else int(lane) & 0xFF) << (8 * (k % 4))
    return len(route), tuple(_as_i32(w) for w in rr), tuple(_as_i32(w) for w in rl)


def unpack_route(length: int, rr: Sequence[int], rl: Sequence[int]) -> Optional[List[Tuple[int, Optional[int]]]]:
    """Inverse of :func:`pack_route`; ``rr`` / ``rl`` are the road / lane words (a single int = word 0 only)."""
    if length < 0:
        return None
    if isinstance(rr, (int, np.integer)):
        rr, rl = (rr,), (rl,)
    rr = [int(w) & 0xFFFFFFFF for w in rr]
    rl = [int(w) & 0xFFFFFFFF for w in rl]
    out = []
    for k in range(length):
        lane = (rl[k // 4] >> (8 * (k % 4))) & 0xFF
        out.append(((rr[k // 4] >> (8 * (k % 4))) & 0xFF, None if lane == 0xFF else lane))
    return out


def _as_i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


@dataclass
class SimState:
    veh_d: np.ndarray  # float64 [ND, E, V]
    veh_i: np.ndarray  # int32   [NI, E, V]
    env_i: np.ndarray  # int32   [NEI, E]
    env_d: np.ndarray  # float64 [NED, E]
    lin: Optional[np.ndarray] = None  # float64 [NLIN, E, V]: LinearVehicle parameters (``vehicle_model`` linear), else None

    @classmethod
    def zeros(cls, num_envs: int, vcap: int, linear: bool = False) -> "SimState":
        return cls(np.zeros((abi.ND, num_envs, vcap), np.float64), np.zeros((abi.NI, num_envs, vcap), np.int32),
                   np.zeros((abi.NEI, num_envs), np.int32), np.zeros((abi.NED, num_envs), np.float64),
                   np.zeros((abi.NLIN, num_envs, vcap), np.float64) if linear else None)

    @property
    def num_envs(self) -> int:
        return self.veh_d.shape[1]

    @property
    def vcap(self) -> int:
        return self.veh_d.shape[2]

    def _map(self, f) -> "SimState":
        return SimState(f(self.veh_d), f(self.veh_i), f(self.env_i), f(self.env_d), None if self.lin is None else f(self.lin))

    def copy(self) -> "SimState":
        return self._map(lambda a: a.copy())

    def contiguous(self) -> "SimState":
        return self._map(np.ascontiguousarray)

    def slice_envs(self, lo: int, hi: int) -> "SimState":
        return self._map(lambda a: a[:, lo:hi].copy())

    def select_envs(self, idx) -> "SimState":
        idx = np.asarray(idx)
        return self._map(lambda a: a[:, idx].copy())

    def n_vehicles(self) -> np.ndarray:
        return self.env_i[abi.EI_NVEH]

    def live_mask(self) -> np.ndarray:
        return np.arange(self.vcap)[None, :] < self.env_i[abi.EI_NVEH][:, None]

    def set_vehicle(self, e: int, s: int, *, x, y, heading, speed, lane, target_lane=None, target_speed=None,
                    timer=0.0, delta=4.0, mdp=False, controlled=False, crashed=False, speed_index=0,
                    route=None, steering=0.0, accel=0.0, impact=None, yielding=False, yield_timer=0, agent=0,
                    linear=None) -> None:
        d, i = self.veh_d, self.veh_i
        d[abi.D_X, e, s], d[abi.D_Y, e, s] = x, y
        d[abi.D_HEADING, e, s], d[abi.D_SPEED, e, s] = heading, speed
        d[abi.D_STEERING, e, s], d[abi.D_ACCEL, e, s] = steering, accel
        d[abi.D_TARGET_SPEED, e, s] = speed if target_speed is None else target_speed
        d[abi.D_TIMER, e, s], d[abi.D_DELTA, e, s] = timer, delta
        flags = (abi.FL_MDP if mdp else 0) | (abi.FL_CONTROLLED if controlled else 0) | (abi.FL_CRASHED if crashed else 0)
        flags |= (int(agent) << abi.FL_AGENT_SHIFT) & abi.FL_AGENT_MASK
        if impact is not None:
            d[abi.D_IMPACT_X, e, s], d[abi.D_IMPACT_Y, e, s] = impact
            flags |= abi.FL_HAS_IMPACT
        else:
            d[abi.D_IMPACT_X, e, s] = d[abi.D_IMPACT_Y, e, s] = 0.0
        if yielding:
            flags |= abi.FL_YIELDING
        i[abi.I_LANE, e, s] = lane
        i[abi.I_TARGET_LANE, e, s] = lane if target_lane is None else target_lane
        i[abi.I_FLAGS, e, s] = flags
        i[abi.I_SPEED_INDEX, e, s] = speed_index
        i[abi.I_ROUTE_LEN, e, s], rr, rl = pack_route(route)
        for w in range(abi.ROUTE_WORDS):
            i[abi.I_ROUTE_ROAD_WORDS[w], e, s], i[abi.I_ROUTE_LANE_WORDS[w], e, s] = rr[w], rl[w]
        i[abi.I_YIELD_TIMER, e, s] = yield_timer
        if self.lin is not None:  # ACCELERATION_PARAMETERS + STEERING_PARAMETERS of a LinearVehicle (zeros for the controlled ones)
            self.lin[:, e, s] = 0.0 if linear is None else np.asarray(linear, np.float64)
