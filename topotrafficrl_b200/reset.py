"""Host-driven reset of the intersection scene (``IntersectionEnv._make_vehicles``, reference
intersection_env.py:251-318), expressed over device primitives: spawn attempts and warm-up sub-steps run in
the CUDA library; the host only draws the random numbers -- from the same numpy ``Generator(PCG64)`` stream,
in the same order, as the reference -- and places the ego.  With one Generator per env seeded like
``gymnasium.Env.reset(seed)`` the post-reset state is the reference's own (tests/golden/intersection_reset.npz).

Reset is not on the per-step hot path: after the first reset, finished envs
restart from the reset pool inside the step kernel.
"""
from __future__ import annotations

from typing import List, Protocol, Sequence

import numpy as np

from . import abi
from .road import NetworkTable, RoadNetwork
from .state import SimState


class ResetBackend(Protocol):
    num_envs: int

    def spawn(self, draws, longitudinal: float, position_deviation: float, speed_deviation: float,
              spawn_probability: float, go_straight: bool) -> np.ndarray: ...

    def substep_none(self) -> None: ...

    def get_state(self) -> SimState: ...

    def set_state(self, st: SimState) -> None: ...


def draw_spawn(rng: np.random.Generator, spawn_probability: float, rec: abi.SpawnDraw):
    """Consume the draws of one ``_spawn_vehicle`` call up to (and speculatively including) the behaviour
    draw; returns the generator state to restore if the spawn is rejected by the 15 m rule
    (the reference only draws DELTA for accepted vehicles: intersection_env.py:342-346)."""
    rec.u_spawn = rng.uniform()
    rec.entry = rec.exit = 0
    rec.n_pos = rec.n_speed = 0.0
    rec.delta = 4.0
    if rec.u_spawn > spawn_probability:
        return None
    route = rng.choice(range(4), size=2, replace=False)
    rec.entry, rec.exit = int(route[0]), int(route[1])
    rec.n_pos = rng.normal()
    rec.n_speed = rng.normal()
    saved = rng.bit_generator.state
    rec.delta = rng.uniform(low=3.5, high=4.5)
    return saved


def reset_intersection(backend: ResetBackend, rngs: Sequence[np.random.Generator], net: RoadNetwork, table: NetworkTable,
                       config: dict, cfg: abi.Config) -> SimState:
    E = backend.num_envs
    if int(config.get("controlled_vehicles", 1)) != 1:
        raise NotImplementedError("multi-agent intersection is outside the round-1 hot path (SURVEY.md section 8f, N3)")
    n_vehicles = int(config["initial_vehicle_count"])
    sim_freq = int(config["simulation_frequency"])
    st0 = SimState.zeros(E, backend.get_state().vcap)
    backend.set_state(st0)
    draws = (abi.SpawnDraw * E)()

    def attempt(longitudinal, position_deviation, speed_deviation, p, go_straight):
        saved = [draw_spawn(rngs[e], p, draws[e]) for e in range(E)]
        accepted = backend.spawn(draws, float(longitudinal), position_deviation, speed_deviation, p, go_straight)
        for e in range(E):
            if saved[e] is not None and not accepted[e]:
                rngs[e].bit_generator.state = saved[e]

    for t in range(n_vehicles - 1):
        attempt(np.linspace(0, 80, n_vehicles)[t], 1.0, 1.0, 0.6, False)
    for _ in range(3 * sim_freq):
        backend.substep_none()
    attempt(60, 0.1, 0.0, 1.0, True)  # challenger vehicle

    # ego: MDPVehicle on (o0, ir0, 0) at s = 60 + 5 N(1, 1), speed = speed_limit, route to the destination
    st = backend.get_state()
    ego_key = ("o0", "ir0", 0)
    ego_lane = net.get_lane(ego_key)
    lane_flat = table.flat(ego_key)
    ts = np.array([cfg.target_speeds[k] for k in range(cfg.n_target_speeds)])
    for e in range(E):
        destination = config["destination"] or "o" + str(rngs[e].integers(1, 4))
        pos = ego_lane.position(60 + 5 * rngs[e].normal(1), 0)
        speed = float(ego_lane.speed_limit)
        x = (speed - ts[0]) / (ts[-1] - ts[0])
        sidx = int(np.clip(np.round(x * (ts.size - 1)), 0, ts.size - 1))  # MDPVehicle.speed_to_index
        route = [(table.road_index_of[(f, t_)], i) for f, t_, i in net.plan_route(ego_key, destination)]
        n = int(st.env_i[abi.EI_NVEH, e])
        if n >= st.vcap:
            raise RuntimeError("vehicle capacity exceeded during reset")
        st.set_vehicle(e, n, x=float(pos[0]), y=float(pos[1]), heading=float(ego_lane.heading_at(60)), speed=speed,
                       lane=lane_flat, target_speed=float(ts[sidx]), speed_index=sidx, mdp=True, controlled=True, route=route)
        # "prevent early collisions": list.remove() while iterating skips the element after each removal (:313-318)
        order: List[int] = list(range(n + 1))
        ego_slot = n
        i = 0
        while i < len(order):
            s = order[i]
            if s != ego_slot:
                d = np.linalg.norm(np.array([st.veh_d[abi.D_X, e, s] - pos[0], st.veh_d[abi.D_Y, e, s] - pos[1]]))
                if d < 20:
                    order.pop(i)
            i += 1
        _compact(st, e, order)
        st.env_i[abi.EI_NVEH, e] = len(order)
        st.env_i[abi.EI_EGO, e] = order.index(ego_slot)
        st.env_i[abi.EI_STEPS, e] = 0
        st.env_d[abi.ED_TIME, e] = 0.0
        st.env_d[abi.ED_RETURN, e] = 0.0
        st.env_i[abi.EI_DONE, e] = 0
    backend.set_state(st)
    return st


def _compact(st: SimState, e: int, order: List[int]) -> None:
    vd = st.veh_d[:, e, :].copy()
    vi = st.veh_i[:, e, :].copy()
    st.veh_d[:, e, :] = 0
    st.veh_i[:, e, :] = 0
    for dst, src in enumerate(order):
        st.veh_d[:, e, dst] = vd[:, src]
        st.veh_i[:, e, dst] = vi[:, src]
