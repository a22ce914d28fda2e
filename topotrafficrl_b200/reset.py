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


def draw_behavior(rng: np.random.Generator, linear: bool):
    """``randomize_behavior()``: IDMVehicle draws DELTA ~ U[3.5, 4.5] (behavior.py:66-69); LinearVehicle draws uniform(size=3)
    and uniform(size=2) for its acceleration / steering parameters (behavior.py:402-410).  Returns (delta, five uniforms | None)."""
    if not linear:
        return float(rng.uniform(low=3.5, high=4.5)), None
    return 4.0, np.concatenate([rng.uniform(size=3), rng.uniform(size=2)])


def linear_parameters(u) -> np.ndarray:
    """ACCELERATION_RANGE[0] + ua (ACCELERATION_RANGE[1] - ACCELERATION_RANGE[0]), same for the steering pair."""
    lo, hi = np.array(abi.LINEAR_RANGE_LO), np.array(abi.LINEAR_RANGE_HI)
    return lo + np.asarray(u) * (hi - lo)


def draw_spawn(rng: np.random.Generator, spawn_probability: float, rec: abi.SpawnDraw, linear: bool = False):
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
    rec.delta, u = draw_behavior(rng, linear)
    for k in range(5):
        rec.lin_u[k] = 0.0 if u is None else float(u[k])
    return saved


def reset_intersection(backend: ResetBackend, rngs: Sequence[np.random.Generator], net: RoadNetwork, table: NetworkTable,
                       config: dict, cfg: abi.Config) -> SimState:
    E = backend.num_envs
    n_controlled = int(config.get("controlled_vehicles", 1))
    if not 1 <= n_controlled <= abi.MAX_CONTROLLED:
        raise NotImplementedError(f"controlled_vehicles must be in 1..{abi.MAX_CONTROLLED}")
    n_vehicles = int(config["initial_vehicle_count"])
    sim_freq = int(config["simulation_frequency"])
    linear = cfg.vehicle_model == abi.VEHICLE_LINEAR
    st0 = SimState.zeros(E, backend.get_state().vcap, linear=linear)
    backend.set_state(st0)
    draws = (abi.SpawnDraw * E)()

    def attempt(longitudinal, position_deviation, speed_deviation, p, go_straight):
        saved = [draw_spawn(rngs[e], p, draws[e], linear) for e in range(E)]
        accepted = backend.spawn(draws, float(longitudinal), position_deviation, speed_deviation, p, go_straight)
        for e in range(E):
            if saved[e] is not None and not accepted[e]:
                rngs[e].bit_generator.state = saved[e]

    for t in range(n_vehicles - 1):
        attempt(np.linspace(0, 80, n_vehicles)[t], 1.0, 1.0, 0.6, False)
    for _ in range(3 * sim_freq):
        backend.substep_none()
    attempt(60, 0.1, 0.0, 1.0, True)  # challenger vehicle

    # controlled vehicles: MDPVehicle k on (o<k%4>, ir<k%4>, 0) at s = 60 + 5 N(1, 1), speed = speed_limit, route to the destination
    st = backend.get_state()
    ts = np.array([cfg.target_speeds[k] for k in range(cfg.n_target_speeds)])
    for e in range(E):
        for ego_id in range(n_controlled):
            ego_key = (f"o{ego_id % 4}", f"ir{ego_id % 4}", 0)
            ego_lane = net.get_lane(ego_key)
            destination = config["destination"] or "o" + str(rngs[e].integers(1, 4))
            pos = ego_lane.position(60 + 5 * rngs[e].normal(1), 0)
            heading = float(ego_lane.heading_at(60))
            speed = float(ego_lane.speed_limit)
            x = (speed - ts[0]) / (ts[-1] - ts[0])
            sidx = int(np.clip(np.round(x * (ts.size - 1)), 0, ts.size - 1))  # MDPVehicle.speed_to_index
            lane_key = net.get_closest_lane_index(pos, heading)                # RoadObject.__init__ objects.py:45-50
            route = [(table.road_index_of[(f, t_)], i) for f, t_, i in net.plan_route(lane_key, destination)]
            n = int(st.env_i[abi.EI_NVEH, e])
            if n >= st.vcap:
                raise RuntimeError("vehicle capacity exceeded during reset")
            st.set_vehicle(e, n, x=float(pos[0]), y=float(pos[1]), heading=heading, speed=speed,
                           lane=table.flat(lane_key), target_speed=float(ts[sidx]), speed_index=sidx, mdp=True, controlled=True,
                           route=route, agent=ego_id)
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
        flags = st.veh_i[abi.I_FLAGS, e, :int(st.env_i[abi.EI_NVEH, e])]
        first = np.nonzero((flags & abi.FL_CONTROLLED) != 0)[0]
        first = [int(k) for k in first if ((int(flags[k]) & abi.FL_AGENT_MASK) >> abi.FL_AGENT_SHIFT) == 0]
        st.env_i[abi.EI_EGO, e] = first[0]
        st.env_i[abi.EI_STEPS, e] = 0
        st.env_d[abi.ED_TIME, e] = 0.0
        st.env_d[abi.ED_RETURN, e] = 0.0
        st.env_i[abi.EI_DONE, e] = 0
    backend.set_state(st)
    return st


# --------------------------------------------------------------------------------------------------
# scripted scenes: RoundaboutEnv / UTurnEnv place a fixed cast of vehicles with a few random draws
# --------------------------------------------------------------------------------------------------
class _Cast:
    """Builds one env's vehicle list the way the reference constructors do (``RoadObject.make_on_lane`` objects.py:67-89,
    ``RoadObject.__init__`` :45-50 closest-lane lookup, ``ControlledVehicle.__init__`` controller.py:35-48,
    ``MDPVehicle.__init__`` :262-293, ``IDMVehicle.__init__`` behavior.py:48-64, ``plan_route_to`` controller.py:71-87)."""

    def __init__(self, st: SimState, e: int, net: RoadNetwork, table: NetworkTable, cfg: abi.Config) -> None:
        self.st, self.e, self.net, self.table = st, e, net, table
        self.ts = np.array([cfg.target_speeds[k] for k in range(cfg.n_target_speeds)])
        self.n = 0

    def _route(self, lane_key, destination):
        if destination is None:
            return None
        cache = self.net.__dict__.setdefault("_planned_routes", {})  # the BFS result depends on (lane, destination) only
        key = (lane_key, destination)
        if key not in cache:
            cache[key] = [(self.table.road_index_of[(f, t)], i) for f, t, i in self.net.plan_route(lane_key, destination)]
        return cache[key]

    def ego(self, position, heading: float, speed: float, destination) -> None:
        lane_key = self.net.get_closest_lane_index(position, heading)
        ts = self.ts
        x = (speed - ts[0]) / (ts[-1] - ts[0])
        sidx = int(np.clip(np.round(x * (ts.size - 1)), 0, ts.size - 1))
        self.st.set_vehicle(self.e, self.n, x=float(position[0]), y=float(position[1]), heading=float(heading), speed=float(speed),
                            lane=self.table.flat(lane_key), target_speed=float(ts[sidx]), speed_index=sidx, mdp=True,
                            controlled=True, route=self._route(lane_key, destination))
        self.st.env_i[abi.EI_EGO, self.e] = self.n
        self.n += 1

    def idm_on_lane(self, lane_index, longitudinal: float, speed: float, destination, delta: float = 4.0, linear=None) -> None:
        """An ``other_vehicles_type`` vehicle; ``linear``: its LinearVehicle parameters (None with IDM traffic)."""
        lane = self.net.get_lane(lane_index)
        position, heading = lane.position(longitudinal, 0), float(lane.heading_at(longitudinal))
        lane_key = self.net.get_closest_lane_index(position, heading)
        self.st.set_vehicle(self.e, self.n, x=float(position[0]), y=float(position[1]), heading=heading, speed=float(speed),
                            lane=self.table.flat(lane_key), timer=float((np.sum(position) * np.pi) % 1.0), delta=float(delta),
                            route=self._route(lane_key, destination), linear=linear)
        self.n += 1

    def done(self) -> None:
        self.st.env_i[abi.EI_NVEH, self.e] = self.n


def reset_roundabout(rngs: Sequence[np.random.Generator], net: RoadNetwork, table: NetworkTable, config: dict, cfg: abi.Config,
                     vcap: int) -> SimState:
    """``RoundaboutEnv._make_vehicles`` (roundabout_env.py:326-387) for one Generator per env, draws in the reference's order."""
    linear = cfg.vehicle_model == abi.VEHICLE_LINEAR
    st = SimState.zeros(len(rngs), vcap, linear=linear)
    position_deviation = speed_deviation = 2
    destinations = ["exr", "sxr", "nxr"]

    def behavior(rng):  # vehicle.randomize_behavior()
        delta, u = draw_behavior(rng, linear)
        return dict(delta=delta, linear=None if u is None else linear_parameters(u))

    for e, rng in enumerate(rngs):
        cast = _Cast(st, e, net, table, cfg)
        ego_lane = net.get_lane(("ser", "ses", 0))
        cast.ego(ego_lane.position(125, 0), ego_lane.heading_at(140), 8, "nxs")
        # incoming vehicle
        lon = 5 + rng.normal() * position_deviation
        speed = 16 + rng.normal() * speed_deviation
        if config["incoming_vehicle_destination"] is not None:
            destination = destinations[config["incoming_vehicle_destination"]]
        else:
            destination = str(rng.choice(destinations))
        cast.idm_on_lane(("we", "sx", 1), lon, speed, destination, **behavior(rng))
        # other vehicles
        for i in list(range(1, 2)) + list(range(-1, 0)):
            lon = 20 * i + rng.normal() * position_deviation
            speed = 16 + rng.normal() * speed_deviation
            destination = str(rng.choice(destinations))
            cast.idm_on_lane(("we", "sx", 0), lon, speed, destination, **behavior(rng))
        # entering vehicle
        lon = 50 + rng.normal() * position_deviation
        speed = 16 + rng.normal() * speed_deviation
        destination = str(rng.choice(destinations))
        cast.idm_on_lane(("eer", "ees", 0), lon, speed, destination, **behavior(rng))
        cast.done()
    return st


def reset_uturn(rngs: Sequence[np.random.Generator], net: RoadNetwork, table: NetworkTable, config: dict, cfg: abi.Config,
                vcap: int) -> SimState:
    """``UTurnEnv._make_vehicles`` (u_turn_env.py:173-271)."""
    linear = cfg.vehicle_model == abi.VEHICLE_LINEAR
    st = SimState.zeros(len(rngs), vcap, linear=linear)
    position_deviation = speed_deviation = 2
    for e, rng in enumerate(rngs):
        cast = _Cast(st, e, net, table, cfg)
        ego_lane = net.get_lane(("a", "b", 0))
        cast.ego(ego_lane.position(0, 0), 0.0, 16, "d")
        script = [(("a", "b", 0), 25, 13.5, True), (("a", "b", 1), 56, 14.5, False), (("b", "c", 1), 0.5, 4.5, False),
                  (("b", "c", 0), 17.5, 5.5, False), (("c", "d", 0), 1, 3.5, False), (("c", "d", 1), 30, 5.5, False)]
        for lane_index, lon0, speed0, randomize in script:
            lon = lon0 + rng.normal() * position_deviation
            speed = speed0 + rng.normal() * speed_deviation
            delta, u = draw_behavior(rng, linear) if randomize else (4.0, None)  # randomize_behavior (behavior.py:66-69, :402-410)
            params = None if not linear else (linear_parameters(u) if u is not None else np.array(abi.LINEAR_DEFAULTS))
            cast.idm_on_lane(lane_index, lon, speed, "d", delta=delta, linear=params)
        cast.done()
    return st


def _compact(st: SimState, e: int, order: List[int]) -> None:
    arrays = [st.veh_d, st.veh_i] + ([st.lin] if st.lin is not None else [])
    for a in arrays:
        old = a[:, e, :].copy()
        a[:, e, :] = 0
        for dst, src in enumerate(order):
            a[:, e, dst] = old[:, src]
