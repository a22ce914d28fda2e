"""Scene builders: road networks, device configs and initial traffic for the hot-path configurations.

* intersection scene = ``IntersectionEnv._make_road`` (reference intersection_env.py:141-249): 4 corners x
  (incoming straight 100 m, right-turn arc R=9, left-turn arc R=13, crossing straight, exit straight 100 m).
* highway scene = ``RoadNetwork.straight_road_network`` (road.py:291-321) populated like
  ``Vehicle.create_random`` (kinematics.py:50-104): the reference ships no HighwayEnv, so BASELINE
  configs 2/3 are synthetic scenes composed from those reference primitives.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi
from .road import AbstractLane, CircularLane, LineType, NetworkTable, RoadNetwork, SineLane, StraightLane
from .state import SimState

# --------------------------------------------------------------------------------------------------
# default configs (same keys as the reference: abstract.py:94-109, intersection_env.py:20-58)
# --------------------------------------------------------------------------------------------------
BASE_CONFIG = {
    "observation": {"type": "Kinematics"},
    "action": {"type": "DiscreteMetaAction"},
    "simulation_frequency": 15,
    "policy_frequency": 1,
    "other_vehicles_type": "ttrl_env.vehicle.behavior.IDMVehicle",
    "screen_width": 600, "screen_height": 150, "centering_position": [0.3, 0.5], "scaling": 5.5,
    "show_trajectories": False, "render_agent": True, "offscreen_rendering": False,
    "manual_control": False, "real_time_rendering": False,
}

INTERSECTION_CONFIG = dict(BASE_CONFIG, **{
    "observation": {
        "type": "Kinematics", "vehicles_count": 15,
        "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
        "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]},
        "absolute": True, "flatten": False, "observe_intentions": False,
    },
    "action": {"type": "DiscreteMetaAction", "longitudinal": True, "lateral": False, "target_speeds": [0, 4.5, 9]},
    "duration": 13, "destination": "o1", "controlled_vehicles": 1, "initial_vehicle_count": 10,
    "spawn_probability": 0.6, "screen_width": 600, "screen_height": 600, "centering_position": [0.5, 0.6],
    "scaling": 5.5 * 1.3, "collision_reward": -5, "high_speed_reward": 1, "arrived_reward": 1,
    "reward_speed_range": [7.0, 9.0], "normalize_reward": False, "offroad_terminal": False,
})

# Synthetic highway (BASELINE configs 2 and 3).  Reward template: u_turn_env.py:20-71.
HIGHWAY_CONFIG = dict(BASE_CONFIG, **{
    "observation": {
        "type": "Kinematics", "vehicles_count": 15,
        "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
        "absolute": False, "order": "sorted",
    },
    "action": {"type": "DiscreteMetaAction"},
    "lanes_count": 4, "vehicles_count": 50, "vehicles_density": 2.0, "ego_spacing": 2.0,
    "road_length": 10000, "speed_limit": 30,
    "duration": 40, "collision_reward": -1.0, "left_lane_reward": 0.1, "high_speed_reward": 0.4,
    "reward_speed_range": [20, 30], "normalize_reward": True, "offroad_terminal": False,
})

# MultiAgentIntersectionEnv.default_config (intersection_env.py:372-394)
MULTI_AGENT_INTERSECTION_CONFIG = dict(INTERSECTION_CONFIG, **{
    "action": {"type": "MultiAgentAction",
               "action_config": {"type": "DiscreteMetaAction", "lateral": False, "longitudinal": True}},
    "observation": {"type": "MultiAgentObservation", "observation_config": {"type": "Kinematics"}},
    "controlled_vehicles": 2,
})

# RoundaboutEnv.default_config (roundabout_env.py:13-41)
ROUNDABOUT_CONFIG = dict(BASE_CONFIG, **{
    "observation": {"type": "Kinematics", "absolute": True,
                    "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-15, 15], "vy": [-15, 15]}},
    "action": {"type": "DiscreteMetaAction", "target_speeds": [0, 8, 16]},
    "incoming_vehicle_destination": None, "collision_reward": -1, "high_speed_reward": 0.2, "right_lane_reward": 0,
    "lane_change_reward": -0.05, "screen_width": 600, "screen_height": 600, "centering_position": [0.5, 0.6],
    "duration": 11, "normalize_reward": True,
})

# UTurnEnv.default_config (u_turn_env.py:18-37)
UTURN_CONFIG = dict(BASE_CONFIG, **{
    "observation": {"type": "TimeToCollision", "horizon": 16},
    "action": {"type": "DiscreteMetaAction", "target_speeds": [8, 16, 24]},
    "screen_width": 789, "screen_height": 289, "duration": 10, "collision_reward": -1.0, "left_lane_reward": 0.1,
    "high_speed_reward": 0.4, "reward_speed_range": [8, 24], "normalize_reward": True, "offroad_terminal": False,
})

# IDM/MOBIL class constants: behavior.py:20-46; the intersection scene overrides three of them
# process-wide (intersection_env.py:258-261) -- here they are per-scene constants.
IDM_DEFAULT = dict(acc_max=6.0, comfort_acc_max=3.0, comfort_acc_min=-5.0, distance_wanted=5.0 + 5.0,
                   time_wanted=1.5, politeness=0.0, lane_change_min_acc_gain=0.2,
                   lane_change_max_braking_imposed=2.0, lane_change_delay=1.0)
IDM_INTERSECTION = dict(IDM_DEFAULT, distance_wanted=7.0, comfort_acc_max=6.0, comfort_acc_min=-3.0)


# --------------------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------------------
def make_intersection_network() -> RoadNetwork:
    """4-way intersection.  Node names: (o|ir|il)+corner, corner 0 south, 1 west, 2 north, 3 east."""
    w = AbstractLane.DEFAULT_WIDTH
    r_right = w + 5
    r_left = r_right + w
    outer = r_right + w / 2
    access = 50 + 50
    n, c, s = LineType.NONE, LineType.CONTINUOUS, LineType.STRIPED
    net = RoadNetwork()
    for k in range(4):
        angle = np.radians(90 * k)
        horizontal = k % 2
        prio = 3 if horizontal else 1
        rot = np.array([[np.cos(angle), -np.sin(angle)], [np.sin(angle), np.cos(angle)]])
        o, ir = f"o{k}", f"ir{k}"
        # incoming
        net.add_lane(o, ir, StraightLane(rot @ np.array([w / 2, access + outer]), rot @ np.array([w / 2, outer]),
                                         line_types=[s, c], priority=prio, speed_limit=10))
        # right turn
        net.add_lane(ir, f"il{(k - 1) % 4}",
                     CircularLane(rot @ np.array([outer, outer]), r_right, angle + np.radians(180),
                                  angle + np.radians(270), line_types=[n, c], priority=prio, speed_limit=10))
        # left turn
        net.add_lane(ir, f"il{(k + 1) % 4}",
                     CircularLane(rot @ np.array([-r_left + w / 2, r_left - w / 2]), r_left, angle + np.radians(0),
                                  angle + np.radians(-90), clockwise=False, line_types=[n, n],
                                  priority=prio - 1, speed_limit=10))
        # straight across
        net.add_lane(ir, f"il{(k + 2) % 4}",
                     StraightLane(rot @ np.array([w / 2, outer]), rot @ np.array([w / 2, -outer]),
                                  line_types=[s, n], priority=prio, speed_limit=10))
        # exit
        ex_start = rot @ np.flip([w / 2, access + outer], axis=0)
        ex_end = rot @ np.flip([w / 2, outer], axis=0)
        net.add_lane(f"il{(k - 1) % 4}", f"o{(k - 1) % 4}",
                     StraightLane(ex_end, ex_start, line_types=[n, c], priority=prio, speed_limit=10))
    return net


def make_roundabout_network() -> RoadNetwork:
    """``RoundaboutEnv._make_road`` (roundabout_env.py:76-324): two-lane ring of 8 arcs (radii 20 / 24 m) and four
    access roads, each a straight lane joined to the ring by a sine lane.  Nodes: (s)outh/(e)ast/(n)orth/(w)est,
    (e)ntry/e(x)it, (r)oad/(s)ine."""
    center, radius, alpha = [0, 0], 20, 24
    net = RoadNetwork()
    radii = [radius, radius + 4]
    n, c, s = LineType.NONE, LineType.CONTINUOUS, LineType.STRIPED
    line = [[c, s], [n, c]]
    ring = [("se", "ex", 90 - alpha, alpha), ("ex", "ee", alpha, -alpha), ("ee", "nx", -alpha, -90 + alpha),
            ("nx", "ne", -90 + alpha, -90 - alpha), ("ne", "wx", -90 - alpha, -180 + alpha),
            ("wx", "we", -180 + alpha, -180 - alpha), ("we", "sx", 180 - alpha, 90 + alpha), ("sx", "se", 90 + alpha, 90 - alpha)]
    for lane in [0, 1]:
        for _from, _to, a0, a1 in ring:
            net.add_lane(_from, _to, CircularLane(center, radii[lane], np.deg2rad(a0), np.deg2rad(a1), clockwise=False,
                                                  line_types=line[lane]))
    access, dev, a = 170, 85, 5
    delta_st = 0.2 * dev
    delta_en = dev - delta_st
    w = 2 * np.pi / dev
    net.add_lane("ser", "ses", StraightLane([2, access], [2, dev / 2], line_types=(s, c)))
    net.add_lane("ses", "se", SineLane([2 + a, dev / 2], [2 + a, dev / 2 - delta_st], a, w, -np.pi / 2, line_types=(c, c)))
    net.add_lane("sx", "sxs", SineLane([-2 - a, -dev / 2 + delta_en], [-2 - a, dev / 2], a, w, -np.pi / 2 + w * delta_en, line_types=(c, c)))
    net.add_lane("sxs", "sxr", StraightLane([-2, dev / 2], [-2, access], line_types=(n, c)))
    net.add_lane("eer", "ees", StraightLane([access, -2], [dev / 2, -2], line_types=(s, c)))
    net.add_lane("ees", "ee", SineLane([dev / 2, -2 - a], [dev / 2 - delta_st, -2 - a], a, w, -np.pi / 2, line_types=(c, c)))
    net.add_lane("ex", "exs", SineLane([-dev / 2 + delta_en, 2 + a], [dev / 2, 2 + a], a, w, -np.pi / 2 + w * delta_en, line_types=(c, c)))
    net.add_lane("exs", "exr", StraightLane([dev / 2, 2], [access, 2], line_types=(n, c)))
    net.add_lane("ner", "nes", StraightLane([-2, -access], [-2, -dev / 2], line_types=(s, c)))
    net.add_lane("nes", "ne", SineLane([-2 - a, -dev / 2], [-2 - a, -dev / 2 + delta_st], a, w, -np.pi / 2, line_types=(c, c)))
    net.add_lane("nx", "nxs", SineLane([2 + a, dev / 2 - delta_en], [2 + a, -dev / 2], a, w, -np.pi / 2 + w * delta_en, line_types=(c, c)))
    net.add_lane("nxs", "nxr", StraightLane([2, -dev / 2], [2, -access], line_types=(n, c)))
    net.add_lane("wer", "wes", StraightLane([-access, 2], [-dev / 2, 2], line_types=(s, c)))
    net.add_lane("wes", "we", SineLane([-dev / 2, 2 + a], [-dev / 2 + delta_st, 2 + a], a, w, -np.pi / 2, line_types=(c, c)))
    net.add_lane("wx", "wxs", SineLane([dev / 2 - delta_en, -2 - a], [-dev / 2, -2 - a], a, w, -np.pi / 2 + w * delta_en, line_types=(c, c)))
    net.add_lane("wxs", "wxr", StraightLane([-dev / 2, -2], [-access, -2], line_types=(n, c)))
    return net


def make_uturn_network(length: float = 128) -> RoadNetwork:
    """``UTurnEnv._make_road`` (u_turn_env.py:83-171): two lanes a->b, a counter-clockwise half circle b->c, two lanes c->d."""
    net = RoadNetwork()
    w = StraightLane.DEFAULT_WIDTH
    net.add_lane("c", "d", StraightLane([length, w], [0, w], line_types=(LineType.CONTINUOUS_LINE, LineType.STRIPED)))
    net.add_lane("c", "d", StraightLane([length, 0], [0, 0], line_types=(LineType.NONE, LineType.CONTINUOUS_LINE)))
    center = [length, w + 20]
    radius, alpha = 20, 0
    radii = [radius, radius + w]
    n, c, s = LineType.NONE, LineType.CONTINUOUS, LineType.STRIPED
    line = [[c, s], [n, c]]
    for lane in [0, 1]:
        net.add_lane("b", "c", CircularLane(center, radii[lane], np.deg2rad(90 - alpha), np.deg2rad(-90 + alpha),
                                            clockwise=False, line_types=line[lane]))
    offset = 2 * radius
    net.add_lane("a", "b", StraightLane([0, ((2 * w + offset) - w)], [length, ((2 * w + offset) - w)],
                                        line_types=(LineType.CONTINUOUS_LINE, LineType.STRIPED)))
    net.add_lane("a", "b", StraightLane([0, (2 * w + offset)], [length, (2 * w + offset)],
                                        line_types=(LineType.NONE, LineType.CONTINUOUS_LINE)))
    return net


def intersection_exit_predicate(_from: str, _to: str) -> bool:
    return "il" in _from and "o" in _to


def make_highway_network(lanes: int = 4, length: float = 10000, speed_limit: float = 30) -> RoadNetwork:
    return RoadNetwork.straight_road_network(lanes=lanes, length=length, speed_limit=speed_limit)


def intersection_spawn_routes(net: RoadNetwork, table: NetworkTable):
    """Route table for ``_spawn_vehicle`` (intersection_env.py:331-345): entry lane per corner and, per
    (entry, exit) pair, the road indices that follow the entry lane in ``plan_route_to``'s BFS path."""
    spawn_lane = np.zeros(4, np.int32)
    rlen = np.zeros((4, 4), np.int32)
    rroad = np.zeros((4, 4, abi.ROUTE_CAP), np.int32)
    for a in range(4):
        idx = (f"o{a}", f"ir{a}", 0)
        spawn_lane[a] = table.flat(idx)
        for b in range(4):
            # a == b (a controlled vehicle of the multi-agent env whose destination is its own arm): 5 roads
            route = net.plan_route(idx, f"o{b}")[1:]
            if len(route) + 1 > abi.ROUTE_CAP:
                raise ValueError("planned route exceeds the device route capacity")
            rlen[a, b] = len(route)
            for k, (f, t, _) in enumerate(route):
                rroad[a, b, k] = table.road_index_of[(f, t)]
    return spawn_lane, rlen, rroad


# --------------------------------------------------------------------------------------------------
# dict config -> device config
# --------------------------------------------------------------------------------------------------
def _resolve_observation(obs_cfg: dict) -> dict:
    if obs_cfg["type"] == "MultiAgentObservation":
        return obs_cfg["observation_config"]
    return obs_cfg


def build_config(table: NetworkTable, config: dict, scene: str, ego_lanes_count: int = 1) -> abi.Config:
    """Translate a reference-style env config dict into the POD the device uses.

    Unknown observation / action types raise ``ValueError`` like the reference factories
    (observation.py:793, action.py:344)."""
    cfg = abi.Config()
    cfg.n_lanes, cfg.n_roads, cfg.n_nodes = table.n_lanes, table.n_roads, table.n_nodes
    cfg.simulation_frequency = float(config["simulation_frequency"])
    cfg.policy_frequency = float(config["policy_frequency"])
    cfg.duration = float(config["duration"])
    if scene not in ("intersection", "highway", "roundabout", "u-turn"):
        raise ValueError(f"unknown scene {scene!r}")
    idm = IDM_INTERSECTION if scene == "intersection" else IDM_DEFAULT
    for k, v in idm.items():
        setattr(cfg, k, float(v))
    cfg.regulated = 1 if scene == "intersection" else 0
    # the class of the surrounding traffic (intersection_env.py:257, roundabout_env.py:337, u_turn_env.py:203)
    ovt = str(config.get("other_vehicles_type", "ttrl_env.vehicle.behavior.IDMVehicle")).rsplit(".", 1)[-1]
    if ovt == "LinearVehicle" and scene != "highway":
        # behavior.py:350-558: controllers linear in per-vehicle parameters, TIME_WANTED 2.5 (class attribute :373)
        cfg.vehicle_model = abi.VEHICLE_LINEAR
        cfg.time_wanted = 2.5
        for k in range(abi.NLIN):
            cfg.lin_lo[k], cfg.lin_hi[k], cfg.lin_default[k] = abi.LINEAR_RANGE_LO[k], abi.LINEAR_RANGE_HI[k], abi.LINEAR_DEFAULTS[k]
    elif ovt != "IDMVehicle":
        raise NotImplementedError(f"other_vehicles_type {ovt!r}: IDMVehicle and LinearVehicle traffic run on the device "
                                  "(AggressiveVehicle / DefensiveVehicle / MultipleModelVehicle would silently change the dynamics)")
    cfg.controlled_vehicles = int(config.get("controlled_vehicles", 1))
    if not 1 <= cfg.controlled_vehicles <= abi.MAX_CONTROLLED:
        # ego k starts on arm k % 4 (intersection_env.py:287-289): a fifth ego would be placed onto the first one
        raise NotImplementedError(f"controlled_vehicles must be in 1..{abi.MAX_CONTROLLED}")
    if cfg.controlled_vehicles > 1 and scene != "intersection":
        raise NotImplementedError("several controlled vehicles exist in the intersection scene only (intersection_env.py:372)")

    # ---- action -------------------------------------------------------------------------------
    act = config["action"]
    if act["type"] == "MultiAgentAction":
        act = act["action_config"]
    if act["type"] != "DiscreteMetaAction":
        if act["type"] in ("ContinuousAction", "DiscreteAction"):
            raise NotImplementedError(f"action type {act['type']} is outside the B200 hot path (SURVEY.md section 2 row 14)")
        raise ValueError("Unknown action type")
    longitudinal, lateral = act.get("longitudinal", True), act.get("lateral", True)
    if longitudinal and lateral:
        cfg.action_mode = abi.ACT_ALL
    elif longitudinal:
        cfg.action_mode = abi.ACT_LONGI
    elif lateral:
        cfg.action_mode = abi.ACT_LAT
    else:
        raise ValueError("At least longitudinal or lateral actions must be included")
    ts = act.get("target_speeds")
    ts = np.linspace(20, 30, 3) if ts is None else np.asarray(ts, dtype=np.float64)
    if len(ts) > abi.MAX_TARGET_SPEEDS:
        raise ValueError("too many target speeds")
    cfg.n_target_speeds = len(ts)
    for k, v in enumerate(ts):
        cfg.target_speeds[k] = float(v)

    # ---- observation --------------------------------------------------------------------------
    obs = _resolve_observation(config["observation"])
    otype = obs["type"]
    if otype == "Kinematics":
        cfg.obs_type = abi.OBS_KINEMATICS
        features = obs.get("features") or ["presence", "x", "y", "vx", "vy"]
        cfg.obs_vehicles = int(obs.get("vehicles_count", 5))
        cfg.absolute = int(bool(obs.get("absolute", False)))
        order = obs.get("order", "sorted")
        cfg.order = abi.ORDER_SHUFFLED if order == "shuffled" else abi.ORDER_SORTED
        cfg.see_behind = int(bool(obs.get("see_behind", False)))
        cfg.normalize = int(bool(obs.get("normalize", True)))
        cfg.clip = int(bool(obs.get("clip", True)))
        frange = obs.get("features_range")
        if not frange:  # lazily derived from the ego's road in the reference (observation.py:213-225)
            frange = {"x": [-5.0 * 40.0, 5.0 * 40.0],
                      "y": [-AbstractLane.DEFAULT_WIDTH * ego_lanes_count, AbstractLane.DEFAULT_WIDTH * ego_lanes_count],
                      "vx": [-2 * 40.0, 2 * 40.0], "vy": [-2 * 40.0, 2 * 40.0]}
    elif otype == "OccupancyGrid":
        cfg.obs_type = abi.OBS_GRID
        features = obs.get("features")
        if features is None:
            features = ["presence", "vx", "vy", "on_road"]
        if obs.get("absolute", False):
            raise NotImplementedError()  # like the reference (observation.py:357-358)
        grid_size = np.array(obs.get("grid_size") or [[-5.5 * 5, 5.5 * 5], [-5.5 * 5, 5.5 * 5]], dtype=np.float64)
        grid_step = np.array(obs.get("grid_step") or [5, 5], dtype=np.float64)
        shape = np.asarray(np.floor((grid_size[:, 1] - grid_size[:, 0]) / grid_step), dtype=np.uint8)
        cfg.grid_w, cfg.grid_h = int(shape[0]), int(shape[1])
        for k in range(2):
            cfg.grid_min[k], cfg.grid_max[k], cfg.grid_step[k] = grid_size[k, 0], grid_size[k, 1], grid_step[k]
        cfg.align_to_vehicle_axes = int(bool(obs.get("align_to_vehicle_axes", False)))
        cfg.as_image = int(bool(obs.get("as_image", False)))
        if cfg.as_image:
            raise NotImplementedError("as_image grids are outside the B200 hot path")
        cfg.clip = int(bool(obs.get("clip", True)))
        cfg.obs_vehicles = 0
        frange = obs.get("features_range")
        if not frange:
            frange = {"vx": [-2 * 40.0, 2 * 40.0], "vy": [-2 * 40.0, 2 * 40.0]}
        if "x" in frange:
            cfg.grid_has_xrange = 1
            cfg.grid_xrange[0], cfg.grid_xrange[1] = float(frange["x"][0]), float(frange["x"][1])
        if "y" in frange:
            cfg.grid_has_yrange = 1
            cfg.grid_yrange[0], cfg.grid_yrange[1] = float(frange["y"][0]), float(frange["y"][1])
    elif otype == "TimeToCollision":  # observation.py:114-151; horizon in seconds, one time cell per policy step
        cfg.obs_type = abi.OBS_TTC
        cfg.ttc_steps = int(obs.get("horizon", 10) * config["policy_frequency"])
        max_lanes = max(int(table.roads[r].n_lanes) for r in range(table.n_roads))
        if cfg.ttc_steps < 1 or cfg.n_target_speeds * max_lanes * cfg.ttc_steps > abi.MAX_TTC_CELLS:
            raise ValueError("TimeToCollision grid too large for the device scratch")
        cfg.obs_vehicles = 0
        features, frange = [], {}
    else:
        known = ("KinematicsGoal", "GrayscaleObservation", "AttributesObservation",
                 "MultiAgentObservation", "TupleObservation", "LidarObservation", "ExitObservation")
        if otype in known:
            raise NotImplementedError(f"observation type {otype} is outside the B200 hot path (SURVEY.md section 2 row 13)")
        raise ValueError("Unknown observation type")
    if len(features) > abi.MAX_FEATURES:
        raise ValueError("too many observation features")
    cfg.n_features = len(features)
    for k, name in enumerate(features):
        if name not in abi.FEATURES:
            raise NotImplementedError(f"observation feature {name!r} is outside the B200 hot path")
        cfg.features[k] = abi.FEATURES[name]
        if name in frange:
            cfg.has_range[k] = 1
            cfg.range_lo[k], cfg.range_hi[k] = float(frange[name][0]), float(frange[name][1])

    # ---- reward / termination -------------------------------------------------------------------
    cfg.speed_index_den = 1.0
    if scene == "intersection":
        cfg.reward_type = abi.REWARD_INTERSECTION
        cfg.arrived_reward = float(config.get("arrived_reward", 0))
        cfg.spawn_enabled = 1
        cfg.spawn_probability = float(config["spawn_probability"])
    elif scene == "roundabout":
        cfg.reward_type = abi.REWARD_ROUNDABOUT
        cfg.lane_change_reward = float(config.get("lane_change_reward", 0))
        cfg.speed_index_den = float(np.linspace(20, 30, 3).size - 1)  # MDPVehicle.DEFAULT_TARGET_SPEEDS.size - 1 (controller.py:259)
    else:
        cfg.reward_type = abi.REWARD_HIGHWAY
        cfg.lane_reward = float(config.get("left_lane_reward", 0))
    cfg.collision_reward = float(config.get("collision_reward", 0))
    cfg.high_speed_reward = float(config.get("high_speed_reward", 0))
    cfg.reward_speed_lo, cfg.reward_speed_hi = (float(x) for x in config.get("reward_speed_range", [0.0, 1.0]))
    cfg.normalize_reward = int(bool(config.get("normalize_reward", False)))
    cfg.offroad_terminal = int(bool(config.get("offroad_terminal", False)))
    return cfg


def obs_shape(cfg: abi.Config) -> tuple:
    """Shape of ONE controlled vehicle's observation (the ``space()`` of the reference's observation types)."""
    if cfg.obs_type == abi.OBS_GRID:
        return (cfg.n_features, cfg.grid_w, cfg.grid_h)
    if cfg.obs_type == abi.OBS_TTC:
        return (3, 3, cfg.ttc_steps)
    return (cfg.obs_vehicles, cfg.n_features)


def merged_config(default: dict, overrides: Optional[dict]) -> dict:
    """``AbstractEnv.configure`` semantics: a SHALLOW ``dict.update`` (abstract.py:111-113)."""
    cfg = copy.deepcopy(default)
    if overrides:
        cfg.update(copy.deepcopy(overrides))
    return cfg


# --------------------------------------------------------------------------------------------------
# highway traffic generator (vectorised over envs)
# --------------------------------------------------------------------------------------------------
def make_highway_state(num_envs: int, config: dict, seed: int = 0, first_env: int = 0, vcap: Optional[int] = None) -> SimState:
    """Initial traffic for the synthetic highway, following ``Vehicle.create_random``'s placement rule
    (kinematics.py:91-103): each new vehicle goes ``offset * U[0.9, 1.1]`` ahead of the furthest vehicle,
    ``offset = spacing * (12 + speed) * exp(-5/40 * lanes)``, on a uniformly random lane, at a speed
    ``U[0.7, 0.8] * speed_limit``; slot 0 is the MDPVehicle ego (speed 25), the rest are IDMVehicles with
    ``DELTA ~ U[3.5, 4.5]`` (behavior.py:66-69) and ``timer = (x + y) * pi mod 1`` (behavior.py:64).
    The stream is keyed by the GLOBAL env index so a shard of envs is identical to the same envs of a
    larger run."""
    lanes = int(config["lanes_count"])
    n = int(config["vehicles_count"])
    vcap = vcap or n
    st = SimState.zeros(num_envs, vcap)
    speed_limit = float(config["speed_limit"])
    density = float(config["vehicles_density"])
    width = AbstractLane.DEFAULT_WIDTH
    lane_factor = np.exp(-5 / 40 * lanes)
    for e in range(num_envs):
        rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, first_env + e])))
        lane_ids = rng.integers(0, lanes, size=n)
        speeds = rng.uniform(0.7 * speed_limit, 0.8 * speed_limit, size=n)
        jitter = rng.uniform(0.9, 1.1, size=n)
        deltas = rng.uniform(3.5, 4.5, size=n)
        speeds[0] = 25.0
        x_prev = 0.0
        for s in range(n):
            spacing = float(config["ego_spacing"]) if s == 0 else 1.0 / density
            offset = spacing * (12 + 1.0 * speeds[s]) * lane_factor
            x0 = 3 * offset if s == 0 else x_prev
            x0 += offset * jitter[s]
            x_prev = x0
            y0 = lane_ids[s] * width
            if s == 0:
                st.set_vehicle(e, s, x=x0, y=y0, heading=0.0, speed=25.0, lane=int(lane_ids[s]), target_speed=25.0,
                               speed_index=1, mdp=True, controlled=True)
            else:
                st.set_vehicle(e, s, x=x0, y=y0, heading=0.0, speed=speeds[s], lane=int(lane_ids[s]),
                               timer=float(((x0 + y0) * np.pi) % 1.0), delta=float(deltas[s]))
        st.env_i[abi.EI_NVEH, e] = n
        st.env_i[abi.EI_EGO, e] = 0
    return st


# --------------------------------------------------------------------------------------------------
# device-side reset parameters (include/ttrl_b200.h: ttrl_reset_params)
# --------------------------------------------------------------------------------------------------
def highway_reset_params(config: dict) -> abi.ResetParams:
    rp = abi.ResetParams()
    rp.scene = 0
    rp.n_vehicles = int(config["vehicles_count"])
    rp.lanes = int(config["lanes_count"])
    rp.speed_limit = float(config["speed_limit"])
    rp.density = float(config["vehicles_density"])
    rp.ego_spacing = float(config["ego_spacing"])
    rp.ego_speed = 25.0
    return rp


def intersection_reset_params(config: dict) -> abi.ResetParams:
    """``IntersectionEnv._make_vehicles`` constants (intersection_env.py:251-318)."""
    rp = abi.ResetParams()
    rp.scene = 1
    n = int(config["initial_vehicle_count"])
    if not 1 <= n <= abi.MAX_SPAWN_ATTEMPTS:
        raise ValueError("initial_vehicle_count out of range")
    rp.n_vehicles = n
    for t, v in enumerate(np.linspace(0, 80, n)):
        rp.spawn_longitudinal[t] = float(v)
    rp.ego_entry = 0
    dest = config.get("destination")
    rp.destination = -1 if dest is None else int(str(dest)[1:])
    rp.warmup_substeps = 3 * int(config["simulation_frequency"])
    rp.ego_longitudinal, rp.ego_longitudinal_std = 60.0, 5.0
    return rp


# (lane the member is made on, longitudinal, speed, randomize_behavior, destinations) -- roundabout_env.py:326-387
_ROUNDABOUT_DESTINATIONS = ["exr", "sxr", "nxr"]
_ROUNDABOUT_CAST = [(("we", "sx", 1), 5, 16, True), (("we", "sx", 0), 20 * 1, 16, True), (("we", "sx", 0), 20 * -1, 16, True),
                    (("eer", "ees", 0), 50, 16, True)]
# u_turn_env.py:173-271 (only vehicle 1 randomises its behaviour)
_UTURN_CAST = [(("a", "b", 0), 25, 13.5, True), (("a", "b", 1), 56, 14.5, False), (("b", "c", 1), 0.5, 4.5, False),
               (("b", "c", 0), 17.5, 5.5, False), (("c", "d", 0), 1, 3.5, False), (("c", "d", 1), 30, 5.5, False)]


def cast_reset_params(scene: str, net: RoadNetwork, table: NetworkTable, config: dict) -> abi.ResetParams:
    """Device-side reset of the scripted scenes (``ttrl_reset_params.scene == 2``): the cast of ``RoundaboutEnv._make_vehicles``
    / ``UTurnEnv._make_vehicles`` and the route table ``plan_route_to`` would produce from any road to each destination."""
    rp = abi.ResetParams()
    rp.scene = 2
    if scene == "roundabout":
        destinations = list(_ROUNDABOUT_DESTINATIONS) + ["nxs"]
        ego = (("ser", "ses", 0), 125.0, 8.0, 140.0, destinations.index("nxs"))
        fixed = config.get("incoming_vehicle_destination")
        members = []
        for k, (lane, lon, speed, rnd) in enumerate(_ROUNDABOUT_CAST):
            dest = [fixed] if (k == 0 and fixed is not None) else [0, 1, 2]
            members.append((lane, lon, 2.0, speed, 2.0, rnd, dest))
    elif scene == "u-turn":
        destinations = ["d"]
        ego = (("a", "b", 0), 0.0, 16.0, 0.0, 0)
        members = [(lane, lon, 2.0, speed, 2.0, rnd, [0]) for lane, lon, speed, rnd in _UTURN_CAST]
    else:
        raise ValueError(f"no scripted cast for scene {scene!r}")
    assert len(destinations) <= abi.CAST_DEST and 1 + len(members) <= abi.MAX_CAST
    m = rp.cast[0]
    m.lane, m.mdp, m.n_dest, m.randomize = table.flat(ego[0]), 1, 1, 0
    m.dest[0] = ego[4]
    m.longitudinal, m.speed, m.heading_longitudinal = ego[1], ego[2], ego[3]
    for k, (lane, lon, lon_std, speed, speed_std, rnd, dest) in enumerate(members, start=1):
        m = rp.cast[k]
        m.lane, m.mdp, m.n_dest, m.randomize = table.flat(lane), 0, len(dest), int(rnd)
        for j, d in enumerate(dest):
            m.dest[j] = int(d)
        m.longitudinal, m.longitudinal_std, m.speed, m.speed_std = float(lon), float(lon_std), float(speed), float(speed_std)
    rp.n_vehicles = 1 + len(members)
    for r, (_from, _to) in enumerate(table.road_keys):
        for d, name in enumerate(destinations):
            route = net.plan_route((_from, _to, 0), name)[1:]
            if len(route) + 1 > abi.ROUTE_CAP:
                raise ValueError("planned route exceeds the device route capacity")
            rp.cast_route_len[r][d] = len(route)
            for j, (f, t, _) in enumerate(route):
                rp.cast_route_road[r][d][j] = table.road_index_of[(f, t)]
    return rp
