"""ctypes mirror of ``include/ttrl_b200.h`` (POD structs, field indices, flag bits).

Keep in lock-step with the header; ``tests/test_abi.py`` checks the struct sizes against the
compiled library (``ttrl_abi_sizeof``).
"""
from __future__ import annotations

import ctypes as C

MAX_LANES = 64
MAX_ROADS = 64
MAX_NODES = 64
MAX_TARGET_SPEEDS = 8
MAX_FEATURES = 8
ROUTE_CAP = 12
ROUTE_WORDS = 3
MAX_CONTROLLED = 4

LANE_STRAIGHT, LANE_CIRCULAR, LANE_SINE = 0, 1, 2

# observation features (vehicle/kinematics.py:237-261 of the reference)
FEATURES = {"presence": 0, "x": 1, "y": 2, "vx": 3, "vy": 4, "cos_h": 5, "sin_h": 6, "heading": 7, "on_road": 8}
OBS_KINEMATICS, OBS_GRID, OBS_TTC = 0, 1, 2
MAX_TTC_CELLS = 2048
ORDER_SORTED, ORDER_SHUFFLED = 0, 1
ACT_ALL, ACT_LONGI, ACT_LAT = 0, 1, 2
REWARD_INTERSECTION, REWARD_HIGHWAY, REWARD_ROUNDABOUT = 0, 1, 2

# vehicle SoA fields
D_X, D_Y, D_HEADING, D_SPEED, D_STEERING, D_ACCEL, D_TARGET_SPEED, D_TIMER, D_DELTA, D_IMPACT_X, D_IMPACT_Y = range(11)
ND = 11
(I_LANE, I_TARGET_LANE, I_FLAGS, I_SPEED_INDEX, I_ROUTE_LEN, I_ROUTE_ROAD, I_ROUTE_LANE, I_YIELD_TIMER,
 I_ROUTE_ROAD1, I_ROUTE_ROAD2, I_ROUTE_LANE1, I_ROUTE_LANE2) = range(12)
NI = 12
I_ROUTE_ROAD_WORDS = (I_ROUTE_ROAD, I_ROUTE_ROAD1, I_ROUTE_ROAD2)  # route entry k: byte k % 4 of word k // 4
I_ROUTE_LANE_WORDS = (I_ROUTE_LANE, I_ROUTE_LANE1, I_ROUTE_LANE2)
FL_MDP, FL_CRASHED, FL_HAS_IMPACT, FL_YIELDING, FL_CONTROLLED = 1, 2, 4, 8, 16
FL_AGENT_SHIFT, FL_AGENT_MASK = 8, 0x700  # index of the vehicle in env.controlled_vehicles
EI_NVEH, EI_STEPS, EI_ROAD_STEPS, EI_EGO, EI_EPISODE, EI_DONE = range(6)
NEI = 6
ED_TIME, ED_RETURN = range(2)
NED = 2

QNET_MLP, QNET_EGO_ATTENTION, QNET_DUELING = 0, 1, 2
VEHICLE_IDM, VEHICLE_LINEAR = 0, 1
# LinearVehicle parameter block [NLIN][E][V] (behavior.py:353-367): ACCELERATION_PARAMETERS, STEERING_PARAMETERS
NLIN = 5
# class attributes of LinearVehicle: ACCELERATION_PARAMETERS + STEERING_PARAMETERS = [KP_HEADING, KP_HEADING * KP_LATERAL]
# (controller.py:24-32), ACCELERATION_RANGE = [0.5 p, 1.5 p], STEERING_RANGE = p -+ [0.07, 1.5]
LINEAR_DEFAULTS = (0.3, 0.3, 2.0, 1 / 0.2, (1 / 0.2) * (1 / 0.6))
LINEAR_RANGE_LO = (0.5 * 0.3, 0.5 * 0.3, 0.5 * 2.0, 1 / 0.2 - 0.07, (1 / 0.2) * (1 / 0.6) - 1.5)
LINEAR_RANGE_HI = (1.5 * 0.3, 1.5 * 0.3, 1.5 * 2.0, 1 / 0.2 + 0.07, (1 / 0.2) * (1 / 0.6) + 1.5)


class Lane(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("road", C.c_int32), ("lane_id", C.c_int32), ("priority", C.c_int32),
        ("forbidden", C.c_int32), ("is_exit", C.c_int32), ("cache_col", C.c_int32), ("pad1", C.c_int32),
        ("ax", C.c_double), ("ay", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
        ("heading", C.c_double), ("length", C.c_double), ("width", C.c_double), ("speed_limit", C.c_double),
        ("radius", C.c_double), ("start_phase", C.c_double), ("end_phase", C.c_double), ("cdir", C.c_double),
        ("amplitude", C.c_double), ("pulsation", C.c_double), ("phase", C.c_double), ("pad2", C.c_double),
    ]


class Road(C.Structure):
    _fields_ = [("from_node", C.c_int32), ("to_node", C.c_int32), ("first_lane", C.c_int32), ("n_lanes", C.c_int32)]


class Config(C.Structure):
    _fields_ = [
        ("n_lanes", C.c_int32), ("n_roads", C.c_int32), ("n_nodes", C.c_int32), ("pad0", C.c_int32),
        ("simulation_frequency", C.c_double), ("policy_frequency", C.c_double), ("duration", C.c_double),
        ("regulated", C.c_int32), ("pad1", C.c_int32),
        ("acc_max", C.c_double), ("comfort_acc_max", C.c_double), ("comfort_acc_min", C.c_double),
        ("distance_wanted", C.c_double), ("time_wanted", C.c_double),
        ("politeness", C.c_double), ("lane_change_min_acc_gain", C.c_double),
        ("lane_change_max_braking_imposed", C.c_double), ("lane_change_delay", C.c_double),
        ("n_target_speeds", C.c_int32), ("action_mode", C.c_int32),
        ("target_speeds", C.c_double * MAX_TARGET_SPEEDS),
        ("obs_type", C.c_int32), ("obs_vehicles", C.c_int32), ("n_features", C.c_int32), ("absolute", C.c_int32),
        ("order", C.c_int32), ("see_behind", C.c_int32), ("normalize", C.c_int32), ("clip", C.c_int32),
        ("features", C.c_int32 * MAX_FEATURES),
        ("has_range", C.c_int32 * MAX_FEATURES),
        ("range_lo", C.c_double * MAX_FEATURES), ("range_hi", C.c_double * MAX_FEATURES),
        ("grid_has_xrange", C.c_int32), ("grid_has_yrange", C.c_int32), ("grid_w", C.c_int32), ("grid_h", C.c_int32),
        ("align_to_vehicle_axes", C.c_int32), ("as_image", C.c_int32), ("ttc_steps", C.c_int32), ("pad3", C.c_int32),
        ("grid_xrange", C.c_double * 2), ("grid_yrange", C.c_double * 2),
        ("grid_min", C.c_double * 2), ("grid_max", C.c_double * 2), ("grid_step", C.c_double * 2),
        ("reward_type", C.c_int32), ("normalize_reward", C.c_int32), ("offroad_terminal", C.c_int32), ("vehicle_model", C.c_int32),
        ("collision_reward", C.c_double), ("high_speed_reward", C.c_double), ("arrived_reward", C.c_double),
        ("lane_reward", C.c_double), ("reward_speed_lo", C.c_double), ("reward_speed_hi", C.c_double),
        ("spawn_enabled", C.c_int32), ("controlled_vehicles", C.c_int32),
        ("spawn_probability", C.c_double),
        ("lane_change_reward", C.c_double), ("speed_index_den", C.c_double),
        ("lin_lo", C.c_double * 5), ("lin_hi", C.c_double * 5), ("lin_default", C.c_double * 5),
    ]


class SpawnDraw(C.Structure):
    _fields_ = [("u_spawn", C.c_double), ("entry", C.c_int32), ("exit", C.c_int32),
                ("n_pos", C.c_double), ("n_speed", C.c_double), ("delta", C.c_double), ("lin_u", C.c_double * 5)]


MAX_SPAWN_ATTEMPTS = 32
MAX_CAST = 8
CAST_DEST = 4
AUTORESET_OFF, AUTORESET_POOL, AUTORESET_DEVICE, AUTORESET_DEVICE_ASYNC = 0, 1, 2, 3


class CastMember(C.Structure):
    _fields_ = [("lane", C.c_int32), ("mdp", C.c_int32), ("n_dest", C.c_int32), ("randomize", C.c_int32),
                ("dest", C.c_int32 * CAST_DEST),
                ("longitudinal", C.c_double), ("longitudinal_std", C.c_double), ("speed", C.c_double), ("speed_std", C.c_double),
                ("heading_longitudinal", C.c_double)]


class ResetParams(C.Structure):
    _fields_ = [("scene", C.c_int32), ("n_vehicles", C.c_int32), ("lanes", C.c_int32), ("ego_entry", C.c_int32),
                ("destination", C.c_int32), ("warmup_substeps", C.c_int32), ("pad0", C.c_int32), ("pad1", C.c_int32),
                ("speed_limit", C.c_double), ("density", C.c_double), ("ego_spacing", C.c_double), ("ego_speed", C.c_double),
                ("ego_longitudinal", C.c_double), ("ego_longitudinal_std", C.c_double),
                ("spawn_longitudinal", C.c_double * MAX_SPAWN_ATTEMPTS),
                ("cast", CastMember * MAX_CAST),
                ("cast_route_len", (C.c_uint8 * CAST_DEST) * MAX_ROADS),
                ("cast_route_road", ((C.c_uint8 * ROUTE_CAP) * CAST_DEST) * MAX_ROADS)]


class EpisodeStats(C.Structure):
    _fields_ = [("episodes", C.c_double), ("total_return", C.c_double), ("total_length", C.c_double),
                ("crashes", C.c_double), ("arrivals", C.c_double), ("total_speed", C.c_double),
                ("vehicle_steps", C.c_double), ("env_steps", C.c_double), ("spawn_capacity_rejects", C.c_double),
                ("sync_resets", C.c_double)]


ABI_VERSION = 2
# ttrl_sim_set_info_outputs: float64 [NINFO][E]
INFO_SPEED, INFO_CRASHED, INFO_REWARDS, NINFO = 0, 1, 2, 6
# keys of info["rewards"] per reward type, in the reference's dict order
REWARD_KEYS = {
    REWARD_INTERSECTION: ("collision_reward", "high_speed_reward", "arrived_reward", "on_road_reward"),   # intersection_env.py:94-104
    REWARD_HIGHWAY: ("collision_reward", "left_lane_reward", "high_speed_reward", "on_road_reward"),       # u_turn_env.py:60-71
    REWARD_ROUNDABOUT: ("collision_reward", "high_speed_reward", "lane_change_reward", "on_road_reward"),  # roundabout_env.py:57-64
}


LOSS_L2, LOSS_L1, LOSS_SMOOTH_L1 = 0, 1, 2


class DqnDesc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("h1", C.c_int32), ("h2", C.c_int32), ("n_actions", C.c_int32), ("batch", C.c_int32),
                ("loss", C.c_int32), ("double_q", C.c_int32), ("gamma", C.c_float)]


class QnetDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("n_entities", C.c_int32), ("n_features", C.c_int32), ("n_actions", C.c_int32),
        ("n_hidden", C.c_int32), ("hidden", C.c_int32 * 4),
        ("embed_layers", C.c_int32), ("embed", C.c_int32 * 4),
        ("feature_size", C.c_int32), ("heads", C.c_int32),
        ("out_layers", C.c_int32), ("out_hidden", C.c_int32 * 4),
        ("presence_feature_idx", C.c_int32),
    ]
