/*
 * ttrl_b200.h -- C ABI of the B200-native batched simulator for the TopoTrafficRL per-step hot path.
 *
 * The reference (pure Python, /root/reference) has no FFI: its boundary is the gymnasium Env protocol
 * (ttrl_env/envs/common/abstract.py:188-250) and the agent protocol (ttrl_agent/agents/deep_q_network/
 * abstract.py:65-83).  This header is the C-ABI drop-in that sits directly under that Python boundary:
 * the host package (topotrafficrl_b200/) mirrors the reference classes and binds these entry points
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross the ABI;
 *  - every function returns 0 on success, non-zero on failure; ttrl_last_error() gives the message;
 *  - "dev" pointers are device pointers on the sim's device (borrowed, never freed by the library);
 *    "host" pointers are ordinary host memory;
 *  - no internal threads; kernels are launched on the stream passed by the caller (0 = legacy default);
 *  - E = number of env instances, V = vehicle slots per env (capacity), L = lanes.
 *
 * All simulation arithmetic is IEEE float64, like the reference (numpy float64 scalars throughout);
 * observations and rewards are emitted as float32 like the reference's (observation.py:275).
 */
#ifndef TTRL_B200_H
#define TTRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Road network: flat lane table in RoadNetwork.graph insertion order (road/road.py:55-71 iterates the
 * dict in that order and np.argmin keeps the first minimum, so the order is part of the semantics).
 * ---------------------------------------------------------------------------------------------- */
enum { TTRL_LANE_STRAIGHT = 0, TTRL_LANE_CIRCULAR = 1, TTRL_LANE_SINE = 2 };

typedef struct ttrl_lane {
    int32_t kind;        /* TTRL_LANE_* (road/lane.py:159, :236, :311) */
    int32_t road;        /* index into the road table: the (_from,_to) pair this lane belongs to */
    int32_t lane_id;     /* index of the lane inside its road (third element of a LaneIndex) */
    int32_t priority;    /* lane.priority (regulation.py:73-76) */
    int32_t forbidden;   /* lane.forbidden (lane.py:111) */
    int32_t is_exit;     /* "il" in _from and "o" in _to (intersection_env.py:352-353, :366-367) */
    int32_t cache_col;   /* reserved (pass 0): the library numbers the curved lanes here (column of its per-vehicle cache) */
    int32_t pad1;
    double ax, ay;       /* straight/sine: start; circular: center */
    double dx, dy;       /* straight/sine: unit direction (lane.py:190); lateral = (-dy, dx) */
    double heading;      /* straight/sine: arctan2 of end-start (lane.py:185) */
    double length;       /* lane.length */
    double width;        /* lane.width */
    double speed_limit;  /* lane.speed_limit */
    double radius, start_phase, end_phase, cdir; /* circular: cdir = +1 clockwise / -1 (lane.py:333) */
    double amplitude, pulsation, phase;          /* sine (lane.py:264-266) */
    double pad2;
} ttrl_lane;

typedef struct ttrl_road {
    int32_t from_node, to_node; /* node ids (strings in the reference, interned by the host) */
    int32_t first_lane;         /* flat index of lane_id 0 of this road */
    int32_t n_lanes;            /* len(graph[_from][_to]) */
} ttrl_road;

#define TTRL_MAX_LANES 64
#define TTRL_MAX_ROADS 64
#define TTRL_MAX_NODES 64
#define TTRL_MAX_TARGET_SPEEDS 8
#define TTRL_MAX_FEATURES 8
#define TTRL_ROUTE_CAP 12      /* route entries per vehicle: 3 words x 4 bytes (roundabout-v0 plans up to 10 roads, road.py:159-188) */
#define TTRL_ROUTE_WORDS 3
#define TTRL_MAX_CONTROLLED 4  /* controlled vehicles (agents) per env: MultiAgentIntersectionEnv places ego k on arm k % 4 */

/* Observation feature ids (vehicle/kinematics.py:237-261, the subset the hot configs use). */
enum { TTRL_F_PRESENCE = 0, TTRL_F_X = 1, TTRL_F_Y = 2, TTRL_F_VX = 3, TTRL_F_VY = 4,
       TTRL_F_COS_H = 5, TTRL_F_SIN_H = 6, TTRL_F_HEADING = 7, TTRL_F_ON_ROAD = 8 };
enum { TTRL_OBS_KINEMATICS = 0, TTRL_OBS_GRID = 1,
       TTRL_OBS_TTC = 2 /* TimeToCollisionObservation (observation.py:114-151 over finite_mdp.compute_ttc_grid :104-163): 3 x 3 x ttc_steps */ };
#define TTRL_MAX_TTC_CELLS 2048  /* target speeds x lanes of the ego's road x ttc_steps */
enum { TTRL_ORDER_SORTED = 0, TTRL_ORDER_SHUFFLED = 1 };
/* DiscreteMetaAction tables (envs/common/action.py:204-211) */
enum { TTRL_ACT_ALL = 0 /* 0 LANE_LEFT 1 IDLE 2 LANE_RIGHT 3 FASTER 4 SLOWER */,
       TTRL_ACT_LONGI = 1 /* 0 SLOWER 1 IDLE 2 FASTER */,
       TTRL_ACT_LAT = 2 /* 0 LANE_LEFT 1 IDLE 2 LANE_RIGHT */ };
enum { TTRL_VEHICLE_IDM = 0, TTRL_VEHICLE_LINEAR = 1 };
enum { TTRL_REWARD_INTERSECTION = 0 /* intersection_env.py:61-104 (mean over the controlled vehicles) */,
       TTRL_REWARD_HIGHWAY = 1 /* u_turn_env.py:39-71: UTurnEnv, and the same template on the synthetic highway */,
       TTRL_REWARD_ROUNDABOUT = 2 /* roundabout_env.py:43-64: speed-index and lane-change terms */ };

typedef struct ttrl_config {
    /* network sizes */
    int32_t n_lanes, n_roads, n_nodes, pad0;
    /* timing (abstract.py:97-98, :252-256) */
    double simulation_frequency, policy_frequency, duration;
    /* RegulatedRoad (regulation.py:14-15, :28-32): regulated != 0 enables enforce_road_rules */
    int32_t regulated, pad1;
    /* IDM / MOBIL class constants (behavior.py:20-46; intersection_env.py:258-261 overrides) */
    double acc_max, comfort_acc_max, comfort_acc_min, distance_wanted, time_wanted;
    double politeness, lane_change_min_acc_gain, lane_change_max_braking_imposed, lane_change_delay;
    /* MDPVehicle (controller.py:259, :317-344) + DiscreteMetaAction (action.py:199-260) */
    int32_t n_target_speeds, action_mode;
    double target_speeds[TTRL_MAX_TARGET_SPEEDS];
    /* observation (observation.py:154-275 Kinematics, :278-498 OccupancyGrid) */
    int32_t obs_type, obs_vehicles, n_features, absolute, order, see_behind, normalize, clip;
    int32_t features[TTRL_MAX_FEATURES];
    /* per-feature normalisation range; has_range[i]==0 -> feature i is not in features_range */
    int32_t has_range[TTRL_MAX_FEATURES];
    double range_lo[TTRL_MAX_FEATURES], range_hi[TTRL_MAX_FEATURES];
    /* OccupancyGrid: features_range of "x"/"y" even when x/y are not observed features (observation.py:375-392) */
    int32_t grid_has_xrange, grid_has_yrange, grid_w, grid_h, align_to_vehicle_axes, as_image;
    int32_t ttc_steps;   /* TimeToCollision: int(horizon * policy_frequency) time cells of 1 / policy_frequency seconds */
    int32_t pad3;
    double grid_xrange[2], grid_yrange[2];
    double grid_min[2], grid_max[2], grid_step[2];
    /* reward / termination */
    int32_t reward_type, normalize_reward, offroad_terminal;
    /* class of the surrounding traffic, config["other_vehicles_type"] (intersection_env.py:257, roundabout_env.py:337):
     * TTRL_VEHICLE_IDM = IDMVehicle (behavior.py:16-348), TTRL_VEHICLE_LINEAR = LinearVehicle (behavior.py:350-558: acceleration
     * and steering linear in per-vehicle parameters, TIME_WANTED 2.5 -- the host passes that as time_wanted) */
    int32_t vehicle_model;
    double collision_reward, high_speed_reward, arrived_reward, lane_reward;
    double reward_speed_lo, reward_speed_hi;
    /* IntersectionEnv spawn / clear (intersection_env.py:320-362) */
    int32_t spawn_enabled;
    /* number of controlled vehicles K (config["controlled_vehicles"], intersection_env.py:285-307); 0 is read as 1.
     * K > 1 = MultiAgentAction / MultiAgentObservation (action.py:301-331, observation.py:587-603): every per-env
     * action / observation / shuffle-permutation array of this ABI holds K consecutive entries per env. */
    int32_t controlled_vehicles;
    double spawn_probability;
    /* RoundaboutEnv (roundabout_env.py:43-64): weight of `action in [0, 2]`, and the denominator of the speed-index
     * term, MDPVehicle.DEFAULT_TARGET_SPEEDS.size - 1 (controller.py:259) */
    double lane_change_reward, speed_index_den;
    /* LinearVehicle.randomize_behavior (behavior.py:402-410): parameter k = lin_lo[k] + u * (lin_hi[k] - lin_lo[k]) with
     * ACCELERATION_RANGE / STEERING_RANGE (behavior.py:359-371) in TTRL_LIN_* order; lin_default = the class attributes */
    double lin_lo[5], lin_hi[5], lin_default[5];
} ttrl_config;

/* ------------------------------------------------------------------------------------------------
 * Vehicle state, struct-of-arrays.  Field f of slot s of env e lives at  buf[(f*E + e)*V + s].
 * Slots 0..n_vehicles[e]-1 are live, in Road.vehicles list order (order is semantic: road.py:461-478).
 * ---------------------------------------------------------------------------------------------- */
enum { TTRL_D_X = 0, TTRL_D_Y, TTRL_D_HEADING, TTRL_D_SPEED, TTRL_D_STEERING, TTRL_D_ACCEL,
       TTRL_D_TARGET_SPEED, TTRL_D_TIMER, TTRL_D_DELTA, TTRL_D_IMPACT_X, TTRL_D_IMPACT_Y, TTRL_ND = 11 };
enum { TTRL_I_LANE = 0, TTRL_I_TARGET_LANE, TTRL_I_FLAGS, TTRL_I_SPEED_INDEX, TTRL_I_ROUTE_LEN,
       TTRL_I_ROUTE_ROAD, TTRL_I_ROUTE_LANE, TTRL_I_YIELD_TIMER,
       TTRL_I_ROUTE_ROAD1, TTRL_I_ROUTE_ROAD2, TTRL_I_ROUTE_LANE1, TTRL_I_ROUTE_LANE2, TTRL_NI = 12 };
/* TTRL_I_FLAGS bits */
enum { TTRL_FL_MDP = 1 /* MDPVehicle (else IDMVehicle) */, TTRL_FL_CRASHED = 2, TTRL_FL_HAS_IMPACT = 4,
       TTRL_FL_YIELDING = 8 /* is_yielding attribute is True */, TTRL_FL_CONTROLLED = 16 /* in env.controlled_vehicles */,
       TTRL_FL_AGENT_SHIFT = 8, TTRL_FL_AGENT_MASK = 0x700 /* index of the vehicle in env.controlled_vehicles */ };
/* TTRL_I_ROUTE_LEN: -1 = route is None, else number of remaining entries (<= TTRL_ROUTE_CAP).
 * Route entry k (road index / lane id, 0xFF = None) is byte k % 4 of word k / 4; the words of the road bytes are
 * TTRL_I_ROUTE_ROAD, _ROAD1, _ROAD2, those of the lane bytes TTRL_I_ROUTE_LANE, _LANE1, _LANE2. */

/* Per-env scalars: ibuf[f*E + e], dbuf[f*E + e] */
enum { TTRL_EI_NVEH = 0, TTRL_EI_STEPS /* env.steps, abstract.py:273 */, TTRL_EI_ROAD_STEPS /* RegulatedRoad.steps */,
       TTRL_EI_EGO /* slot of controlled_vehicles[0]; agents k >= 1 are found by their TTRL_FL_AGENT bits */, TTRL_EI_EPISODE, TTRL_EI_DONE, TTRL_NEI = 6 };
enum { TTRL_ED_TIME = 0 /* env.time, abstract.py:239 */, TTRL_ED_RETURN, TTRL_NED = 2 };

/* Spawn draws injected for parity (the reference draws from numpy PCG64: intersection_env.py:328-346,
 * behavior.py:66-69).  One record per env per step; u_spawn > spawn_probability means "no spawn". */
typedef struct ttrl_spawn_draw {
    double u_spawn;      /* np_random.uniform()                      (:328) */
    int32_t entry, exit; /* np_random.choice(range(4), 2, False)      (:331) */
    double n_pos, n_speed; /* np_random.normal() x2                   (:338, :340) */
    double delta;        /* IDMVehicle.randomize_behavior: np_random.uniform(3.5, 4.5)            (behavior.py:67) */
    double lin_u[5];     /* LinearVehicle.randomize_behavior: np_random.uniform(size=3), then (size=2) (behavior.py:402-410) */
} ttrl_spawn_draw;

/* LinearVehicle parameters, one block per vehicle slot next to the state arrays (sims with vehicle_model == TTRL_VEHICLE_LINEAR
 * only): field f of slot s of env e at buf[(f*E + e)*V + s].  ACCELERATION_PARAMETERS[0..2], STEERING_PARAMETERS[0..1]
 * (behavior.py:353-357; class defaults 0.3, 0.3, 2.0 and KP_HEADING, KP_HEADING * KP_LATERAL). */
enum { TTRL_LIN_ACC0 = 0, TTRL_LIN_ACC1, TTRL_LIN_ACC2, TTRL_LIN_STEER0, TTRL_LIN_STEER1, TTRL_NLIN = 5 };

/* Episode statistics accumulated on device (maps to Evaluation.after_all_episodes, trainer/evaluation.py:325-333) */
typedef struct ttrl_episode_stats {
    double episodes, total_return, total_length, crashes, arrivals, total_speed, vehicle_steps, env_steps;
    /* spawn attempts that found a free place on the road but no free vehicle slot (the reference's vehicle list has no
     * capacity): non-zero means the slot capacity `vcap` is too small for the workload and the episodes DIFFER from the
     * reference's from that point on */
    double spawn_capacity_rejects;
    /* TTRL_AUTORESET_DEVICE_ASYNC: finished envs whose next episode was not ready in the shadow ring and that were reset by
     * the packed second launch on the step's critical path instead (same episode, only slower) */
    double sync_resets;
} ttrl_episode_stats;

/* Device-side reset: fresh episodes generated on the GPU with counter-based (Philox) draws
 * keyed by (seed, global env, episode, draw), following the reference's procedures:
 *   scene 0: synthetic highway, Vehicle.create_random's placement rule (kinematics.py:50-104) + randomize_behavior;
 *   scene 1: IntersectionEnv._make_vehicles (intersection_env.py:251-318): staggered spawn attempts, warm-up
 *            sub-steps, challenger, MDPVehicle ego with its route, pruning of vehicles within 20 m of the ego;
 *   scene 2: a scripted cast -- RoundaboutEnv._make_vehicles (roundabout_env.py:326-387), UTurnEnv._make_vehicles
 *            (u_turn_env.py:173-271): every member is made on a lane at longitudinal + N(0, 1) * std with speed + N(0, 1) *
 *            std (RoadObject.make_on_lane objects.py:67-89), takes the closest lane at that pose (objects.py:45-50), plans
 *            a route to a fixed or uniformly drawn destination (controller.py:71-87) and optionally draws its DELTA. */
#define TTRL_MAX_SPAWN_ATTEMPTS 32
#define TTRL_MAX_CAST 8
#define TTRL_CAST_DEST 4
typedef struct ttrl_cast_member {
    int32_t lane;        /* flat index of the lane the member is made on */
    int32_t mdp;         /* 1: the controlled MDPVehicle (no draws); 0: IDMVehicle */
    int32_t n_dest;      /* 0: route None; 1: dest[0]; > 1: np_random.choice over dest[0..n_dest-1] */
    int32_t randomize;   /* randomize_behavior(): DELTA ~ U[3.5, 4.5] (behavior.py:66-69), else the class default 4.0 */
    int32_t dest[TTRL_CAST_DEST];  /* destination ids = columns of cast_route_* */
    double longitudinal, longitudinal_std, speed, speed_std;
    double heading_longitudinal;   /* heading = lane.heading_at(this) (the roundabout's ego: position at 125, heading at 140) */
} ttrl_cast_member;
typedef struct ttrl_reset_params {
    int32_t scene;            /* 0 highway, 1 intersection, 2 scripted cast (roundabout, u-turn) */
    int32_t n_vehicles;       /* highway: vehicles per env incl. the ego; intersection: initial_vehicle_count */
    int32_t lanes;            /* highway: lanes of road 0 */
    int32_t ego_entry;        /* intersection: corner of ego 0's start lane (o<k>, ir<k>, 0); ego j starts on corner (k + j) % 4 */
    int32_t destination;      /* intersection: exit corner 0..3, or -1 = drawn per episode ("destination": None) */
    int32_t warmup_substeps;  /* intersection: 3 * simulation_frequency (intersection_env.py:267-274) */
    int32_t pad0, pad1;
    double speed_limit, density, ego_spacing, ego_speed;          /* highway */
    double ego_longitudinal, ego_longitudinal_std;                /* intersection: 60 + 5 * N(1, 1) -> 60, 5 */
    double spawn_longitudinal[TTRL_MAX_SPAWN_ATTEMPTS];           /* intersection: np.linspace(0, 80, n)[t] */
    /* scene 2: the cast (n_vehicles members, list order) and, per (road of the member's closest lane, destination id), the
     * roads plan_route_to appends after the member's own lane (BFS of road.py:159-188 done on the host) */
    ttrl_cast_member cast[TTRL_MAX_CAST];
    uint8_t cast_route_len[TTRL_MAX_ROADS][TTRL_CAST_DEST];
    uint8_t cast_route_road[TTRL_MAX_ROADS][TTRL_CAST_DEST][TTRL_ROUTE_CAP];
} ttrl_reset_params;

typedef struct ttrl_sim ttrl_sim;
typedef struct ttrl_qnet ttrl_qnet;

const char* ttrl_last_error(void);
void ttrl_set_error(const char* msg);
int ttrl_abi_version(void);
/* sizeof of the POD structs as compiled (0 lane, 1 road, 2 config, 3 spawn_draw, 4 episode_stats, 5 qnet_desc, 6 reset_params,
 * 7 cast_member): binding self-check */
int ttrl_abi_sizeof(int which);

/* Create E env instances with V vehicle slots each on CUDA device `device`.
 * Replaces: AbstractEnv.__init__/configure (abstract.py:44-113) + _make_road (intersection_env.py:141-249,
 * road.py:291-321) for a batch.  `node_roads`/`node_first`: CSR adjacency "roads starting at node n" in
 * graph[_to].keys() order (road.py:126). */
int ttrl_sim_create(const ttrl_config* cfg, const ttrl_lane* lanes, const ttrl_road* roads,
                    const int32_t* node_first /* n_nodes+1 */, const int32_t* node_roads,
                    int num_envs, int vcap, int device, ttrl_sim** out);
int ttrl_sim_destroy(ttrl_sim* sim);
/* IntersectionEnv spawn routes: entry lane per corner and, per (entry, exit), the roads that plan_route_to
 * (controller.py:71-87, BFS of road.py:159-188 done on the host) appends after the entry lane. */
int ttrl_sim_set_spawn_routes(ttrl_sim* sim, const int32_t* spawn_lane /*4*/, const int32_t* route_len /*16*/,
                              const int32_t* route_road /*16*TTRL_ROUTE_CAP*/);
int ttrl_sim_num_envs(const ttrl_sim* sim);
int ttrl_sim_vcap(const ttrl_sim* sim);
int ttrl_sim_obs_size(const ttrl_sim* sim); /* floats per env (K observations of the configured shape) */
int ttrl_sim_num_agents(const ttrl_sim* sim); /* K = controlled vehicles per env */
/* Per-agent results of the last ttrl_sim_step: info["agents_rewards"] / info["agents_terminated"]
 * (IntersectionEnv._info intersection_env.py:121-129; what MultiAgentWrapper.step returns, abstract.py:432-441).
 * float32[E*K] / uint8[E*K] device buffers owned by the library, valid for the sim's lifetime. */
int ttrl_sim_agent_outputs(ttrl_sim* sim, const float** agent_reward_dev, const uint8_t** agent_terminated_dev);
/* Redirect those per-agent results into caller-owned device buffers (borrowed; NULL = back to the library's own). */
int ttrl_sim_set_agent_outputs(ttrl_sim* sim, float* agent_reward_dev, uint8_t* agent_terminated_dev);

/* The step's `info` dict for a batch (AbstractEnv._info abstract.py:169-186, evaluated like the reference right after the
 * reward and BEFORE IntersectionEnv.step's clear / spawn and before any autoreset), one value per key per env, float64:
 *   info_dev[TTRL_INFO_SPEED * E + e]     info["speed"]    = self.vehicle.speed (controlled_vehicles[0])
 *   info_dev[TTRL_INFO_CRASHED * E + e]   info["crashed"]  = self.vehicle.crashed (0 / 1)
 *   info_dev[(TTRL_INFO_REWARDS + k) * E + e], k = 0..3: info["rewards"], the dict of _rewards(action) in its own key order
 *     (mean over the controlled vehicles, intersection_env.py:67-76):
 *       TTRL_REWARD_INTERSECTION: collision_reward, high_speed_reward, arrived_reward, on_road_reward (intersection_env.py:94-104)
 *       TTRL_REWARD_HIGHWAY:      collision_reward, left_lane_reward, high_speed_reward, on_road_reward (u_turn_env.py:60-71)
 *       TTRL_REWARD_ROUNDABOUT:   collision_reward, high_speed_reward, lane_change_reward, on_road_reward (roundabout_env.py:57-64)
 * final_obs_dev float32[E * obs_size]: gymnasium's info["final_observation"] -- with autoreset on, an env that finished in this
 * step has its first observation of the NEXT episode in obs_dev and its terminal observation (what the reference's step()
 * returned) here; rows of envs that did not finish are left untouched.  Either pointer may be NULL (= not wanted).
 * Borrowed device buffers, used by ttrl_sim_step on any stream (the host-buffer step has its own: ttrl_sim_host_info_buffers). */
enum { TTRL_INFO_SPEED = 0, TTRL_INFO_CRASHED = 1, TTRL_INFO_REWARDS = 2, TTRL_NINFO = 6 };
int ttrl_sim_set_info_outputs(ttrl_sim* sim, double* info_dev /* [TTRL_NINFO][E] */, float* final_obs_dev /* [E][obs_size] */);
/* page-locked double info[TTRL_NINFO][E] and float final_obs[E][obs_size], filled by every ttrl_sim_step_pinned /
 * ttrl_sim_step_host call after ttrl_sim_host_info_buffers was called once (it switches the two outputs on for the host path) */
int ttrl_sim_host_info_buffers(ttrl_sim* sim, double** info, float** final_obs);

/* Resync / golden capture (host buffers, layout above).  Replaces direct attribute access on
 * Vehicle objects (kinematics.py:34-48, controller.py:35-48, behavior.py:48-64). */
int ttrl_sim_set_state(ttrl_sim* sim, const double* veh_d, const int32_t* veh_i,
                       const int32_t* env_i, const double* env_d);
int ttrl_sim_get_state(ttrl_sim* sim, double* veh_d, int32_t* veh_i, int32_t* env_i, double* env_d);
/* The LinearVehicle parameter block of the state / of the reset pool (host double[TTRL_NLIN * E * V] resp. [TTRL_NLIN * pool_size * V]);
 * fail unless the sim was created with vehicle_model == TTRL_VEHICLE_LINEAR.  Call the pool form after ttrl_sim_set_reset_pool. */
int ttrl_sim_set_linear_params(ttrl_sim* sim, const double* lin);
int ttrl_sim_get_linear_params(ttrl_sim* sim, double* lin);
int ttrl_sim_set_reset_pool_linear_params(ttrl_sim* sim, const double* lin);

/* Store the current state as reset pool entry `slot` (autoreset source; reset itself stays on the host:
 * AbstractEnv.reset, abstract.py:188-214).  Env e restarts from pool entry (e + k*E) mod pool_size at its k-th reset. */
int ttrl_sim_set_reset_pool(ttrl_sim* sim, int pool_size, const double* veh_d, const int32_t* veh_i,
                            const int32_t* env_i, const double* env_d);
/* autoreset mode of ttrl_sim_step: 0 = off, 1 = restart finished envs from the reset pool, 2 = device-side reset
 * (ttrl_sim_set_reset_params must have been called). */
enum { TTRL_AUTORESET_OFF = 0, TTRL_AUTORESET_POOL = 1, TTRL_AUTORESET_DEVICE = 2,
       /* same episodes as TTRL_AUTORESET_DEVICE, bit for bit (the draws are keyed by (seed, global env, episode)), but for scenes
        * whose reset runs warm-up sub-steps (intersection: 45) every env's NEXT episode is generated ahead of time into a shadow
        * buffer by a kernel on a side stream, overlapping the following env-steps; a finished env restarts from its shadow inside
        * the step kernel.  Envs that finish again before their shadow is ready take the synchronous path. */
       TTRL_AUTORESET_DEVICE_ASYNC = 3 };
int ttrl_sim_set_autoreset(ttrl_sim* sim, int mode);
int ttrl_sim_set_reset_params(ttrl_sim* sim, const ttrl_reset_params* params);
/* Device-side reset of every env (mask_dev == NULL) or of the envs with mask_dev[e] != 0; replaces
 * AbstractEnv.reset (abstract.py:188-214) for a batch.  Uses the seed of ttrl_sim_seed. */
int ttrl_sim_reset(ttrl_sim* sim, const uint8_t* mask_dev, void* stream);

/* One simulation sub-step for every env: [ego meta-action if steps % F == 0] -> Road.act -> (regulation)
 * -> Road.step.  Replaces one iteration of AbstractEnv._simulate (abstract.py:257-273).
 * actions_dev: int32[E] or NULL (no action, like action=None). */
int ttrl_sim_substep(ttrl_sim* sim, const int32_t* actions_dev /* int32[E*K] */, void* stream);

/* One env.step() for every env.  Replaces AbstractEnv.step (abstract.py:224-250) +
 * IntersectionEnv.step's clear/spawn (intersection_env.py:135-139).
 * actions_dev int32[E*K]; obs_dev float32[E*obs_size]; reward_dev float32[E]; terminated_dev/truncated_dev uint8[E]. */
int ttrl_sim_step(ttrl_sim* sim, const int32_t* actions_dev, float* obs_dev, float* reward_dev,
                  uint8_t* terminated_dev, uint8_t* truncated_dev, void* stream);

/* Same call with HOST buffers: H2D of actions, the step, D2H of obs/reward/flags, synchronised on return.
 * This is the reference-facing end-to-end form (one gym step for E envs from numpy arrays). */
int ttrl_sim_step_host(ttrl_sim* sim, const int32_t* actions_host, float* obs_host, float* reward_host,
                       uint8_t* terminated_host, uint8_t* truncated_host);

/* Zero-copy form of the host-buffer step: the library's own page-locked staging buffers (int32 actions[E], float
 * obs[E*obs_size], float reward[E], uint8 terminated[E], uint8 truncated[E]; valid for the sim's lifetime) are
 * handed to the caller, who writes the actions there, calls ttrl_sim_step_pinned and reads the results in place
 * (they are overwritten by the next call). */
int ttrl_sim_host_buffers(ttrl_sim* sim, int32_t** actions, float** obs, float** reward, uint8_t** terminated, uint8_t** truncated);
int ttrl_sim_step_pinned(ttrl_sim* sim, int use_actions);
/* page-locked float agent_reward[E*K], uint8 agent_terminated[E*K]; filled by every ttrl_sim_step_pinned call */
int ttrl_sim_host_agent_buffers(ttrl_sim* sim, float** agent_reward, uint8_t** agent_terminated);

/* Observation only (observation_type.observe() at reset: abstract.py:210); also fills the info buffer of
 * ttrl_sim_set_info_outputs (speed, crashed, _rewards with no action: the reset's info dict, abstract.py:211). */
int ttrl_sim_observe(ttrl_sim* sim, float* obs_dev, void* stream);

/* Parity hooks: feed the oracle's RNG draws (spawn decisions; Kinematics "shuffled" permutations). */
int ttrl_sim_inject_spawn(ttrl_sim* sim, const ttrl_spawn_draw* draws_host /* E records or NULL to clear */);
int ttrl_sim_inject_shuffle(ttrl_sim* sim, const int32_t* perm_host /* E*K*(obs_vehicles-1) or NULL */);
/* seed != 0 enables device-side (Philox) draws keyed by (seed, first_global_env + e, episode, step): the spawn draws and, for
 * Kinematics order == "shuffled" without an injected permutation, the row shuffle of every observation (a Fisher-Yates
 * permutation per (env, agent, step), np_random.shuffle(obs[1:]) observation.py:272-273).  Without a seed and without an
 * injected permutation the rows of a "shuffled" observation come out in the pre-shuffle order (Road.vehicles list order,
 * road.py:418-447 with sort=False) and the caller applies its own permutation (the single-env front end does, from the
 * env's numpy stream). */
int ttrl_sim_seed(ttrl_sim* sim, uint64_t seed, int64_t first_global_env);
/* spawn outcome of the last step (1 = a vehicle was appended), int32[E] to host */
int ttrl_sim_spawn_accepted(ttrl_sim* sim, int32_t* accepted_host);
/* Host-driven reset primitive: ONE _spawn_vehicle attempt with explicit arguments, as _make_vehicles calls it
 * (intersection_env.py:265-283).  Synchronous. */
int ttrl_sim_spawn(ttrl_sim* sim, const ttrl_spawn_draw* draws_host, double longitudinal, double position_deviation,
                   double speed_deviation, double spawn_probability, int go_straight, int32_t* accepted_host);

int ttrl_sim_read_stats(ttrl_sim* sim, ttrl_episode_stats* out_host, int reset_after_read);
/* number of kernel launches issued by this sim so far (bench.py's gpu_launches) */
int64_t ttrl_sim_launch_count(const ttrl_sim* sim);

/* ------------------------------------------------------------------------------------------------
 * DQN Q-network rollout.  Replaces AbstractDQNAgent.act (deep_q_network/abstract.py:65-83) ->
 * DQNAgent.get_batch_state_action_values (pytorch.py:79-80) -> EpsilonGreedy.update/sample
 * (exploration/epsilon_greedy.py:32-48, exploration/abstract.py:20-25) for a batch of E observations.
 * ---------------------------------------------------------------------------------------------- */
enum { TTRL_QNET_MLP = 0 /* models.py:50-76 */, TTRL_QNET_EGO_ATTENTION = 1 /* models.py:237-312 */,
       TTRL_QNET_DUELING = 2 /* models.py:79-104 */ };

typedef struct ttrl_qnet_desc {
    int32_t type;
    int32_t n_entities, n_features;   /* observation (V, Fe) */
    int32_t n_actions;
    int32_t n_hidden;                 /* MLP / dueling base: number of hidden layers */
    int32_t hidden[4];                /* their sizes */
    int32_t embed_layers, embed[4];   /* ego-attention: embedding MLP sizes (ego and others share shapes) */
    int32_t feature_size, heads;      /* EgoAttention */
    int32_t out_layers, out_hidden[4];/* output_layer MLP */
    int32_t presence_feature_idx;
} ttrl_qnet_desc;

/* weights_host: all parameters as one float32 blob in the order documented in INTEGRATION.md
 * (reference state_dict order: models.py:56-59, :163-166, :249-255). */
int ttrl_qnet_create(const ttrl_qnet_desc* desc, const float* weights_host, int64_t n_weights,
                     int device, ttrl_qnet** out);
int ttrl_qnet_destroy(ttrl_qnet* q);
/* Replace all parameters (same blob layout as ttrl_qnet_create): the training driver refreshes the rollout network after
 * optimiser steps (value_net of DQNAgent, pytorch.py:15-31).  on_device != 0: `weights` is a device pointer and the copy
 * is enqueued on `stream`, ordered before later ttrl_qnet_act calls on that stream. */
int ttrl_qnet_set_weights(ttrl_qnet* q, const float* weights, int64_t n_weights, int on_device, void* stream);
/* Arithmetic of the forward pass.  FP32 (default): CUDA-core fp32 FMA, the parity path (actions bit-exact vs the
 * reference's torch CPU fp32 forward up to summation order).  TENSOR: the hidden GEMMs on the tcgen05 tensor cores
 * with BF16x3 split operands and FP32 accumulation in TMEM (MultiLayerPerceptron, two hidden layers); Q-values agree
 * to ~1e-5, the greedy action can differ on near-ties. */
enum { TTRL_QNET_MODE_FP32 = 0, TTRL_QNET_MODE_TENSOR = 1 };
int ttrl_qnet_set_mode(ttrl_qnet* q, int mode);
/* Q-values + epsilon-greedy action for E observations.  q_dev may be NULL.  epsilon<=0 -> Greedy. */
int ttrl_qnet_act(ttrl_qnet* q, const float* obs_dev, int num_envs, double epsilon, uint64_t seed,
                  uint64_t step, int32_t* actions_dev, float* q_dev, void* stream);
/* Same with caller-supplied exploration uniforms u_dev double[E] (np_random.choice draws one U[0,1) per call). */
int ttrl_qnet_act_injected(ttrl_qnet* q, const float* obs_dev, int num_envs, double epsilon, const double* u_dev,
                           int32_t* actions_dev, float* q_dev, void* stream);
int64_t ttrl_qnet_launch_count(const ttrl_qnet* q);

/* ------------------------------------------------------------------------------------------------
 * DQN update (training driver) for the MultiLayerPerceptron model with two hidden layers.
 * Replaces DQNAgent.compute_bellman_residual (deep_q_network/pytorch.py:41-73: value_net(s), value_net(s'), target_net(s'),
 * double-DQN target, loss) + DQNAgent.step_optimizer (pytorch.py:32-39: backward, clamp to [-1, 1], Adam) for one minibatch
 * gathered from a device-resident replay memory.  Parameters are six float32 device tensors in torch's state_dict order and
 * layout: layers.0.weight [h1][n_in], layers.0.bias, layers.1.weight [h2][h1], layers.1.bias, predict.weight [A][h2], predict.bias.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ttrl_dqn ttrl_dqn;
enum { TTRL_LOSS_L2 = 0, TTRL_LOSS_L1 = 1, TTRL_LOSS_SMOOTH_L1 = 2 }; /* loss_function_factory: F.mse_loss / l1_loss / smooth_l1_loss */
typedef struct ttrl_dqn_desc {
    int32_t n_in, h1, h2, n_actions;  /* flattened observation size, hidden widths (<= 128), actions (<= 7) */
    int32_t batch;                    /* minibatch size (any; processed 64 rows at a time) */
    int32_t loss;                     /* TTRL_LOSS_* */
    int32_t double_q;                 /* config["double"] (pytorch.py:52-60) */
    float gamma;
} ttrl_dqn_desc;
int ttrl_dqn_create(const ttrl_dqn_desc* desc, int device, ttrl_dqn** out);
int ttrl_dqn_destroy(ttrl_dqn* q);
int64_t ttrl_dqn_num_params(const ttrl_dqn* q);
int64_t ttrl_dqn_launch_count(const ttrl_dqn* q);
/* Flat gradient (parameter order above, float32[num_params]) and mean loss of the minibatch `idx_dev` (int64[batch] rows of the
 * replay memory: state / next_state float32[capacity][n_in], action int64, reward float32, terminal uint8).  One launch. */
int ttrl_dqn_grad(ttrl_dqn* q, const float* const* value_params /* 6 device pointers */, const float* const* target_params,
                  const float* state_dev, const float* next_state_dev, const int64_t* action_dev, const float* reward_dev,
                  const uint8_t* terminal_dev, const int64_t* idx_dev, float* grad_dev, float* loss_dev, void* stream);
/* grad * grad_scale (1 / world size after an all-reduce), clamp to +-grad_clamp, torch.optim.Adam's update number `step` (1-based) of
 * the six tensors and their exp_avg / exp_avg_sq; rollout_blob_dev (or NULL): the weight blob of a ttrl_qnet of the same
 * architecture, refreshed in the same pass (same layout as ttrl_qnet_set_weights).  One launch. */
int ttrl_dqn_adam(ttrl_dqn* q, float* const* params, float* const* exp_avg, float* const* exp_avg_sq, const float* grad_dev, int64_t step,
                  double lr, double beta1, double beta2, double eps, double weight_decay, double grad_clamp, double grad_scale,
                  float* rollout_blob_dev, void* stream);
/* device pointer of a ttrl_qnet's float32 weight blob (layout of ttrl_qnet_create) for ttrl_dqn_adam's rollout_blob_dev.
 * Once it has been handed out the library assumes the caller's kernels write the blob in place: the tensor-core MLP path
 * then rebuilds its derived weight image (one small extra launch) on every ttrl_qnet_act instead of only after
 * ttrl_qnet_set_weights. */
float* ttrl_qnet_weights_dev(ttrl_qnet* q);

#ifdef __cplusplus
}
#endif
#endif /* TTRL_B200_H */
