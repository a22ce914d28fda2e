/*
 * ttrl_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, float64, sequential) of the
 * TopoTrafficRL per-step hot path, one function per reference function, in the reference's own order of
 * operations.  It is the checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline / --impl reference leg) and is never linked, imported or executed by the product package.
 *
 * Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
 * against outputs of the reference itself, run unmodified in the build container through oracle/ref_shim.py;
 * the vectors live in tests/golden/ with the generating script (tests/golden/make_golden.py) and are
 * checked by tests/test_oracle_vs_reference.py.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Data formats (lane table, SoA state, config) are the ones declared in include/ttrl_b200.h.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/ttrl_b200.h"

#define VMAX 256
#define PI 3.141592653589793 /* np.pi */

typedef struct {
    double x, y, heading, speed, steering, accel, target_speed, timer, delta, impact_x, impact_y;
    int lane, target_lane, flags, speed_index, route_len, yield_timer;
    int route_road[TTRL_ROUTE_CAP], route_lane[TTRL_ROUTE_CAP];
    double lin[TTRL_NLIN]; /* LinearVehicle.ACCELERATION_PARAMETERS[3] + STEERING_PARAMETERS[2] (behavior.py:353-357) */
} veh_t;

typedef struct {
    int n, steps, road_steps, ego, episode, done;
    int egos[TTRL_MAX_CONTROLLED]; /* slots of env.controlled_vehicles; egos[0] == ego */
    double time, ret;
    veh_t v[VMAX];
} env_t;

typedef struct orc_scene {
    ttrl_config cfg;
    ttrl_lane lanes[TTRL_MAX_LANES];
    ttrl_road roads[TTRL_MAX_ROADS];
    int node_first[TTRL_MAX_NODES + 1];
    int node_roads[TTRL_MAX_ROADS];
    /* reset pool (autoreset source), SoA like the state buffers */
    int pool_size, pool_E, pool_V;
    double* pool_vd; int32_t* pool_vi; int32_t* pool_ei; double* pool_ed;
    int autoreset;
    /* spawn route table: route for (entry lane, exit node) pairs, filled by the host (BFS there) */
    int spawn_lane[4];
    int spawn_route_len[4][4];
    int spawn_route_road[4][4][TTRL_ROUTE_CAP];
    /* LinearVehicle parameter blocks [TTRL_NLIN][E][V] of the state arrays of the next calls / of the pool (NULL: class defaults) */
    double* lin; const double* pool_lin;
} orc_scene;

/* ----------------------------------------------------------------------------------------------
 * utils.py
 * -------------------------------------------------------------------------------------------- */
/* utils.py:48-54 */
static double not_zero(double x) {
    const double eps = 1e-2;
    if (fabs(x) > eps) return x;
    else if (x >= 0) return eps;
    else return -eps;
}
/* Python/numpy floored modulo for floats (utils.py:58 uses `%`) */
static double py_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) { if ((b < 0) != (m < 0)) m += b; }
    else m = copysign(0.0, b);
    return m;
}
/* utils.py:57-58 */
static double wrap_to_pi(double x) { return py_mod(x + PI, 2 * PI) - PI; }
/* utils.py:29-31 */
static double lmap(double v, double x0, double x1, double y0, double y1) {
    return y0 + (v - x0) * (y1 - y0) / (x1 - x0);
}
static double clipd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

double orc_wrap_to_pi(double x) { return wrap_to_pi(x); }
double orc_not_zero(double x) { return not_zero(x); }

/* ----------------------------------------------------------------------------------------------
 * road/lane.py
 * -------------------------------------------------------------------------------------------- */
/* StraightLane.local_coordinates lane.py:209-213; SineLane :282-286; CircularLane :355-362 */
static void lane_local(const ttrl_lane* l, double px, double py, double* s, double* r) {
    if (l->kind == TTRL_LANE_CIRCULAR) {
        double dx = px - l->ax, dy = py - l->ay;
        double phi = atan2(dy, dx);
        phi = l->start_phase + wrap_to_pi(phi - l->start_phase);
        double rad = sqrt(dx * dx + dy * dy);
        *s = l->cdir * (phi - l->start_phase) * l->radius;
        *r = l->cdir * (l->radius - rad);
    } else {
        double dx = px - l->ax, dy = py - l->ay;
        double lon = dx * l->dx + dy * l->dy;
        double lat = dx * (-l->dy) + dy * l->dx;
        if (l->kind == TTRL_LANE_SINE) lat = lat - l->amplitude * sin(l->pulsation * lon + l->phase);
        *s = lon; *r = lat;
    }
}
/* StraightLane.position lane.py:196-201; SineLane :268-273; CircularLane :341-345 */
static void lane_position(const ttrl_lane* l, double s, double r, double* px, double* py) {
    if (l->kind == TTRL_LANE_CIRCULAR) {
        double phi = l->cdir * s / l->radius + l->start_phase;
        double rr = l->radius - r * l->cdir;
        *px = l->ax + rr * cos(phi);
        *py = l->ay + rr * sin(phi);
    } else {
        if (l->kind == TTRL_LANE_SINE) r = r + l->amplitude * sin(l->pulsation * s + l->phase);
        *px = l->ax + s * l->dx + r * (-l->dy);
        *py = l->ay + s * l->dy + r * l->dx;
    }
}
/* heading_at: lane.py:203-204, :275-280, :347-350 */
static double lane_heading_at(const ttrl_lane* l, double s) {
    if (l->kind == TTRL_LANE_CIRCULAR) {
        double phi = l->cdir * s / l->radius + l->start_phase;
        return phi + PI / 2 * l->cdir;
    }
    if (l->kind == TTRL_LANE_SINE)
        return l->heading + atan(l->amplitude * l->pulsation * cos(l->pulsation * s + l->phase));
    return l->heading;
}
/* AbstractLane.on_lane lane.py:80-102 (with known s, r) */
static int lane_on_lane_sr(const ttrl_lane* l, double s, double r, double margin) {
    return fabs(r) <= l->width / 2 + margin && -5.0 <= s && s < l->length + 5.0;
}
/* AbstractLane.is_reachable_from lane.py:104-118 */
static int lane_reachable_sr(const ttrl_lane* l, double s, double r) {
    if (l->forbidden) return 0;
    return fabs(r) <= 2 * l->width && 0 <= s && s < l->length + 5.0;
}
/* AbstractLane.distance lane.py:127-130 */
static double lane_distance(const ttrl_lane* l, double px, double py) {
    double s, r; lane_local(l, px, py, &s, &r);
    return fabs(r) + fmax(s - l->length, 0) + fmax(0 - s, 0);
}
/* AbstractLane.distance_with_heading lane.py:132-143 + local_angle :145-147 */
static double lane_distance_with_heading(const ttrl_lane* l, double px, double py, double heading) {
    double s, r; lane_local(l, px, py, &s, &r);
    double angle = fabs(wrap_to_pi(heading - lane_heading_at(l, s)));
    return fabs(r) + fmax(s - l->length, 0) + fmax(0 - s, 0) + 1.0 * angle;
}

void orc_lane_local(const ttrl_lane* l, double px, double py, double* out) { lane_local(l, px, py, out, out + 1); }
void orc_lane_position(const ttrl_lane* l, double s, double r, double* out) { lane_position(l, s, r, out, out + 1); }
double orc_lane_heading_at(const ttrl_lane* l, double s) { return lane_heading_at(l, s); }
double orc_lane_distance_with_heading(const ttrl_lane* l, double px, double py, double h) {
    return lane_distance_with_heading(l, px, py, h);
}

/* ----------------------------------------------------------------------------------------------
 * road/road.py -- RoadNetwork
 * -------------------------------------------------------------------------------------------- */
/* RoadNetwork.get_closest_lane_index road.py:55-71 (np.argmin: first minimum wins) */
static int closest_lane(const orc_scene* sc, double px, double py, double heading) {
    int best = 0; double bd = 0;
    for (int l = 0; l < sc->cfg.n_lanes; ++l) {
        double d = lane_distance_with_heading(&sc->lanes[l], px, py, heading);
        if (l == 0 || d < bd) { bd = d; best = l; }
    }
    return best;
}
int orc_closest_lane(const orc_scene* sc, double px, double py, double h) { return closest_lane(sc, px, py, h); }

/* RoadNetwork.next_lane_given_next_road road.py:138-157; next_id = -1 means None */
static int next_lane_given_next_road(const orc_scene* sc, int road, int id, int next_road, int next_id,
                                     double px, double py, double* dist) {
    const ttrl_road* nr = &sc->roads[next_road];
    if (sc->roads[road].n_lanes == nr->n_lanes) {
        if (next_id < 0) next_id = id;
    } else {
        double bd = 0; int b = 0;
        for (int l = 0; l < nr->n_lanes; ++l) {
            double d = lane_distance(&sc->lanes[nr->first_lane + l], px, py);
            if (l == 0 || d < bd) { bd = d; b = l; }
        }
        next_id = b;
    }
    *dist = lane_distance(&sc->lanes[nr->first_lane + next_id], px, py);
    return next_id;
}

static void route_pop(veh_t* v) {
    for (int k = 1; k < v->route_len; ++k) { v->route_road[k - 1] = v->route_road[k]; v->route_lane[k - 1] = v->route_lane[k]; }
    v->route_len--;
}

/* RoadNetwork.next_lane road.py:73-136.  The route list is mutated in place (route.pop(0), :100). */
static int next_lane(const orc_scene* sc, veh_t* v, int cur, double px, double py) {
    const ttrl_lane* cl = &sc->lanes[cur];
    int road = cl->road, id = cl->lane_id;
    int to = sc->roads[road].to_node;
    int next_road = -1, next_id = -1;
    if (v->route_len > 0) {
        if (v->route_road[0] == road) route_pop(v);
        if (v->route_len > 0 && sc->roads[v->route_road[0]].from_node == to) {
            next_road = v->route_road[0];
            next_id = v->route_lane[0];
        }
    }
    double lon, lat, qx, qy;
    lane_local(cl, px, py, &lon, &lat);
    lane_position(cl, lon, 0, &qx, &qy);
    if (next_road < 0) {
        int a = sc->node_first[to], b = sc->node_first[to + 1];
        if (a == b) return cur; /* KeyError: _to has no outgoing road (:129-130) */
        double bd = 0; int br = -1, bid = 0;
        for (int k = a; k < b; ++k) {
            double d; int nid = next_lane_given_next_road(sc, road, id, sc->node_roads[k], -1, qx, qy, &d);
            if (k == a || d < bd) { bd = d; br = sc->node_roads[k]; bid = nid; }
        }
        next_road = br; next_id = bid;
    } else {
        double d; next_id = next_lane_given_next_road(sc, road, id, next_road, next_id, qx, qy, &d);
    }
    return sc->roads[next_road].first_lane + next_id;
}

/* RoadNetwork.side_lanes road.py:200-211: id-1 then id+1 */
static int side_lanes(const orc_scene* sc, int lane, int out[2]) {
    const ttrl_lane* l = &sc->lanes[lane];
    int n = 0;
    if (l->lane_id > 0) out[n++] = lane - 1;
    if (l->lane_id < sc->roads[l->road].n_lanes - 1) out[n++] = lane + 1;
    return n;
}

/* RoadNetwork.position_heading_along_route road.py:323-362 with lateral = 0.
 * route given as (road,lane_id|-1)[len]; `cur` is the vehicle's current lane. */
static void position_heading_along_route(const orc_scene* sc, const int* rroad, const int* rlane, int rlen,
                                         double lon, int cur, double* px, double* py, double* h) {
    int k = 0;
    int cur_id = sc->lanes[cur].lane_id; /* :345-350: cur id is always < its own road's lane count */
#define ROUTE_HEAD(k_) (sc->roads[rroad[k_]].first_lane + (rlane[k_] < 0 ? cur_id : rlane[k_]))
    int li = ROUTE_HEAD(k);
    while (rlen - k > 1 && lon > sc->lanes[li].length) {
        lon -= sc->lanes[li].length;
        k++;
        li = ROUTE_HEAD(k);
    }
#undef ROUTE_HEAD
    lane_position(&sc->lanes[li], lon, 0, px, py);
    *h = lane_heading_at(&sc->lanes[li], lon);
}

/* ----------------------------------------------------------------------------------------------
 * vehicle/objects.py, vehicle/kinematics.py
 * -------------------------------------------------------------------------------------------- */
/* RoadObject.lane_distance_to objects.py:182-197 (lane defaults to self.lane) */
static double lane_distance_to(const orc_scene* sc, const veh_t* self, const veh_t* other) {
    const ttrl_lane* l = &sc->lanes[self->lane];
    double s0, r0, s1, r1;
    lane_local(l, other->x, other->y, &s1, &r1);
    lane_local(l, self->x, self->y, &s0, &r0);
    return s1 - s0;
}
/* RoadObject.on_road objects.py:199-202 */
static int on_road(const orc_scene* sc, const veh_t* v) {
    double s, r; lane_local(&sc->lanes[v->lane], v->x, v->y, &s, &r);
    return lane_on_lane_sr(&sc->lanes[v->lane], s, r, 0.0);
}

/* Road.neighbour_vehicles road.py:480-513 */
static void neighbour_vehicles(const orc_scene* sc, const env_t* e, int vi, int lane, int* front, int* rear) {
    const ttrl_lane* l = &sc->lanes[lane];
    double s, r0; lane_local(l, e->v[vi].x, e->v[vi].y, &s, &r0);
    double s_front = 0, s_rear = 0; int vf = -1, vr = -1;
    for (int j = 0; j < e->n; ++j) {
        if (j == vi) continue;
        double sv, lv; lane_local(l, e->v[j].x, e->v[j].y, &sv, &lv);
        if (!lane_on_lane_sr(l, sv, lv, 1.0)) continue;
        if (s <= sv && (vf < 0 || sv <= s_front)) { s_front = sv; vf = j; }
        if (sv < s && (vr < 0 || sv > s_rear)) { s_rear = sv; vr = j; }
    }
    *front = vf; *rear = vr;
}

/* ----------------------------------------------------------------------------------------------
 * vehicle/controller.py
 * -------------------------------------------------------------------------------------------- */
/* ControlledVehicle.follow_road controller.py:135-143 + AbstractLane.after_end lane.py:120-125 */
static void follow_road(const orc_scene* sc, veh_t* v) {
    const ttrl_lane* tl = &sc->lanes[v->target_lane];
    double s, r; lane_local(tl, v->x, v->y, &s, &r);
    if (s > tl->length - 5.0 / 2) v->target_lane = next_lane(sc, v, v->target_lane, v->x, v->y);
}
/* ControlledVehicle.steering_control controller.py:145-187 (constants :24-32) */
static double steering_control(const orc_scene* sc, const veh_t* v, int target_lane) {
    const double TAU_PURSUIT = 0.5 * 0.2, KP_LATERAL = 1 / 0.6, KP_HEADING = 1 / 0.2, LENGTH = 5.0;
    const double MAX_STEER = PI / 3;
    const ttrl_lane* tl = &sc->lanes[target_lane];
    double s, r; lane_local(tl, v->x, v->y, &s, &r);
    double lane_next = s + v->speed * TAU_PURSUIT;
    double lane_future_heading = lane_heading_at(tl, lane_next);
    if (sc->cfg.vehicle_model == TTRL_VEHICLE_LINEAR && !(v->flags & TTRL_FL_MDP)) {
        /* LinearVehicle.steering_control = np.dot(STEERING_PARAMETERS, steering_features) behavior.py:466-500 */
        double f0 = wrap_to_pi(lane_future_heading - v->heading) * LENGTH / not_zero(v->speed);
        double f1 = -r * LENGTH / (not_zero(v->speed) * not_zero(v->speed)); /* not_zero(speed) ** 2 */
        return v->lin[TTRL_LIN_STEER0] * f0 + v->lin[TTRL_LIN_STEER1] * f1;
    }
    double lateral_speed_command = -KP_LATERAL * r;
    double heading_command = asin(clipd(lateral_speed_command / not_zero(v->speed), -1, 1));
    double heading_ref = lane_future_heading + clipd(heading_command, -PI / 4, PI / 4);
    double heading_rate_command = KP_HEADING * wrap_to_pi(heading_ref - v->heading);
    double slip_angle = asin(clipd(LENGTH / 2 / not_zero(v->speed) * heading_rate_command, -1, 1));
    double steering_angle = atan(2 * tan(slip_angle));
    return clipd(steering_angle, -MAX_STEER, MAX_STEER);
}
double orc_steering_control(const orc_scene* sc, double x, double y, double h, double speed, int target_lane) {
    veh_t v; memset(&v, 0, sizeof v); v.x = x; v.y = y; v.heading = h; v.speed = speed;
    return steering_control(sc, &v, target_lane);
}
/* MDPVehicle.speed_to_index controller.py:326-344 (np.round = round-half-even = rint) */
static int speed_to_index(const orc_scene* sc, double speed) {
    int n = sc->cfg.n_target_speeds;
    const double* ts = sc->cfg.target_speeds;
    double x = (speed - ts[0]) / (ts[n - 1] - ts[0]);
    return (int)clipd(rint(x * (n - 1)), 0, n - 1);
}
int orc_speed_to_index(const orc_scene* sc, double speed) { return speed_to_index(sc, speed); }

enum { A_NONE = 0, A_IDLE, A_LANE_LEFT, A_LANE_RIGHT, A_FASTER, A_SLOWER };
/* DiscreteMetaAction tables envs/common/action.py:204-211 */
static int decode_action(const orc_scene* sc, int a) {
    static const int all[5] = {A_LANE_LEFT, A_IDLE, A_LANE_RIGHT, A_FASTER, A_SLOWER};
    static const int lon[3] = {A_SLOWER, A_IDLE, A_FASTER};
    static const int lat[3] = {A_LANE_LEFT, A_IDLE, A_LANE_RIGHT};
    if (a < 0) return A_NONE;
    if (sc->cfg.action_mode == TTRL_ACT_ALL) return all[a];
    if (sc->cfg.action_mode == TTRL_ACT_LONGI) return lon[a];
    return lat[a];
}
/* ControlledVehicle.act controller.py:89-133 (FASTER/SLOWER never reach it from MDPVehicle) */
static void controlled_act(const orc_scene* sc, veh_t* v, int action) {
    const double KP_A = 1 / 0.6, MAX_STEER = PI / 3;
    follow_road(sc, v);
    if (action == A_LANE_RIGHT || action == A_LANE_LEFT) {
        const ttrl_lane* tl = &sc->lanes[v->target_lane];
        int n = sc->roads[tl->road].n_lanes;
        int id = tl->lane_id + (action == A_LANE_RIGHT ? 1 : -1);
        if (id < 0) id = 0;
        if (id > n - 1) id = n - 1;
        int cand = sc->roads[tl->road].first_lane + id;
        double s, r; lane_local(&sc->lanes[cand], v->x, v->y, &s, &r);
        if (lane_reachable_sr(&sc->lanes[cand], s, r)) v->target_lane = cand;
    }
    double steering = steering_control(sc, v, v->target_lane);
    double accel = KP_A * (v->target_speed - v->speed); /* speed_control :189-198 */
    v->steering = clipd(steering, -MAX_STEER, MAX_STEER);
    v->accel = accel;
}
/* MDPVehicle.act controller.py:295-315 */
static void mdp_act(const orc_scene* sc, veh_t* v, int action) {
    if (action == A_FASTER) v->speed_index = speed_to_index(sc, v->speed) + 1;
    else if (action == A_SLOWER) v->speed_index = speed_to_index(sc, v->speed) - 1;
    else { controlled_act(sc, v, action); return; }
    int n = sc->cfg.n_target_speeds;
    if (v->speed_index < 0) v->speed_index = 0;
    if (v->speed_index > n - 1) v->speed_index = n - 1;
    v->target_speed = sc->cfg.target_speeds[v->speed_index];
    controlled_act(sc, v, A_NONE);
}

/* ----------------------------------------------------------------------------------------------
 * vehicle/behavior.py -- IDMVehicle
 * -------------------------------------------------------------------------------------------- */
/* IDMVehicle.desired_gap behavior.py:192-217 (projected=True) */
static double desired_gap(const orc_scene* sc, const veh_t* ego, const veh_t* front) {
    double d0 = sc->cfg.distance_wanted, tau = sc->cfg.time_wanted;
    double ab = -sc->cfg.comfort_acc_max * sc->cfg.comfort_acc_min;
    double ec = cos(ego->heading), es = sin(ego->heading);
    double fc = cos(front->heading), fs = sin(front->heading);
    double dvx = ego->speed * ec - front->speed * fc, dvy = ego->speed * es - front->speed * fs;
    double dv = dvx * ec + dvy * es;
    return d0 + ego->speed * tau + ego->speed * dv / (2 * sqrt(ab));
}
/* IDMVehicle.acceleration behavior.py:150-190.  `self_delta` is SELF's DELTA even when ego is another vehicle. */
static double idm_acceleration(const orc_scene* sc, double self_delta, const veh_t* ego, const veh_t* front) {
    if (!ego) return 0;
    double ego_target_speed = ego->target_speed;
    ego_target_speed = clipd(ego_target_speed, 0, sc->lanes[ego->lane].speed_limit);
    double acc = sc->cfg.comfort_acc_max *
                 (1 - pow(fmax(ego->speed, 0) / fabs(not_zero(ego_target_speed)), self_delta));
    if (front) {
        double d = lane_distance_to(sc, ego, front);
        double q = desired_gap(sc, ego, front) / not_zero(d);
        acc -= sc->cfg.comfort_acc_max * (q * q); /* np.power(x, 2) */
    }
    return acc;
}
/* LinearVehicle.acceleration behavior.py:416-464: np.dot(self.ACCELERATION_PARAMETERS, acceleration_features(ego, front)) */
static double linear_acceleration(const orc_scene* sc, const veh_t* self, const veh_t* ego, const veh_t* front) {
    double vt = 0, dv = 0, dp = 0;
    if (ego) {
        vt = ego->target_speed - ego->speed;
        double d_safe = sc->cfg.distance_wanted + fmax(ego->speed, 0) * sc->cfg.time_wanted;
        if (front) {
            double d = lane_distance_to(sc, ego, front);
            dv = fmin(front->speed - ego->speed, 0);
            dp = fmin(d - d_safe, 0);
        }
    }
    return self->lin[TTRL_LIN_ACC0] * vt + self->lin[TTRL_LIN_ACC1] * dv + self->lin[TTRL_LIN_ACC2] * dp;
}
/* self.acceleration(ego_vehicle, front_vehicle): the class of `self` decides (other_vehicles_type) */
static double acceleration(const orc_scene* sc, const veh_t* self, const veh_t* ego, const veh_t* front) {
    if (sc->cfg.vehicle_model == TTRL_VEHICLE_LINEAR) return linear_acceleration(sc, self, ego, front);
    return idm_acceleration(sc, self->delta, ego, front);
}
double orc_idm_acceleration(const orc_scene* sc, double self_delta, const double* ego6, const double* front6) {
    /* ego6/front6 = x, y, heading, speed, target_speed, lane */
    veh_t e, f; memset(&e, 0, sizeof e); memset(&f, 0, sizeof f);
    e.x = ego6[0]; e.y = ego6[1]; e.heading = ego6[2]; e.speed = ego6[3]; e.target_speed = ego6[4]; e.lane = (int)ego6[5];
    if (front6) { f.x = front6[0]; f.y = front6[1]; f.heading = front6[2]; f.speed = front6[3]; f.target_speed = front6[4]; f.lane = (int)front6[5]; }
    return idm_acceleration(sc, self_delta, &e, front6 ? &f : 0);
}

/* IDMVehicle.mobil behavior.py:265-324 */
static int mobil(const orc_scene* sc, const env_t* e, int vi, int cand) {
    const veh_t* self = &e->v[vi];
    int np_, nf_;
    neighbour_vehicles(sc, e, vi, cand, &np_, &nf_);
    const veh_t* new_preceding = np_ >= 0 ? &e->v[np_] : 0;
    const veh_t* new_following = nf_ >= 0 ? &e->v[nf_] : 0;
    double new_following_a = acceleration(sc, self, new_following, new_preceding);
    double new_following_pred_a = acceleration(sc, self, new_following, self);
    if (new_following_pred_a < -sc->cfg.lane_change_max_braking_imposed) return 0;
    int op_, of_;
    neighbour_vehicles(sc, e, vi, self->lane, &op_, &of_);
    const veh_t* old_preceding = op_ >= 0 ? &e->v[op_] : 0;
    const veh_t* old_following = of_ >= 0 ? &e->v[of_] : 0;
    double self_pred_a = acceleration(sc, self, self, new_preceding);
    if (self->route_len > 0 && self->route_lane[0] >= 0) {
        int want = self->route_lane[0] - sc->lanes[self->target_lane].lane_id;
        int dir = sc->lanes[cand].lane_id - sc->lanes[self->target_lane].lane_id;
        int sw = (want > 0) - (want < 0), sd = (dir > 0) - (dir < 0);
        if (sd != sw) return 0;
        else if (self_pred_a < -sc->cfg.lane_change_max_braking_imposed) return 0;
    } else {
        double self_a = acceleration(sc, self, self, old_preceding);
        double old_following_a = acceleration(sc, self, old_following, self);
        double old_following_pred_a = acceleration(sc, self, old_following, old_preceding);
        double jerk = self_pred_a - self_a +
                      sc->cfg.politeness * (new_following_pred_a - new_following_a + old_following_pred_a - old_following_a);
        if (jerk < sc->cfg.lane_change_min_acc_gain) return 0;
    }
    return 1;
}

/* IDMVehicle.change_lane_policy behavior.py:219-263 */
static void change_lane_policy(const orc_scene* sc, env_t* e, int vi) {
    veh_t* self = &e->v[vi];
    if (self->lane != self->target_lane) {
        if (sc->lanes[self->lane].road == sc->lanes[self->target_lane].road) {
            for (int j = 0; j < e->n; ++j) {
                const veh_t* v = &e->v[j];
                if (j != vi && v->lane != self->target_lane && v->target_lane == self->target_lane) {
                    double d = lane_distance_to(sc, self, v);
                    double d_star = desired_gap(sc, self, v);
                    if (0 < d && d < d_star) { self->target_lane = self->lane; break; }
                }
            }
        }
        return;
    }
    if (!(sc->cfg.lane_change_delay < self->timer)) return; /* utils.do_every utils.py:25-26 */
    self->timer = 0;
    int sl[2]; int n = side_lanes(sc, self->lane, sl);
    for (int k = 0; k < n; ++k) {
        double s, r; lane_local(&sc->lanes[sl[k]], self->x, self->y, &s, &r);
        if (!lane_reachable_sr(&sc->lanes[sl[k]], s, r)) continue;
        if (fabs(self->speed) < 1) continue;
        if (mobil(sc, e, vi, sl[k])) self->target_lane = sl[k];
    }
}

/* IDMVehicle.act behavior.py:93-137 */
static void idm_act(const orc_scene* sc, env_t* e, int vi) {
    veh_t* self = &e->v[vi];
    const double MAX_STEER = PI / 3;
    if (self->flags & TTRL_FL_CRASHED) return;
    follow_road(sc, self);
    change_lane_policy(sc, e, vi);
    double steering = clipd(steering_control(sc, self, self->target_lane), -MAX_STEER, MAX_STEER);
    int f, r;
    neighbour_vehicles(sc, e, vi, self->lane, &f, &r);
    double acc = acceleration(sc, self, self, f >= 0 ? &e->v[f] : 0);
    if (self->lane != self->target_lane) {
        neighbour_vehicles(sc, e, vi, self->target_lane, &f, &r);
        double tacc = acceleration(sc, self, self, f >= 0 ? &e->v[f] : 0);
        acc = fmin(acc, tacc);
    }
    acc = clipd(acc, -sc->cfg.acc_max, sc->cfg.acc_max);
    self->steering = steering;
    self->accel = acc;
}

/* Vehicle.step kinematics.py:130-153 + clip_actions :155-168 + on_state_update :170-177;
 * IDMVehicle.step behavior.py:139-148 adds timer += dt first. */
static void vehicle_step(const orc_scene* sc, veh_t* v, double dt) {
    if (!(v->flags & TTRL_FL_MDP)) v->timer += dt;
    if (v->flags & TTRL_FL_CRASHED) { v->steering = 0; v->accel = -1.0 * v->speed; }
    if (v->speed > 40.0) v->accel = fmin(v->accel, 1.0 * (40.0 - v->speed));
    else if (v->speed < -40.0) v->accel = fmax(v->accel, 1.0 * (-40.0 - v->speed));
    double delta_f = v->steering;
    double beta = atan(1.0 / 2 * tan(delta_f));
    double vx = v->speed * cos(v->heading + beta), vy = v->speed * sin(v->heading + beta);
    v->x += vx * dt; v->y += vy * dt;
    if (v->flags & TTRL_FL_HAS_IMPACT) {
        v->x += v->impact_x; v->y += v->impact_y;
        v->flags |= TTRL_FL_CRASHED;
        v->flags &= ~TTRL_FL_HAS_IMPACT;
        v->impact_x = 0; v->impact_y = 0;
    }
    v->heading += v->speed * sin(beta) / (5.0 / 2) * dt;
    v->speed += v->accel * dt;
    v->lane = closest_lane(sc, v->x, v->y, v->heading);
}

/* ----------------------------------------------------------------------------------------------
 * collisions: objects.py:91-137, :168-180; utils.py:175-239
 * -------------------------------------------------------------------------------------------- */
/* RoadObject.polygon objects.py:168-180: 5 points (closed), LENGTH 5 WIDTH 2 */
static void polygon(const veh_t* v, double p[5][2]) {
    static const double pts[4][2] = {{-2.5, -1.0}, {-2.5, +1.0}, {+2.5, +1.0}, {+2.5, -1.0}};
    double c = cos(v->heading), s = sin(v->heading);
    for (int k = 0; k < 4; ++k) {
        p[k][0] = (c * pts[k][0] + (-s) * pts[k][1]) + v->x;
        p[k][1] = (s * pts[k][0] + c * pts[k][1]) + v->y;
    }
    p[4][0] = p[0][0]; p[4][1] = p[0][1];
}
/* utils.project_polygon utils.py:175-183 */
static void project_polygon(double p[5][2], double nx, double ny, double* mn, double* mx) {
    for (int k = 0; k < 5; ++k) {
        double pr = p[k][0] * nx + p[k][1] * ny;
        if (k == 0 || pr < *mn) *mn = pr;
        if (k == 0 || pr > *mx) *mx = pr;
    }
}
/* utils.interval_distance utils.py:186-191 */
static double interval_distance(double min_a, double max_a, double min_b, double max_b) {
    return min_a < min_b ? min_b - max_a : min_a - max_b;
}
/* utils.are_polygons_intersecting utils.py:194-239 */
static void are_polygons_intersecting(double a[5][2], double b[5][2], double dax, double day, double dbx, double dby,
                                      int* intersecting, int* will_intersect, double* tx, double* ty) {
    int inter = 1, will = 1;
    double min_distance = INFINITY, axx = 0, axy = 0;
    for (int poly = 0; poly < 2; ++poly) {
        double (*pg)[2] = poly == 0 ? a : b;
        for (int k = 0; k < 4; ++k) {
            double nx = -pg[k + 1][1] + pg[k][1], ny = pg[k + 1][0] - pg[k][0];
            double nn = sqrt(nx * nx + ny * ny);
            nx /= nn; ny /= nn;
            double min_a, max_a, min_b, max_b;
            project_polygon(a, nx, ny, &min_a, &max_a);
            project_polygon(b, nx, ny, &min_b, &max_b);
            if (interval_distance(min_a, max_a, min_b, max_b) > 0) inter = 0;
            double vp = nx * (dax - dbx) + ny * (day - dby);
            if (vp < 0) min_a += vp; else max_a += vp;
            double distance = interval_distance(min_a, max_a, min_b, max_b);
            if (distance > 0) will = 0;
            if (!inter && !will) break;
            if (fabs(distance) < min_distance) {
                min_distance = fabs(distance);
                double cax = (((a[0][0] + a[1][0]) + a[2][0]) + a[3][0]) / 4, cay = (((a[0][1] + a[1][1]) + a[2][1]) + a[3][1]) / 4;
                double cbx = (((b[0][0] + b[1][0]) + b[2][0]) + b[3][0]) / 4, cby = (((b[0][1] + b[1][1]) + b[2][1]) + b[3][1]) / 4;
                double ddx = cax - cbx, ddy = cay - cby;
                if (ddx * nx + ddy * ny > 0) { axx = nx; axy = ny; } else { axx = -nx; axy = -ny; }
            }
        }
    }
    *intersecting = inter; *will_intersect = will;
    if (will) { *tx = min_distance * axx; *ty = min_distance * axy; } else { *tx = 0; *ty = 0; }
}
/* RoadObject.handle_collisions objects.py:91-119 + _is_colliding :121-137 (vehicles only: all solid/collidable) */
static void handle_collisions(veh_t* self, veh_t* other, double dt) {
    const double diag = sqrt(5.0 * 5.0 + 2.0 * 2.0);
    double dx = other->x - self->x, dy = other->y - self->y;
    if (sqrt(dx * dx + dy * dy) > (diag + diag) / 2 + self->speed * dt) return;
    double a[5][2], b[5][2];
    polygon(self, a); polygon(other, b);
    double sc_ = cos(self->heading), ss = sin(self->heading), oc = cos(other->heading), os = sin(other->heading);
    int inter, will; double tx, ty;
    are_polygons_intersecting(a, b, self->speed * sc_ * dt, self->speed * ss * dt, other->speed * oc * dt,
                              other->speed * os * dt, &inter, &will, &tx, &ty);
    if (will) {
        self->impact_x = tx / 2; self->impact_y = ty / 2; self->flags |= TTRL_FL_HAS_IMPACT;
        other->impact_x = -tx / 2; other->impact_y = -ty / 2; other->flags |= TTRL_FL_HAS_IMPACT;
    }
    if (inter) { self->flags |= TTRL_FL_CRASHED; other->flags |= TTRL_FL_CRASHED; }
}
void orc_polygons_intersecting(const double* va /* x,y,h,speed */, const double* vb, double dt, double* out4) {
    veh_t a, b; memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
    a.x = va[0]; a.y = va[1]; a.heading = va[2]; a.speed = va[3];
    b.x = vb[0]; b.y = vb[1]; b.heading = vb[2]; b.speed = vb[3];
    double pa[5][2], pb[5][2]; polygon(&a, pa); polygon(&b, pb);
    int inter, will; double tx, ty;
    are_polygons_intersecting(pa, pb, a.speed * cos(a.heading) * dt, a.speed * sin(a.heading) * dt,
                              b.speed * cos(b.heading) * dt, b.speed * sin(b.heading) * dt, &inter, &will, &tx, &ty);
    out4[0] = inter; out4[1] = will; out4[2] = tx; out4[3] = ty;
}

/* ----------------------------------------------------------------------------------------------
 * road/regulation.py
 * -------------------------------------------------------------------------------------------- */
/* utils.point_in_rotated_rectangle utils.py:75-91 (+ point_in_rectangle :61-72), literal rotation R(+angle) */
static int point_in_rotated_rectangle(double px, double py, double cx, double cy, double length, double width, double angle) {
    double c = cos(angle), s = sin(angle);
    double dx = px - cx, dy = py - cy;
    double rx = c * dx + (-s) * dy, ry = s * dx + c * dy;
    return -length / 2 <= rx && rx <= length / 2 && -width / 2 <= ry && ry <= width / 2;
}
/* utils.has_corner_inside utils.py:158-172 + rect_corners :126-155 (4 corners + centre + 4 midpoints) */
static int has_corner_inside(const double r1[5], const double r2[5]) {
    double hl = r1[2] / 2, hw = r1[3] / 2;
    double pts[9][2] = {{-hl - 0, 0 - hw}, {-hl + 0, 0 + hw}, {+hl + 0, 0 + hw}, {+hl - 0, 0 - hw}, {0, 0},
                        {-hl, -0.0}, {hl, 0}, {-0.0, -hw}, {0, hw}};
    double c = cos(r1[4]), s = sin(r1[4]);
    for (int k = 0; k < 9; ++k) {
        double px = (c * pts[k][0] + (-s) * pts[k][1]) + r1[0];
        double py = (s * pts[k][0] + c * pts[k][1]) + r1[1];
        if (point_in_rotated_rectangle(px, py, r2[0], r2[1], r2[2], r2[3], r2[4])) return 1;
    }
    return 0;
}
/* utils.rotated_rectangles_intersect utils.py:113-123; rect = (cx, cy, length, width, angle) */
static int rotated_rectangles_intersect(const double r1[5], const double r2[5]) {
    return has_corner_inside(r1, r2) || has_corner_inside(r2, r1);
}
int orc_rotated_rectangles_intersect(const double* r1, const double* r2) { return rotated_rectangles_intersect(r1, r2); }

/* ControlledVehicle.predict_trajectory_constant_speed controller.py:236-253 at t = 0.25..2.75 (regulation.py:87) */
static void predict_trajectory(const orc_scene* sc, const veh_t* v, double px[11], double py[11], double ph[11]) {
    double s, r; lane_local(&sc->lanes[v->lane], v->x, v->y, &s, &r);
    int rroad[TTRL_ROUTE_CAP], rlane[TTRL_ROUTE_CAP], rlen;
    if (v->route_len > 0) {
        rlen = v->route_len;
        for (int k = 0; k < rlen; ++k) { rroad[k] = v->route_road[k]; rlane[k] = v->route_lane[k]; }
    } else { rlen = 1; rroad[0] = sc->lanes[v->lane].road; rlane[0] = sc->lanes[v->lane].lane_id; }
    for (int k = 0; k < 11; ++k) {
        double t = 0.25 + k * 0.25; /* np.arange(0.25, 3, 0.25): start + k*step */
        position_heading_along_route(sc, rroad, rlane, rlen, s + v->speed * t, v->lane, &px[k], &py[k], &ph[k]);
    }
}
/* RegulatedRoad.is_conflict_possible regulation.py:80-103 */
static int is_conflict_possible(double x1[11], double y1[11], double h1[11], double x2[11], double y2[11], double h2[11]) {
    for (int k = 0; k < 11; ++k) {
        double dx = x2[k] - x1[k], dy = y2[k] - y1[k];
        if (sqrt(dx * dx + dy * dy) > 5.0) continue;
        double r1[5] = {x1[k], y1[k], 1.5 * 5.0, 0.9 * 2.0, h1[k]};
        double r2[5] = {x2[k], y2[k], 1.5 * 5.0, 0.9 * 2.0, h2[k]};
        if (rotated_rectangles_intersect(r1, r2)) return 1;
    }
    return 0;
}
/* RoadObject.front_distance_to objects.py:204-205 */
static double front_distance_to(const veh_t* a, const veh_t* b) {
    return cos(a->heading) * (b->x - a->x) + sin(a->heading) * (b->y - a->y);
}
/* RegulatedRoad.enforce_road_rules regulation.py:34-62 + respect_priorities :64-78 */
static void enforce_road_rules(const orc_scene* sc, env_t* e) {
    static _Thread_local double px[VMAX][11], py[VMAX][11], ph[VMAX][11];
    for (int i = 0; i < e->n; ++i) {
        veh_t* v = &e->v[i];
        if (v->flags & TTRL_FL_YIELDING) {
            if (v->yield_timer >= 0.0 * 2) { /* YIELD_DURATION * REGULATION_FREQUENCY */
                v->target_speed = sc->lanes[v->lane].speed_limit;
                v->flags &= ~TTRL_FL_YIELDING;
            } else v->yield_timer += 1;
        }
    }
    for (int i = 0; i < e->n; ++i) predict_trajectory(sc, &e->v[i], px[i], py[i], ph[i]);
    for (int i = 0; i < e->n - 1; ++i)
        for (int j = i + 1; j < e->n; ++j) {
            if (!is_conflict_possible(px[i], py[i], ph[i], px[j], py[j], ph[j])) continue;
            veh_t *v1 = &e->v[i], *v2 = &e->v[j], *y;
            int p1 = sc->lanes[v1->lane].priority, p2 = sc->lanes[v2->lane].priority;
            if (p1 > p2) y = v2;
            else if (p1 < p2) y = v1;
            else y = front_distance_to(v1, v2) > front_distance_to(v2, v1) ? v1 : v2;
            if (!(y->flags & TTRL_FL_MDP)) {
                y->target_speed = 0;
                y->flags |= TTRL_FL_YIELDING;
                y->yield_timer = 0;
            }
        }
}

/* ----------------------------------------------------------------------------------------------
 * one simulation sub-step: AbstractEnv._simulate body abstract.py:257-273; Road.act road.py:461-464;
 * RegulatedRoad.step regulation.py:28-32; Road.step road.py:466-478
 * -------------------------------------------------------------------------------------------- */
static int n_agents(const orc_scene* sc) { return sc->cfg.controlled_vehicles < 1 ? 1 : sc->cfg.controlled_vehicles; }
static void env_substep(const orc_scene* sc, env_t* e, const int32_t* actions /* K raw ids or NULL */) {
    int F = (int)floor(sc->cfg.simulation_frequency / sc->cfg.policy_frequency);
    double dt = 1 / sc->cfg.simulation_frequency;
    if (actions && e->steps % F == 0) {
        /* DiscreteMetaAction.act action.py:259-260 on every controlled vehicle, in order (MultiAgentAction.act :320-323) */
        for (int k = 0; k < n_agents(sc); ++k)
            if (actions[k] >= 0) mdp_act(sc, &e->v[e->egos[k]], decode_action(sc, actions[k]));
    }
    for (int i = 0; i < e->n; ++i) {
        if (e->v[i].flags & TTRL_FL_MDP) mdp_act(sc, &e->v[i], A_NONE);
        else idm_act(sc, e, i);
    }
    if (sc->cfg.regulated) {
        e->road_steps += 1;
        if (e->road_steps % (int)(1 / dt / 2) == 0) enforce_road_rules(sc, e);
    }
    for (int i = 0; i < e->n; ++i) vehicle_step(sc, &e->v[i], dt);
    for (int i = 0; i < e->n; ++i)
        for (int j = i + 1; j < e->n; ++j) handle_collisions(&e->v[i], &e->v[j], dt);
    e->steps += 1;
}

/* ----------------------------------------------------------------------------------------------
 * observations: envs/common/observation.py
 * -------------------------------------------------------------------------------------------- */
/* Vehicle.to_dict kinematics.py:237-261 restricted to the hot features */
static double feature_of(const veh_t* v, int f) {
    switch (f) {
        case TTRL_F_PRESENCE: return 1;
        case TTRL_F_X: return v->x;
        case TTRL_F_Y: return v->y;
        case TTRL_F_VX: return v->speed * cos(v->heading);
        case TTRL_F_VY: return v->speed * sin(v->heading);
        case TTRL_F_COS_H: return cos(v->heading);
        case TTRL_F_SIN_H: return sin(v->heading);
        case TTRL_F_HEADING: return v->heading;
    }
    return 0;
}
static int is_relative_feature(int f) { return f == TTRL_F_X || f == TTRL_F_Y || f == TTRL_F_VX || f == TTRL_F_VY; }

/* KinematicObservation.observe observation.py:233-275 (+ normalize_obs :206-231, Road.close_objects_to road.py:418-447).
 * "shuffled" order is applied by the caller (host RNG); rows here are in sorted / list order. */
static void observe_kinematics(const orc_scene* sc, const env_t* e, int observer, float* obs) {
    const ttrl_config* c = &sc->cfg;
    int V = c->obs_vehicles, Fe = c->n_features;
    const veh_t* ego = &e->v[observer];
    int idx[VMAX]; double key[VMAX]; int m = 0;
    for (int j = 0; j < e->n; ++j) {
        const veh_t* v = &e->v[j];
        double dx = v->x - ego->x, dy = v->y - ego->y;
        if (!(sqrt(dx * dx + dy * dy) < 200.0)) continue; /* PERCEPTION_DISTANCE abstract.py:41 */
        if (j == observer) continue;
        double d = lane_distance_to(sc, ego, v);
        if (!(c->see_behind || -2 * 5.0 < d)) continue;
        idx[m] = j; key[m] = fabs(d); m++;
    }
    if (c->order == TTRL_ORDER_SORTED) { /* stable insertion sort == Python sorted() */
        for (int a = 1; a < m; ++a) {
            int ia = idx[a]; double ka = key[a]; int b = a - 1;
            while (b >= 0 && key[b] > ka) { idx[b + 1] = idx[b]; key[b + 1] = key[b]; b--; }
            idx[b + 1] = ia; key[b + 1] = ka;
        }
    }
    if (m > V - 1) m = V - 1;
    for (int row = 0; row < V; ++row) {
        for (int k = 0; k < Fe; ++k) {
            double val = 0;
            if (row == 0 || row - 1 < m) {
                const veh_t* v = row == 0 ? ego : &e->v[idx[row - 1]];
                int f = c->features[k];
                val = feature_of(v, f);
                if (row > 0 && !c->absolute && is_relative_feature(f)) val -= feature_of(ego, f);
                if (c->normalize && c->has_range[k]) {
                    val = lmap(val, c->range_lo[k], c->range_hi[k], -1, 1);
                    if (c->clip) val = clipd(val, -1, 1);
                }
            }
            obs[row * Fe + k] = (float)val;
        }
    }
}

/* OccupancyGridObservation.observe observation.py:353-412 (+ normalize :336-351, pos_to_index :414-434,
 * fill_road_layer_by_lanes :453-483).  Relative (absolute=False) only, like the reference (:357-358). */
static void observe_grid(const orc_scene* sc, const env_t* e, int observer, float* obs) {
    const ttrl_config* c = &sc->cfg;
    int W = c->grid_w, H = c->grid_h, Fe = c->n_features;
    const veh_t* ego = &e->v[observer];
    double* grid = (double*)malloc(sizeof(double) * Fe * W * H);
    for (int k = 0; k < Fe * W * H; ++k) grid[k] = NAN;
    double ca = cos(ego->heading), sa = sin(ego->heading);
    for (int layer = 0; layer < Fe; ++layer) {
        int f = c->features[layer];
        if (f == TTRL_F_ON_ROAD) {
            double spacing = fmin(c->grid_step[0], c->grid_step[1]);
            for (int li = 0; li < c->n_lanes; ++li) {
                const ttrl_lane* l = &sc->lanes[li];
                double origin, r0; lane_local(l, ego->x, ego->y, &origin, &r0);
                double start = origin - 100, stop = origin + 100;
                int n = (int)ceil((stop - start) / spacing); /* np.arange length */
                for (int w = 0; w < n; ++w) {
                    double wp = clipd(start + w * spacing, 0, l->length);
                    double px, py; lane_position(l, wp, 0, &px, &py);
                    px -= ego->x; py -= ego->y; /* pos_to_index relative=False */
                    if (c->align_to_vehicle_axes) { double qx = ca * px + sa * py, qy = -sa * px + ca * py; px = qx; py = qy; }
                    int ci = (int)floor((px - c->grid_min[0]) / c->grid_step[0]);
                    int cj = (int)floor((py - c->grid_min[1]) / c->grid_step[1]);
                    if (0 <= ci && ci < W && 0 <= cj && cj < H) grid[(layer * W + ci) * H + cj] = 1;
                }
            }
            continue;
        }
        for (int j = e->n - 1; j >= 0; --j) { /* df[::-1]: reverse list order, earliest vehicle wins */
            const veh_t* v = &e->v[j];
            /* to_dict(origin=observer): x, y, vx, vy relative (kinematics.py:257-260) -- the observer itself too */
            double x = v->x - ego->x, y = v->y - ego->y;
            /* normalize() then recover (observation.py:348-392): lmap round trip when x / y in features_range */
            if (c->grid_has_xrange) { x = lmap(x, c->grid_xrange[0], c->grid_xrange[1], -1, 1); x = lmap(x, -1, 1, c->grid_xrange[0], c->grid_xrange[1]); }
            if (c->grid_has_yrange) { y = lmap(y, c->grid_yrange[0], c->grid_yrange[1], -1, 1); y = lmap(y, -1, 1, c->grid_yrange[0], c->grid_yrange[1]); }
            double px = x, py = y;
            if (c->align_to_vehicle_axes) { double qx = ca * px + sa * py, qy = -sa * px + ca * py; px = qx; py = qy; }
            int ci = (int)floor((px - c->grid_min[0]) / c->grid_step[0]);
            int cj = (int)floor((py - c->grid_min[1]) / c->grid_step[1]);
            if (!(0 <= ci && ci < W && 0 <= cj && cj < H)) continue;
            double val = feature_of(v, f);
            if (is_relative_feature(f)) val -= feature_of(ego, f);
            if (c->has_range[layer]) val = lmap(val, c->range_lo[layer], c->range_hi[layer], -1, 1);
            grid[(layer * W + ci) * H + cj] = val;
        }
    }
    for (int k = 0; k < Fe * W * H; ++k) {
        double v = grid[k];
        if (c->clip && !isnan(v)) v = clipd(v, -1, 1);
        if (isnan(v)) v = 0;
        obs[k] = (float)v;
    }
    free(grid);
}
/* RoadNetwork.is_connected_road road.py:231-276 (same_lane = False: roads only).  route = (road)[rlen] from position k. */
static int is_connected_road(const orc_scene* sc, int r1, int r2, const int* rroad, int rlen, int depth) {
    /* is_same_road(lane 2, lane 1) or is_leading_to_road(lane 2, lane 1) */
    if (r2 == r1 || sc->roads[r2].to_node == sc->roads[r1].from_node) return 1;
    if (depth > 0) {
        if (rlen > 0 && rroad[0] == r1) return is_connected_road(sc, r1, r2, rroad + 1, rlen - 1, depth);
        else if (rlen > 0 && sc->roads[rroad[0]].from_node == sc->roads[r1].to_node)
            return is_connected_road(sc, rroad[0], r2, rroad + 1, rlen - 1, depth - 1);
        else {
            int to = sc->roads[r1].to_node, any = 0;
            for (int k = sc->node_first[to]; k < sc->node_first[to + 1]; ++k)
                if (is_connected_road(sc, sc->node_roads[k], r2, rroad, rlen, depth - 1)) any = 1;
            return any;
        }
    }
    return 0;
}
/* compute_ttc_grid finite_mdp.py:104-163 + TimeToCollisionObservation.observe observation.py:114-151 */
static void observe_ttc(const orc_scene* sc, const env_t* e, int observer, float* obs) {
    const ttrl_config* c = &sc->cfg;
    const veh_t* ego = &e->v[observer];
    int S = c->n_target_speeds, H = c->ttc_steps;
    int re = sc->lanes[ego->lane].road, L = sc->roads[re].n_lanes;
    double tq = 1 / c->policy_frequency;
    double* grid = (double*)calloc((size_t)S * L * H, sizeof(double));
    for (int si = 0; si < S; ++si) {
        double ego_speed = c->target_speeds[si];
        for (int j = 0; j < e->n; ++j) {
            const veh_t* other = &e->v[j];
            if (j == observer || ego_speed == other->speed) continue;
            double margin = 5.0 / 2 + 5.0 / 2;
            double ms[3] = {0, -margin, margin}, costs[3] = {1, 0.5, 0.5};
            for (int mi = 0; mi < 3; ++mi) {
                double distance = lane_distance_to(sc, ego, other) + ms[mi];
                double other_projected_speed = other->speed * (cos(other->heading) * cos(ego->heading) + sin(other->heading) * sin(ego->heading));
                double ttc = distance / not_zero(ego_speed - other_projected_speed);
                if (ttc < 0) continue;
                int ro = sc->lanes[other->lane].road;
                if (!is_connected_road(sc, re, ro, ego->route_road, ego->route_len < 0 ? 0 : ego->route_len, 3)) continue;
                int lane_lo, lane_hi;
                if (sc->roads[ro].n_lanes == L) { lane_lo = lane_hi = sc->lanes[other->lane].lane_id; }
                else { lane_lo = 0; lane_hi = L - 1; }
                double q = ttc / tq;
                if (!(q < (double)H)) continue; /* int(q) >= H: outside the grid (also keeps the casts defined) */
                int times[2] = {(int)q, (int)ceil(q)};
                for (int w = 0; w < 2; ++w) {
                    int time = times[w];
                    if (!(0 <= time && time < H)) continue;
                    for (int l = lane_lo; l <= lane_hi; ++l) {
                        double* g = &grid[((size_t)si * L + l) * H + time];
                        if (costs[mi] > *g) *g = costs[mi];
                    }
                }
            }
        }
    }
    /* padding with ones across lanes, repeated first / last rows across speeds, 3 x 3 window (observation.py:137-151) */
    int lid = sc->lanes[ego->lane].lane_id;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int t = 0; t < H; ++t) {
                int so = ego->speed_index - 1 + a;
                if (so < 0) so = 0;
                if (so > S - 1) so = S - 1;
                int lo = lid - 1 + b;
                double v = (lo >= 0 && lo < L) ? grid[((size_t)so * L + lo) * H + t] : 1.0;
                obs[(a * 3 + b) * H + t] = (float)v;
            }
    free(grid);
}
static int obs_single(const orc_scene* sc) {
    const ttrl_config* c = &sc->cfg;
    if (c->obs_type == TTRL_OBS_TTC) return 9 * c->ttc_steps;
    return c->obs_type == TTRL_OBS_GRID ? c->n_features * c->grid_w * c->grid_h : c->obs_vehicles * c->n_features;
}
/* observation_type.observe(); MultiAgentObservation.observe observation.py:602-603: one block per controlled vehicle */
static void observe(const orc_scene* sc, const env_t* e, float* obs) {
    for (int k = 0; k < n_agents(sc); ++k) {
        if (sc->cfg.obs_type == TTRL_OBS_TTC) observe_ttc(sc, e, e->egos[k], obs + (size_t)k * obs_single(sc));
        else if (sc->cfg.obs_type == TTRL_OBS_GRID) observe_grid(sc, e, e->egos[k], obs + (size_t)k * obs_single(sc));
        else observe_kinematics(sc, e, e->egos[k], obs + (size_t)k * obs_single(sc));
    }
}
static int obs_size(const orc_scene* sc) { return n_agents(sc) * obs_single(sc); }

/* ----------------------------------------------------------------------------------------------
 * reward / termination
 * -------------------------------------------------------------------------------------------- */
/* IntersectionEnv.has_arrived intersection_env.py:364-369 */
static int has_arrived(const orc_scene* sc, const veh_t* v) {
    const ttrl_lane* l = &sc->lanes[v->lane];
    if (!l->is_exit) return 0;
    double s, r; lane_local(l, v->x, v->y, &s, &r);
    return s >= 25;
}
/* IntersectionEnv._agent_reward(s) intersection_env.py:78-104; highway: u_turn_env.py:39-71 template */
static double agent_reward(const orc_scene* sc, const veh_t* v, int raw_action) {
    const ttrl_config* c = &sc->cfg;
    double crashed = (v->flags & TTRL_FL_CRASHED) ? 1 : 0;
    if (c->reward_type == TTRL_REWARD_ROUNDABOUT) { /* RoundaboutEnv._reward / _rewards roundabout_env.py:43-64 */
        double hs = (double)v->speed_index / c->speed_index_den; /* MDPVehicle.get_speed_index / (DEFAULT_TARGET_SPEEDS.size - 1) */
        double lc = (raw_action == 0 || raw_action == 2) ? 1 : 0; /* action in [0, 2] */
        double onr0 = on_road(sc, v);
        double reward = 0 + c->collision_reward * crashed + c->high_speed_reward * hs + c->lane_change_reward * lc + 0 * onr0;
        if (c->normalize_reward) reward = lmap(reward, c->collision_reward, c->high_speed_reward, 0, 1);
        reward *= onr0;
        return reward;
    }
    double scaled = lmap(v->speed, c->reward_speed_lo, c->reward_speed_hi, 0, 1);
    double hs = clipd(scaled, 0, 1);
    double onr = on_road(sc, v);
    if (c->reward_type == TTRL_REWARD_INTERSECTION) {
        double arrived = has_arrived(sc, v);
        double reward = 0 + c->collision_reward * crashed + c->high_speed_reward * hs + c->arrived_reward * arrived + 0 * onr;
        reward = arrived ? c->arrived_reward : reward;
        reward *= onr;
        if (c->normalize_reward) reward = lmap(reward, c->collision_reward, c->arrived_reward, 0, 1);
        return reward;
    } else {
        const ttrl_lane* l = &sc->lanes[v->lane];
        int nl = sc->roads[l->road].n_lanes;
        double lane_term = (double)l->lane_id / (double)(nl - 1 > 1 ? nl - 1 : 1);
        double reward = 0 + c->collision_reward * crashed + c->lane_reward * lane_term + c->high_speed_reward * hs + 0 * onr;
        if (c->normalize_reward) reward = lmap(reward, c->collision_reward, c->high_speed_reward + c->lane_reward, 0, 1);
        reward *= onr;
        return reward;
    }
}
/* _is_terminated intersection_env.py:106-111 / u_turn_env.py:73-74 (single controlled vehicle) */
static int is_terminated(const orc_scene* sc, const env_t* e) {
    const veh_t* v = &e->v[e->ego];
    int crashed = (v->flags & TTRL_FL_CRASHED) != 0;
    if (sc->cfg.reward_type == TTRL_REWARD_ROUNDABOUT) return crashed; /* roundabout_env.py:66-67 */
    if (sc->cfg.reward_type == TTRL_REWARD_INTERSECTION) {
        /* any(crashed) or all(has_arrived) over controlled_vehicles or (offroad_terminal and not self.vehicle.on_road) */
        int any_crashed = 0, all_arrived = 1;
        for (int k = 0; k < n_agents(sc); ++k) {
            const veh_t* a = &e->v[e->egos[k]];
            if (a->flags & TTRL_FL_CRASHED) any_crashed = 1;
            if (!has_arrived(sc, a)) all_arrived = 0;
        }
        return any_crashed || all_arrived || (sc->cfg.offroad_terminal && !on_road(sc, v));
    }
    return crashed || (sc->cfg.offroad_terminal && !on_road(sc, v));
}
static void find_agents(env_t* e) {
    for (int i = 0; i < e->n; ++i)
        if ((e->v[i].flags & TTRL_FL_CONTROLLED) && (e->v[i].flags & TTRL_FL_AGENT_MASK))
            e->egos[(e->v[i].flags & TTRL_FL_AGENT_MASK) >> TTRL_FL_AGENT_SHIFT] = i;
}

/* ----------------------------------------------------------------------------------------------
 * IntersectionEnv clear / spawn: intersection_env.py:320-362
 * -------------------------------------------------------------------------------------------- */
static void clear_vehicles(const orc_scene* sc, env_t* e) {
    int w = 0, new_ego = e->ego;
    for (int i = 0; i < e->n; ++i) {
        const veh_t* v = &e->v[i];
        const ttrl_lane* l = &sc->lanes[v->lane];
        int keep = (v->flags & TTRL_FL_CONTROLLED) != 0;
        if (!keep) {
            int leaving = 0;
            if (l->is_exit) { double s, r; lane_local(l, v->x, v->y, &s, &r); leaving = s >= l->length - 4 * 5.0; }
            keep = !(leaving || v->route_len < 0);
        }
        if (keep) { if (i == e->ego) new_ego = w; if (w != i) e->v[w] = e->v[i]; w++; }
    }
    e->n = w; e->ego = new_ego; e->egos[0] = new_ego;
    find_agents(e);
}
/* returns 1 if a vehicle was appended.  longitudinal/deviations as in _spawn_vehicle's signature. */
static int spawn_vehicle(const orc_scene* sc, env_t* e, const ttrl_spawn_draw* d, double longitudinal,
                         double position_deviation, double speed_deviation, double spawn_probability, int go_straight, int vcap) {
    if (d->u_spawn > spawn_probability) return 0;
    int entry = d->entry, exit_ = go_straight ? (d->entry + 2) % 4 : d->exit;
    const ttrl_lane* l = &sc->lanes[sc->spawn_lane[entry]];
    double lon = longitudinal + 5 + d->n_pos * position_deviation;
    double speed = 8 + d->n_speed * speed_deviation;
    veh_t v; memset(&v, 0, sizeof v);
    lane_position(l, lon, 0, &v.x, &v.y);
    v.heading = lane_heading_at(l, lon);
    v.speed = speed;
    v.lane = closest_lane(sc, v.x, v.y, v.heading);     /* RoadObject.__init__ objects.py:45-50 */
    v.target_lane = v.lane;                              /* controller.py:46 */
    v.target_speed = speed;                              /* controller.py:47 (`or self.speed`: same value when 0) */
    v.timer = py_mod((v.x + v.y) * PI, 1.0);             /* behavior.py:64 */
    for (int j = 0; j < e->n; ++j) {
        double dx = e->v[j].x - v.x, dy = e->v[j].y - v.y;
        if (sqrt(dx * dx + dy * dy) < 15) return 0;
    }
    if (e->n >= vcap) return 0; /* capacity guard (not in the reference; documented in DESIGN.md) */
    /* plan_route_to controller.py:71-87: route = [lane_index] + BFS path roads (precomputed by the host) */
    v.route_len = 1 + sc->spawn_route_len[entry][exit_];
    v.route_road[0] = sc->lanes[v.lane].road; v.route_lane[0] = sc->lanes[v.lane].lane_id;
    for (int k = 0; k < sc->spawn_route_len[entry][exit_]; ++k) { v.route_road[1 + k] = sc->spawn_route_road[entry][exit_][k]; v.route_lane[1 + k] = -1; }
    v.delta = d->delta;                                  /* randomize_behavior behavior.py:66-69 */
    if (sc->cfg.vehicle_model == TTRL_VEHICLE_LINEAR)    /* LinearVehicle.randomize_behavior behavior.py:402-410 */
        for (int k = 0; k < TTRL_NLIN; ++k) v.lin[k] = sc->cfg.lin_lo[k] + d->lin_u[k] * (sc->cfg.lin_hi[k] - sc->cfg.lin_lo[k]);
    e->v[e->n++] = v;
    return 1;
}

/* ----------------------------------------------------------------------------------------------
 * SoA <-> env_t
 * -------------------------------------------------------------------------------------------- */
static void load_env(env_t* e, const double* vd, const int32_t* vi, const int32_t* ei, const double* ed, int E, int V, int ie, const double* lin) {
    e->n = ei[TTRL_EI_NVEH * E + ie]; e->steps = ei[TTRL_EI_STEPS * E + ie]; e->road_steps = ei[TTRL_EI_ROAD_STEPS * E + ie];
    e->ego = ei[TTRL_EI_EGO * E + ie]; e->episode = ei[TTRL_EI_EPISODE * E + ie]; e->done = ei[TTRL_EI_DONE * E + ie];
    e->time = ed[TTRL_ED_TIME * E + ie]; e->ret = ed[TTRL_ED_RETURN * E + ie];
    for (int s = 0; s < e->n; ++s) {
        veh_t* v = &e->v[s];
#define D(f) vd[((size_t)(f) * E + ie) * V + s]
#define I(f) vi[((size_t)(f) * E + ie) * V + s]
        v->x = D(TTRL_D_X); v->y = D(TTRL_D_Y); v->heading = D(TTRL_D_HEADING); v->speed = D(TTRL_D_SPEED);
        v->steering = D(TTRL_D_STEERING); v->accel = D(TTRL_D_ACCEL); v->target_speed = D(TTRL_D_TARGET_SPEED);
        v->timer = D(TTRL_D_TIMER); v->delta = D(TTRL_D_DELTA); v->impact_x = D(TTRL_D_IMPACT_X); v->impact_y = D(TTRL_D_IMPACT_Y);
        v->lane = I(TTRL_I_LANE); v->target_lane = I(TTRL_I_TARGET_LANE); v->flags = I(TTRL_I_FLAGS);
        v->speed_index = I(TTRL_I_SPEED_INDEX); v->route_len = I(TTRL_I_ROUTE_LEN); v->yield_timer = I(TTRL_I_YIELD_TIMER);
        for (int k = 0; k < TTRL_NLIN; ++k) v->lin[k] = lin ? lin[((size_t)k * E + ie) * V + s] : 0;
        const uint32_t rr[TTRL_ROUTE_WORDS] = {(uint32_t)I(TTRL_I_ROUTE_ROAD), (uint32_t)I(TTRL_I_ROUTE_ROAD1), (uint32_t)I(TTRL_I_ROUTE_ROAD2)};
        const uint32_t rl[TTRL_ROUTE_WORDS] = {(uint32_t)I(TTRL_I_ROUTE_LANE), (uint32_t)I(TTRL_I_ROUTE_LANE1), (uint32_t)I(TTRL_I_ROUTE_LANE2)};
        for (int k = 0; k < TTRL_ROUTE_CAP; ++k) {
            v->route_road[k] = (rr[k >> 2] >> (8 * (k & 3))) & 0xFF;
            int b = (rl[k >> 2] >> (8 * (k & 3))) & 0xFF; v->route_lane[k] = b == 0xFF ? -1 : b;
        }
#undef D
#undef I
    }
    for (int k = 0; k < TTRL_MAX_CONTROLLED; ++k) e->egos[k] = e->ego;
    find_agents(e);
}
static void store_env(const env_t* e, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int V, int ie, double* lin) {
    ei[TTRL_EI_NVEH * E + ie] = e->n; ei[TTRL_EI_STEPS * E + ie] = e->steps; ei[TTRL_EI_ROAD_STEPS * E + ie] = e->road_steps;
    ei[TTRL_EI_EGO * E + ie] = e->ego; ei[TTRL_EI_EPISODE * E + ie] = e->episode; ei[TTRL_EI_DONE * E + ie] = e->done;
    ed[TTRL_ED_TIME * E + ie] = e->time; ed[TTRL_ED_RETURN * E + ie] = e->ret;
    for (int s = 0; s < V; ++s) {
        veh_t z; memset(&z, 0, sizeof z);
        const veh_t* v = s < e->n ? &e->v[s] : &z;
#define D(f) vd[((size_t)(f) * E + ie) * V + s]
#define I(f) vi[((size_t)(f) * E + ie) * V + s]
        D(TTRL_D_X) = v->x; D(TTRL_D_Y) = v->y; D(TTRL_D_HEADING) = v->heading; D(TTRL_D_SPEED) = v->speed;
        D(TTRL_D_STEERING) = v->steering; D(TTRL_D_ACCEL) = v->accel; D(TTRL_D_TARGET_SPEED) = v->target_speed;
        D(TTRL_D_TIMER) = v->timer; D(TTRL_D_DELTA) = v->delta; D(TTRL_D_IMPACT_X) = v->impact_x; D(TTRL_D_IMPACT_Y) = v->impact_y;
        I(TTRL_I_LANE) = v->lane; I(TTRL_I_TARGET_LANE) = v->target_lane; I(TTRL_I_FLAGS) = v->flags;
        I(TTRL_I_SPEED_INDEX) = v->speed_index; I(TTRL_I_ROUTE_LEN) = s < e->n ? v->route_len : 0; I(TTRL_I_YIELD_TIMER) = v->yield_timer;
        uint32_t rr[TTRL_ROUTE_WORDS] = {0, 0, 0}, rl[TTRL_ROUTE_WORDS] = {0, 0, 0};
        for (int k = 0; k < TTRL_ROUTE_CAP; ++k) {
            int on = s < e->n && k < v->route_len;
            rr[k >> 2] |= (uint32_t)((on ? v->route_road[k] : 0) & 0xFF) << (8 * (k & 3));
            rl[k >> 2] |= (uint32_t)((on ? (v->route_lane[k] < 0 ? 0xFF : v->route_lane[k]) : 0) & 0xFF) << (8 * (k & 3));
        }
        I(TTRL_I_ROUTE_ROAD) = (int32_t)rr[0]; I(TTRL_I_ROUTE_LANE) = (int32_t)rl[0];
        I(TTRL_I_ROUTE_ROAD1) = (int32_t)rr[1]; I(TTRL_I_ROUTE_LANE1) = (int32_t)rl[1];
        I(TTRL_I_ROUTE_ROAD2) = (int32_t)rr[2]; I(TTRL_I_ROUTE_LANE2) = (int32_t)rl[2];
        if (lin) for (int k = 0; k < TTRL_NLIN; ++k) lin[((size_t)k * E + ie) * V + s] = v->lin[k];
#undef D
#undef I
    }
}

/* ----------------------------------------------------------------------------------------------
 * public (ctypes) entry points
 * -------------------------------------------------------------------------------------------- */
orc_scene* orc_scene_create(const ttrl_config* cfg, const ttrl_lane* lanes, const ttrl_road* roads,
                            const int32_t* node_first, const int32_t* node_roads) {
    orc_scene* sc = (orc_scene*)calloc(1, sizeof(orc_scene));
    sc->cfg = *cfg;
    memcpy(sc->lanes, lanes, sizeof(ttrl_lane) * cfg->n_lanes);
    memcpy(sc->roads, roads, sizeof(ttrl_road) * cfg->n_roads);
    memcpy(sc->node_first, node_first, sizeof(int32_t) * (cfg->n_nodes + 1));
    memcpy(sc->node_roads, node_roads, sizeof(int32_t) * node_first[cfg->n_nodes]);
    return sc;
}
void orc_scene_destroy(orc_scene* sc) {
    if (!sc) return;
    free(sc->pool_vd); free(sc->pool_vi); free(sc->pool_ei); free(sc->pool_ed);
    free(sc);
}
void orc_scene_set_spawn_routes(orc_scene* sc, const int32_t* spawn_lane, const int32_t* route_len, const int32_t* route_road) {
    for (int a = 0; a < 4; ++a) {
        sc->spawn_lane[a] = spawn_lane[a];
        for (int b = 0; b < 4; ++b) {
            sc->spawn_route_len[a][b] = route_len[a * 4 + b];
            for (int k = 0; k < TTRL_ROUTE_CAP; ++k) sc->spawn_route_road[a][b][k] = route_road[(a * 4 + b) * TTRL_ROUTE_CAP + k];
        }
    }
}
void orc_scene_set_reset_pool(orc_scene* sc, int pool_size, int V, const double* vd, const int32_t* vi, const int32_t* ei, const double* ed) {
    free(sc->pool_vd); free(sc->pool_vi); free(sc->pool_ei); free(sc->pool_ed);
    sc->pool_size = pool_size; sc->pool_E = pool_size; sc->pool_V = V;
    size_t nd = (size_t)TTRL_ND * pool_size * V, ni = (size_t)TTRL_NI * pool_size * V;
    sc->pool_vd = (double*)malloc(nd * sizeof(double)); memcpy(sc->pool_vd, vd, nd * sizeof(double));
    sc->pool_vi = (int32_t*)malloc(ni * sizeof(int32_t)); memcpy(sc->pool_vi, vi, ni * sizeof(int32_t));
    sc->pool_ei = (int32_t*)malloc(sizeof(int32_t) * TTRL_NEI * pool_size); memcpy(sc->pool_ei, ei, sizeof(int32_t) * TTRL_NEI * pool_size);
    sc->pool_ed = (double*)malloc(sizeof(double) * TTRL_NED * pool_size); memcpy(sc->pool_ed, ed, sizeof(double) * TTRL_NED * pool_size);
}
void orc_scene_set_autoreset(orc_scene* sc, int on) { sc->autoreset = on; }
/* parameter blocks [TTRL_NLIN][E][V] that go with the state arrays of the following calls / with the reset pool */
void orc_scene_set_linear_params(orc_scene* sc, double* lin, const double* pool_lin) { sc->lin = lin; sc->pool_lin = pool_lin; }
int orc_obs_size(const orc_scene* sc) { return obs_size(sc); }

/* F x substep is what _simulate does; this runs ONE substep for E envs (debug / per-substep parity). */
void orc_substep(const orc_scene* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int V, const int32_t* actions, int threads) {
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
    for (int ie = 0; ie < E; ++ie) {
        env_t e; load_env(&e, vd, vi, ei, ed, E, V, ie, sc->lin);
        env_substep(sc, &e, actions ? actions + (size_t)ie * n_agents(sc) : NULL);
        store_env(&e, vd, vi, ei, ed, E, V, ie, sc->lin);
    }
}

void orc_observe(const orc_scene* sc, const double* vd, const int32_t* vi, const int32_t* ei, const double* ed, int E, int V, float* obs, int threads) {
    int osz = obs_size(sc);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
    for (int ie = 0; ie < E; ++ie) {
        env_t e; load_env(&e, vd, vi, ei, ed, E, V, ie, sc->lin);
        observe(sc, &e, obs + (size_t)ie * osz);
    }
}

/* One env.step() for E envs: AbstractEnv.step abstract.py:224-250 then IntersectionEnv.step's clear + spawn
 * (intersection_env.py:135-139).  draws == NULL -> no spawn attempt.  stats[8] accumulates ttrl_episode_stats. */
void orc_step_agents(const orc_scene* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int V, const int32_t* actions,
              float* obs, float* reward, uint8_t* terminated, uint8_t* truncated, const ttrl_spawn_draw* draws,
              int32_t* spawn_accepted, double* stats, int threads, float* agent_reward_out, uint8_t* agent_term_out);
void orc_step(const orc_scene* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int V, const int32_t* actions,
              float* obs, float* reward, uint8_t* terminated, uint8_t* truncated, const ttrl_spawn_draw* draws,
              int32_t* spawn_accepted, double* stats, int threads) {
    orc_step_agents(sc, vd, vi, ei, ed, E, V, actions, obs, reward, terminated, truncated, draws, spawn_accepted, stats, threads, NULL, NULL);
}
/* same + per-agent outputs: info["agents_rewards"], info["agents_terminated"] (intersection_env.py:121-129), [E][K] */
void orc_step_agents(const orc_scene* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int V, const int32_t* actions,
              float* obs, float* reward, uint8_t* terminated, uint8_t* truncated, const ttrl_spawn_draw* draws,
              int32_t* spawn_accepted, double* stats, int threads, float* agent_reward_out, uint8_t* agent_term_out) {
    int F = (int)floor(sc->cfg.simulation_frequency / sc->cfg.policy_frequency);
    int osz = obs_size(sc);
    double st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4) reduction(+ : st[:8])
    for (int ie = 0; ie < E; ++ie) {
        env_t e; load_env(&e, vd, vi, ei, ed, E, V, ie, sc->lin);
        e.time += 1 / sc->cfg.policy_frequency;
        const int K = n_agents(sc);
        const int32_t* act = actions ? actions + (size_t)ie * K : NULL;
        for (int f = 0; f < F; ++f) { env_substep(sc, &e, act); st[6] += e.n; }
        observe(sc, &e, obs + (size_t)ie * osz);
        /* IntersectionEnv._reward intersection_env.py:61-65: sum(agent rewards) / len(controlled_vehicles) */
        double r = 0;
        for (int k = 0; k < K; ++k) {
            const veh_t* a = &e.v[e.egos[k]];
            double rk = agent_reward(sc, a, act ? act[k] : -1);
            r = r + rk;
            if (agent_reward_out) agent_reward_out[(size_t)ie * K + k] = (float)rk;
            if (agent_term_out) agent_term_out[(size_t)ie * K + k] = (uint8_t)(((a->flags & TTRL_FL_CRASHED) != 0) || has_arrived(sc, a)); /* :113-115 */
        }
        r = r / K;
        int term = is_terminated(sc, &e), trunc = e.time >= sc->cfg.duration;
        reward[ie] = (float)r; terminated[ie] = (uint8_t)term; truncated[ie] = (uint8_t)trunc;
        e.ret += r; st[7] += 1; st[5] += e.v[e.ego].speed;
        if (sc->cfg.spawn_enabled) {
            clear_vehicles(sc, &e);
            int acc = 0;
            if (draws) acc = spawn_vehicle(sc, &e, &draws[ie], 0, 1.0, 1.0, sc->cfg.spawn_probability, 0, V);
            if (spawn_accepted) spawn_accepted[ie] = acc;
        }
        if (term || trunc) {
            e.done = 1;
            st[0] += 1; st[1] += e.ret; st[2] += e.time * sc->cfg.policy_frequency;
            st[3] += (e.v[e.ego].flags & TTRL_FL_CRASHED) ? 1 : 0;
            st[4] += (sc->cfg.reward_type == TTRL_REWARD_INTERSECTION && has_arrived(sc, &e.v[e.ego])) ? 1 : 0;
            if (sc->autoreset && sc->pool_size > 0) {
                int episode = e.episode + 1;
                int slot = (int)(((long long)ie + (long long)episode * E) % sc->pool_size);
                load_env(&e, sc->pool_vd, sc->pool_vi, sc->pool_ei, sc->pool_ed, sc->pool_E, sc->pool_V, slot, sc->pool_lin);
                e.episode = episode; e.done = 0;
                observe(sc, &e, obs + (size_t)ie * osz);
            }
        }
        store_env(&e, vd, vi, ei, ed, E, V, ie, sc->lin);
    }
    if (stats) for (int k = 0; k < 8; ++k) stats[k] += st[k];
}

/* Host-driven reset primitive (IntersectionEnv._make_vehicles uses _spawn_vehicle with other arguments,
 * intersection_env.py:265-283): spawn attempt with explicit longitudinal / deviations / probability. */
void orc_spawn(const orc_scene* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int V, const ttrl_spawn_draw* draws,
               double longitudinal, double position_deviation, double speed_deviation, double spawn_probability, int go_straight,
               int32_t* accepted) {
    for (int ie = 0; ie < E; ++ie) {
        env_t e; load_env(&e, vd, vi, ei, ed, E, V, ie, sc->lin);
        int acc = spawn_vehicle(sc, &e, &draws[ie], longitudinal, position_deviation, speed_deviation, spawn_probability, go_straight, V);
        if (accepted) accepted[ie] = acc;
        store_env(&e, vd, vi, ei, ed, E, V, ie, sc->lin);
    }
}
