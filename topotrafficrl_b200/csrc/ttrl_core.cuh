// ttrl_core.cuh -- device-side simulation logic of the B200 batched simulator.
//
// One CTA ("team" of T threads, T = 32 for up to 64 vehicle slots) advances one env instance; the env's whole
// state lives in shared memory for the F sub-steps of an env-step (one HBM round trip per env-step).  The logic
// is written as barrier-separated PHASES over TASK LISTS behind an `Exec` policy:
//   * ex.par(f)      f(slot) for every slot 0..V-1           (strided over the team) ; barrier
//   * ex.parn(n, f)  f(task) for task 0..n-1                 (strided over the team) ; barrier
//   * ex.any(n, f)   OR-reduction of f(task)
//   * ex.sat_pairs   8 lanes per candidate pair, one SAT axis each, shuffle-reduced
//   * on the device Exec = one CTA; in tests/emu a host Exec runs the tasks of a phase one after the other, so
//     the very same code is validated against the CPU oracle in a container without a GPU (NOT a product path).
//
// Why task lists (ncu, profiles/r1a_*): the first version ran rare heavy work (MOBIL, SAT) inside per-vehicle
// threads, so on average 10 of 32 lanes were active and 320 KB of SASS thrashed the instruction cache.  Here the
// rare work is queued in shared memory and executed with full lanes: MOBIL = neighbour-query tasks + IDM
// evaluation tasks + one decision task per vehicle; collisions = half-ring candidate scan with an exact
// squared-distance guard -> pair queue -> 8 axis-lanes per pair.  Heavy math (pow, atan2, asin ...) exists ONCE
// in the binary behind __noinline__ functions.
//
// A per-env table SR[v][l] = (s, r) of every vehicle's local coordinates in every lane is rebuilt after each
// integration (it is the by-product of the closest-lane search, reference road.py:55-71) together with a
// per-lane bitmask of the vehicles that are "on" the lane (lane.py:80-102, margin 1); they serve all neighbour
// scans, lane distances and controllers of the next sub-step from shared memory.
//
// Reference citations are relative to /root/reference.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/ttrl_b200.h"

#if defined(__CUDACC__)
#define TT_HD __host__ __device__ __forceinline__
#define TT_HDN inline __host__ __device__ __noinline__
#else
#define TT_HD inline
#define TT_HDN inline
#endif

namespace ttrl {

constexpr double kPi = 3.141592653589793;
constexpr int kPred = 11;            // regulation.py:87: np.arange(0.25, 3, 0.25)
constexpr double kVehLength = 5.0;   // kinematics.py:21
constexpr double kVehWidth = 2.0;    // kinematics.py:23
constexpr int kStatFields = 10;      // ttrl_episode_stats
struct alignas(16) d2 { double x, y; };  // 16-byte pair (one LDS.128 on the device)
struct alignas(8) f2 { float x, y; };     // 8-byte pair (one LDS.64)

// Read-only scene description, resident in global memory (L1/L2 hot: ~12 KB).
struct SceneDev {
    ttrl_config cfg;
    alignas(16) ttrl_lane lanes[TTRL_MAX_LANES];  // copied to shared memory with 16-byte vector loads (make_ctx)
    ttrl_road roads[TTRL_MAX_ROADS];
    int32_t node_first[TTRL_MAX_NODES + 1];
    int32_t node_roads[TTRL_MAX_ROADS];
    int32_t spawn_lane[4];
    int32_t spawn_route_len[16];
    int32_t spawn_route_road[16 * TTRL_ROUTE_CAP];
    ttrl_reset_params rp;  // device-side reset parameters (have_rp != 0)
    int32_t have_rp;
    int32_t n_curved;    // curved lanes = columns of the per-vehicle coordinate cache (assign_cache_columns)
    int8_t curved_lane[TTRL_MAX_LANES];  // cache column -> lane index
    int32_t arc_tasks;   // closest-lane search through the curved-lane task list (use_arc_tasks), else one loop per vehicle
    int32_t pad_arc;
    int32_t F;           // sub-steps per env-step (abstract.py:254-256)
    int32_t reg_period;  // int(1/dt/REGULATION_FREQUENCY) (regulation.py:30)
    double dt;           // 1/simulation_frequency
};

// Per-env state in shared memory (struct of arrays over the V slots) + task scratch.
template <int V>
struct alignas(16) EnvState {
    // MOBIL batch: vehicles whose lane-change timer fired, per round.  Small scenes (a handful of vehicles per env, the
    // shared-memory footprint of an env decides how many an SM holds) take rounds of 4: 9 x 4 evaluation slots instead of 9 x 16
    static constexpr int MB = V <= 32 ? 4 : 16;
    static constexpr int PQ = 2 * V;            // collision candidate pair queue
    static constexpr int WQ = V / 2;            // will-intersect pair list
    d2 pos[V];   // x, y
    d2 cs[V];    // cos(heading), sin(heading)
    d2 imp[V];   // pending impact
    d2 wt[WQ];   // translation of will-intersect pairs
    double h[V], v[V];
    double steer[V], acc[V], tspeed[V], timer[V], delta[V];
    double thr2[V];      // collision pre-check guard: (diag + v dt)^2 (1 + 1e-12), or -1 when diag + v dt < 0
    double acc2[V];      // IDM acceleration w.r.t. the target lane (vehicles changing lane); dead after the integration: its
                         // storage and tsteer's then hold the collision scan's float pre-filter (collide_all)
    double tsteer[V];    // tan(steering command) of this sub-step (by-product of steering_control, used by integrate); follows acc2
    double mq_a[9 * MB];  // MOBIL IDM evaluations of the batch
    int32_t lane[V], tlane[V], flags[V], sidx[V], rlen[V], ytimer[V];
    uint32_t rroad[TTRL_ROUTE_WORDS][V], rlanew[TTRL_ROUTE_WORDS][V];  // route entry k: byte k % 4 of word k / 4 (road index / lane id)
    int32_t tl_old[V], mark[V];
    int32_t best[V];     // collision: largest will-intersect partner index
    int32_t chg[V];      // vehicles in an ongoing lane change (IDM w.r.t. target lane too)
    int32_t mob_veh[V];  // vehicles whose MOBIL timer fired this sub-step
    int32_t pairq[PQ];   // lo | hi << 16
    int32_t wpair[WQ];
    int16_t fo[V];       // front neighbour on the own lane
    uint32_t bmask[(V + 31) / 32];  // vehicles in an ongoing lane change on the same road (phase B), list order = bit order
    int16_t mq_f[3 * MB], mq_r[3 * MB];  // MOBIL neighbour queries: own lane, left candidate, right candidate
    int32_t n, steps, road_steps, ego, episode, done, flag0, flag1;
    int32_t egos[TTRL_MAX_CONTROLLED];  // slots of env.controlled_vehicles (egos[0] == ego)
    int32_t n_chg, n_mob, n_pair, n_w, overflow, n_arc;  // n_arc: queued (vehicle, curved lane) projection tasks of this sub-step
    double time, ret;
};

// P_ = 1 ("plain" scenes: straight lanes only, no regulation, no spawn / clear, highway reward -- BASELINE configs 1
// and 2) compiles the curved-lane geometry, the regulation and the intersection reward out of the step loop: the kernel
// is bound by instruction-cache miss traffic, so code that is never executed should not sit between code that is.
// P_ = 2 ("multi"): the number of controlled vehicles is read from the config at run time (MultiAgentIntersectionEnv); the
// single-agent profiles 0 / 1 compile the per-agent loops and lookups out for the same reason.
// P_ = 3 ("linear"): general single-agent scenes whose traffic is LinearVehicle -- its controllers are compiled into this
// profile only, for the same reason.  P_ = 4 ("any"): everything decided at run time (several agents, linear traffic if
// c.lin is set): the single-team utility kernels (k_substep, k_observe, k_spawn, k_reset), which are not performance critical.
template <int V_, int P_ = 0>
struct EnvCtx {
    static constexpr int V = V_;
    static constexpr int P = P_;
    static constexpr bool kPlain = P_ == 1;
    static constexpr bool kMulti = P_ == 2 || P_ == 4;
    static constexpr int W = (V_ + 31) / 32;  // mask words per lane
    EnvState<V_>* st;
    const SceneDev* sc;
    const ttrl_lane* lanes;  // lane table (shared-memory copy)
    d2* SR;                  // [V][NC] (longitudinal, lateral) local coordinates of vehicle v in CURVED lane column k (sr_of)
    int NC;                  // columns of SR: the curved lanes of the network (plain profile: every lane)
    uint32_t* lmask;         // [L][W] bit v of lane l: vehicle v is on lane l with margin 1 (road.py:503)
    double* pred;            // [4][V] regulation predictions of one time slice (x, y, cos h, sin h); null if not regulated
    uint32_t* pbits;         // [V (V - 1) / 2 bits] pairs in conflict (regulation); null if not regulated
    float* obs_s;            // staging for one Kinematics observation
    uint32_t* perm_s;        // [2 (obs_vehicles - 1)] device-drawn row shuffle: random keys, then their ranks
    double* lin;             // [TTRL_NLIN][V] LinearVehicle parameters (vehicle_model == TTRL_VEHICLE_LINEAR), else null
    int32_t* cell;           // OccupancyGrid per-cell winner (W*H ints)
    int L;
    int vcap;                // storage capacity (slots per env in HBM), <= V
    double gap_den;          // 2 sqrt(-COMFORT_ACC_MAX COMFORT_ACC_MIN) (behavior.py:214-216)
    double tan_max_steer;    // tan(MAX_STEERING_ANGLE)
};

// LinearVehicle traffic in this env? (compile-time for the step profiles, run-time for the "any" profile)
template <class C> TT_HD bool is_linear(const C& c) { return C::P == 3 || (C::P == 4 && c.lin != nullptr); }

// ------------------------------------------------------------------------------------------------
// scalar helpers: ttrl_env/utils.py
// ------------------------------------------------------------------------------------------------
// a / b for a finite non-zero b, without the division's slow path when a is exactly zero.  The fp64 division is inlined as a
// reciprocal refinement whose fast path excludes tiny numerators (|a| < 2^-120, zero included); the call it falls back to
// (__cuda_sm20_div_rn_f64_full, ~70 instructions) was 7.4 % of the highway kernel's instructions: a vehicle centred on its lane has a
// lateral error of exactly 0, a steering command of exactly 0 and a slip term of exactly 0 in every sub-step.  0 * b has the sign of 0 / b.
TT_HD double div_z(double a, double b) { return a == 0.0 ? a * b : a / b; }
TT_HD double not_zero(double x) {  // utils.py:48-54
    const double eps = 1e-2;
    if (fabs(x) > eps) return x;
    return x >= 0 ? eps : -eps;
}
// Python floored float modulo by 2*pi (utils.py:58).  Fast path for 0 <= a < 2pi (the common case); the
// general path is the exact remainder through one fma.
TT_HD double mod_2pi(double a) {
    const double b = 2 * kPi;
    if (a >= 0 && a < b) return a;
    // one turn below / above: the quotient floor(a / b) is exactly -1 / +1 there (the 1e-6 margins keep clear of the values
    // where the division could round across an integer), so the general path's fma(-q, b, a) is a + b / a - b rounded once
    double m;
    if (a < 0 && a > -(b - 1e-6)) m = a + b;
    else if (a >= b && a < 2 * b - 1e-6) m = a - b;
    else {
        const double q = floor(a / b);
        m = fma(-q, b, a);
    }
    if (m < 0) m += b;
    else if (m >= b) m -= b;
    return m;
}
TT_HD double wrap_to_pi(double x) { return mod_2pi(x + kPi) - kPi; }  // utils.py:57-58
TT_HD double py_mod1(double a) {  // a % 1.0 for behavior.py:64
    double m = a - floor(a);
    return (m >= 1.0) ? 0.0 : m;
}
TT_HD double lmap(double v, double x0, double x1, double y0, double y1) {  // utils.py:29-31
    return y0 + (v - x0) * (y1 - y0) / (x1 - x0);
}
TT_HD double clipd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// counter-based RNG (Philox-4x32-10) of every device-side draw: spawns, resets, observation shuffles
TT_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
TT_HD void philox4x32(uint32_t ctr[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
TT_HD double u01(uint32_t hi, uint32_t lo) {  // 53-bit uniform in [0,1)
    const uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)x * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------------------
// lane geometry: ttrl_env/road/lane.py.  The curved kinds live behind __noinline__ functions so that
// atan2 / sin / cos exist once in the binary.
// ------------------------------------------------------------------------------------------------
// The curved-lane projection exists twice: inlined into the closest-lane loop (table_row_and_closest: every vehicle x every
// lane x every sub-step; without the call the compiler overlaps the atan2 / sqrt chain with the neighbouring lanes' work:
// intersection -5 %, multi-agent -8 %, roundabout -16 %, profiles/r1l_lane_loop.txt) and as ONE out-of-line copy for the rare
// call sites.  TT_LANE_UNROLL > 1 unrolls that loop (measured slower: 2 -> +1 %, 4 -> +6 .. +17 %).
#ifndef TT_LANE_UNROLL
#define TT_LANE_UNROLL 1
#endif
// Call boundaries (profiles/r1m_inline.txt).  A __noinline__ function exists once in the binary (less instruction fetch, what
// the first kernels needed most) but is a scheduling barrier for the code around the call.  With the fetch traffic down, inlining
// the controllers pays: steering + IDM inlined: highway 1.188 -> 1.120 ms/step; + the regulation predictions / pair tests:
// intersection 3.29 -> 2.97.  the curved-lane position / sine heading: another 1.3 %.  -DTT_NOINL_STEER / _IDM / _REG / _POS restore the calls.
#if defined(TT_NOINL_STEER)
#define TT_STEER TT_HDN
#else
#define TT_STEER TT_HD
#endif
#if defined(TT_NOINL_IDM)
#define TT_IDM TT_HDN
#else
#define TT_IDM TT_HD
#endif
#if defined(TT_NOINL_REG)
#define TT_REG TT_HDN
#else
#define TT_REG TT_HD
#endif
#if defined(TT_NOINL_POS)
#define TT_POS TT_HDN
#else
#define TT_POS TT_HD
#endif
#define TT_PRAGMA_(x) _Pragma(#x)
#define TT_PRAGMA(x) TT_PRAGMA_(x)
TT_HD void lane_local_curved_inl(const ttrl_lane& l, double px, double py, double& s, double& r) {
    double dx = px - l.ax, dy = py - l.ay;
    if (l.kind == TTRL_LANE_CIRCULAR) {  // lane.py:355-362
        double phi = atan2(dy, dx);
        phi = l.start_phase + wrap_to_pi(phi - l.start_phase);
        double rad = sqrt(dx * dx + dy * dy);
        s = l.cdir * (phi - l.start_phase) * l.radius;
        r = l.cdir * (l.radius - rad);
    } else {  // sine lane.py:282-286
        double lon = dx * l.dx + dy * l.dy;
        double lat = dx * (-l.dy) + dy * l.dx;
        s = lon;
        r = lat - l.amplitude * sin(l.pulsation * lon + l.phase);
    }
}
TT_HDN void lane_local_curved(const ttrl_lane& l, double px, double py, double& s, double& r) { lane_local_curved_inl(l, px, py, s, r); }
TT_HD void lane_local(const ttrl_lane& l, double px, double py, double& s, double& r) {
    if (l.kind != TTRL_LANE_STRAIGHT) { lane_local_curved(l, px, py, s, r); return; }
    double dx = px - l.ax, dy = py - l.ay;  // lane.py:209-213
    s = dx * l.dx + dy * l.dy;
    r = dx * (-l.dy) + dy * l.dx;
}
TT_POS void lane_position_curved(const ttrl_lane& l, double s, double r, double& px, double& py) {
    if (l.kind == TTRL_LANE_CIRCULAR) {  // lane.py:341-345
        double phi = l.cdir * s / l.radius + l.start_phase;
        double rr = l.radius - r * l.cdir;
        px = l.ax + rr * cos(phi);
        py = l.ay + rr * sin(phi);
    } else {  // sine lane.py:268-273
        r = r + l.amplitude * sin(l.pulsation * s + l.phase);
        px = l.ax + s * l.dx + r * (-l.dy);
        py = l.ay + s * l.dy + r * l.dx;
    }
}
TT_HD void lane_position(const ttrl_lane& l, double s, double r, double& px, double& py) {
    if (l.kind != TTRL_LANE_STRAIGHT) { lane_position_curved(l, s, r, px, py); return; }
    px = l.ax + s * l.dx + r * (-l.dy);  // lane.py:196-201
    py = l.ay + s * l.dy + r * l.dx;
}
TT_POS double lane_heading_sine(const ttrl_lane& l, double s) {  // lane.py:275-280
    return l.heading + atan(l.amplitude * l.pulsation * cos(l.pulsation * s + l.phase));
}
TT_HD double lane_heading_at(const ttrl_lane& l, double s) {  // lane.py:203-204, :347-350
    if (l.kind == TTRL_LANE_CIRCULAR) {
        double phi = l.cdir * s / l.radius + l.start_phase;
        return phi + kPi / 2 * l.cdir;
    }
    if (l.kind == TTRL_LANE_SINE) return lane_heading_sine(l, s);
    return l.heading;
}
TT_HD bool lane_on_lane(const ttrl_lane& l, double s, double r, double margin) {  // lane.py:80-102
    return fabs(r) <= l.width / 2 + margin && -kVehLength <= s && s < l.length + kVehLength;
}
TT_HD bool lane_reachable(const ttrl_lane& l, double s, double r) {  // lane.py:104-118
    if (l.forbidden) return false;
    return fabs(r) <= 2 * l.width && 0 <= s && s < l.length + kVehLength;
}
TT_HD double lane_distance_sr(const ttrl_lane& l, double s, double r) {  // lane.py:127-130
    return fabs(r) + fmax(s - l.length, 0.0) + fmax(0.0 - s, 0.0);
}
TT_HD double lane_distance(const ttrl_lane& l, double px, double py) {
    double s, r;
    lane_local(l, px, py, s, r);
    return lane_distance_sr(l, s, r);
}

// scene-profile aware forms for the hot call sites
template <class C> TT_HD void lane_local_c(const ttrl_lane& l, double px, double py, double& s, double& r) {
    if (C::kPlain) { const double dx = px - l.ax, dy = py - l.ay; s = dx * l.dx + dy * l.dy; r = dx * (-l.dy) + dy * l.dx; }
    else lane_local(l, px, py, s, r);
}
template <class C> TT_HD double lane_heading_at_c(const ttrl_lane& l, double s) { return C::kPlain ? l.heading : lane_heading_at(l, s); }

// ------------------------------------------------------------------------------------------------
// table access
// ------------------------------------------------------------------------------------------------
TT_HDN d2 sr_curved_slow(const ttrl_lane& ln, d2 p) {
    double s, r;
    lane_local_curved_inl(ln, p.x, p.y, s, r);
    return d2{s, r};
}
// (s, r) of vehicle v in lane l.  Straight lanes: 6 flops from the position (the same expression as the closest-lane
// search: bit-identical); curved lanes: the cached result of that search (one atan2 / sin per vehicle, lane and sub-step).
// Caching only the curved lanes is what lets an SM hold more envs: the full V x L table was the largest shared-memory
// array of the intersection scene (24 x 20 x 16 B = 7.7 KB of 15.9 KB per env; 8 curved lanes: 3 KB; highway: none).
// The plain profile (highway: 4 lanes, register-limited occupancy) keeps the full table: there it costs no resident env
// and saves the recomputation (measured: 1.154 -> 1.188 ms/step without it).
template <class C> TT_HD d2 sr_of(C& c, int v, int l) {
    if (C::kPlain) return c.SR[v * c.NC + l];
    const ttrl_lane& ln = c.lanes[l];
    if (ln.kind == TTRL_LANE_STRAIGHT) {
        const d2 p = c.st->pos[v];
        const double dx = p.x - ln.ax, dy = p.y - ln.ay;
        return d2{dx * ln.dx + dy * ln.dy, dx * (-ln.dy) + dy * ln.dx};
    }
    const d2 cached = c.SR[v * c.NC + ln.cache_col];
    // a NaN lateral marks an entry the closest-lane search skipped (the vehicle is nowhere near that lane: closest_lane_tasks);
    // the rare reader of such an entry projects on demand
    if (cached.y != cached.y) return sr_curved_slow(ln, c.st->pos[v]);
    return cached;
}
template <class C> TT_HD double S_(C& c, int v, int l) { return sr_of(c, v, l).x; }
template <class C> TT_HD double R_(C& c, int v, int l) { return sr_of(c, v, l).y; }
// columns of the curved-lane cache (host side: ttrl_sim_create / the test emulator); returns their number
inline int assign_cache_columns(ttrl_lane* lanes, int n_lanes, int8_t* curved_lane) {
    int nc = 0;
    for (int l = 0; l < n_lanes; ++l) {
        if (lanes[l].kind == TTRL_LANE_STRAIGHT) { lanes[l].cache_col = -1; continue; }
        curved_lane[nc] = (int8_t)l;
        lanes[l].cache_col = nc++;
    }
    return nc;
}

// Rebuild row v of the table and return the closest lane: RoadNetwork.get_closest_lane_index
// (road.py:55-71, np.argmin keeps the FIRST minimum) over distance_with_heading (lane.py:132-147).
// Also returns the bitmask of the lanes the vehicle is "on" with margin 1 (neighbour_vehicles road.py:503).
template <class C>
TT_HD int table_row_and_closest(C& c, int v, uint64_t& on_mask) {
    const double px = c.st->pos[v].x, py = c.st->pos[v].y, hd = c.st->h[v];
    int best = 0;
    double bd = 0;
    uint64_t m = 0;
    TT_PRAGMA(unroll TT_LANE_UNROLL)
    for (int l = 0; l < c.L; ++l) {
        const ttrl_lane& ln = c.lanes[l];
        double s, r;
        if (C::kPlain) { const double dx = px - ln.ax, dy = py - ln.ay; s = dx * ln.dx + dy * ln.dy; r = dx * (-ln.dy) + dy * ln.dx; c.SR[v * c.NC + l] = d2{s, r}; }
        else if (ln.kind == TTRL_LANE_STRAIGHT) { const double dx = px - ln.ax, dy = py - ln.ay; s = dx * ln.dx + dy * ln.dy; r = dx * (-ln.dy) + dy * ln.dx; }
        else { lane_local_curved_inl(ln, px, py, s, r); c.SR[v * c.NC + ln.cache_col] = d2{s, r}; }
        if (lane_on_lane(ln, s, r, 1.0)) m |= 1ull << l;
        double ang = fabs(wrap_to_pi(hd - lane_heading_at_c<C>(ln, s)));
        double d = lane_distance_sr(ln, s, r) + 1.0 * ang;
        if (l == 0 || d < bd) { bd = d; best = l; }
    }
    on_mask = m;
    return best;
}
template <class C, class Exec>
TT_HD void publish_on_mask(C& c, Exec& ex, int v, uint64_t m) {
    while (m) {
#if defined(__CUDA_ARCH__)
        const int l = __ffsll((long long)m) - 1;
#else
        const int l = __builtin_ctzll(m);
#endif
        m &= m - 1;
        ex.atomic_or(&c.lmask[l * C::W + (v >> 5)], 1u << (v & 31));
    }
}
// ------------------------------------------------------------------------------------------------
// The closest-lane search of the sub-step, as three phases (non-plain profiles).  The curved-lane projection (atan2 / sin) is the
// most expensive part of a sub-step and is only needed for the few (vehicle, lane) pairs where the vehicle can be ON the lane
// (|r| <= width / 2 + 1) or the lane can be the argmin (|r| <= the best straight-lane distance; d >= |r| exactly, every other term
// of lane.py:127-147 being non-negative).  |r| costs one sqrt for an arc (|radius - rad|), a bound for a sine lane (|lat| - amplitude):
//   S  per vehicle: all straight lanes (argmin candidate, on-lane bits); curved lanes: the bound -> task queue, or NaN into the cache
//   T  per task:    the projection -> cache, on-lane bit                      (full lanes: ~1 trip instead of one iteration per arc)
//   M  per vehicle: distance of its queued lanes from the cache, merged into the argmin as (distance, lane index) pairs --
//                   np.argmin's first minimum (road.py:55-71) whatever the evaluation order.
// Scratch: candidate bits in mark[], straight-lane best distance / lane in acc2[] / tl_old[], the queue in mq_a (all idle here).
// ------------------------------------------------------------------------------------------------
template <class C> TT_HD uint16_t* arc_queue(C& c) { return reinterpret_cast<uint16_t*>(c.st->mq_a); }
template <class C> constexpr int arc_queue_cap() { return 9 * EnvState<C::V>::MB * (int)(sizeof(double) / sizeof(uint16_t)); }

template <class C, class Exec>
TT_HD void arc_project(C& c, Exec& ex, int v, int col) {
    const int l = c.sc->curved_lane[col];
    const ttrl_lane& ln = c.lanes[l];
    double s, r;
    lane_local_curved_inl(ln, c.st->pos[v].x, c.st->pos[v].y, s, r);
    c.SR[v * c.NC + col] = d2{s, r};
    if (lane_on_lane(ln, s, r, 1.0)) ex.atomic_or(&c.lmask[l * C::W + (v >> 5)], 1u << (v & 31));
}
template <class C, class Exec>
TT_HD void closest_lane_straight(C& c, Exec& ex, int v) {
    auto* st = c.st;
    const double px = st->pos[v].x, py = st->pos[v].y, hd = st->h[v];
    int best = -1;
    double bd = 0;
    uint64_t m = 0;
    for (int l = 0; l < c.L; ++l) {
        const ttrl_lane& ln = c.lanes[l];
        if (ln.kind != TTRL_LANE_STRAIGHT) continue;
        const double dx = px - ln.ax, dy = py - ln.ay;
        const double s = dx * ln.dx + dy * ln.dy, r = dx * (-ln.dy) + dy * ln.dx;
        if (lane_on_lane(ln, s, r, 1.0)) m |= 1ull << l;
        const double d = lane_distance_sr(ln, s, r) + 1.0 * fabs(wrap_to_pi(hd - ln.heading));
        if (best < 0 || d < bd) { bd = d; best = l; }
    }
    publish_on_mask(c, ex, v, m);
    uint32_t cand = 0;
    for (int col = 0; col < c.NC; ++col) {
        const ttrl_lane& ln = c.lanes[c.sc->curved_lane[col]];
        const double dx = px - ln.ax, dy = py - ln.ay;
        double lb;  // lower bound of |r|
        if (ln.kind == TTRL_LANE_CIRCULAR) lb = fabs(ln.radius - sqrt(dx * dx + dy * dy));
        else lb = fmax(fabs(dx * (-ln.dy) + dy * ln.dx) - fabs(ln.amplitude), 0.0);
        if (best < 0 || lb <= bd || lb <= ln.width / 2 + 1.0) {
            cand |= 1u << col;
            const int q = ex.atomic_add(&st->n_arc, 1);
            if (q < arc_queue_cap<C>()) arc_queue(c)[q] = (uint16_t)(v | (col << 8));
            else arc_project(c, ex, v, col);  // queue full: project right here
        } else {
            c.SR[v * c.NC + col] = d2{0.0, NAN};
        }
    }
    st->mark[v] = (int32_t)cand;
    st->acc2[v] = bd;
    st->tl_old[v] = best;
}
template <class C>
TT_HD void closest_lane_merge(C& c, int v) {
    auto* st = c.st;
    const double hd = st->h[v];
    double bd = st->acc2[v];
    int best = st->tl_old[v];
    uint32_t cand = (uint32_t)st->mark[v];
    while (cand) {
#if defined(__CUDA_ARCH__)
        const int col = __ffs((int)cand) - 1;
#else
        const int col = __builtin_ctz(cand);
#endif
        cand &= cand - 1;
        const int l = c.sc->curved_lane[col];
        const ttrl_lane& ln = c.lanes[l];
        const d2 sr = c.SR[v * c.NC + col];
        const double d = lane_distance_sr(ln, sr.x, sr.y) + 1.0 * fabs(wrap_to_pi(hd - lane_heading_at(ln, sr.x)));
        if (best < 0 || d < bd || (d == bd && l < best)) { bd = d; best = l; }
    }
    st->lane[v] = best;
    st->mark[v] = 0;
}
// the search for every vehicle of the env (after positions changed): replaces table_row_and_closest per vehicle
template <class C, class Exec>
TT_HD void closest_lane_tasks(C& c, Exec& ex) {
    auto* st = c.st;
    const int n = st->n;
    ex.parn(n, [&](int t) { closest_lane_straight(c, ex, t); });
    const int nq = st->n_arc < arc_queue_cap<C>() ? st->n_arc : arc_queue_cap<C>();
    ex.parn(nq, [&](int q) { const int rec = arc_queue(c)[q]; arc_project(c, ex, rec & 0xFF, rec >> 8); });
    ex.parn(n, [&](int t) { closest_lane_merge(c, t); });
}

// When the task form pays (measured, profiles/r2_step_scheduling.txt item 14): several curved lanes whose |r| bounds separate
// them -- the intersection's eight turn arcs (1.53 -> 1.47 ms/step).  Not for a ring (the roundabout's arcs share centre and radius, so
// a vehicle on the ring is a candidate of all of them: 2.01 -> 3.08 ms) nor for two arcs (u-turn: 0.71 -> 0.81, the extra phases cost
// more than they save).  Host side (ttrl_sim_create / the test emulator).
inline int use_arc_tasks(const ttrl_lane* lanes, int n_lanes, int n_curved) {
    if (n_curved < 4 || n_curved > 32) return 0;  // (candidate sets are 32-bit)
    for (int a = 0; a < n_lanes; ++a) {
        if (lanes[a].kind == TTRL_LANE_STRAIGHT) continue;
        if (lanes[a].kind != TTRL_LANE_CIRCULAR) return 0;  // sine lanes: the amplitude bound is too weak to separate them
        for (int b = a + 1; b < n_lanes; ++b)
            if (lanes[b].kind == TTRL_LANE_CIRCULAR && fabs(lanes[a].ax - lanes[b].ax) < 1e-9 && fabs(lanes[a].ay - lanes[b].ay) < 1e-9 &&
                fabs(lanes[a].radius - lanes[b].radius) < 0.5 * (lanes[a].width + lanes[b].width))
                return 0;                                    // two arcs of one circle
    }
    return 1;
}

// Rebuild the whole table + lane masks from positions (after load / compaction).
template <class C, class Exec>
TT_HD void rebuild_tables(C& c, Exec& ex) {
    ex.parn(c.L * C::W, [&](int k) { c.lmask[k] = 0; });
    ex.parn(c.st->n, [&](int t) {
        uint64_t m;
        (void)table_row_and_closest(c, t, m);
        publish_on_mask(c, ex, t, m);
    });
}

// ------------------------------------------------------------------------------------------------
// routes (packed: entry k in byte k; lane byte 0xFF = None)
// ------------------------------------------------------------------------------------------------
TT_HD int route_road_at(uint32_t w, int k) { return (w >> (8 * k)) & 0xFF; }
TT_HD int route_lane_at(uint32_t w, int k) { int b = (w >> (8 * k)) & 0xFF; return b == 0xFF ? -1 : b; }
// a whole route in registers (up to TTRL_ROUTE_CAP entries)
struct Route { uint32_t r[TTRL_ROUTE_WORDS], l[TTRL_ROUTE_WORDS]; };
TT_HD int route_road_at(const Route& q, int k) { return route_road_at(k < 4 ? q.r[0] : k < 8 ? q.r[1] : q.r[2], k & 3); }
TT_HD int route_lane_at(const Route& q, int k) { return route_lane_at(k < 4 ? q.l[0] : k < 8 ? q.l[1] : q.l[2], k & 3); }
template <class S> TT_HD Route route_of(const S* st, int i) {
    Route q;
    for (int w = 0; w < TTRL_ROUTE_WORDS; ++w) { q.r[w] = st->rroad[w][i]; q.l[w] = st->rlanew[w][i]; }
    return q;
}
template <class S> TT_HD void route_store(S* st, int i, const Route& q) {
    for (int w = 0; w < TTRL_ROUTE_WORDS; ++w) { st->rroad[w][i] = q.r[w]; st->rlanew[w][i] = q.l[w]; }
}
TT_HD void route_set(Route& q, int k, int road, int lane /* < 0: None */) {
    q.r[k >> 2] |= (uint32_t)(road & 0xFF) << (8 * (k & 3));
    q.l[k >> 2] |= (uint32_t)(lane < 0 ? 0xFF : lane & 0xFF) << (8 * (k & 3));
}
// route.pop(0) (road.py:100)
template <class S> TT_HD void route_pop(S* st, int i) {
    st->rroad[0][i] = (st->rroad[0][i] >> 8) | (st->rroad[1][i] << 24);
    st->rroad[1][i] = (st->rroad[1][i] >> 8) | (st->rroad[2][i] << 24);
    st->rroad[2][i] >>= 8;
    st->rlanew[0][i] = (st->rlanew[0][i] >> 8) | (st->rlanew[1][i] << 24);
    st->rlanew[1][i] = (st->rlanew[1][i] >> 8) | (st->rlanew[2][i] << 24);
    st->rlanew[2][i] >>= 8;
    st->rlen[i] -= 1;
}

// RoadNetwork.next_lane_given_next_road road.py:138-157 (next_id < 0 means None)
template <class C>
TT_HD int next_lane_given_next_road(C& c, int road, int id, int next_road, int next_id, double px, double py, double& dist) {
    const ttrl_road& nr = c.sc->roads[next_road];
    if (c.sc->roads[road].n_lanes == nr.n_lanes) {
        if (next_id < 0) next_id = id;
    } else {
        double bd = 0;
        int b = 0;
        for (int l = 0; l < nr.n_lanes; ++l) {
            double d = lane_distance(c.lanes[nr.first_lane + l], px, py);
            if (l == 0 || d < bd) { bd = d; b = l; }
        }
        next_id = b;
    }
    dist = lane_distance(c.lanes[nr.first_lane + next_id], px, py);
    return next_id;
}

// RoadNetwork.next_lane road.py:73-136; mutates the vehicle's route (route.pop(0), :100)
template <class C>
TT_HDN int next_lane(C& c, int i, int cur) {
    auto* st = c.st;
    const ttrl_lane& cl = c.lanes[cur];
    const int road = cl.road, id = cl.lane_id;
    const int to = c.sc->roads[road].to_node;
    int next_road = -1, next_id = -1;
    if (st->rlen[i] > 0) {
        if (route_road_at(st->rroad[0][i], 0) == road) route_pop(st, i);
        if (st->rlen[i] > 0 && c.sc->roads[route_road_at(st->rroad[0][i], 0)].from_node == to) {
            next_road = route_road_at(st->rroad[0][i], 0);
            next_id = route_lane_at(st->rlanew[0][i], 0);
        }
    }
    double lon = S_(c, i, cur), qx, qy;
    lane_position(cl, lon, 0.0, qx, qy);
    if (next_road < 0) {
        const int a = c.sc->node_first[to], b = c.sc->node_first[to + 1];
        if (a == b) return cur;  // KeyError branch (:129-130)
        double bd = 0;
        int br = -1, bid = 0;
        for (int k = a; k < b; ++k) {
            double d;
            int nid = next_lane_given_next_road(c, road, id, c.sc->node_roads[k], -1, qx, qy, d);
            if (k == a || d < bd) { bd = d; br = c.sc->node_roads[k]; bid = nid; }
        }
        next_road = br;
        next_id = bid;
    } else {
        double d;
        next_id = next_lane_given_next_road(c, road, id, next_road, next_id, qx, qy, d);
    }
    return c.sc->roads[next_road].first_lane + next_id;
}

// ControlledVehicle.follow_road controller.py:135-143 (+ after_end lane.py:120-125)
template <class C>
TT_HD void follow_road(C& c, int i) {
    const int tl = c.st->tlane[i];
    if (S_(c, i, tl) > c.lanes[tl].length - kVehLength / 2) c.st->tlane[i] = next_lane(c, i, tl);
}

// ControlledVehicle.steering_control controller.py:145-187, clipped to +-MAX_STEERING_ANGLE (controller.py:131).
// tan(asin(w)) is evaluated as w / sqrt((1-w)(1+w)) (same value to a few ulp, no asin/tan pair), and the tangent of
// the returned command comes out as a by-product for Vehicle.step's beta = atan(tan(delta) / 2).
template <class C>
TT_STEER double steering_control(C& c, int i, int target_lane, double& tan_steer) {
    const double TAU_PURSUIT = 0.5 * 0.2, KP_LATERAL = 1 / 0.6, KP_HEADING = 1 / 0.2;
    const double MAX_STEER = kPi / 3;
    const ttrl_lane& tl = c.lanes[target_lane];
    const d2 sr = sr_of(c, i, target_lane);
    const double speed = c.st->v[i];
    double lane_next = sr.x + speed * TAU_PURSUIT;
    double lane_future_heading = lane_heading_at_c<C>(tl, lane_next);
    if (is_linear(c) && !(c.st->flags[i] & TTRL_FL_MDP)) {
        // LinearVehicle.steering_control / steering_features (behavior.py:466-500): linear in STEERING_PARAMETERS; clipped to
        // +-MAX_STEERING_ANGLE by IDMVehicle.act (behavior.py:114-116)
        const double f0 = div_z(wrap_to_pi(lane_future_heading - c.st->h[i]) * kVehLength, not_zero(speed));
        const double nz = not_zero(speed);
        const double f1 = div_z(-sr.y * kVehLength, nz * nz);
        double steering_angle = c.lin[TTRL_LIN_STEER0 * C::V + i] * f0 + c.lin[TTRL_LIN_STEER1 * C::V + i] * f1;
        steering_angle = clipd(steering_angle, -MAX_STEER, MAX_STEER);
        tan_steer = tan(steering_angle);
        return steering_angle;
    }
    double lateral_speed_command = -KP_LATERAL * sr.y;
    double heading_command = asin(clipd(div_z(lateral_speed_command, not_zero(speed)), -1.0, 1.0));
    double heading_ref = lane_future_heading + clipd(heading_command, -kPi / 4, kPi / 4);
    double heading_rate_command = KP_HEADING * wrap_to_pi(heading_ref - c.st->h[i]);
    const double w = clipd(kVehLength / 2 / not_zero(speed) * heading_rate_command, -1.0, 1.0);  // sin(slip_angle)
    const double tan2 = 2 * div_z(w, sqrt((1 - w) * (1 + w)));                                      // 2 tan(slip_angle)
    double steering_angle = atan(tan2);
    if (steering_angle > MAX_STEER) { tan_steer = c.tan_max_steer; return MAX_STEER; }
    if (steering_angle < -MAX_STEER) { tan_steer = -c.tan_max_steer; return -MAX_STEER; }
    tan_steer = tan2;
    return steering_angle;
}

// MDPVehicle.speed_to_index controller.py:326-344 (np.round -> round half to even -> rint)
TT_HD int speed_to_index(const ttrl_config& cfg, double speed) {
    const int n = cfg.n_target_speeds;
    double x = (speed - cfg.target_speeds[0]) / (cfg.target_speeds[n - 1] - cfg.target_speeds[0]);
    return (int)clipd(rint(x * (n - 1)), 0.0, (double)(n - 1));
}

enum { A_NONE = 0, A_IDLE, A_LANE_LEFT, A_LANE_RIGHT, A_FASTER, A_SLOWER };
// number of controlled vehicles (plain scenes: always one)
template <class C> TT_HD int n_agents(const C& c) {
    if (!C::kMulti) return 1;
    const int k = c.sc->cfg.controlled_vehicles;
    return k < 1 ? 1 : (k > TTRL_MAX_CONTROLLED ? TTRL_MAX_CONTROLLED : k);
}
TT_HD int agent_of(int flags) { return (flags & TTRL_FL_AGENT_MASK) >> TTRL_FL_AGENT_SHIFT; }
TT_HD int decode_action(const ttrl_config& cfg, int a) {  // action.py:204-211
    if (a < 0) return A_NONE;
    if (cfg.action_mode == TTRL_ACT_ALL) return a == 0 ? A_LANE_LEFT : a == 1 ? A_IDLE : a == 2 ? A_LANE_RIGHT : a == 3 ? A_FASTER : A_SLOWER;
    if (cfg.action_mode == TTRL_ACT_LONGI) return a == 0 ? A_SLOWER : a == 1 ? A_IDLE : A_FASTER;
    return a == 0 ? A_LANE_LEFT : a == 1 ? A_IDLE : A_LANE_RIGHT;
}

// ControlledVehicle.act controller.py:89-133, WITHOUT its controller tail: the steering / speed commands of
// an act() call are overwritten by the next call, so only the last call of a sub-step (Road.act) computes them
// (controlled_commands, phase C).  The target-lane side effects of every call are kept.
template <class C>
TT_HD void controlled_act_lanes(C& c, int i, int action) {
    auto* st = c.st;
    follow_road(c, i);
    if (action == A_LANE_RIGHT || action == A_LANE_LEFT) {
        const ttrl_lane& tl = c.lanes[st->tlane[i]];
        const ttrl_road& rd = c.sc->roads[tl.road];
        int id = tl.lane_id + (action == A_LANE_RIGHT ? 1 : -1);
        id = id < 0 ? 0 : (id > rd.n_lanes - 1 ? rd.n_lanes - 1 : id);
        const int cand = rd.first_lane + id;
        if (lane_reachable(c.lanes[cand], S_(c, i, cand), R_(c, i, cand))) st->tlane[i] = cand;
    }
}
// MDPVehicle.act controller.py:295-315 (lane / speed-index side effects)
template <class C>
TT_HD void mdp_act_lanes(C& c, int i, int action) {
    auto* st = c.st;
    const ttrl_config& cfg = c.sc->cfg;
    if (action == A_FASTER) st->sidx[i] = speed_to_index(cfg, st->v[i]) + 1;
    else if (action == A_SLOWER) st->sidx[i] = speed_to_index(cfg, st->v[i]) - 1;
    else { controlled_act_lanes(c, i, action); return; }
    const int n = cfg.n_target_speeds;
    st->sidx[i] = st->sidx[i] < 0 ? 0 : (st->sidx[i] > n - 1 ? n - 1 : st->sidx[i]);
    st->tspeed[i] = cfg.target_speeds[st->sidx[i]];
    controlled_act_lanes(c, i, A_NONE);
}

// ------------------------------------------------------------------------------------------------
// IDM / MOBIL: ttrl_env/vehicle/behavior.py
// ------------------------------------------------------------------------------------------------
// Road.neighbour_vehicles road.py:480-513 served from the table; candidates come from the lane's on-lane
// bitmask in increasing index order, which keeps the reference's tie rules (front: a later vehicle with an
// equal s replaces, `<=` :507; rear: the earlier one stays, `>` :510).
template <class C>
TT_HD void neighbours(C& c, int i, int lane, int& front, int& rear) {
    const double s = S_(c, i, lane);
    double sf = 0, sr = 0;
    int f = -1, r = -1;
#pragma unroll 1
    for (int w = 0; w < C::W; ++w) {
        uint32_t m = c.lmask[lane * C::W + w];
        if ((i >> 5) == w) m &= ~(1u << (i & 31));
        while (m) {
#if defined(__CUDA_ARCH__)
            const int b = __ffs((int)m) - 1;
#else
            const int b = __builtin_ctz(m);
#endif
            m &= m - 1;
            const int j = w * 32 + b;
            const double sv = S_(c, j, lane);
            if (s <= sv) { if (f < 0 || sv <= sf) { sf = sv; f = j; } }
            else if (r < 0 || sv > sr) { sr = sv; r = j; }
        }
    }
    front = f;
    rear = r;
}
// IDMVehicle.desired_gap behavior.py:192-217
template <class C>
TT_HD double desired_gap(C& c, int ego, int front) {
    const ttrl_config& cfg = c.sc->cfg;
    auto* st = c.st;
    const d2 e = st->cs[ego], f = st->cs[front];
    const double ve = st->v[ego], vf = st->v[front];
    const double dvx = ve * e.x - vf * f.x, dvy = ve * e.y - vf * f.y;
    const double dv = dvx * e.x + dvy * e.y;
    return cfg.distance_wanted + ve * cfg.time_wanted + div_z(ve * dv, c.gap_den);
}
// IDMVehicle.acceleration behavior.py:150-190 as vehicle `self` evaluates it for `ego` (its own DELTA / parameters, also when
// ego is another vehicle: MOBIL); LinearVehicle.acceleration behavior.py:416-464 when the traffic is linear
template <class C>
TT_IDM double idm_acceleration(C& c, int self, int ego, int front) {
    if (ego < 0) return 0.0;
    const ttrl_config& cfg = c.sc->cfg;
    auto* st = c.st;
    const int le = st->lane[ego];
    if (is_linear(c)) {
        // acceleration_features: vt = target_speed - speed, dv = min(front.speed - speed, 0), dp = min(d - d_safe, 0) with
        // d_safe = DISTANCE_WANTED + max(speed, 0) TIME_WANTED; np.dot(ACCELERATION_PARAMETERS, [vt, dv, dp])
        const double vt = st->tspeed[ego] - st->v[ego];
        double dv = 0, dp = 0;
        if (front >= 0) {
            const double d_safe = cfg.distance_wanted + fmax(st->v[ego], 0.0) * cfg.time_wanted;
            const double d = S_(c, front, le) - S_(c, ego, le);
            dv = fmin(st->v[front] - st->v[ego], 0.0);
            dp = fmin(d - d_safe, 0.0);
        }
        return (c.lin[TTRL_LIN_ACC0 * C::V + self] * vt + c.lin[TTRL_LIN_ACC1 * C::V + self] * dv) + c.lin[TTRL_LIN_ACC2 * C::V + self] * dp;
    }
    const double self_delta = st->delta[self];
    const double ts = clipd(st->tspeed[ego], 0.0, c.lanes[le].speed_limit);
    // np.power(x, delta) for x >= 0, delta in [3.5, 4.5] as exp(delta log x): a few ulp from pow, a third of its cost
    double acc = cfg.comfort_acc_max * (1 - exp(self_delta * log(div_z(fmax(st->v[ego], 0.0), fabs(not_zero(ts))))));
    if (front >= 0) {
        const double d = S_(c, front, le) - S_(c, ego, le);  // lane_distance_to objects.py:182-197
        const double q = desired_gap(c, ego, front) / not_zero(d);
        acc -= cfg.comfort_acc_max * (q * q);
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// ACT.  Reproduces the reference's sequential Road.act (road.py:461-464) with parallel phases.  The only
// cross-vehicle write->read coupling inside act() is the abort rule of change_lane_policy
// (behavior.py:229-244), which reads OTHER vehicles' target_lane_index: vehicles earlier in the list have
// already acted (new value), later ones have not (old value).
//   A   every vehicle: follow_road; ego: meta-action side effects; classify: ongoing lane change (-> phase B)
//       or lane-change timer fired (-> MOBIL batch)
//   M   MOBIL over the fired vehicles: neighbour-query tasks, IDM-evaluation tasks, one decision per vehicle
//   B   vehicles in an ongoing lane change, in list order: abort test against the mixed old/new view
//   C   every vehicle: steering + own-lane front neighbour; then IDM-evaluation tasks (own lane for all,
//       target lane for the vehicles changing lane)
// ------------------------------------------------------------------------------------------------
template <class C, class Exec>
TT_HD void act_phase_a(C& c, Exec& ex, int i, const int32_t* actions) {
    auto* st = c.st;
    st->mark[i] = 0;
    if (st->flags[i] & TTRL_FL_MDP) {
        // DiscreteMetaAction.act (action.py:259-260; MultiAgentAction.act :320-323 for K agents) runs BEFORE Road.act:
        // its target-lane change is the "old" value every other vehicle sees; then Road.act -> MDPVehicle.act(None)
        int first_action = A_NONE;
        if (actions) {  // non-null on the first sub-step of an env-step only
            if (n_agents(c) == 1) { if (i == st->ego) first_action = decode_action(c.sc->cfg, actions[0]); }
            else if (st->flags[i] & TTRL_FL_CONTROLLED) first_action = decode_action(c.sc->cfg, actions[agent_of(st->flags[i])]);
        }
        if (first_action != A_NONE) mdp_act_lanes(c, i, first_action);
        st->tl_old[i] = st->tlane[i];
        mdp_act_lanes(c, i, A_NONE);
        return;
    }
    st->tl_old[i] = st->tlane[i];
    if (st->flags[i] & TTRL_FL_CRASHED) return;  // behavior.py:102-103
    follow_road(c, i);
    const int ln = st->lane[i];
    if (ln != st->tlane[i]) {  // ongoing change: behavior.py:229-244, resolved in phase B
        if (c.lanes[ln].road == c.lanes[st->tlane[i]].road) ex.atomic_or(&st->bmask[i >> 5], 1u << (i & 31));
        return;
    }
    if (!(c.sc->cfg.lane_change_delay < st->timer[i])) return;  // utils.do_every utils.py:25-26
    st->timer[i] = 0;
    // side_lanes road.py:200-211 (id-1 then id+1), reachable (lane.py:104-118), abs(speed) >= 1 (behavior.py:257)
    const ttrl_lane& l = c.lanes[ln];
    const int nl = c.sc->roads[l.road].n_lanes;
    int valid = 0;
    if (!(fabs(st->v[i]) < 1)) {
        if (l.lane_id > 0 && lane_reachable(c.lanes[ln - 1], S_(c, i, ln - 1), R_(c, i, ln - 1))) valid |= 1;
        if (l.lane_id < nl - 1 && lane_reachable(c.lanes[ln + 1], S_(c, i, ln + 1), R_(c, i, ln + 1))) valid |= 2;
    }
    if (valid) st->mob_veh[ex.atomic_add(&st->n_mob, 1)] = i | (valid << 16);
}

// IDMVehicle.mobil behavior.py:265-324 for the vehicles of one batch (ranks base .. base+nb-1 of mob_veh).
template <class C, class Exec>
TT_HD void mobil_batch(C& c, Exec& ex, int base, int nb) {
    auto* st = c.st;
    const ttrl_config& cfg = c.sc->cfg;
    // M1: neighbour queries: q = 0 own lane, 1 left candidate, 2 right candidate
    ex.parn(3 * nb, [&](int k) {
        const int b = k / 3, q = k - 3 * b;
        const int rec = st->mob_veh[base + b], i = rec & 0xFFFF, valid = rec >> 16;
        int f = -1, r = -1;
        if (q == 0 || (valid >> (q - 1)) & 1) neighbours(c, i, st->lane[i] + (q == 0 ? 0 : q == 1 ? -1 : 1), f, r);
        st->mq_f[k] = (int16_t)f;
        st->mq_r[k] = (int16_t)r;
    });
    // M2: IDM evaluations.  e = 0 self_a, 1 old_following_a, 2 old_following_pred_a; 3+3k+{0,1,2} =
    // new_following_a, new_following_pred_a, self_pred_a of candidate k
    ex.parn(9 * nb, [&](int k) {
        const int b = k / 9, e = k - 9 * b;
        const int rec = st->mob_veh[base + b], i = rec & 0xFFFF, valid = rec >> 16;
        const int old_prec = st->mq_f[3 * b], old_foll = st->mq_r[3 * b];
        int ego = -1, front = -1;
        if (e == 0) { ego = i; front = old_prec; }
        else if (e == 1) { ego = old_foll; front = i; }
        else if (e == 2) { ego = old_foll; front = old_prec; }
        else {
            const int kk = (e - 3) / 3, w = (e - 3) - 3 * kk;
            if ((valid >> kk) & 1) {
                const int new_prec = st->mq_f[3 * b + 1 + kk], new_foll = st->mq_r[3 * b + 1 + kk];
                if (w == 0) { ego = new_foll; front = new_prec; }
                else if (w == 1) { ego = new_foll; front = i; }
                else { ego = i; front = new_prec; }
            }
        }
        st->mq_a[k] = idm_acceleration(c, i, ego, front);
    });
    // M3: decisions, candidates in side_lanes order; a later accepted candidate overwrites (behavior.py:250-263)
    ex.parn(nb, [&](int b) {
        const int rec = st->mob_veh[base + b], i = rec & 0xFFFF, valid = rec >> 16;
        const double* a = st->mq_a + 9 * b;
        const double self_a = a[0], old_following_a = a[1], old_following_pred_a = a[2];
        for (int k = 0; k < 2; ++k) {
            if (!((valid >> k) & 1)) continue;
            const int cand = st->lane[i] + (k == 0 ? -1 : 1);
            const double new_following_a = a[3 + 3 * k], new_following_pred_a = a[4 + 3 * k], self_pred_a = a[5 + 3 * k];
            if (new_following_pred_a < -cfg.lane_change_max_braking_imposed) continue;
            if (st->rlen[i] > 0 && route_lane_at(st->rlanew[0][i], 0) >= 0) {
                const int cur_t = c.lanes[st->tlane[i]].lane_id;
                const int want = route_lane_at(st->rlanew[0][i], 0) - cur_t, dir = c.lanes[cand].lane_id - cur_t;
                const int sw = (want > 0) - (want < 0), sd = (dir > 0) - (dir < 0);
                if (sd != sw) continue;
                if (self_pred_a < -cfg.lane_change_max_braking_imposed) continue;
            } else {
                const double jerk = self_pred_a - self_a +
                                    cfg.politeness * (new_following_pred_a - new_following_a + old_following_pred_a - old_following_a);
                if (jerk < cfg.lane_change_min_acc_gain) continue;
            }
            st->tlane[i] = cand;
        }
    });
}

// abort predicate of vehicle i against vehicle j (behavior.py:232-243)
template <class C>
TT_HD bool abort_pred(C& c, int i, int j) {
    auto* st = c.st;
    if (j == i) return false;
    const int ti = st->tlane[i];
    const int view = j < i ? st->tlane[j] : st->tl_old[j];
    if (!(st->lane[j] != ti && view == ti)) return false;
    const int li = st->lane[i];
    const double d = S_(c, j, li) - S_(c, i, li);
    const double d_star = desired_gap(c, i, j);
    return 0 < d && d < d_star;
}

// phase C1: steering for every vehicle (+ speed control for the ego), own-lane front neighbour
template <class C, class Exec>
TT_HD void act_phase_c1(C& c, Exec& ex, int i) {
    auto* st = c.st;
    const int fl = st->flags[i];
    st->fo[i] = -1;
    if (!(fl & TTRL_FL_MDP) && (fl & TTRL_FL_CRASHED)) return;  // IDMVehicle.act returns early when crashed
    const int tl = st->tlane[i];
    double tan_steer;
    st->steer[i] = steering_control(c, i, tl, tan_steer);  // clipped: controller.py:131, behavior.py:114-116
    st->tsteer[i] = tan_steer;
    if (fl & TTRL_FL_MDP) {
        st->acc[i] = (1 / 0.6) * (st->tspeed[i] - st->v[i]);  // speed_control controller.py:189-198
        return;
    }
    int f, r;
    neighbours(c, i, st->lane[i], f, r);
    st->fo[i] = (int16_t)f;
    if (st->lane[i] != tl) st->chg[ex.atomic_add(&st->n_chg, 1)] = i;
}
// phase C2 task k: k < n -> IDM acceleration of vehicle k on its own lane; k >= n -> of a lane-changing vehicle
// w.r.t. the front vehicle on its target lane (behavior.py:118-137)
template <class C>
TT_HD void act_phase_c2(C& c, int k) {
    auto* st = c.st;
    const int n = st->n;
    int i, front;
    if (k < n) {
        i = k;
        if (st->flags[i] & (TTRL_FL_MDP | TTRL_FL_CRASHED)) return;
        front = st->fo[i];
    } else {
        i = st->chg[k - n];
        int r;
        neighbours(c, i, st->tlane[i], front, r);
    }
    const double a = idm_acceleration(c, i, i, front);
    if (k < n) st->acc[i] = a; else st->acc2[i] = a;
}

// ------------------------------------------------------------------------------------------------
// INTEGRATE: IDMVehicle.step behavior.py:139-148 -> Vehicle.step kinematics.py:130-153, clip_actions :155-168,
// on_state_update :170-177 (closest lane via the rebuilt table row)
// ------------------------------------------------------------------------------------------------
template <class C, class Exec>
TT_HD void integrate(C& c, Exec& ex, int i) {
    auto* st = c.st;
    const double dt = c.sc->dt;
    int fl = st->flags[i];
    double steer = st->steer[i], acc = st->acc[i];
    if (!(fl & TTRL_FL_MDP)) {
        st->timer[i] += dt;
        if (!(fl & TTRL_FL_CRASHED)) {  // tail of IDMVehicle.act (behavior.py:130-137)
            if (st->lane[i] != st->tlane[i]) acc = fmin(acc, st->acc2[i]);
            acc = clipd(acc, -c.sc->cfg.acc_max, c.sc->cfg.acc_max);
        }
    }
    const double speed = st->v[i];
    if (fl & TTRL_FL_CRASHED) { steer = 0; acc = -1.0 * speed; }
    if (speed > 40.0) acc = fmin(acc, 1.0 * (40.0 - speed));
    else if (speed < -40.0) acc = fmax(acc, 1.0 * (-40.0 - speed));
    st->steer[i] = steer;
    st->acc[i] = acc;
    // beta = arctan(1/2 tan(delta)) (kinematics.py:143): only cos/sin of beta and of heading+beta are needed
    const double tb = (fl & TTRL_FL_CRASHED) ? 0.0 : 0.5 * st->tsteer[i];
    const double cbeta = 1.0 / sqrt(1.0 + tb * tb), sbeta = tb * cbeta;
    const double hd = st->h[i];
    const d2 hc = st->cs[i];
    const double vx = speed * (hc.x * cbeta - hc.y * sbeta), vy = speed * (hc.y * cbeta + hc.x * sbeta);
    double px = st->pos[i].x + vx * dt, py = st->pos[i].y + vy * dt;
    if (fl & TTRL_FL_HAS_IMPACT) {
        px += st->imp[i].x;
        py += st->imp[i].y;
        fl = (fl | TTRL_FL_CRASHED) & ~TTRL_FL_HAS_IMPACT;
        st->imp[i] = d2{0, 0};
    }
    const double nh = hd + div_z(speed * sbeta, kVehLength / 2) * dt;
    const double nv = speed + acc * dt;
    st->pos[i] = d2{px, py};
    st->h[i] = nh;
    st->v[i] = nv;
    double sn, cn;
    sincos(nh, &sn, &cn);
    st->cs[i] = d2{cn, sn};
    st->flags[i] = fl;
    if (C::kPlain || !c.sc->arc_tasks) {  // (else: closest_lane_tasks, after every vehicle has moved)
        uint64_t m;
        st->lane[i] = table_row_and_closest(c, i, m);
        publish_on_mask(c, ex, i, m);
    }
    // collision pre-check guard (objects.py:123-126): radius (diag_i + diag_j)/2 + speed dt of the LOWER index
    const double diag = sqrt(kVehLength * kVehLength + kVehWidth * kVehWidth);
    const double thr = (diag + diag) / 2 + nv * dt;
    st->thr2[i] = thr < 0 ? -1.0 : thr * thr * (1.0 + 1e-12);
    st->best[i] = -1;
}

// ------------------------------------------------------------------------------------------------
// COLLIDE: Road.step's pair loop road.py:474-476 -> handle_collisions objects.py:91-137 -> SAT utils.py:175-239.
// In the reference the pair (i<j) loop assigns `impact` to both members, so for vehicle k the LAST assignment
// comes from its will-intersect partner with the largest index (pairs involving k are visited in increasing
// partner order); `crashed` is an OR.
//   K1  half-ring scan: vehicle k tests partners k+1 .. k+n/2 (mod n) -> candidate pair queue.  The exact test
//       `norm(dp) > thr` is guarded by a squared-distance comparison that is conservative by 1e-12 relative.
//   K2  8 lanes per candidate pair, one separating axis each (ex.sat_pairs); crashed flags OR-ed, will-intersect
//       pairs appended to a list with their translation.
//   K3  best[k] = max will-intersect partner;  K4  the winning pair writes the impact.
// Queue overflow (pathological pile-ups) falls back to the plain per-vehicle loop `collide_serial`.
// ------------------------------------------------------------------------------------------------
struct Poly { double p[5][2]; };
template <class C>
TT_HD void vehicle_polygon(C& c, int i, Poly& o) {  // objects.py:168-180
    const double cs = c.st->cs[i].x, sn = c.st->cs[i].y, px = c.st->pos[i].x, py = c.st->pos[i].y;
    const double hx[4] = {-kVehLength / 2, -kVehLength / 2, +kVehLength / 2, +kVehLength / 2};
    const double hy[4] = {-kVehWidth / 2, +kVehWidth / 2, +kVehWidth / 2, -kVehWidth / 2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        o.p[k][0] = (cs * hx[k] + (-sn) * hy[k]) + px;
        o.p[k][1] = (sn * hx[k] + cs * hy[k]) + py;
    }
    o.p[4][0] = o.p[0][0];
    o.p[4][1] = o.p[0][1];
}
TT_HD void project_polygon(const Poly& g, double nx, double ny, double& mn, double& mx) {  // utils.py:175-183
    mn = mx = g.p[0][0] * nx + g.p[0][1] * ny;
#pragma unroll
    for (int k = 1; k < 4; ++k) {  // the closing 5th point repeats the first: it cannot change min / max
        const double pr = g.p[k][0] * nx + g.p[k][1] * ny;
        if (pr < mn) mn = pr;
        if (pr > mx) mx = pr;
    }
}
TT_HD double interval_distance(double min_a, double max_a, double min_b, double max_b) {  // utils.py:186-191
    return min_a < min_b ? min_b - max_a : min_a - max_b;
}
// One separating-axis test of utils.are_polygons_intersecting (utils.py:194-239): axis = edge (axis & 3) of
// polygon a (axis < 4) or b.  fl bit0: separated now, bit1: separated after the relative displacement.
// The reference's early `break` only skips work once both flags are already false, and its running
// `min_distance` matters only when will_intersect stays true (no break happened), so the 8 axes are independent:
//   intersecting = no axis has bit0, will_intersect = no axis has bit1, translation = first axis of minimal absd.
struct AxisRes { double absd, nx, ny; int fl; };
template <class C>
TT_HD AxisRes sat_axis(C& c, int ia, int ib, int axis) {
    auto* st = c.st;
    const double dt = c.sc->dt;
    Poly a, b;
    vehicle_polygon(c, ia, a);
    vehicle_polygon(c, ib, b);
    const double rvx = st->v[ia] * st->cs[ia].x * dt - st->v[ib] * st->cs[ib].x * dt;
    const double rvy = st->v[ia] * st->cs[ia].y * dt - st->v[ib] * st->cs[ib].y * dt;
    const Poly& pg = axis < 4 ? a : b;
    const int k = axis & 3;
    double nx = -pg.p[k + 1][1] + pg.p[k][1], ny = pg.p[k + 1][0] - pg.p[k][0];
    const double nn = sqrt(nx * nx + ny * ny);
    nx /= nn;
    ny /= nn;
    double min_a, max_a, min_b, max_b;
    project_polygon(a, nx, ny, min_a, max_a);
    project_polygon(b, nx, ny, min_b, max_b);
    AxisRes res;
    res.fl = interval_distance(min_a, max_a, min_b, max_b) > 0 ? 1 : 0;
    const double vp = nx * rvx + ny * rvy;
    if (vp < 0) min_a += vp; else max_a += vp;
    const double distance = interval_distance(min_a, max_a, min_b, max_b);
    if (distance > 0) res.fl |= 2;
    res.absd = fabs(distance);
    // orientation by the centre difference (utils.py:233-236)
    const double cax = (((a.p[0][0] + a.p[1][0]) + a.p[2][0]) + a.p[3][0]) / 4, cay = (((a.p[0][1] + a.p[1][1]) + a.p[2][1]) + a.p[3][1]) / 4;
    const double cbx = (((b.p[0][0] + b.p[1][0]) + b.p[2][0]) + b.p[3][0]) / 4, cby = (((b.p[0][1] + b.p[1][1]) + b.p[2][1]) + b.p[3][1]) / 4;
    if ((cax - cbx) * nx + (cay - cby) * ny > 0) { res.nx = nx; res.ny = ny; } else { res.nx = -nx; res.ny = -ny; }
    return res;
}
// exact pre-check objects.py:123-126 for the pair (lo < hi): true = the pair needs the SAT
template <class C>
TT_HD bool collide_candidate(C& c, int lo, int hi) {
    auto* st = c.st;
    const d2 a = st->pos[lo], b = st->pos[hi];
    const double dx = b.x - a.x, dy = b.y - a.y;
    if (fma(dy, dy, dx * dx) > st->thr2[lo]) return false;  // conservative guard (fma only here)
    const double diag = sqrt(kVehLength * kVehLength + kVehWidth * kVehWidth);
    const double thr = (diag + diag) / 2 + st->v[lo] * c.sc->dt;
    return !(sqrt(dx * dx + dy * dy) > thr);
}
// plain per-vehicle loop (fallback on queue overflow; same results)
template <class C>
TT_HDN void collide_serial(C& c, int k) {
    auto* st = c.st;
    const int n = st->n;
    bool crashed = false, has_imp = false;
    double ix = 0, iy = 0;
    for (int p = 0; p < n; ++p) {
        if (p == k) continue;
        const int lo = k < p ? k : p, hi = k < p ? p : k;
        if (!collide_candidate(c, lo, hi)) continue;
        bool inter = true, will = true;
        double best = INFINITY, tx = 0, ty = 0;
        for (int ax = 0; ax < 8; ++ax) {
            const AxisRes r = sat_axis(c, lo, hi, ax);
            if (r.fl & 1) inter = false;
            if (r.fl & 2) will = false;
            if (r.absd < best) { best = r.absd; tx = r.nx; ty = r.ny; }
        }
        if (will) {  // self.impact = +t/2, other.impact = -t/2 (objects.py:110-111)
            has_imp = true;
            ix = lo == k ? best * tx / 2 : -(best * tx) / 2;
            iy = lo == k ? best * ty / 2 : -(best * ty) / 2;
        }
        if (inter) crashed = true;
    }
    if (has_imp) { st->imp[k] = d2{ix, iy}; st->flags[k] |= TTRL_FL_HAS_IMPACT; }
    if (crashed) st->flags[k] |= TTRL_FL_CRASHED;
}
template <class C>
TT_HD void advance_substep_counters(C& c) {
    c.st->steps += 1;
    if (!C::kPlain && c.sc->cfg.regulated) c.st->road_steps += 1;
}
template <class C, class Exec>
TT_HD void collide_all(C& c, Exec& ex) {
    auto* st = c.st;
    const int n = st->n;
    // K0: float copies for the pre-filter.  |x_hi - x_lo| > radius_lo + 1 cm in float implies the exact squared-distance guard
    // rejects the pair (float rounding of positions up to 1e4 m is < 1 mm): the filter can only skip pairs the guard would skip.
    // (x, guard radius) per vehicle as float pairs, the list written twice back to back (entries k and n + k) so that the half-ring
    // scan indexes k + m without wrapping.  Storage: acc2 and tsteer, both dead between the integration and the next controls.
    static_assert(offsetof(EnvState<C::V>, tsteer) == offsetof(EnvState<C::V>, acc2) + sizeof(double) * C::V, "acc2 and tsteer must be adjacent");
    f2* xr = reinterpret_cast<f2*>(st->acc2);
    ex.parn(n, [&](int k) {
        const f2 v = f2{(float)st->pos[k].x, st->thr2[k] < 0 ? -1.0f : (float)sqrt(st->thr2[k]) * 1.000001f + 0.01f};
        xr[k] = v;
        xr[n + k] = v;
    });
    // K1 (vehicle 0's thread also advances the sub-step counters: nothing below reads them, and the barrier that ends this loop
    // publishes them -- one team barrier per sub-step less than a separate update)
    if (n == 0) { if (ex.first()) advance_substep_counters(c); ex.sync(); return; }
    ex.parn(n, [&](int k) {
        if (k == 0) advance_substep_counters(c);
        const int half = n / 2;
        const int mmax = (2 * half == n && k >= half) ? half - 1 : half;  // even n: the diametral pair is visited once
        const f2 me = xr[k];
        for (int q = k + 1; q <= k + mmax; ++q) {
            const f2 other = xr[q];
            if (fabsf(other.x - me.x) > (q < n ? me.y : other.y)) continue;  // the guard radius is the lower index's
            const int p = q < n ? q : q - n;
            const int lo = k < p ? k : p, hi = k < p ? p : k;
            if (!collide_candidate(c, lo, hi)) continue;
            const int slot = ex.atomic_add(&st->n_pair, 1);
            if (slot < EnvState<C::V>::PQ) st->pairq[slot] = lo | (hi << 16);
            else st->overflow = 1;
        }
    });
    const int np = st->n_pair;
    if (np == 0) return;
    if (!st->overflow) {
        // K2
        ex.sat_pairs(np,
            [&](int q, int axis) { const int pr = st->pairq[q]; return sat_axis(c, pr & 0xFFFF, pr >> 16, axis); },
            [&](int q, bool inter, bool will, double absd, double nx, double ny) {
                const int pr = st->pairq[q], lo = pr & 0xFFFF, hi = pr >> 16;
                if (inter) { ex.atomic_or((uint32_t*)&st->flags[lo], (uint32_t)TTRL_FL_CRASHED); ex.atomic_or((uint32_t*)&st->flags[hi], (uint32_t)TTRL_FL_CRASHED); }
                if (will) {
                    const int slot = ex.atomic_add(&st->n_w, 1);
                    if (slot < EnvState<C::V>::WQ) { st->wpair[slot] = pr; st->wt[slot] = d2{absd * nx, absd * ny}; }
                    else st->overflow = 1;
                }
            });
    }
    if (st->overflow) {  // uniform: shared memory
        ex.parn(n, [&](int k) { collide_serial(c, k); });
        return;
    }
    const int nw = st->n_w;
    if (nw == 0) return;
    // K3 / K4
    ex.parn(nw, [&](int q) {
        const int pr = st->wpair[q], lo = pr & 0xFFFF, hi = pr >> 16;
        ex.atomic_max(&st->best[lo], hi);
        ex.atomic_max(&st->best[hi], lo);
    });
    ex.parn(nw, [&](int q) {
        const int pr = st->wpair[q], lo = pr & 0xFFFF, hi = pr >> 16;
        const d2 t = st->wt[q];
        if (st->best[lo] == hi) { st->imp[lo] = d2{t.x / 2, t.y / 2}; st->flags[lo] |= TTRL_FL_HAS_IMPACT; }
        if (st->best[hi] == lo) { st->imp[hi] = d2{-t.x / 2, -t.y / 2}; st->flags[hi] |= TTRL_FL_HAS_IMPACT; }
    });
}

// ------------------------------------------------------------------------------------------------
// REGULATE: RegulatedRoad.enforce_road_rules regulation.py:34-103
// ------------------------------------------------------------------------------------------------
// RoadNetwork.position_heading_along_route road.py:323-362 with lateral 0
template <class C>
TT_HD void position_heading_along_route(C& c, Route rt, int rlen, double lon, int cur_lane,
                                        double& px, double& py, double& ph) {
    const int cur_id = c.lanes[cur_lane].lane_id;
    int k = 0;
    auto head = [&]() {  // the walk consumes the local copy of the route: its head is always entry 0
        const int lid = route_lane_at(rt.l[0], 0);
        return c.sc->roads[route_road_at(rt.r[0], 0)].first_lane + (lid < 0 ? cur_id : lid);
    };
    int li = head();
    while (rlen - k > 1 && lon > c.lanes[li].length) {
        lon -= c.lanes[li].length;
        ++k;
        rt.r[0] = (rt.r[0] >> 8) | (rt.r[1] << 24); rt.r[1] = (rt.r[1] >> 8) | (rt.r[2] << 24); rt.r[2] >>= 8;
        rt.l[0] = (rt.l[0] >> 8) | (rt.l[1] << 24); rt.l[1] = (rt.l[1] >> 8) | (rt.l[2] << 24); rt.l[2] >>= 8;
        li = head();
    }
    lane_position(c.lanes[li], lon, 0.0, px, py);
    ph = lane_heading_at(c.lanes[li], lon);
}
// Regulation runs as TIME SLICES: for each of the 11 prediction times, every vehicle's predicted pose is computed
// once (5 doubles: x, y, heading's cos / sin ... per vehicle, instead of 33 for all times at once -- the shared-memory
// footprint of an env decides how many envs an SM holds) and every pair not yet in conflict is tested at that time.
// phase R0: un-yield (regulation.py:38-45)
template <class C>
TT_HD void regulate_unyield(C& c, int i) {
    auto* st = c.st;
    st->mark[i] = 0;
    if (st->flags[i] & TTRL_FL_YIELDING) {
        if (st->ytimer[i] >= 0.0 * 2) {
            st->tspeed[i] = c.lanes[st->lane[i]].speed_limit;
            st->flags[i] &= ~TTRL_FL_YIELDING;
        } else st->ytimer[i] += 1;
    }
}
// phase R1(k): ControlledVehicle.predict_trajectory_constant_speed (controller.py:236-253) at time 0.25 (k + 1)
template <class C>
TT_REG void regulate_predict(C& c, int i, int k) {
    auto* st = c.st;
    const int ln = st->lane[i];
    const double s0 = S_(c, i, ln);
    Route rt = route_of(st, i);
    int rlen = st->rlen[i];
    if (rlen <= 0) {  // `self.route or [self.lane_index]`
        rlen = 1;
        rt.r[0] = (uint32_t)c.lanes[ln].road;
        rt.l[0] = (uint32_t)c.lanes[ln].lane_id;
    }
    const double t = 0.25 + k * 0.25;
    double px, py, ph, sn, cn;
    position_heading_along_route(c, rt, rlen, s0 + st->v[i] * t, ln, px, py, ph);
    sincos(ph, &sn, &cn);
    c.pred[0 * C::V + i] = px;
    c.pred[1 * C::V + i] = py;
    c.pred[2 * C::V + i] = cn;
    c.pred[3 * C::V + i] = sn;
}
// utils.point_in_rotated_rectangle utils.py:75-91 with precomputed cos/sin of the rectangle angle
TT_HD bool point_in_rotated_rectangle(double px, double py, double cx, double cy, double length, double width, double cs, double sn) {
    const double dx = px - cx, dy = py - cy;
    const double rx = cs * dx + (-sn) * dy, ry = sn * dx + cs * dy;
    return -length / 2 <= rx && rx <= length / 2 && -width / 2 <= ry && ry <= width / 2;
}
// utils.has_corner_inside utils.py:158-172 (4 corners, centre, 4 edge midpoints: rect_corners :126-155)
TT_HD bool has_corner_inside(double x1, double y1, double c1, double s1, double x2, double y2, double c2, double s2, double len, double wid) {
    const double hl = len / 2, hw = wid / 2;
    const double qx[9] = {-hl, -hl, +hl, +hl, 0, -hl, hl, 0, 0};
    const double qy[9] = {-hw, +hw, +hw, -hw, 0, 0, 0, -hw, hw};
    for (int k = 0; k < 9; ++k) {
        const double px = (c1 * qx[k] + (-s1) * qy[k]) + x1;
        const double py = (s1 * qx[k] + c1 * qy[k]) + y1;
        if (point_in_rotated_rectangle(px, py, x2, y2, len, wid, c2, s2)) return true;
    }
    return false;
}
// one time slice of RegulatedRoad.is_conflict_possible (regulation.py:80-103) for the pair (i, j)
template <class C>
TT_REG bool conflict_at_slice(C& c, int i, int j) {
    const double xi = c.pred[0 * C::V + i], yi = c.pred[1 * C::V + i], c1 = c.pred[2 * C::V + i], s1 = c.pred[3 * C::V + i];
    const double xj = c.pred[0 * C::V + j], yj = c.pred[1 * C::V + j], c2 = c.pred[2 * C::V + j], s2 = c.pred[3 * C::V + j];
    const double len = 1.5 * kVehLength, wid = 0.9 * kVehWidth;
    const double dx = xj - xi, dy = yj - yi;
    if (sqrt(dx * dx + dy * dy) > kVehLength) return false;
    return has_corner_inside(xi, yi, c1, s1, xj, yj, c2, s2, len, wid) || has_corner_inside(xj, yj, c2, s2, xi, yi, c1, s1, len, wid);
}
// pair index q -> (i < j), row-major over the strict upper triangle of an n x n matrix
TT_HD void pair_of(int q, int n, int& i, int& j) {
    int r = (int)((2.0f * n - 1.0f - sqrtf((2.0f * n - 1.0f) * (2.0f * n - 1.0f) - 8.0f * q)) * 0.5f);
    if (r < 0) r = 0;
    if (r > n - 2) r = n - 2;
    while (r > 0 && r * (2 * n - r - 1) / 2 > q) --r;
    while ((r + 1) * (2 * n - r - 2) / 2 <= q) ++r;
    i = r;
    j = r + 1 + (q - r * (2 * n - r - 1) / 2);
}
// phase R3: the yielding vehicle of a conflicting pair (respect_priorities regulation.py:64-78).  The writes of the
// reference (:59-62) are idempotent flags, so pair order is irrelevant; they are merged in regulate_apply.
template <class C>
TT_HD void regulate_yield(C& c, int i, int j) {
    auto* st = c.st;
    const int p1 = c.lanes[st->lane[i]].priority, p2 = c.lanes[st->lane[j]].priority;
    int y;
    if (p1 > p2) y = j;
    else if (p1 < p2) y = i;
    else {
        const d2 pi_ = st->pos[i], pj = st->pos[j];
        const double f12 = st->cs[i].x * (pj.x - pi_.x) + st->cs[i].y * (pj.y - pi_.y);  // objects.py:204-205
        const double f21 = st->cs[j].x * (pi_.x - pj.x) + st->cs[j].y * (pi_.y - pj.y);
        y = f12 > f21 ? i : j;
    }
    if (!(st->flags[y] & TTRL_FL_MDP)) st->mark[y] = 1;  // benign same-value race
}
template <class C>
TT_HD void regulate_apply(C& c, int i) {
    auto* st = c.st;
    if (st->mark[i]) {
        st->tspeed[i] = 0;
        st->flags[i] |= TTRL_FL_YIELDING;
        st->ytimer[i] = 0;
    }
    st->mark[i] = 0;
}

// ------------------------------------------------------------------------------------------------
// one simulation sub-step (AbstractEnv._simulate body abstract.py:257-273)
// ------------------------------------------------------------------------------------------------
// In four parts -- lane decisions (follow_road, meta-action, MOBIL, abort rule), controls (steering, IDM, regulation), integration
// (+ closest lane), collisions -- so that a multi-env CTA can place alignment barriers between them (k_step: Exec::align_at<k>);
// env_substep is the four back to back.
template <class C, class Exec>
TT_HD void substep_lanes(C& c, Exec& ex, const int32_t* actions) {  // inlined on purpose: a call boundary here costs ~30 % (ctx spills)
    // `actions`: this env's raw action ids (one per controlled vehicle) or null (action=None)
    auto* st = c.st;
    const SceneDev* sc = c.sc;
    using ES = EnvState<C::V>;
    const int n = st->n;
    // ego meta-action on the first sub-step of an env-step: DiscreteMetaAction.act action.py:259-260
    const int32_t* first_actions = (actions && st->steps % sc->F == 0) ? actions : nullptr;
    ex.parn(n, [&](int t) {
        if (t == 0) { st->n_chg = 0; st->n_pair = 0; st->n_w = 0; st->overflow = 0; }
        act_phase_a(c, ex, t, first_actions);
    });
    {   // MOBIL over the vehicles whose timer fired (uniform: counters live in shared memory)
        const int nm = st->n_mob;
        for (int base = 0; base < nm; base += ES::MB) mobil_batch(c, ex, base, nm - base < ES::MB ? nm - base : ES::MB);
    }
    {   // phase B, in list order (bit order of bmask)
#pragma unroll 1
        for (int w = 0; w < C::W; ++w) {
            uint32_t m = st->bmask[w];
            while (m) {
#if defined(__CUDA_ARCH__)
                const int i = w * 32 + __ffs((int)m) - 1;
#else
                const int i = w * 32 + __builtin_ctz(m);
#endif
                m &= m - 1;
                const bool ab = ex.any(n, [&](int j) { return abort_pred(c, i, j); });
                if (ab) { if (ex.first()) st->tlane[i] = st->lane[i]; ex.sync(); }
            }
        }
    }
}
template <class C, class Exec>
TT_HD void substep_controls(C& c, Exec& ex) {
    auto* st = c.st;
    const SceneDev* sc = c.sc;
    const int n = st->n;
    ex.parn(n, [&](int t) {
        if (t == 0) { st->n_mob = 0; for (int w = 0; w < C::W; ++w) st->bmask[w] = 0; }
        act_phase_c1(c, ex, t);
    });
    ex.parn(n + st->n_chg, [&](int k) { act_phase_c2(c, k); });
    if (!C::kPlain && sc->cfg.regulated) {  // RegulatedRoad.step regulation.py:28-32
        const int rs = st->road_steps + 1;
        if (rs % sc->reg_period == 0) {
            const int np = n * (n - 1) / 2;
            ex.parn(n, [&](int t) { regulate_unyield(c, t); });
            ex.parn((np + 31) / 32, [&](int w) { c.pbits[w] = 0; });
            for (int k = 0; k < kPred; ++k) {
                ex.parn(n, [&](int t) { regulate_predict(c, t, k); });
                ex.parn(np, [&](int q) {
                    if ((c.pbits[q >> 5] >> (q & 31)) & 1) return;  // already in conflict at an earlier time
                    int i, j;
                    pair_of(q, n, i, j);
                    if (conflict_at_slice(c, i, j)) ex.atomic_or(&c.pbits[q >> 5], 1u << (q & 31));
                });
            }
            ex.parn(np, [&](int q) {
                if (!((c.pbits[q >> 5] >> (q & 31)) & 1)) return;
                int i, j;
                pair_of(q, n, i, j);
                regulate_yield(c, i, j);
            });
            ex.parn(n, [&](int t) { regulate_apply(c, t); });
        }
    }
}
template <class C, class Exec>
TT_HD void substep_integrate(C& c, Exec& ex) {
    const int n = c.st->n;
    ex.parn(c.L * C::W, [&](int k) { c.lmask[k] = 0; if (k == 0) c.st->n_arc = 0; });
    ex.parn(n, [&](int t) { integrate(c, ex, t); });
    if (!C::kPlain && c.sc->arc_tasks) closest_lane_tasks(c, ex);
}
template <class C, class Exec>
TT_HD void substep_collide(C& c, Exec& ex) {
    auto* st = c.st;
    collide_all(c, ex);  // advances st->steps / st->road_steps; ends on a team barrier on every path
    (void)st;
}
template <class C, class Exec>
TT_HD void env_substep(C& c, Exec& ex, const int32_t* actions) {
    substep_lanes(c, ex, actions);
    substep_controls(c, ex);
    substep_integrate(c, ex);
    substep_collide(c, ex);
}
// the same for the teams of a multi-env CTA (`active`: this team has an env): every alignment barrier is ONE instruction
// that active and idle teams reach alike (see env_reset_lockstep)
template <class C, class Exec>
TT_HD void env_substep_lockstep(C& c, Exec& ex, bool active, const int32_t* actions) {
    ex.template align_at<0>();
    if (active) substep_lanes(c, ex, actions);
    ex.template align_at<1>();
    if (active) substep_controls(c, ex);
    ex.template align_at<2>();
    if (active) substep_integrate(c, ex);
    ex.template align_at<3>();
    if (active) substep_collide(c, ex);
}

// ------------------------------------------------------------------------------------------------
// observations: envs/common/observation.py
// ------------------------------------------------------------------------------------------------
template <class C>
TT_HD double feature_of(C& c, int i, int f) {  // Vehicle.to_dict kinematics.py:237-261
    auto* st = c.st;
    switch (f) {
        case TTRL_F_PRESENCE: return 1.0;
        case TTRL_F_X: return st->pos[i].x;
        case TTRL_F_Y: return st->pos[i].y;
        case TTRL_F_VX: return st->v[i] * st->cs[i].x;
        case TTRL_F_VY: return st->v[i] * st->cs[i].y;
        case TTRL_F_COS_H: return st->cs[i].x;
        case TTRL_F_SIN_H: return st->cs[i].y;
        case TTRL_F_HEADING: return st->h[i];
    }
    return 0.0;
}
TT_HD bool is_relative_feature(int f) { return f == TTRL_F_X || f == TTRL_F_Y || f == TTRL_F_VX || f == TTRL_F_VY; }

// KinematicObservation.observe observation.py:233-275 + Road.close_objects_to road.py:418-447.
// Stable sort by rank counting: rank(j) = #{k : key_k < key_j or (key_k == key_j and k < j)}.
// `inv_perm` (or null): row permutation for order == "shuffled" (np_random.shuffle(obs[1:]), :272-273).
template <class C, class Exec>
TT_HD void observe_kinematics(C& c, Exec& ex, float* out, const int32_t* inv_perm, int ego) {
    auto* st = c.st;
    const ttrl_config& cfg = c.sc->cfg;
    const int Vo = cfg.obs_vehicles, Fe = cfg.n_features;
    const int le = st->lane[ego];
    ex.parn(Vo * Fe, [&](int k) { c.obs_s[k] = 0.0f; });
    ex.par([&](int t) {
        int cand = 0;
        if (t < st->n && t != ego) {
            const double dx = st->pos[t].x - st->pos[ego].x, dy = st->pos[t].y - st->pos[ego].y;
            if (sqrt(dx * dx + dy * dy) < 5.0 * 40.0) {  // PERCEPTION_DISTANCE abstract.py:41
                const double d = S_(c, t, le) - S_(c, ego, le);
                if (cfg.see_behind || -2 * kVehLength < d) cand = 1;
            }
        }
        st->mark[t] = cand;
    });
    ex.parn(st->n, [&](int t) {
        int row = -1;
        if (t == ego) row = 0;
        else if (st->mark[t]) {
            const double kt = fabs(S_(c, t, le) - S_(c, ego, le));
            int rank = 0;
            for (int k = 0; k < st->n; ++k) {
                if (!st->mark[k] || k == t) continue;
                if (cfg.order == TTRL_ORDER_SORTED) {
                    const double kk = fabs(S_(c, k, le) - S_(c, ego, le));
                    if (kk < kt || (kk == kt && k < t)) ++rank;
                } else if (k < t) ++rank;
            }
            if (rank < Vo - 1) row = 1 + (inv_perm ? inv_perm[rank] : rank);
        }
        if (row >= 0) {
            for (int k = 0; k < Fe; ++k) {
                const int f = cfg.features[k];
                double val = feature_of(c, t, f);
                if (row > 0 && !cfg.absolute && is_relative_feature(f)) val -= feature_of(c, ego, f);
                if (cfg.normalize && cfg.has_range[k]) {
                    val = lmap(val, cfg.range_lo[k], cfg.range_hi[k], -1.0, 1.0);
                    if (cfg.clip) val = clipd(val, -1.0, 1.0);
                }
                c.obs_s[row * Fe + k] = (float)val;
            }
        }
    });
    ex.parn(Vo * Fe, [&](int k) { out[k] = c.obs_s[k]; });
}

// OccupancyGridObservation.observe observation.py:353-412.  The reference writes vehicles in REVERSE list order
// per layer, so the EARLIEST vehicle of a cell wins: per-cell winner = min slot index (atomicMin in smem).
template <class C>
TT_HD bool grid_cell_of(C& c, int ego, double px, double py, int& ci, int& cj) {  // pos_to_index :414-434 (already relative)
    const ttrl_config& cfg = c.sc->cfg;
    if (cfg.align_to_vehicle_axes) {
        const double ca = c.st->cs[ego].x, sa = c.st->cs[ego].y;
        const double qx = ca * px + sa * py, qy = -sa * px + ca * py;
        px = qx;
        py = qy;
    }
    const double fi = floor((px - cfg.grid_min[0]) / cfg.grid_step[0]);
    const double fj = floor((py - cfg.grid_min[1]) / cfg.grid_step[1]);
    if (!(fi >= 0 && fi < cfg.grid_w && fj >= 0 && fj < cfg.grid_h)) return false;
    ci = (int)fi;
    cj = (int)fj;
    return true;
}
template <class C, class Exec>
TT_HD void observe_grid(C& c, Exec& ex, float* out, int ego) {
    auto* st = c.st;
    const ttrl_config& cfg = c.sc->cfg;
    const int W = cfg.grid_w, H = cfg.grid_h, Fe = cfg.n_features;
    ex.parn(W * H, [&](int k) { c.cell[k] = 0x7fffffff; });
    ex.parn(Fe * W * H, [&](int k) { out[k] = 0.0f; });
    ex.par([&](int t) {
        st->mark[t] = -1;
        if (t < st->n) {
            double x = st->pos[t].x - st->pos[ego].x, y = st->pos[t].y - st->pos[ego].y;
            if (cfg.grid_has_xrange) { x = lmap(x, cfg.grid_xrange[0], cfg.grid_xrange[1], -1.0, 1.0); x = lmap(x, -1.0, 1.0, cfg.grid_xrange[0], cfg.grid_xrange[1]); }
            if (cfg.grid_has_yrange) { y = lmap(y, cfg.grid_yrange[0], cfg.grid_yrange[1], -1.0, 1.0); y = lmap(y, -1.0, 1.0, cfg.grid_yrange[0], cfg.grid_yrange[1]); }
            int ci, cj;
            if (grid_cell_of(c, ego, x, y, ci, cj)) {
                st->mark[t] = ci * H + cj;
                ex.atomic_min(&c.cell[ci * H + cj], t);
            }
        }
    });
    ex.parn(st->n, [&](int t) {
        const int cellid = st->mark[t];
        if (cellid >= 0 && c.cell[cellid] == t) {
            for (int layer = 0; layer < Fe; ++layer) {
                const int f = cfg.features[layer];
                if (f == TTRL_F_ON_ROAD) continue;
                double val = feature_of(c, t, f);
                if (is_relative_feature(f)) val -= feature_of(c, ego, f);
                if (cfg.has_range[layer]) val = lmap(val, cfg.range_lo[layer], cfg.range_hi[layer], -1.0, 1.0);
                if (cfg.clip) val = clipd(val, -1.0, 1.0);
                out[layer * W * H + cellid] = (float)val;
            }
        }
    });
    // on_road layer: fill_road_layer_by_lanes :453-483 (waypoints every min(grid_step) within +-100 m)
    for (int layer = 0; layer < Fe; ++layer) {
        if (cfg.features[layer] != TTRL_F_ON_ROAD) continue;
        const double spacing = fmin(cfg.grid_step[0], cfg.grid_step[1]);
        const int nw = (int)ceil(200.0 / spacing) + 1;  // task bound; the exact per-lane count is applied below
        ex.parn(c.L * nw, [&](int k) {
            const int li = k / nw, w = k - li * nw;
            const ttrl_lane& l = c.lanes[li];
            const double origin = S_(c, ego, li);
            const double start = origin - 100.0, stop = origin + 100.0;
            if (w >= (int)ceil((stop - start) / spacing)) return;
            const double wp = clipd(start + w * spacing, 0.0, l.length);
            double px, py;
            lane_position(l, wp, 0.0, px, py);
            int ci, cj;
            if (grid_cell_of(c, ego, px - st->pos[ego].x, py - st->pos[ego].y, ci, cj)) out[layer * W * H + ci * H + cj] = 1.0f;
        });
    }
}
// RoadNetwork.is_connected_road road.py:231-276 with same_lane = False: only the ROADS of the two lanes matter.  D = remaining
// search depth; the "route starts at the current road" case keeps the depth and consumes the route, so it is a loop.
template <int D, class C>
TT_HDN bool is_connected_road(C& c, int r1, int r2, Route rt, int rlen) {
    const SceneDev* sc = c.sc;
    for (;;) {
        if (r2 == r1 || sc->roads[r2].to_node == sc->roads[r1].from_node) return true;  // is_same_road / is_leading_to_road(lane 2, lane 1)
        if (D == 0) return false;
        if (rlen > 0 && route_road_at(rt.r[0], 0) == r1) {  // route starts at the current road: skip it (:255-259)
            rt.r[0] = (rt.r[0] >> 8) | (rt.r[1] << 24); rt.r[1] = (rt.r[1] >> 8) | (rt.r[2] << 24); rt.r[2] >>= 8;
            --rlen;
            continue;
        }
        break;
    }
    if (D > 0) {
        constexpr int D1 = D > 0 ? D - 1 : 0;
        if (rlen > 0 && sc->roads[route_road_at(rt.r[0], 0)].from_node == sc->roads[r1].to_node) {  // follow the route (:260-264)
            const int nr = route_road_at(rt.r[0], 0);
            rt.r[0] = (rt.r[0] >> 8) | (rt.r[1] << 24); rt.r[1] = (rt.r[1] >> 8) | (rt.r[2] << 24); rt.r[2] >>= 8;
            return is_connected_road<D1>(c, nr, r2, rt, rlen - 1);
        }
        const int to = sc->roads[r1].to_node;  // every road leaving the end node (:265-275)
        for (int k = sc->node_first[to]; k < sc->node_first[to + 1]; ++k)
            if (is_connected_road<D1>(c, sc->node_roads[k], r2, rt, rlen)) return true;
    }
    return false;
}

// TimeToCollisionObservation.observe observation.py:114-151 over compute_ttc_grid finite_mdp.py:104-163.  The grid
// [speeds][lanes of the ego's road][time] only holds the costs 0 / 0.5 / 1 and is a running maximum: kept as 0 / 1 / 2
// in shared-memory ints (`cell`) under atomicMax.  Then the 3 x 3 x T window around the ego's speed index and lane is cut
// out of the grid padded with ones across the lanes and with repeated first / last rows across the speeds.
template <class C, class Exec>
TT_HD void observe_ttc(C& c, Exec& ex, float* out, int ego) {
    auto* st = c.st;
    const ttrl_config& cfg = c.sc->cfg;
    const int S = cfg.n_target_speeds, H = cfg.ttc_steps, n = st->n;
    const int le = st->lane[ego], re = c.lanes[le].road, L = c.sc->roads[re].n_lanes;
    const double tq = 1 / cfg.policy_frequency;
    ex.parn(S * L * H, [&](int k) { c.cell[k] = 0; });
    ex.parn(n, [&](int t) {  // connectivity of every other vehicle's road: independent of the speed row and the margin
        int m = -2;          // -2 not connected, -1 every lane, else the lane id
        if (t != ego) {
            const int ro = c.lanes[st->lane[t]].road;
            if (is_connected_road<3>(c, re, ro, route_of(st, ego), st->rlen[ego] < 0 ? 0 : st->rlen[ego]))
                m = c.sc->roads[ro].n_lanes == L ? c.lanes[st->lane[t]].lane_id : -1;
        }
        st->mark[t] = m;
    });
    ex.parn(S * n * 3, [&](int k) {
        const int si = k / (n * 3), rem = k - si * n * 3, t = rem / 3, mi = rem - 3 * t;
        if (t == ego || st->mark[t] == -2) return;
        const double ego_speed = cfg.target_speeds[si];  // index_to_speed controller.py:317-324
        if (ego_speed == st->v[t]) return;
        const double margin = kVehLength / 2 + kVehLength / 2;
        const double mm = mi == 0 ? 0.0 : (mi == 1 ? -margin : margin);
        const double distance = (S_(c, t, le) - S_(c, ego, le)) + mm;
        const double other_projected_speed = st->v[t] * (st->cs[t].x * st->cs[ego].x + st->cs[t].y * st->cs[ego].y);
        const double ttc = distance / not_zero(ego_speed - other_projected_speed);
        if (ttc < 0) return;
        const double q = ttc / tq;
        if (!(q < (double)H)) return;  // both time cells are beyond the horizon
        const int cost = mi == 0 ? 2 : 1;
        const int times[2] = {(int)q, (int)ceil(q)};
        for (int w = 0; w < 2; ++w) {
            const int time = times[w];
            if (time < 0 || time >= H) continue;
            if (st->mark[t] >= 0) ex.atomic_max(&c.cell[(si * L + st->mark[t]) * H + time], cost);
            else for (int l = 0; l < L; ++l) ex.atomic_max(&c.cell[(si * L + l) * H + time], cost);
        }
    });
    const int sidx = st->sidx[ego], lid = c.lanes[le].lane_id;
    ex.parn(9 * H, [&](int k) {
        const int a = k / (3 * H), b = (k / H) % 3, time = k % H;
        int so = sidx - 1 + a;
        so = so < 0 ? 0 : (so > S - 1 ? S - 1 : so);
        const int lo = lid - 1 + b;
        float val = 1.0f;
        if (lo >= 0 && lo < L) val = 0.5f * (float)c.cell[(so * L + lo) * H + time];
        out[k] = val;
    });
}

TT_HD int obs_single_size(const ttrl_config& cfg) {  // floats of ONE controlled vehicle's observation
    if (cfg.obs_type == TTRL_OBS_TTC) return 9 * cfg.ttc_steps;
    return cfg.obs_type == TTRL_OBS_GRID ? cfg.n_features * cfg.grid_w * cfg.grid_h : cfg.obs_vehicles * cfg.n_features;
}
// observation_type.observe(): one observation per controlled vehicle (MultiAgentObservation observation.py:587-603),
// K consecutive blocks in `out`; `inv_perm` (or null) holds K consecutive row permutations.
// np_random.shuffle(obs[1:]) (observation.py:272-273) with device draws: row k of the unshuffled observation gets a 32-bit
// Philox key keyed by (seed, global env, episode, env.steps, agent, k); its rank among the keys is its new row -- a uniform
// random permutation, the same for any team size.  Result: inv_perm[k] in c.perm_s[m ..].
template <class C, class Exec>
TT_HD const int32_t* draw_shuffle(C& c, Exec& ex, uint64_t seed, int64_t genv, int agent) {
    const int m = c.sc->cfg.obs_vehicles - 1;
    auto* st = c.st;
    ex.parn(m, [&](int k) {
        uint32_t c4[4] = {(uint32_t)genv, (uint32_t)((uint64_t)genv >> 32), (uint32_t)st->episode ^ ((uint32_t)st->steps * 0x9E3779B1u),
                          0x53000000u | ((uint32_t)agent << 16) | (uint32_t)k};
        philox4x32(c4, (uint32_t)seed, (uint32_t)(seed >> 32));
        c.perm_s[k] = c4[0];
    });
    ex.parn(m, [&](int k) {
        const uint32_t key = c.perm_s[k];
        int rank = 0;
        for (int j = 0; j < m; ++j) { const uint32_t kj = c.perm_s[j]; rank += (kj < key || (kj == key && j < k)) ? 1 : 0; }
        c.perm_s[m + k] = (uint32_t)rank;
    });
    return reinterpret_cast<const int32_t*>(c.perm_s + m);
}

// observation_type.observe(): one observation per controlled vehicle (MultiAgentObservation observation.py:587-603),
// K consecutive blocks in `out`; `inv_perm` (or null) holds K consecutive row permutations; without one, seed != 0 draws
// the "shuffled" order on the device.
template <class C, class Exec>
TT_HD void observe(C& c, Exec& ex, float* out, const int32_t* inv_perm, uint64_t seed = 0, int64_t genv = 0) {
    const ttrl_config& cfg = c.sc->cfg;
    const int K = n_agents(c);
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const int ego = K == 1 ? c.st->ego : c.st->egos[k];
        if (!C::kPlain && cfg.obs_type == TTRL_OBS_TTC) observe_ttc(c, ex, out + (size_t)k * obs_single_size(cfg), ego);
        else if (cfg.obs_type == TTRL_OBS_GRID) observe_grid(c, ex, out + (size_t)k * obs_single_size(cfg), ego);
        else {
            const int32_t* perm = inv_perm ? inv_perm + k * (cfg.obs_vehicles - 1) : nullptr;
            if (!perm && seed != 0 && cfg.order == TTRL_ORDER_SHUFFLED && cfg.obs_vehicles > 1) perm = draw_shuffle(c, ex, seed, genv, k);
            observe_kinematics(c, ex, out + (size_t)k * obs_single_size(cfg), perm, ego);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// reward / termination (single controlled vehicle)
// ------------------------------------------------------------------------------------------------
template <class C>
TT_HD bool has_arrived(C& c, int i) {  // intersection_env.py:364-369
    const int ln = c.st->lane[i];
    return c.lanes[ln].is_exit && S_(c, i, ln) >= 25;
}
template <class C>
TT_HD bool on_road(C& c, int i) {  // objects.py:199-202
    const int ln = c.st->lane[i];
    return lane_on_lane(c.lanes[ln], S_(c, i, ln), R_(c, i, ln), 0.0);
}
// `rc` (or null): the four entries of _agent_rewards / _rewards in the reference's key order (include/ttrl_b200.h: TTRL_INFO_REWARDS)
template <class C>
TT_HD double agent_reward(C& c, int i, int raw_action, double* rc = nullptr) {  // intersection_env.py:78-104 / u_turn_env.py:39-71 / roundabout_env.py:43-64
    const ttrl_config& cfg = c.sc->cfg;
    auto* st = c.st;
    const double crashed = (st->flags[i] & TTRL_FL_CRASHED) ? 1.0 : 0.0;
    if (!C::kPlain && cfg.reward_type == TTRL_REWARD_ROUNDABOUT) {
        // MDPVehicle.get_speed_index(vehicle) / (DEFAULT_TARGET_SPEEDS.size - 1); `action in [0, 2]`
        const double hs = (double)st->sidx[i] / cfg.speed_index_den;
        const double lc = (raw_action == 0 || raw_action == 2) ? 1.0 : 0.0;
        const double onr = on_road(c, i) ? 1.0 : 0.0;
        if (rc) { rc[0] = crashed; rc[1] = hs; rc[2] = lc; rc[3] = onr; }
        double reward = 0 + cfg.collision_reward * crashed + cfg.high_speed_reward * hs + cfg.lane_change_reward * lc + 0 * onr;
        if (cfg.normalize_reward) reward = lmap(reward, cfg.collision_reward, cfg.high_speed_reward, 0.0, 1.0);
        reward *= onr;
        return reward;
    }
    const double hs = clipd(lmap(st->v[i], cfg.reward_speed_lo, cfg.reward_speed_hi, 0.0, 1.0), 0.0, 1.0);
    const double onr = on_road(c, i) ? 1.0 : 0.0;
    if (!C::kPlain && cfg.reward_type == TTRL_REWARD_INTERSECTION) {
        const bool arrived = has_arrived(c, i);
        if (rc) { rc[0] = crashed; rc[1] = hs; rc[2] = arrived ? 1.0 : 0.0; rc[3] = onr; }
        double reward = 0 + cfg.collision_reward * crashed + cfg.high_speed_reward * hs + cfg.arrived_reward * (arrived ? 1.0 : 0.0) + 0 * onr;
        reward = arrived ? cfg.arrived_reward : reward;
        reward *= onr;
        if (cfg.normalize_reward) reward = lmap(reward, cfg.collision_reward, cfg.arrived_reward, 0.0, 1.0);
        return reward;
    }
    const ttrl_lane& l = c.lanes[st->lane[i]];
    const int nl = c.sc->roads[l.road].n_lanes;
    const double lane_term = (double)l.lane_id / (double)(nl - 1 > 1 ? nl - 1 : 1);
    if (rc) { rc[0] = crashed; rc[1] = lane_term; rc[2] = hs; rc[3] = onr; }
    double reward = 0 + cfg.collision_reward * crashed + cfg.lane_reward * lane_term + cfg.high_speed_reward * hs + 0 * onr;
    if (cfg.normalize_reward) reward = lmap(reward, cfg.collision_reward, cfg.high_speed_reward + cfg.lane_reward, 0.0, 1.0);
    reward *= onr;
    return reward;
}
template <class C>
TT_HD bool is_terminated(C& c) {  // intersection_env.py:106-111 / u_turn_env.py:73-74 / roundabout_env.py:66-67
    const int e = c.st->ego;
    const bool crashed = (c.st->flags[e] & TTRL_FL_CRASHED) != 0;
    if (!C::kPlain && c.sc->cfg.reward_type == TTRL_REWARD_ROUNDABOUT) return crashed;
    const bool off = c.sc->cfg.offroad_terminal && !on_road(c, e);  // self.vehicle = controlled_vehicles[0]
    if (!C::kPlain && c.sc->cfg.reward_type == TTRL_REWARD_INTERSECTION) {
        const int K = n_agents(c);
        if (K == 1) return crashed || has_arrived(c, e) || off;
        bool any_crashed = false, all_arrived = true;  // any(crashed) or all(has_arrived) over the controlled vehicles
        for (int k = 0; k < K; ++k) {
            any_crashed = any_crashed || (c.st->flags[c.st->egos[k]] & TTRL_FL_CRASHED) != 0;
            all_arrived = all_arrived && has_arrived(c, c.st->egos[k]);
        }
        return any_crashed || all_arrived || off;
    }
    return crashed || off;
}

// AbstractEnv._info (abstract.py:169-186) of the current state without a step (reset: abstract.py:211): speed, crashed and
// _rewards(action) with `raw_actions` (or null: no action) -- by one thread
template <class C>
TT_HD void write_info(C& c, double* info, int E, int e, const int32_t* raw_actions) {
    auto* st = c.st;
    const int K = n_agents(c), ego = st->ego;
    double rc[4] = {0, 0, 0, 0};
    for (int k = 0; k < K; ++k) {
        double rk4[4];
        (void)agent_reward(c, K == 1 ? ego : st->egos[k], raw_actions ? raw_actions[k] : -1, rk4);
        for (int q = 0; q < 4; ++q) rc[q] = rc[q] + rk4[q];
    }
    info[(size_t)TTRL_INFO_SPEED * E + e] = st->v[ego];
    info[(size_t)TTRL_INFO_CRASHED * E + e] = (st->flags[ego] & TTRL_FL_CRASHED) ? 1.0 : 0.0;
    for (int q = 0; q < 4; ++q) info[(size_t)(TTRL_INFO_REWARDS + q) * E + e] = rc[q] / K;
}

// ------------------------------------------------------------------------------------------------
// IntersectionEnv._clear_vehicles / _spawn_vehicle intersection_env.py:320-362
// ------------------------------------------------------------------------------------------------
struct SlotRegs {
    double d[13];
    double lin[TTRL_NLIN];
    int32_t i[6];
    Route rt;
};
template <class C>
TT_HD void slot_read(C& c, int t, SlotRegs& r) {
    auto* st = c.st;
    r.d[0] = st->pos[t].x; r.d[1] = st->pos[t].y; r.d[2] = st->h[t]; r.d[3] = st->v[t]; r.d[4] = st->cs[t].x; r.d[5] = st->cs[t].y;
    r.d[6] = st->steer[t]; r.d[7] = st->acc[t]; r.d[8] = st->tspeed[t]; r.d[9] = st->timer[t]; r.d[10] = st->delta[t];
    r.d[11] = st->imp[t].x; r.d[12] = st->imp[t].y;
    r.i[0] = st->lane[t]; r.i[1] = st->tlane[t]; r.i[2] = st->flags[t]; r.i[3] = st->sidx[t]; r.i[4] = st->rlen[t]; r.i[5] = st->ytimer[t];
    r.rt = route_of(st, t);
    if (is_linear(c)) for (int k = 0; k < TTRL_NLIN; ++k) r.lin[k] = c.lin[k * C::V + t];
}
template <class C>
TT_HD void slot_write(C& c, int t, const SlotRegs& r) {
    auto* st = c.st;
    st->pos[t] = d2{r.d[0], r.d[1]}; st->h[t] = r.d[2]; st->v[t] = r.d[3]; st->cs[t] = d2{r.d[4], r.d[5]};
    st->steer[t] = r.d[6]; st->acc[t] = r.d[7]; st->tspeed[t] = r.d[8]; st->timer[t] = r.d[9]; st->delta[t] = r.d[10];
    st->imp[t] = d2{r.d[11], r.d[12]};
    st->lane[t] = r.i[0]; st->tlane[t] = r.i[1]; st->flags[t] = r.i[2]; st->sidx[t] = r.i[3]; st->rlen[t] = r.i[4]; st->ytimer[t] = r.i[5];
    route_store(st, t, r.rt);
    if (is_linear(c)) for (int k = 0; k < TTRL_NLIN; ++k) c.lin[k * C::V + t] = r.lin[k];
}

// slots of the controlled vehicles from their agent bits (after a compaction / reset of a multi-agent env)
template <class C, class Exec>
TT_HD void find_agents(C& c, Exec& ex) {
    auto* st = c.st;
    ex.parn(st->n, [&](int t) {
        if (st->flags[t] & TTRL_FL_CONTROLLED) st->egos[agent_of(st->flags[t])] = t;
    });
    if (ex.first()) st->ego = st->egos[0];
    ex.sync();
}

// Stable compaction of the vehicle list (order preserved like the list comprehension at :357-362).
// NB: table rows / lane masks are NOT moved; callers rebuild them before their next use.
template <class C, class Exec>
TT_HD void clear_vehicles(C& c, Exec& ex) {
    auto* st = c.st;
    ex.par([&](int t) {
        int keep = 0;
        if (t < st->n) {
            keep = (st->flags[t] & TTRL_FL_CONTROLLED) != 0;
            if (!keep) {
                const int ln = st->lane[t];
                const ttrl_lane& l = c.lanes[ln];
                const bool leaving = l.is_exit && S_(c, t, ln) >= l.length - 4 * kVehLength;
                keep = !(leaving || st->rlen[t] < 0);
            }
        }
        st->mark[t] = keep;
    });
    ex.par2(
        [&](int t, SlotRegs& r, int& dst) {
            dst = -1;
            if (st->mark[t]) {
                dst = 0;
                for (int k = 0; k < t; ++k) dst += st->mark[k];
                slot_read(c, t, r);
            }
        },
        [&](int t, SlotRegs& r, int& dst) {
            if (dst >= 0) {
                slot_write(c, dst, r);
                if (t == st->ego) st->flag0 = dst;
            }
            if (t == 0) { int tot = 0; for (int k = 0; k < st->n; ++k) tot += st->mark[k]; st->flag1 = tot; }
        });
    if (ex.first()) { st->ego = st->flag0; st->egos[0] = st->flag0; st->n = st->flag1; }
    ex.sync();
    if (n_agents(c) > 1) find_agents(c, ex);
}

// ControlledVehicle.plan_route_to (controller.py:71-87): [lane_index] + the roads of the host's BFS path (road.py:159-188)
// from the entry corner to exit node o<exit_>; written to slot s, returns the route length.
template <class C>
TT_HD int planned_route(C& c, int ln, int entry, int exit_, int s) {
    const SceneDev* sc = c.sc;
    const int nr = sc->spawn_route_len[entry * 4 + exit_];
    Route q{};
    route_set(q, 0, c.lanes[ln].road, c.lanes[ln].lane_id);
    for (int k = 0; k < nr && k + 1 < TTRL_ROUTE_CAP; ++k) route_set(q, k + 1, sc->spawn_route_road[(entry * 4 + exit_) * TTRL_ROUTE_CAP + k], -1);
    route_store(c.st, s, q);
    return 1 + nr;
}

struct SpawnParams {
    double longitudinal, position_deviation, speed_deviation, spawn_probability;
    int go_straight;
};
// One spawn attempt (_spawn_vehicle :320-348 + make_on_lane objects.py:67-89 + IDMVehicle/ControlledVehicle
// constructors + plan_route_to + randomize_behavior).  Sets st->flag0 = 1 if a vehicle was appended.
template <class C, class Exec>
TT_HD void spawn_vehicle(C& c, Exec& ex, const ttrl_spawn_draw& d, const SpawnParams& sp) {
    auto* st = c.st;
    const SceneDev* sc = c.sc;
    if (ex.first()) { st->flag0 = 0; st->flag1 = 0; }  // flag0: a vehicle was appended; flag1: no free slot (capacity reject)
    ex.sync();
    if (d.u_spawn > sp.spawn_probability) return;  // uniform: d is per-env
    const int entry = d.entry, exit_ = sp.go_straight ? (d.entry + 2) % 4 : d.exit;
    const ttrl_lane& l = c.lanes[sc->spawn_lane[entry]];
    const double lon = sp.longitudinal + 5 + d.n_pos * sp.position_deviation;
    const double speed = 8 + d.n_speed * sp.speed_deviation;
    double px, py;
    lane_position(l, lon, 0.0, px, py);
    const double hd = lane_heading_at(l, lon);
    const bool too_close = ex.any(st->n, [&](int t) {
        const double dx = st->pos[t].x - px, dy = st->pos[t].y - py;
        return sqrt(dx * dx + dy * dy) < 15;
    });
    if (too_close) return;
    if (st->n >= c.vcap) {  // uniform (shared memory).  The reference's list has no capacity: counted in ttrl_episode_stats
        if (ex.first()) st->flag1 = 1;
        ex.sync();
        return;
    }
    if (ex.first()) {
        const int s = st->n;
        st->pos[s] = d2{px, py}; st->h[s] = hd; st->v[s] = speed;
        st->cs[s] = d2{cos(hd), sin(hd)};
        st->steer[s] = 0; st->acc[s] = 0; st->imp[s] = d2{0, 0};
        uint64_t m;
        const int ln = table_row_and_closest(c, s, m);  // RoadObject.__init__ objects.py:45-50
        st->lane[s] = ln; st->tlane[s] = ln;         // controller.py:46
        st->tspeed[s] = speed;                       // controller.py:47
        st->timer[s] = py_mod1((px + py) * kPi);     // behavior.py:64
        st->delta[s] = d.delta;                      // IDMVehicle.randomize_behavior behavior.py:66-69
        if (is_linear(c)) st->delta[s] = 4.0; // LinearVehicle.randomize_behavior (behavior.py:402-410) leaves the class DELTA
        if (is_linear(c))
            for (int k = 0; k < TTRL_NLIN; ++k) c.lin[k * C::V + s] = sc->cfg.lin_lo[k] + d.lin_u[k] * (sc->cfg.lin_hi[k] - sc->cfg.lin_lo[k]);
        st->flags[s] = 0; st->sidx[s] = 0; st->ytimer[s] = 0;
        st->rlen[s] = planned_route(c, ln, entry, exit_, s);
        st->n = s + 1;
        st->flag0 = 1;
    }
    ex.sync();
}

// ------------------------------------------------------------------------------------------------
// global <-> shared state movement (coalesced: consecutive threads move consecutive slots of a field)
// ------------------------------------------------------------------------------------------------
struct GlobalState {
    double* vd;   // [ND][E][V]
    int32_t* vi;  // [NI][E][V]
    int32_t* ei;  // [NEI][E]
    double* ed;   // [NED][E]
    int E;
    int V;        // slots per env in this buffer
    double* lin;  // [TTRL_NLIN][E][V] LinearVehicle parameters, or null (IDM traffic)
};
// CG = true: every load bypasses L1 (ld.global.cg).  Used for the shadow buffers of the asynchronous device reset, which
// another kernel may have written while this one was already running (L1 is not coherent; a neighbouring env's earlier
// load can have cached a line that also holds part of this env's record).
template <bool CG, class T_> TT_HD T_ gload(const T_* p) {
#if defined(__CUDA_ARCH__)
    if (CG) return __ldcg(p);
#endif
    return *p;
}
template <bool CG = false, class C, class Exec>
TT_HD void load_env(C& c, Exec& ex, const GlobalState& g, int e) {
    auto* st = c.st;
    const int V = g.V;
    if (ex.first()) {
        st->n = gload<CG>(&g.ei[TTRL_EI_NVEH * g.E + e]); st->steps = gload<CG>(&g.ei[TTRL_EI_STEPS * g.E + e]);
        st->road_steps = gload<CG>(&g.ei[TTRL_EI_ROAD_STEPS * g.E + e]); st->ego = gload<CG>(&g.ei[TTRL_EI_EGO * g.E + e]);
        st->episode = gload<CG>(&g.ei[TTRL_EI_EPISODE * g.E + e]); st->done = gload<CG>(&g.ei[TTRL_EI_DONE * g.E + e]);
        st->time = gload<CG>(&g.ed[TTRL_ED_TIME * g.E + e]); st->ret = gload<CG>(&g.ed[TTRL_ED_RETURN * g.E + e]);
        st->n_chg = st->n_mob = st->n_pair = st->n_w = st->overflow = 0;
        for (int w = 0; w < C::W; ++w) st->bmask[w] = 0;
        st->egos[0] = st->ego;
        for (int k = 1; k < TTRL_MAX_CONTROLLED; ++k) st->egos[k] = st->ego;
    }
    ex.sync();
    const int st_n = gload<CG>(&g.ei[TTRL_EI_NVEH * g.E + e]);
    ex.par([&](int t) {
        if (t < V) {
            const size_t o = (size_t)e * V + t, fs = (size_t)g.E * V;
            st->pos[t] = d2{gload<CG>(&g.vd[TTRL_D_X * fs + o]), gload<CG>(&g.vd[TTRL_D_Y * fs + o])};
            const double hd = gload<CG>(&g.vd[TTRL_D_HEADING * fs + o]);
            double sn, cn;
            sincos(hd, &sn, &cn);
            st->h[t] = hd; st->cs[t] = d2{cn, sn};
            st->v[t] = gload<CG>(&g.vd[TTRL_D_SPEED * fs + o]);
            st->steer[t] = gload<CG>(&g.vd[TTRL_D_STEERING * fs + o]); st->acc[t] = gload<CG>(&g.vd[TTRL_D_ACCEL * fs + o]);
            st->tspeed[t] = gload<CG>(&g.vd[TTRL_D_TARGET_SPEED * fs + o]); st->timer[t] = gload<CG>(&g.vd[TTRL_D_TIMER * fs + o]);
            st->delta[t] = gload<CG>(&g.vd[TTRL_D_DELTA * fs + o]);
            st->imp[t] = d2{gload<CG>(&g.vd[TTRL_D_IMPACT_X * fs + o]), gload<CG>(&g.vd[TTRL_D_IMPACT_Y * fs + o])};
            st->lane[t] = gload<CG>(&g.vi[TTRL_I_LANE * fs + o]); st->tlane[t] = gload<CG>(&g.vi[TTRL_I_TARGET_LANE * fs + o]);
            st->flags[t] = gload<CG>(&g.vi[TTRL_I_FLAGS * fs + o]); st->sidx[t] = gload<CG>(&g.vi[TTRL_I_SPEED_INDEX * fs + o]);
            st->rlen[t] = gload<CG>(&g.vi[TTRL_I_ROUTE_LEN * fs + o]);
            st->rroad[0][t] = (uint32_t)gload<CG>(&g.vi[TTRL_I_ROUTE_ROAD * fs + o]); st->rlanew[0][t] = (uint32_t)gload<CG>(&g.vi[TTRL_I_ROUTE_LANE * fs + o]);
            st->rroad[1][t] = (uint32_t)gload<CG>(&g.vi[TTRL_I_ROUTE_ROAD1 * fs + o]); st->rlanew[1][t] = (uint32_t)gload<CG>(&g.vi[TTRL_I_ROUTE_LANE1 * fs + o]);
            st->rroad[2][t] = (uint32_t)gload<CG>(&g.vi[TTRL_I_ROUTE_ROAD2 * fs + o]); st->rlanew[2][t] = (uint32_t)gload<CG>(&g.vi[TTRL_I_ROUTE_LANE2 * fs + o]);
            st->ytimer[t] = gload<CG>(&g.vi[TTRL_I_YIELD_TIMER * fs + o]);
            if (is_linear(c)) for (int k = 0; k < TTRL_NLIN; ++k) c.lin[k * C::V + t] = g.lin ? gload<CG>(&g.lin[k * fs + o]) : c.sc->cfg.lin_default[k];
            if (C::kMulti && t < st_n && (st->flags[t] & TTRL_FL_CONTROLLED) && (st->flags[t] & TTRL_FL_AGENT_MASK))
                st->egos[(st->flags[t] & TTRL_FL_AGENT_MASK) >> TTRL_FL_AGENT_SHIFT] = t;  // agents k >= 1 (distinct slots: no race)
        } else {
            st->pos[t] = d2{0, 0}; st->cs[t] = d2{1, 0}; st->imp[t] = d2{0, 0};
            st->h[t] = st->v[t] = st->steer[t] = st->acc[t] = 0;
            st->tspeed[t] = st->timer[t] = st->delta[t] = 0;
            st->lane[t] = st->tlane[t] = st->flags[t] = st->sidx[t] = st->rlen[t] = st->ytimer[t] = 0;
            for (int w = 0; w < TTRL_ROUTE_WORDS; ++w) st->rroad[w][t] = st->rlanew[w][t] = 0;
            if (is_linear(c)) for (int k = 0; k < TTRL_NLIN; ++k) c.lin[k * C::V + t] = 0;
        }
        st->mark[t] = 0; st->tl_old[t] = 0; st->acc2[t] = 0; st->tsteer[t] = 0; st->best[t] = -1; st->fo[t] = -1;
        // pre-check guard from the loaded speed (integrate refreshes it every sub-step)
        const double diag = sqrt(kVehLength * kVehLength + kVehWidth * kVehWidth);
        const double thr = (diag + diag) / 2 + st->v[t] * c.sc->dt;
        st->thr2[t] = thr < 0 ? -1.0 : thr * thr * (1.0 + 1e-12);
    });
    // the table is a pure function of positions: rebuild it instead of storing it
    rebuild_tables(c, ex);
}
template <class C, class Exec>
TT_HD void store_env(C& c, Exec& ex, const GlobalState& g, int e) {
    auto* st = c.st;
    const int V = g.V;
    if (ex.first()) {
        g.ei[TTRL_EI_NVEH * g.E + e] = st->n; g.ei[TTRL_EI_STEPS * g.E + e] = st->steps;
        g.ei[TTRL_EI_ROAD_STEPS * g.E + e] = st->road_steps; g.ei[TTRL_EI_EGO * g.E + e] = st->ego;
        g.ei[TTRL_EI_EPISODE * g.E + e] = st->episode; g.ei[TTRL_EI_DONE * g.E + e] = st->done;
        g.ed[TTRL_ED_TIME * g.E + e] = st->time; g.ed[TTRL_ED_RETURN * g.E + e] = st->ret;
    }
    ex.parn(V, [&](int t) {
        const size_t o = (size_t)e * V + t, fs = (size_t)g.E * V;
        if (t >= C::V) {  // the buffer has more slots than this kernel's capacity (small size class): dead by construction
            for (int f = 0; f < TTRL_ND; ++f) g.vd[f * fs + o] = 0;
            for (int f = 0; f < TTRL_NI; ++f) g.vi[f * fs + o] = 0;
            if (g.lin) for (int f = 0; f < TTRL_NLIN; ++f) g.lin[f * fs + o] = 0;
            return;
        }
        const bool live = t < st->n;
        const int rlen = live ? st->rlen[t] : 0;
        // canonical form: dead slots zero, unused route bytes zero (keeps get_state comparable bit for bit)
        auto keep_of = [](int n) { return n >= 4 ? 0xFFFFFFFFu : (n <= 0 ? 0u : ((1u << (8 * n)) - 1u)); };
        const uint32_t keep = live ? keep_of(rlen) : 0u, keep1 = live ? keep_of(rlen - 4) : 0u, keep2 = live ? keep_of(rlen - 8) : 0u;
        g.vd[TTRL_D_X * fs + o] = live ? st->pos[t].x : 0; g.vd[TTRL_D_Y * fs + o] = live ? st->pos[t].y : 0;
        g.vd[TTRL_D_HEADING * fs + o] = live ? st->h[t] : 0; g.vd[TTRL_D_SPEED * fs + o] = live ? st->v[t] : 0;
        g.vd[TTRL_D_STEERING * fs + o] = live ? st->steer[t] : 0; g.vd[TTRL_D_ACCEL * fs + o] = live ? st->acc[t] : 0;
        g.vd[TTRL_D_TARGET_SPEED * fs + o] = live ? st->tspeed[t] : 0; g.vd[TTRL_D_TIMER * fs + o] = live ? st->timer[t] : 0;
        g.vd[TTRL_D_DELTA * fs + o] = live ? st->delta[t] : 0;
        g.vd[TTRL_D_IMPACT_X * fs + o] = live ? st->imp[t].x : 0; g.vd[TTRL_D_IMPACT_Y * fs + o] = live ? st->imp[t].y : 0;
        g.vi[TTRL_I_LANE * fs + o] = live ? st->lane[t] : 0; g.vi[TTRL_I_TARGET_LANE * fs + o] = live ? st->tlane[t] : 0;
        g.vi[TTRL_I_FLAGS * fs + o] = live ? st->flags[t] : 0; g.vi[TTRL_I_SPEED_INDEX * fs + o] = live ? st->sidx[t] : 0;
        g.vi[TTRL_I_ROUTE_LEN * fs + o] = rlen;
        g.vi[TTRL_I_ROUTE_ROAD * fs + o] = (int32_t)(st->rroad[0][t] & keep);
        g.vi[TTRL_I_ROUTE_LANE * fs + o] = (int32_t)(st->rlanew[0][t] & keep);
        g.vi[TTRL_I_ROUTE_ROAD1 * fs + o] = (int32_t)(st->rroad[1][t] & keep1);
        g.vi[TTRL_I_ROUTE_LANE1 * fs + o] = (int32_t)(st->rlanew[1][t] & keep1);
        g.vi[TTRL_I_ROUTE_ROAD2 * fs + o] = (int32_t)(st->rroad[2][t] & keep2);
        g.vi[TTRL_I_ROUTE_LANE2 * fs + o] = (int32_t)(st->rlanew[2][t] & keep2);
        g.vi[TTRL_I_YIELD_TIMER * fs + o] = live ? st->ytimer[t] : 0;
        if (is_linear(c) && g.lin) for (int k = 0; k < TTRL_NLIN; ++k) g.lin[k * fs + o] = live ? c.lin[k * C::V + t] : 0;
    });
}

// ------------------------------------------------------------------------------------------------
// counter-based RNG for device-side spawn draws (Philox-4x32-10), keyed by (seed, global env, step, draw)
// ------------------------------------------------------------------------------------------------
// `counter` = steps | episode << 32 for the per-step spawn, (1 << 63) | episode << 8 | attempt for the reset attempts
TT_HDN void device_spawn_draw(uint64_t seed, int64_t global_env, uint64_t counter, ttrl_spawn_draw& d) {
    const uint32_t ctr = (uint32_t)counter ^ ((uint32_t)(counter >> 32) * 0x9E3779B1u);
    const uint32_t dom = (uint32_t)(counter >> 63) << 4;
    uint32_t a[4] = {(uint32_t)global_env, (uint32_t)((uint64_t)global_env >> 32), ctr, dom | 0u};
    uint32_t b[4] = {(uint32_t)global_env, (uint32_t)((uint64_t)global_env >> 32), ctr, dom | 1u};
    uint32_t e[4] = {(uint32_t)global_env, (uint32_t)((uint64_t)global_env >> 32), ctr, dom | 2u};
    philox4x32(a, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32(b, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32(e, (uint32_t)seed, (uint32_t)(seed >> 32));
    d.u_spawn = u01(a[0], a[1]);
    const double ue = u01(a[2], a[3]);
    d.entry = (int)(ue * 4.0) & 3;                       // choice(range(4), 2, replace=False): ordered pair,
    d.exit = (d.entry + 1 + ((int)(u01(e[0], e[1]) * 3.0) % 3)) % 4;  // uniform over the 12 possibilities
    const double u1 = 1.0 - u01(b[0], b[1]), u2 = u01(b[2], b[3]);    // Box-Muller
    const double rad = sqrt(-2.0 * log(u1));
    d.n_pos = rad * cos(2 * kPi * u2);
    d.n_speed = rad * sin(2 * kPi * u2);
    d.delta = 3.5 + u01(e[2], e[3]) * (4.5 - 3.5);       // behavior.py:66-69
    // LinearVehicle.randomize_behavior's five uniforms (behavior.py:402-410): three more counter blocks
    for (int k = 0; k < 3; ++k) {
        uint32_t f[4] = {(uint32_t)global_env, (uint32_t)((uint64_t)global_env >> 32), ctr, dom | (3u + (uint32_t)k)};
        philox4x32(f, (uint32_t)seed, (uint32_t)(seed >> 32));
        d.lin_u[2 * k] = u01(f[0], f[1]);
        if (2 * k + 1 < 5) d.lin_u[2 * k + 1] = u01(f[2], f[3]);
    }
}


// ------------------------------------------------------------------------------------------------
// device-side reset
// ------------------------------------------------------------------------------------------------
// two uniforms of the reset stream, keyed by (seed, global env, episode, index)
TT_HD void reset_uniforms(uint64_t seed, int64_t genv, int episode, uint32_t idx, double& u0, double& u1) {
    uint32_t c4[4] = {(uint32_t)genv, (uint32_t)((uint64_t)genv >> 32), (uint32_t)episode, 0x52000000u | idx};
    philox4x32(c4, (uint32_t)seed, (uint32_t)(seed >> 32));
    u0 = u01(c4[0], c4[1]);
    u1 = u01(c4[2], c4[3]);
}
TT_HD uint64_t reset_attempt_counter(int episode, int attempt) { return (1ull << 63) | ((uint64_t)(uint32_t)episode << 8) | (uint64_t)attempt; }

template <class C, class Exec>
TT_HD void reset_scalars(C& c, Exec& ex, int episode) {
    auto* st = c.st;
    if (ex.first()) {
        st->n = 0; st->steps = 0; st->road_steps = 0; st->ego = 0; st->episode = episode; st->done = 0;
        for (int k = 0; k < TTRL_MAX_CONTROLLED; ++k) st->egos[k] = 0;
        st->time = 0; st->ret = 0;
        st->n_chg = st->n_mob = st->n_pair = st->n_w = st->overflow = 0;
        for (int w = 0; w < C::W; ++w) st->bmask[w] = 0;
    }
    ex.par([&](int t) {
        st->pos[t] = d2{0, 0}; st->cs[t] = d2{1, 0}; st->imp[t] = d2{0, 0};
        st->h[t] = st->v[t] = st->steer[t] = st->acc[t] = st->tspeed[t] = st->timer[t] = st->delta[t] = 0;
        st->acc2[t] = st->tsteer[t] = 0; st->thr2[t] = 0;
        st->lane[t] = st->tlane[t] = st->flags[t] = st->sidx[t] = st->rlen[t] = st->ytimer[t] = 0;
        for (int w = 0; w < TTRL_ROUTE_WORDS; ++w) st->rroad[w][t] = st->rlanew[w][t] = 0;
        st->mark[t] = 0; st->tl_old[t] = 0; st->best[t] = -1; st->fo[t] = -1;
        if (is_linear(c)) for (int k = 0; k < TTRL_NLIN; ++k) c.lin[k * C::V + t] = 0;
    });
    ex.parn(c.L * C::W, [&](int k) { c.lmask[k] = 0; });
}

// Synthetic highway: Vehicle.create_random's rule (kinematics.py:91-103): every new vehicle goes
// offset * U[0.9, 1.1] ahead of the furthest one, offset = spacing (12 + speed) exp(-5/40 lanes), on a uniformly
// random lane at U[0.7, 0.8] speed_limit; slot 0 is the MDPVehicle ego; IDMVehicle DELTA ~ U[3.5, 4.5]
// (behavior.py:66-69), timer = (x + y) pi mod 1 (behavior.py:64).  Same rule as scenes.make_highway_state.
template <class C, class Exec>
TT_HD void reset_highway(C& c, Exec& ex, uint64_t seed, int64_t genv, int episode) {
    auto* st = c.st;
    const ttrl_reset_params& rp = c.sc->rp;
    const ttrl_config& cfg = c.sc->cfg;
    reset_scalars(c, ex, episode);
    const int n = rp.n_vehicles < c.vcap ? rp.n_vehicles : c.vcap;
    const double lane_factor = exp(-5.0 / 40 * rp.lanes);
    ex.parn(n, [&](int s) {
        double u_lane, u_speed, u_jit, u_delta;
        reset_uniforms(seed, genv, episode, 2u * s, u_lane, u_speed);
        reset_uniforms(seed, genv, episode, 2u * s + 1u, u_jit, u_delta);
        int lid = (int)(u_lane * rp.lanes);
        if (lid > rp.lanes - 1) lid = rp.lanes - 1;
        const double speed = s == 0 ? rp.ego_speed : (0.7 + 0.1 * u_speed) * rp.speed_limit;
        const double spacing = s == 0 ? rp.ego_spacing : 1.0 / rp.density;
        const double offset = spacing * (12 + 1.0 * speed) * lane_factor;
        st->acc2[s] = offset * (0.9 + 0.2 * u_jit) + (s == 0 ? 3 * offset : 0.0);  // advance relative to the previous vehicle
        st->v[s] = speed;
        st->tspeed[s] = speed;
        st->lane[s] = st->tlane[s] = c.sc->roads[0].first_lane + lid;
        st->delta[s] = s == 0 ? 4.0 : 3.5 + u_delta;
        st->rlen[s] = -1;  // route None
        if (s == 0) {
            st->flags[s] = TTRL_FL_MDP | TTRL_FL_CONTROLLED;
            st->sidx[s] = speed_to_index(cfg, speed);
            st->tspeed[s] = cfg.target_speeds[st->sidx[s]];
        }
    });
    if (ex.first()) {  // longitudinal positions: cumulative advances in slot order (the only serial part)
        double x = 0;
        for (int s = 0; s < n; ++s) { x += st->acc2[s]; st->acc2[s] = x; }
        st->n = n;
        st->ego = 0;
    }
    ex.sync();
    ex.parn(n, [&](int s) {
        const double x = st->acc2[s];
        double px, py;
        lane_position(c.lanes[st->lane[s]], x, 0.0, px, py);
        st->pos[s] = d2{px, py};
        st->h[s] = lane_heading_at(c.lanes[st->lane[s]], x);
        st->cs[s] = d2{cos(st->h[s]), sin(st->h[s])};
        st->timer[s] = s == 0 ? 0.0 : py_mod1((px + py) * kPi);
        st->acc2[s] = 0;
    });
    rebuild_tables(c, ex);
}

// IntersectionEnv._make_vehicles (intersection_env.py:251-318)
// In three parts, so that the kernels that reset many envs per CTA can run the warm-up sub-steps of all their envs in
// lockstep (one CTA-wide alignment per sub-step, k_reset_list / k_regen_list): begin, warm-up loop, end.
template <class C, class Exec>
TT_HD void reset_intersection_begin(C& c, Exec& ex, uint64_t seed, int64_t genv, int episode) {
    auto* st = c.st;
    const SceneDev* sc = c.sc;
    const ttrl_reset_params& rp = sc->rp;
    const ttrl_config& cfg = sc->cfg;
    reset_scalars(c, ex, episode);
    // staggered spawn attempts (:265-266), default _spawn_vehicle arguments (spawn_probability 0.6)
    for (int t = 0; t < rp.n_vehicles - 1; ++t) {
        ttrl_spawn_draw d;
        device_spawn_draw(seed, genv, reset_attempt_counter(episode, t), d);
        SpawnParams sp{rp.spawn_longitudinal[t], 1.0, 1.0, 0.6, 0};
        spawn_vehicle(c, ex, d, sp);
    }
    rebuild_tables(c, ex);
}
template <class C, class Exec>
TT_HD void reset_intersection_end(C& c, Exec& ex, uint64_t seed, int64_t genv, int episode) {
    auto* st = c.st;
    const SceneDev* sc = c.sc;
    const ttrl_reset_params& rp = sc->rp;
    const ttrl_config& cfg = sc->cfg;
    {   // challenger vehicle (:276-277)
        ttrl_spawn_draw d;
        device_spawn_draw(seed, genv, reset_attempt_counter(episode, rp.n_vehicles - 1), d);
        SpawnParams sp{60.0, 0.1, 0.0, 1.0, 1};
        spawn_vehicle(c, ex, d, sp);
    }
    if (ex.first()) {
        const int K = n_agents(c);
#pragma unroll 1
        for (int a = 0; a < K; ++a) {
            // ego MDPVehicle a (:286-307) at ego_longitudinal + std * N(1, 1) on (o<k>, ir<k>, 0), k = a % 4, speed = speed_limit
            double u0, u1, ud, unused;
            reset_uniforms(seed, genv, episode, 0x100u + 2u * a, u0, u1);
            reset_uniforms(seed, genv, episode, 0x101u + 2u * a, ud, unused);
            const double z = sqrt(-2.0 * log(1.0 - u0)) * cos(2 * kPi * u1);
            const int entry = (rp.ego_entry + a) % 4;
            const ttrl_lane& el = c.lanes[sc->spawn_lane[entry]];
            double px, py;
            lane_position(el, rp.ego_longitudinal + rp.ego_longitudinal_std * (1.0 + z), 0.0, px, py);
            int s = st->n;
            if (s >= c.vcap) s = c.vcap - 1;  // full: the ego replaces the last vehicle
            st->pos[s] = d2{px, py};
            st->h[s] = lane_heading_at(el, rp.ego_longitudinal);
            st->cs[s] = d2{cos(st->h[s]), sin(st->h[s])};
            st->v[s] = el.speed_limit;
            st->steer[s] = 0; st->acc[s] = 0; st->imp[s] = d2{0, 0}; st->timer[s] = 0; st->delta[s] = 4.0;
            uint64_t m;
            const int ln = table_row_and_closest(c, s, m);  // RoadObject.__init__ objects.py:45-50
            st->lane[s] = ln; st->tlane[s] = ln;
            st->sidx[s] = speed_to_index(cfg, st->v[s]);    // MDPVehicle.__init__ controller.py:283-293
            st->tspeed[s] = cfg.target_speeds[st->sidx[s]];
            st->flags[s] = TTRL_FL_MDP | TTRL_FL_CONTROLLED | (a << TTRL_FL_AGENT_SHIFT);
            st->ytimer[s] = 0;
            int dest = rp.destination;
            if (dest < 0) { dest = 1 + (int)(ud * 3.0); if (dest > 3) dest = 3; }  // "o" + str(integers(1, 4))
            st->rlen[s] = planned_route(c, ln, entry, dest, s);
            const int n0 = s + 1;
            // "prevent early collisions" (:313-318): list.remove() while iterating skips the element after each removal
            int* order = st->chg;
            int len = n0;
            for (int k = 0; k < n0; ++k) order[k] = k;
            for (int i = 0; i < len; ++i) {
                const int v = order[i];
                if (v == s) continue;
                const double dx = st->pos[v].x - px, dy = st->pos[v].y - py;
                if (sqrt(dx * dx + dy * dy) < 20) {
                    for (int k = i; k + 1 < len; ++k) order[k] = order[k + 1];
                    --len;
                }
            }
            for (int dst = 0; dst < len; ++dst) {
                const int src = order[dst];
                if (src != dst) { SlotRegs r; slot_read(c, src, r); slot_write(c, dst, r); }
            }
            st->n = len;
        }
        for (int t = 0; t < st->n; ++t)
            if (st->flags[t] & TTRL_FL_CONTROLLED) st->egos[agent_of(st->flags[t])] = t;
        st->ego = st->egos[0];
        st->steps = 0;
        st->n_chg = 0;
    }
    ex.sync();
    rebuild_tables(c, ex);
}

// Scripted cast: RoundaboutEnv._make_vehicles (roundabout_env.py:326-387), UTurnEnv._make_vehicles (u_turn_env.py:173-271).
// The members are independent of each other, one task each.  Draws of member m: stream index 0x200 + 2m -> the two normals
// (Box-Muller pair: longitudinal, speed), 0x201 + 2m -> destination choice and DELTA.
template <class C, class Exec>
TT_HD void reset_cast(C& c, Exec& ex, uint64_t seed, int64_t genv, int episode) {
    auto* st = c.st;
    const SceneDev* sc = c.sc;
    const ttrl_reset_params& rp = sc->rp;
    const ttrl_config& cfg = sc->cfg;
    reset_scalars(c, ex, episode);
    const int n = rp.n_vehicles < c.vcap ? rp.n_vehicles : c.vcap;
    ex.parn(n, [&](int s) {
        const ttrl_cast_member& m = rp.cast[s];
        double u0, u1, ud, ue;
        reset_uniforms(seed, genv, episode, 0x200u + 2u * s, u0, u1);
        reset_uniforms(seed, genv, episode, 0x201u + 2u * s, ud, ue);
        const double rad = sqrt(-2.0 * log(1.0 - u0));
        const double z_lon = rad * cos(2 * kPi * u1), z_speed = rad * sin(2 * kPi * u1);
        const ttrl_lane& l = c.lanes[m.lane];
        const double lon = m.mdp ? m.longitudinal : m.longitudinal + z_lon * m.longitudinal_std;
        const double speed = m.mdp ? m.speed : m.speed + z_speed * m.speed_std;
        double px, py;
        lane_position(l, lon, 0.0, px, py);                                   // make_on_lane objects.py:67-89
        const double hd = lane_heading_at(l, m.mdp ? m.heading_longitudinal : lon);
        st->pos[s] = d2{px, py}; st->h[s] = hd; st->v[s] = speed;
        st->cs[s] = d2{cos(hd), sin(hd)};
        uint64_t mask;
        const int ln = table_row_and_closest(c, s, mask);                     // RoadObject.__init__ objects.py:45-50
        st->lane[s] = ln; st->tlane[s] = ln;
        st->tspeed[s] = speed;                                                // controller.py:47
        if (m.mdp) {                                                          // MDPVehicle.__init__ controller.py:283-293
            st->flags[s] = TTRL_FL_MDP | TTRL_FL_CONTROLLED;
            st->sidx[s] = speed_to_index(cfg, speed);
            st->tspeed[s] = cfg.target_speeds[st->sidx[s]];
            st->delta[s] = 4.0;
        } else {
            st->timer[s] = py_mod1((px + py) * kPi);                          // behavior.py:64
            st->delta[s] = (m.randomize && !is_linear(c)) ? 3.5 + ue * (4.5 - 3.5) : 4.0;   // behavior.py:66-69
            if (is_linear(c)) {                                               // LinearVehicle: class defaults or randomize_behavior
                double lu[6];
                reset_uniforms(seed, genv, episode, 0x300u + 4u * s, lu[0], lu[1]);
                reset_uniforms(seed, genv, episode, 0x301u + 4u * s, lu[2], lu[3]);
                reset_uniforms(seed, genv, episode, 0x302u + 4u * s, lu[4], lu[5]);
                for (int k = 0; k < TTRL_NLIN; ++k)
                    c.lin[k * C::V + s] = m.randomize ? cfg.lin_lo[k] + lu[k] * (cfg.lin_hi[k] - cfg.lin_lo[k]) : cfg.lin_default[k];
            }
        }
        if (m.n_dest <= 0) { st->rlen[s] = -1; return; }                      // route None
        int k = m.n_dest == 1 ? 0 : (int)(ud * m.n_dest);                     // np_random.choice(destinations)
        if (k > m.n_dest - 1) k = m.n_dest - 1;
        const int dest = m.dest[k], road = c.lanes[ln].road;
        const int nr = rp.cast_route_len[road][dest];                         // plan_route_to controller.py:71-87
        Route q{};
        route_set(q, 0, road, c.lanes[ln].lane_id);
        for (int j = 0; j < nr && j + 1 < TTRL_ROUTE_CAP; ++j) route_set(q, j + 1, rp.cast_route_road[road][dest][j], -1);
        route_store(st, s, q);
        st->rlen[s] = 1 + nr;
    });
    if (ex.first()) {
        st->n = n;
        for (int s = 0; s < n; ++s) if (rp.cast[s].mdp) { st->ego = s; st->egos[0] = s; }
    }
    ex.sync();
    rebuild_tables(c, ex);
}

template <class C, class Exec>
TT_HD void reset_intersection(C& c, Exec& ex, uint64_t seed, int64_t genv, int episode) {
    reset_intersection_begin(c, ex, seed, genv, episode);
    for (int k = 0; k < c.sc->rp.warmup_substeps; ++k) env_substep(c, ex, nullptr);  // (:267-274) road.act(); road.step(1/sf)
    reset_intersection_end(c, ex, seed, genv, episode);
}

template <class C, class Exec>
TT_HDN void env_reset(C& c, Exec& ex, uint64_t seed, int64_t genv, int episode) {  // cold path: keep it out of the step loop's code
    // the plain profile only ever resets highways (ttrl_sim_set_reset_params moves a sim with another reset scene to the general
    // profile): the intersection reset would drag a second copy of the whole sub-step into its kernel
    if (!C::kPlain && c.sc->rp.scene == 1) reset_intersection(c, ex, seed, genv, episode);
    else if (!C::kPlain && c.sc->rp.scene == 2) reset_cast(c, ex, seed, genv, episode);
    else reset_highway(c, ex, seed, genv, episode);
}

// env_reset for the teams of a multi-env CTA, `active` = this team has an env to reset.  EVERY team of the CTA calls it: the
// alignment barrier of the warm-up loop is ONE instruction that active and idle teams reach alike -- a CTA barrier executed
// from different instructions by two sub-warp teams of the same warp deadlocks (tools/probes/bar_divergence.cu).
// `part` of `parts`: the reset may be spread over several launches (k_regen_list: short kernels that the next step does not have
// to wait long for); the state between two parts lives in the record `rec` of `g` -- a sub-step is a pure function of the stored
// state (what the resynced parity tests rely on), so the split changes nothing.  The caller stores the state after each part.
template <class C, class Exec>
TT_HD void env_reset_lockstep(C& c, Exec& ex, bool active, uint64_t seed, int64_t genv, int episode, int part = 0, int parts = 1,
                              const GlobalState* g = nullptr, int rec = 0) {
    if (C::kPlain || c.sc->rp.scene != 1) {  // uniform over the CTA: no warm-up sub-steps, everything in the last part
        if (active && part == parts - 1) env_reset(c, ex, seed, genv, episode);
        return;
    }
    if (active) {
        if (part == 0) reset_intersection_begin(c, ex, seed, genv, episode);
        else load_env(c, ex, *g, rec);
    }
    const int warmup = c.sc->rp.warmup_substeps;
    const int k0 = (int)((long long)warmup * part / parts), k1 = (int)((long long)warmup * (part + 1) / parts);
    for (int k = k0; k < k1; ++k) env_substep_lockstep(c, ex, active, nullptr);
    if (active && part == parts - 1) reset_intersection_end(c, ex, seed, genv, episode);
}

// ------------------------------------------------------------------------------------------------
// one env.step(): AbstractEnv.step abstract.py:224-250 + IntersectionEnv.step intersection_env.py:135-139
// ------------------------------------------------------------------------------------------------
struct StepIO {
    const int32_t* actions;        // [E][K] or null (K = controlled vehicles per env)
    float* agent_reward;           // [E][K] or null: info["agents_rewards"] (intersection_env.py:121-129)
    uint8_t* agent_terminated;     // [E][K] or null: info["agents_terminated"]
    float* obs;                    // [E][obs_size] or null (obs_size = K observations)
    float* reward;                 // [E] or null
    uint8_t* terminated;           // [E] or null
    uint8_t* truncated;            // [E] or null
    const ttrl_spawn_draw* draws;  // [E] injected spawn draws or null (device Philox)
    int32_t* spawn_accepted;       // [E] or null
    const int32_t* inv_perm;       // [E][K][obs_vehicles-1] or null
    double* stats;                 // [kStatFields][E] per-env accumulators
    double* info;                  // [TTRL_NINFO][E] or null: info["speed"], ["crashed"], ["rewards"] (abstract.py:169-186)
    float* final_obs;              // [E][obs_size] or null: terminal observation of the envs that finish in this step
    GlobalState pool;              // reset pool (pool.E == 0: none)
    int32_t* done_list;            // device autoreset with warm-up: finished envs are queued here (done_count) and reset by
    int32_t* done_count;           //   k_reset_list right after the step, packed and phase-aligned; null: reset inside the step
    // asynchronous device reset (TTRL_AUTORESET_DEVICE_ASYNC): the NEXT episode of every env is generated ahead of time into
    // `shadow` by k_regen_list on a side stream; shadow_ready[e] = episode number held by the shadow of env e (-1: none),
    // written after the record (release) and read before it (acquire).  A finished env whose shadow holds episode + 1
    // restarts from it inside the step; otherwise it takes the done_list path.  Either way (e, episode + 2) is queued.
    // The shadows form a ring of `shadow_depth` episodes per env: episode p of env e lives in record (p % depth) * E + e, so an
    // env whose episodes end in quick succession still finds the next one ready (the synchronous fallback costs the latency of
    // a whole reset -- 45 sub-steps -- on the step's critical path even for a single env).  shadow_queued[record] = the episode
    // that record holds or is queued for (-1: none).
    GlobalState shadow;            // shadow.E == 0: off; else shadow_depth * E records
    int32_t* shadow_ready;         // [shadow_depth * E]
    int32_t* shadow_queued;        // [shadow_depth * E]
    int shadow_depth;
    int32_t* regen_list;           // [2 E]: (env, episode to generate) pairs
    int32_t* regen_count;
    int autoreset;
    uint64_t seed;
    int64_t first_global_env;
    int obs_size;
    // Env classification of this step (k_classify; null: CTA b, team t steps env b G + t).  The envs are binned by
    // (size class, regulation phase): cls_list[b * E ..] holds the cls_count[b] envs of bin b, and a launch covers the bins
    // cls_first .. cls_first + cls_n - 1, every CTA taking G envs of ONE bin.  Phase bins keep the teams of a CTA on the same
    // regulation ticks (RegulatedRoad.step regulation.py:28-32 fires every 7th road step, at a phase that depends on the env's
    // episode length so far): the per-sub-step alignment barrier otherwise makes every sub-step wait for the few teams that tick.
    // Size classes let envs that currently hold few vehicles run in a kernel instantiated for a smaller slot capacity.
    const int32_t* cls_list;
    const int32_t* cls_count;
    int cls_first, cls_n;
    // a plain (not binned) launch may cover a sub-range of the envs: env_first .. env_first + env_count - 1 (env_count 0: all).
    // The host-buffer step of scenes with large observations runs the batch as a few such slices on two streams, so that the
    // device-to-host copy of one slice's observations overlaps the kernel of the next (ttrl_sim_step_pinned).
    int env_first, env_count;
};
constexpr int kClsPhases = 8;              // regulation phase bins per size class
constexpr int kClsBins = 2 * kClsPhases;   // size class 0 (small) bins 0..7, class 1 (large) bins 8..15

// env_step in three parts (load, F sub-steps, finish) so that k_step can run the sub-steps of the envs of a CTA in lockstep,
// with the alignment barrier as ONE instruction outside the per-team branches (see env_reset_lockstep).
template <class C, class Exec>
TT_HD void env_step_load(C& c, Exec& ex, const GlobalState& g, const StepIO& io, int e) {
    load_env(c, ex, g, e);
    if (ex.first()) c.st->time += 1 / c.sc->cfg.policy_frequency;
    ex.sync();
}
template <class C, class Exec>
TT_HD void env_step_finish(C& c, Exec& ex, const GlobalState& g, const StepIO& io, int e, double veh_steps) {
    auto* st = c.st;
    const SceneDev* sc = c.sc;
    const ttrl_config& cfg = sc->cfg;
    const int K = n_agents(c);
    const int32_t* actions = io.actions ? io.actions + (size_t)e * K : nullptr;
    float* obs = io.obs ? io.obs + (size_t)e * io.obs_size : nullptr;
    const int32_t* perm = io.inv_perm ? io.inv_perm + (size_t)e * K * (cfg.obs_vehicles - 1) : nullptr;
    if (obs) observe(c, ex, obs, perm, io.seed, io.first_global_env + e);
    // reward / flags / episode accounting by the first thread
    if (ex.first()) {
        const int ego = st->ego;
        double r;
        double rc[4] = {0, 0, 0, 0};  // info["rewards"] (mean over the controlled vehicles, intersection_env.py:67-76)
        if (K == 1) {
            r = agent_reward(c, ego, actions ? actions[0] : -1, io.info ? rc : nullptr);
            if (io.agent_reward) io.agent_reward[e] = (float)r;
            if (io.agent_terminated) io.agent_terminated[e] = ((st->flags[ego] & TTRL_FL_CRASHED) || (!C::kPlain && has_arrived(c, ego))) ? 1 : 0;
        } else {  // sum(agent rewards) / len(controlled_vehicles) (intersection_env.py:61-65)
            double sum = 0;
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                const int v = st->egos[k];
                double rk4[4];
                const double rk = agent_reward(c, v, actions ? actions[k] : -1, rk4);
                sum = sum + rk;
                for (int q = 0; q < 4; ++q) rc[q] = rc[q] + rk4[q];
                if (io.agent_reward) io.agent_reward[(size_t)e * K + k] = (float)rk;
                if (io.agent_terminated) io.agent_terminated[(size_t)e * K + k] = ((st->flags[v] & TTRL_FL_CRASHED) || has_arrived(c, v)) ? 1 : 0;  // _agent_is_terminal :113-115
            }
            r = sum / K;
            for (int q = 0; q < 4; ++q) rc[q] = rc[q] / K;
        }
        if (io.info) {  // AbstractEnv._info abstract.py:169-186: self.vehicle = controlled_vehicles[0]
            const size_t E = (size_t)g.E;
            io.info[TTRL_INFO_SPEED * E + e] = st->v[ego];
            io.info[TTRL_INFO_CRASHED * E + e] = (st->flags[ego] & TTRL_FL_CRASHED) ? 1.0 : 0.0;
            for (int q = 0; q < 4; ++q) io.info[(TTRL_INFO_REWARDS + q) * E + e] = rc[q];
        }
        const bool term = is_terminated(c), trunc = st->time >= cfg.duration;
        if (io.reward) io.reward[e] = (float)r;
        if (io.terminated) io.terminated[e] = term ? 1 : 0;
        if (io.truncated) io.truncated[e] = trunc ? 1 : 0;
        st->ret += r;
        st->done = (term || trunc) ? 1 : 0;
        if (io.stats) {
            double* s = io.stats + e;
            const int E = g.E;
            s[5 * E] += st->v[ego];
            s[6 * E] += veh_steps;
            s[7 * E] += 1;
            if (st->done) {
                s[0 * E] += 1;
                s[1 * E] += st->ret;
                s[2 * E] += st->time * cfg.policy_frequency;
                s[3 * E] += (st->flags[ego] & TTRL_FL_CRASHED) ? 1 : 0;
                s[4 * E] += (!C::kPlain && cfg.reward_type == TTRL_REWARD_INTERSECTION && has_arrived(c, ego)) ? 1 : 0;
            }
        }
    }
    ex.sync();
    // gymnasium's info["final_observation"]: the autoreset below replaces the terminal observation in `obs`
    if (io.final_obs && obs && st->done && io.autoreset != TTRL_AUTORESET_OFF) {  // uniform: st->done is in shared memory
        float* fo = io.final_obs + (size_t)e * io.obs_size;
        ex.parn(io.obs_size, [&](int k) { fo[k] = obs[k]; });  // written by this team above, barrier since
    }
    if (!C::kPlain && cfg.spawn_enabled) {
        clear_vehicles(c, ex);
        ttrl_spawn_draw d;
        bool have = false;
        if (io.draws) { d = io.draws[e]; have = true; }
        else if (io.seed != 0) { device_spawn_draw(io.seed, io.first_global_env + e, (uint64_t)st->steps + ((uint64_t)st->episode << 32), d); have = true; }
        if (have) {
            SpawnParams sp{0.0, 1.0, 1.0, cfg.spawn_probability, 0};
            spawn_vehicle(c, ex, d, sp);
            if (io.spawn_accepted && ex.first()) io.spawn_accepted[e] = st->flag0;
            if (io.stats && ex.first() && st->flag1) io.stats[(size_t)8 * g.E + e] += 1;  // spawn_capacity_rejects
        } else if (io.spawn_accepted && ex.first()) {
            io.spawn_accepted[e] = 0;
        }
        ex.sync();
    }
    if (st->done && io.autoreset == TTRL_AUTORESET_DEVICE && sc->have_rp && io.done_list) {
        // resets with warm-up sub-steps are batched: a lone resetting team would keep its whole CTA (and SM) waiting
        const int next_episode = st->episode + 1;
        bool from_shadow = false;
        const int D = io.shadow_depth > 0 ? io.shadow_depth : 1;
        const int rec = (next_episode % D) * g.E + e;
        if (io.shadow.E > 0) {
            if (ex.first()) st->flag0 = ex.load_acquire(&io.shadow_ready[rec]) == next_episode ? 1 : 0;
            ex.sync();
            from_shadow = st->flag0 != 0;  // uniform: shared memory
            ex.sync();
        }
        if (from_shadow) {
            load_env<true>(c, ex, io.shadow, rec);  // episode, time, counters come with the record (env_reset wrote them)
            if (obs) observe(c, ex, obs, perm, io.seed, io.first_global_env + e);
        }
        if (ex.first()) {
            if (!from_shadow) {
                io.done_list[ex.atomic_add_global(io.done_count, 1)] = e;
                if (io.stats) io.stats[(size_t)9 * g.E + e] += 1;  // sync_resets
            }
            if (io.shadow.E > 0) {
                // queue every episode of the ring's horizon that no record holds or waits for yet: next_episode + 1 .. + D
                // (the last one reuses the record of next_episode, consumed above or stale)
                for (int k = 1; k <= D; ++k) {
                    const int ep = next_episode + k, r = (ep % D) * g.E + e;
                    if (io.shadow_queued[r] == ep) continue;
                    io.shadow_queued[r] = ep;
                    io.shadow_ready[r] = -1;  // invalid until the queued regeneration has finished
                    const int q = ex.atomic_add_global(io.regen_count, 1);
                    io.regen_list[2 * q] = e;
                    io.regen_list[2 * q + 1] = ep;
                }
            }
        }
        ex.sync();
    } else if (st->done && io.autoreset == TTRL_AUTORESET_DEVICE && sc->have_rp) {  // uniform: st->done is in shared memory
        const int episode = st->episode + 1;
        ex.sync();
        env_reset(c, ex, io.seed, io.first_global_env + e, episode);
        if (obs) observe(c, ex, obs, perm, io.seed, io.first_global_env + e);
    } else if (st->done && io.autoreset == TTRL_AUTORESET_POOL && io.pool.E > 0) {
        const int episode = st->episode + 1;
        const int slot = (int)(((long long)e + (long long)episode * g.E) % io.pool.E);
        ex.sync();
        load_env(c, ex, io.pool, slot);
        if (ex.first()) { st->episode = episode; st->done = 0; }
        ex.sync();
        if (obs) observe(c, ex, obs, perm, io.seed, io.first_global_env + e);
    }
    store_env(c, ex, g, e);
}
template <class C, class Exec>
TT_HD void env_step(C& c, Exec& ex, const GlobalState& g, const StepIO& io, int e) {
    env_step_load(c, ex, g, io, e);
    const int32_t* actions = io.actions ? io.actions + (size_t)e * n_agents(c) : nullptr;
    double veh_steps = 0;
    for (int f = 0; f < c.sc->F; ++f) {
        env_substep(c, ex, actions);
        veh_steps += c.st->n;
    }
    env_step_finish(c, ex, g, io, e, veh_steps);
}

}  // namespace ttrl
