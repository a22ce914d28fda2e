// TEST-ONLY sequential emulation of the CUDA kernels' phases (no GPU in the build container).
//
// Compiles topotrafficrl_b200/csrc/ttrl_core.cuh -- the exact device logic -- for the host with an Exec
// policy that runs the V "threads" of each phase one after the other (a __syncthreads() boundary becomes
// the end of the loop).  Lets `pytest -m "not gpu"` check the device code against the CPU oracle and the
// golden vectors before any GPU time is spent.  It is NOT reachable from the product package: the product
// path has no CPU fallback and fails loudly without the CUDA library.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../topotrafficrl_b200/csrc/ttrl_core.cuh"

using namespace ttrl;

struct HostExec {
    int T;  // = V: every slot is a "thread"
    bool first() const { return true; }  // single-thread sections run once
    void sync() {}
    void align() {}
    template <int K> void align_at() {}
    template <class F> void par(F f) { for (int t = 0; t < T; ++t) f(t); }
    template <class F> void parn(int n, F f) { for (int t = 0; t < n; ++t) f(t); }
    template <class F> bool any(int n, F f) { bool r = false; for (int t = 0; t < n; ++t) r = f(t) || r; return r; }
    template <class F1, class F2> void par2(F1 f1, F2 f2) {
        std::vector<SlotRegs> r(T);
        std::vector<int> dst(T);
        for (int t = 0; t < T; ++t) f1(t, r[t], dst[t]);
        for (int t = 0; t < T; ++t) f2(t, r[t], dst[t]);
    }
    // same reduction the device does with shuffles: first axis of minimal absd wins
    template <class FA, class FP> void sat_pairs(int np, FA fa, FP fp) {
        for (int q = 0; q < np; ++q) {
            bool inter = true, will = true;
            double bd = INFINITY, nx = 0, ny = 0;
            for (int axis = 0; axis < 8; ++axis) {
                const AxisRes r = fa(q, axis);
                if (r.fl & 1) inter = false;
                if (r.fl & 2) will = false;
                if (r.absd < bd) { bd = r.absd; nx = r.nx; ny = r.ny; }
            }
            fp(q, inter, will, bd, nx, ny);
        }
    }
    void atomic_min(int32_t* a, int32_t v) { if (v < *a) *a = v; }
    void atomic_max(int32_t* a, int32_t v) { if (v > *a) *a = v; }
    void atomic_or(uint32_t* a, uint32_t v) { *a |= v; }
    int atomic_add(int32_t* a, int32_t v) { const int o = *a; *a += v; return o; }
    int atomic_add_global(int32_t* a, int32_t v) { return atomic_add(a, v); }
    int load_acquire(const int32_t* a) { return *a; }
};

template <int V, int P>
struct HostEnv {
    EnvState<V> st;
    std::vector<d2> SR;
    std::vector<uint32_t> lmask, pbits;
    std::vector<double> pred;
    std::vector<float> obs_s;
    std::vector<uint32_t> perm_s;
    std::vector<double> lin;
    std::vector<int32_t> cell;
    EnvCtx<V, P> c;
    HostEnv(const SceneDev* sc, int vcap) {
        memset(&st, 0, sizeof st);
        const ttrl_config& cfg = sc->cfg;
        SR.assign((size_t)V * cfg.n_lanes, d2{0.0, 0.0});
        lmask.assign((size_t)((V + 31) / 32) * cfg.n_lanes, 0u);
        pred.assign((size_t)4 * V, 0.0);
        pbits.assign((size_t)(V * (V - 1) / 2 + 31) / 32 + 1, 0u);
        obs_s.assign((size_t)(cfg.obs_vehicles * cfg.n_features + 4), 0.f);
        lin.assign((size_t)TTRL_NLIN * V, 0.0);
        perm_s.assign((size_t)(2 * cfg.obs_vehicles + 4), 0u);
        cell.assign((size_t)(cfg.grid_w * cfg.grid_h + 4 + TTRL_MAX_TTC_CELLS), 0);
        c.st = &st; c.sc = sc; c.lanes = sc->lanes; c.SR = SR.data(); c.NC = P == 1 ? cfg.n_lanes : sc->n_curved; c.lmask = lmask.data();
        c.pred = cfg.regulated ? pred.data() : nullptr; c.pbits = cfg.regulated ? pbits.data() : nullptr; c.obs_s = obs_s.data(); c.perm_s = perm_s.data(); c.lin = cfg.vehicle_model == TTRL_VEHICLE_LINEAR ? lin.data() : nullptr; c.cell = cell.data();
        c.L = cfg.n_lanes; c.vcap = vcap;
        c.gap_den = 2 * sqrt(-cfg.comfort_acc_max * cfg.comfort_acc_min);
        c.tan_max_steer = tan(kPi / 3);
    }
};

// scene profile like the library's (ttrl_kern.cu configure): 3 = LinearVehicle traffic, 2 = several controlled vehicles, 0 = general single-agent
// (the "plain" profile 1 only removes code paths that plain scenes never take)
#define DISPATCH_V(vcap, ...)                                        \
    do {                                                             \
        if (vcap <= 32) { constexpr int V = 32; __VA_ARGS__; }       \
        else if (vcap <= 64) { constexpr int V = 64; __VA_ARGS__; }  \
        else if (vcap <= 128) { constexpr int V = 128; __VA_ARGS__; }\
        else { constexpr int V = 256; __VA_ARGS__; }                 \
    } while (0)
#define DISPATCH(vcap, ...)                                                                    \
    do {                                                                                       \
        if (sc->cfg.vehicle_model == TTRL_VEHICLE_LINEAR) { constexpr int P = 3; DISPATCH_V(vcap, __VA_ARGS__); } \
        else if (sc->cfg.controlled_vehicles > 1) { constexpr int P = 2; DISPATCH_V(vcap, __VA_ARGS__); } \
        else { constexpr int P = 0; DISPATCH_V(vcap, __VA_ARGS__); }                           \
    } while (0)

extern "C" {

SceneDev* emu_scene_create(const ttrl_config* cfg, const ttrl_lane* lanes, const ttrl_road* roads, const int32_t* node_first,
                           const int32_t* node_roads) {
    SceneDev* s = (SceneDev*)calloc(1, sizeof(SceneDev));
    s->cfg = *cfg;
    memcpy(s->lanes, lanes, sizeof(ttrl_lane) * cfg->n_lanes);
    s->n_curved = assign_cache_columns(s->lanes, cfg->n_lanes, s->curved_lane);
    s->arc_tasks = getenv("TTRL_EMU_ARC_TASKS") ? atoi(getenv("TTRL_EMU_ARC_TASKS")) : use_arc_tasks(s->lanes, cfg->n_lanes, s->n_curved);
    memcpy(s->roads, roads, sizeof(ttrl_road) * cfg->n_roads);
    memcpy(s->node_first, node_first, sizeof(int32_t) * (cfg->n_nodes + 1));
    memcpy(s->node_roads, node_roads, sizeof(int32_t) * node_first[cfg->n_nodes]);
    s->F = (int)floor(cfg->simulation_frequency / cfg->policy_frequency);
    s->dt = 1 / cfg->simulation_frequency;
    s->reg_period = (int)(1 / s->dt / 2);
    return s;
}
void emu_scene_destroy(SceneDev* s) { free(s); }
// LinearVehicle parameter blocks of the state / pool buffers the following calls operate on ([TTRL_NLIN][E][V]; null: class defaults)
static thread_local double* g_lin = nullptr;
static thread_local double* g_pool_lin = nullptr;
void emu_set_linear_params(double* lin, double* pool_lin) { g_lin = lin; g_pool_lin = pool_lin; }
void emu_scene_set_spawn_routes(SceneDev* s, const int32_t* spawn_lane, const int32_t* route_len, const int32_t* route_road) {
    memcpy(s->spawn_lane, spawn_lane, sizeof(int32_t) * 4);
    memcpy(s->spawn_route_len, route_len, sizeof(int32_t) * 16);
    memcpy(s->spawn_route_road, route_road, sizeof(int32_t) * 16 * TTRL_ROUTE_CAP);
}

void emu_substep(const SceneDev* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int Vs, const int32_t* actions) {
    GlobalState g{vd, vi, ei, ed, E, Vs, g_lin};
    DISPATCH(Vs, {
        HostEnv<V, P> env(sc, Vs);
        HostExec ex{V};
        for (int e = 0; e < E; ++e) {
            load_env(env.c, ex, g, e);
            env_substep(env.c, ex, actions ? actions + (size_t)e * n_agents(env.c) : nullptr);
            store_env(env.c, ex, g, e);
        }
    });
}

void emu_observe(const SceneDev* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int Vs, float* obs, int obs_size,
                 const int32_t* inv_perm) {
    GlobalState g{vd, vi, ei, ed, E, Vs, g_lin};
    DISPATCH(Vs, {
        HostEnv<V, P> env(sc, Vs);
        HostExec ex{V};
        for (int e = 0; e < E; ++e) {
            load_env(env.c, ex, g, e);
            observe(env.c, ex, obs + (size_t)e * obs_size, inv_perm ? inv_perm + (size_t)e * n_agents(env.c) * (sc->cfg.obs_vehicles - 1) : nullptr);
        }
    });
}

void emu_step(const SceneDev* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int Vs, const int32_t* actions,
              float* obs, int obs_size, float* reward, uint8_t* terminated, uint8_t* truncated, const ttrl_spawn_draw* draws,
              int32_t* accepted, const int32_t* inv_perm, double* stats, int pool_size, double* pvd, int32_t* pvi, int32_t* pei,
              double* ped, int autoreset, uint64_t seed, int64_t first_global_env, float* agent_reward, uint8_t* agent_terminated,
              double* info, float* final_obs) {
    GlobalState g{vd, vi, ei, ed, E, Vs, g_lin};
    StepIO io{};
    io.agent_reward = agent_reward; io.agent_terminated = agent_terminated;
    io.info = info; io.final_obs = final_obs;
    io.actions = actions; io.obs = obs; io.reward = reward; io.terminated = terminated; io.truncated = truncated;
    io.draws = draws; io.spawn_accepted = accepted; io.inv_perm = inv_perm; io.stats = stats;
    io.pool = GlobalState{pvd, pvi, pei, ped, pool_size, Vs, g_pool_lin};
    io.autoreset = autoreset; io.seed = seed; io.first_global_env = first_global_env; io.obs_size = obs_size;
    DISPATCH(Vs, {
        HostEnv<V, P> env(sc, Vs);
        HostExec ex{V};
        for (int e = 0; e < E; ++e) env_step(env.c, ex, g, io, e);
    });
}

void emu_spawn(const SceneDev* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int Vs, const ttrl_spawn_draw* draws,
               double longitudinal, double position_deviation, double speed_deviation, double spawn_probability, int go_straight,
               int32_t* accepted) {
    GlobalState g{vd, vi, ei, ed, E, Vs, g_lin};
    SpawnParams sp{longitudinal, position_deviation, speed_deviation, spawn_probability, go_straight};
    DISPATCH(Vs, {
        HostEnv<V, P> env(sc, Vs);
        HostExec ex{V};
        for (int e = 0; e < E; ++e) {
            load_env(env.c, ex, g, e);
            spawn_vehicle(env.c, ex, draws[e], sp);
            if (accepted) accepted[e] = env.st.flag0;
            store_env(env.c, ex, g, e);
        }
    });
}

void emu_scene_set_reset_params(SceneDev* s, const ttrl_reset_params* rp) { s->rp = *rp; s->have_rp = 1; }
void emu_reset(const SceneDev* sc, double* vd, int32_t* vi, int32_t* ei, double* ed, int E, int Vs, uint64_t seed, int64_t first_global_env, int episode) {
    GlobalState g{vd, vi, ei, ed, E, Vs, g_lin};
    DISPATCH(Vs, {
        HostEnv<V, P> env(sc, Vs);
        HostExec ex{V};
        for (int e = 0; e < E; ++e) {
            env_reset(env.c, ex, seed, first_global_env + e, episode);
            store_env(env.c, ex, g, e);
        }
    });
}
void emu_reset_attempt_draw(uint64_t seed, int64_t env, int episode, int attempt, ttrl_spawn_draw* out) {
    device_spawn_draw(seed, env, reset_attempt_counter(episode, attempt), *out);
}
void emu_reset_uniforms(uint64_t seed, int64_t env, int episode, uint32_t idx, double* out) { reset_uniforms(seed, env, episode, idx, out[0], out[1]); }
void emu_device_spawn_draw(uint64_t seed, int64_t env, uint64_t counter, ttrl_spawn_draw* out) { device_spawn_draw(seed, env, counter, *out); }
}
