// ttrl_sim.cu -- sm_100a kernels and the C ABI (include/ttrl_b200.h) of the batched simulator.
//
// Launch shape: one CTA per env instance, blockDim = V threads (V = slot capacity rounded up to 32/64/128/256),
// thread t <-> vehicle slot t.  k_step keeps the env's state in shared memory across the F sub-steps of an
// env-step: one coalesced HBM read and one write of the state per env-step.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "ttrl_core.cuh"

using namespace ttrl;

// ------------------------------------------------------------------------------------------------
// device execution policy: one CTA of T threads ("team") per env; slots / tasks are strided over the team.
// T == 32: the team is one warp, barriers are __syncwarp().
// ------------------------------------------------------------------------------------------------
template <int V, int T>
struct DevExec {
    int tid;
    __device__ __forceinline__ bool first() const { return tid == 0; }
    __device__ __forceinline__ void sync() {
        if (T == 32) __syncwarp(); else __syncthreads();
    }
    template <class F> __device__ __forceinline__ void par(F f) {
#pragma unroll 1
        for (int t = tid; t < V; t += T) f(t);
        sync();
    }
    template <class F> __device__ __forceinline__ void parn(int n, F f) {
#pragma unroll 1
        for (int t = tid; t < n; t += T) f(t);
        sync();
    }
    template <class F> __device__ __forceinline__ bool any(int n, F f) {
        int p = 0;
#pragma unroll 1
        for (int t = tid; t < n; t += T) p |= f(t) ? 1 : 0;
        if (T == 32) return __any_sync(0xffffffffu, p) != 0;
        return __syncthreads_or(p) != 0;
    }
    template <class F1, class F2> __device__ __forceinline__ void par2(F1 f1, F2 f2) {
        constexpr int R = (V + T - 1) / T;
        SlotRegs r[R];
        int dst[R];
#pragma unroll
        for (int k = 0; k < R; ++k) { dst[k] = -1; if (tid + k * T < V) f1(tid + k * T, r[k], dst[k]); }
        sync();
#pragma unroll
        for (int k = 0; k < R; ++k) if (tid + k * T < V) f2(tid + k * T, r[k], dst[k]);
        sync();
    }
    // 8 lanes per candidate pair, one separating axis each; the pair's result is reduced with shuffles:
    // intersecting / will_intersect = no lane reports a separation, translation axis = first axis of minimal absd.
    template <class FA, class FP> __device__ __forceinline__ void sat_pairs(int np, FA fa, FP fp) {
        const int total = np * 8;
#pragma unroll 1
        for (int base = 0; base < total; base += T) {
            const int idx = base + tid;
            const bool act = idx < total;
            const int q = idx >> 3, axis = idx & 7;
            AxisRes r;
            r.absd = INFINITY; r.nx = 0; r.ny = 0; r.fl = 0;
            if (act) r = fa(q, axis);
            const unsigned gbase = (unsigned)(tid & 31) & ~7u;
            const unsigned sep_now = __ballot_sync(0xffffffffu, act && (r.fl & 1));
            const unsigned sep_after = __ballot_sync(0xffffffffu, act && (r.fl & 2));
            const bool inter = ((sep_now >> gbase) & 0xFFu) == 0, will = ((sep_after >> gbase) & 0xFFu) == 0;
            double bd = r.absd;
            int bi = axis;
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            const double nx = __shfl_sync(0xffffffffu, r.nx, (int)gbase + bi), ny = __shfl_sync(0xffffffffu, r.ny, (int)gbase + bi);
            if (act && axis == 0) fp(q, inter, will, bd, nx, ny);
        }
        sync();
    }
    __device__ __forceinline__ void atomic_min(int32_t* a, int32_t v) { atomicMin(a, v); }
    __device__ __forceinline__ void atomic_max(int32_t* a, int32_t v) { atomicMax(a, v); }
    __device__ __forceinline__ void atomic_or(uint32_t* a, uint32_t v) { atomicOr(a, v); }
    __device__ __forceinline__ int atomic_add(int32_t* a, int32_t v) { return atomicAdd(a, v); }
};

// threads per env for a slot capacity
// Threads per env ("team") and minimum resident CTAs per SM (register cap) for a slot capacity.  Tuned on B200
// (profiles/r1c_variants.txt): the kernel is latency-bound (fp64 dependency chains, instruction fetch), so
// resident warps matter more than registers per thread; 2 warps per 64-slot env with >= 10 CTAs/SM is the best point.
#ifndef TT_T64
#define TT_T64 64
#endif
#ifndef TT_MINB64
#define TT_MINB64 10
#endif
#ifndef TT_MINB32
#define TT_MINB32 1
#endif
template <int V> struct TeamOf { static constexpr int T = V / 2, MINB = 1; };
template <> struct TeamOf<32> { static constexpr int T = 32, MINB = TT_MINB32; };
template <> struct TeamOf<64> { static constexpr int T = TT_T64, MINB = TT_MINB64; };

struct SmemLayout {
    int off_lanes, off_SR, off_lmask, off_pred, off_obs, off_cell, total;
};

template <int V>
__device__ __forceinline__ void make_ctx(EnvCtx<V>& c, unsigned char* smem, const SceneDev* sc, const SmemLayout& lay, int vcap) {
    c.st = reinterpret_cast<EnvState<V>*>(smem);
    c.sc = sc;
    c.L = sc->cfg.n_lanes;
    c.vcap = vcap;
    ttrl_lane* lanes_s = reinterpret_cast<ttrl_lane*>(smem + lay.off_lanes);
    // copy the lane table into shared memory (n_lanes * 160 B) with 16-byte vector loads
    {
        const int4* src = reinterpret_cast<const int4*>(sc->lanes);
        int4* dst = reinterpret_cast<int4*>(lanes_s);
        const int n16 = c.L * (int)(sizeof(ttrl_lane) / 16);
        for (int k = threadIdx.x; k < n16; k += blockDim.x) dst[k] = __ldg(src + k);
    }
    c.lanes = lanes_s;
    c.SR = reinterpret_cast<d2*>(smem + lay.off_SR);
    c.lmask = reinterpret_cast<uint32_t*>(smem + lay.off_lmask);
    c.pred = lay.off_pred >= 0 ? reinterpret_cast<double*>(smem + lay.off_pred) : nullptr;
    c.obs_s = reinterpret_cast<float*>(smem + lay.off_obs);
    c.cell = reinterpret_cast<int32_t*>(smem + lay.off_cell);
    c.gap_den = 2 * sqrt(-sc->cfg.comfort_acc_max * sc->cfg.comfort_acc_min);  // behavior.py:214-216
    c.tan_max_steer = tan(kPi / 3);
    __syncthreads();
}

#define TT_KERNEL_PROLOGUE                                     \
    extern __shared__ __align__(16) unsigned char smem[];      \
    constexpr int T = TeamOf<V>::T;                            \
    EnvCtx<V> c;                                               \
    make_ctx<V>(c, smem, sc, lay, g.V);                        \
    DevExec<V, T> ex{(int)threadIdx.x};

template <int V>
__global__ void __launch_bounds__(TeamOf<V>::T, TeamOf<V>::MINB) k_step(const SceneDev* __restrict__ sc, GlobalState g, StepIO io, SmemLayout lay) {
    TT_KERNEL_PROLOGUE
    env_step(c, ex, g, io, (int)blockIdx.x);
}

template <int V>
__global__ void __launch_bounds__(TeamOf<V>::T) k_substep(const SceneDev* __restrict__ sc, GlobalState g, const int32_t* __restrict__ actions, SmemLayout lay) {
    TT_KERNEL_PROLOGUE
    const int e = blockIdx.x;
    load_env(c, ex, g, e);
    env_substep(c, ex, actions ? actions[e] : -1);
    store_env(c, ex, g, e);
}

template <int V>
__global__ void __launch_bounds__(TeamOf<V>::T) k_observe(const SceneDev* __restrict__ sc, GlobalState g, float* __restrict__ obs, int obs_size,
                                                          const int32_t* __restrict__ inv_perm, SmemLayout lay) {
    TT_KERNEL_PROLOGUE
    const int e = blockIdx.x;
    load_env(c, ex, g, e);
    observe(c, ex, obs + (size_t)e * obs_size, inv_perm ? inv_perm + (size_t)e * (sc->cfg.obs_vehicles - 1) : nullptr);
}

template <int V>
__global__ void __launch_bounds__(TeamOf<V>::T) k_spawn(const SceneDev* __restrict__ sc, GlobalState g, const ttrl_spawn_draw* __restrict__ draws,
                                                        SpawnParams sp, int32_t* __restrict__ accepted, SmemLayout lay) {
    TT_KERNEL_PROLOGUE
    const int e = blockIdx.x;
    load_env(c, ex, g, e);
    spawn_vehicle(c, ex, draws[e], sp);
    if (accepted && threadIdx.x == 0) accepted[e] = c.st->flag0;
    ex.sync();
    store_env(c, ex, g, e);
}

// per-env accumulators [kStatFields][E] -> 8 totals (one block per field)
__global__ void k_reduce_stats(const double* __restrict__ stats, int E, double* __restrict__ out) {
    __shared__ double sh[256];
    const int f = blockIdx.x;
    double a = 0;
    for (int e = threadIdx.x; e < E; e += blockDim.x) a += stats[(size_t)f * E + e];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[f] = sh[0];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));               \
    } while (0)

struct ttrl_sim {
    int E = 0, vcap = 0, VT = 0, device = 0, obs_size = 0;
    SceneDev scene_host;
    SceneDev* d_scene = nullptr;
    GlobalState g{};
    GlobalState pool{};
    double* d_stats = nullptr;
    double* d_stats_out = nullptr;
    ttrl_spawn_draw* d_draws = nullptr;
    bool have_draws = false;
    int32_t* d_perm = nullptr;
    bool have_perm = false;
    int32_t* d_accepted = nullptr;
    int autoreset = 0;
    uint64_t seed = 0;
    int64_t first_global_env = 0;
    SmemLayout lay{};
    int64_t launches = 0;
    // host-buffer path
    cudaStream_t hstream = nullptr;
    int32_t* h_actions = nullptr; float* h_obs = nullptr; float* h_reward = nullptr; uint8_t* h_flags = nullptr;
    int32_t* d_actions = nullptr; float* d_obs = nullptr; float* d_reward = nullptr; uint8_t* d_flags = nullptr;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <int V>
static SmemLayout make_layout(const ttrl_config& cfg) {
    SmemLayout l{};
    size_t off = align_up(sizeof(EnvState<V>), 16);
    l.off_lanes = (int)off; off += align_up(sizeof(ttrl_lane) * cfg.n_lanes, 16);
    l.off_SR = (int)off; off += sizeof(d2) * V * cfg.n_lanes;
    l.off_lmask = (int)off; off += align_up(sizeof(uint32_t) * ((V + 31) / 32) * cfg.n_lanes, 16);
    if (cfg.regulated) { l.off_pred = (int)off; off += sizeof(double) * 3 * V * kPred; } else l.off_pred = -1;
    l.off_obs = (int)off; off += align_up(sizeof(float) * (cfg.obs_type == TTRL_OBS_KINEMATICS ? cfg.obs_vehicles * cfg.n_features : 4), 16);
    l.off_cell = (int)off; off += align_up(sizeof(int32_t) * (cfg.obs_type == TTRL_OBS_GRID ? cfg.grid_w * cfg.grid_h : 4), 16);
    l.total = (int)off;
    return l;
}

template <int V>
static int configure_kernels(ttrl_sim* s) {
    s->lay = make_layout<V>(s->scene_host.cfg);
    if (s->lay.total > 227 * 1024) return fail("shared memory footprint of one env exceeds 227 KB");
    CK(cudaFuncSetAttribute(k_step<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->lay.total));
    CK(cudaFuncSetAttribute(k_substep<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->lay.total));
    CK(cudaFuncSetAttribute(k_observe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->lay.total));
    CK(cudaFuncSetAttribute(k_spawn<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->lay.total));
    return 0;
}

#define DISPATCH_V(s, EXPR)                                   \
    switch ((s)->VT) {                                        \
        case 32: { constexpr int V = 32; EXPR; } break;       \
        case 64: { constexpr int V = 64; EXPR; } break;       \
        case 128: { constexpr int V = 128; EXPR; } break;     \
        case 256: { constexpr int V = 256; EXPR; } break;     \
        default: return fail("unsupported slot capacity");    \
    }

extern "C" {

const char* ttrl_last_error(void) { return g_err.c_str(); }
void ttrl_set_error(const char* msg) { g_err = msg ? msg : ""; }
int ttrl_abi_version(void) { return 1; }
int ttrl_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(ttrl_lane);
        case 1: return (int)sizeof(ttrl_road);
        case 2: return (int)sizeof(ttrl_config);
        case 3: return (int)sizeof(ttrl_spawn_draw);
        case 4: return (int)sizeof(ttrl_episode_stats);
        case 5: return (int)sizeof(ttrl_qnet_desc);
    }
    return -1;
}

static int alloc_state(GlobalState& g, int E, int V) {
    g.E = E;
    g.V = V;
    CK(cudaMalloc(&g.vd, sizeof(double) * TTRL_ND * E * V));
    CK(cudaMalloc(&g.vi, sizeof(int32_t) * TTRL_NI * E * V));
    CK(cudaMalloc(&g.ei, sizeof(int32_t) * TTRL_NEI * E));
    CK(cudaMalloc(&g.ed, sizeof(double) * TTRL_NED * E));
    CK(cudaMemset(g.vd, 0, sizeof(double) * TTRL_ND * E * V));
    CK(cudaMemset(g.vi, 0, sizeof(int32_t) * TTRL_NI * E * V));
    CK(cudaMemset(g.ei, 0, sizeof(int32_t) * TTRL_NEI * E));
    CK(cudaMemset(g.ed, 0, sizeof(double) * TTRL_NED * E));
    return 0;
}
static void free_state(GlobalState& g) {
    cudaFree(g.vd); cudaFree(g.vi); cudaFree(g.ei); cudaFree(g.ed);
    g = GlobalState{};
}
static int upload_state(GlobalState& g, const double* vd, const int32_t* vi, const int32_t* ei, const double* ed) {
    CK(cudaMemcpy(g.vd, vd, sizeof(double) * TTRL_ND * g.E * g.V, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g.vi, vi, sizeof(int32_t) * TTRL_NI * g.E * g.V, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g.ei, ei, sizeof(int32_t) * TTRL_NEI * g.E, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g.ed, ed, sizeof(double) * TTRL_NED * g.E, cudaMemcpyHostToDevice));
    return 0;
}

int ttrl_sim_create(const ttrl_config* cfg, const ttrl_lane* lanes, const ttrl_road* roads, const int32_t* node_first,
                    const int32_t* node_roads, int num_envs, int vcap, int device, ttrl_sim** out) {
    if (!cfg || !lanes || !roads || !node_first || !node_roads || !out) return fail("null argument");
    if (num_envs <= 0 || vcap <= 0 || vcap > 256) return fail("num_envs must be > 0 and 0 < vcap <= 256");
    if (cfg->n_lanes > TTRL_MAX_LANES || cfg->n_roads > TTRL_MAX_ROADS || cfg->n_nodes > TTRL_MAX_NODES) return fail("network too large");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("no CUDA device: the simulator has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("bad device index");
    CK(cudaSetDevice(device));
    ttrl_sim* s = new ttrl_sim();
    s->E = num_envs; s->vcap = vcap; s->device = device;
    s->VT = vcap <= 32 ? 32 : vcap <= 64 ? 64 : vcap <= 128 ? 128 : 256;
    memset(&s->scene_host, 0, sizeof(SceneDev));
    s->scene_host.cfg = *cfg;
    memcpy(s->scene_host.lanes, lanes, sizeof(ttrl_lane) * cfg->n_lanes);
    memcpy(s->scene_host.roads, roads, sizeof(ttrl_road) * cfg->n_roads);
    memcpy(s->scene_host.node_first, node_first, sizeof(int32_t) * (cfg->n_nodes + 1));
    memcpy(s->scene_host.node_roads, node_roads, sizeof(int32_t) * node_first[cfg->n_nodes]);
    s->scene_host.F = (int)floor(cfg->simulation_frequency / cfg->policy_frequency);  // abstract.py:254-256
    s->scene_host.dt = 1 / cfg->simulation_frequency;
    s->scene_host.reg_period = (int)(1 / s->scene_host.dt / 2);                        // regulation.py:30
    s->obs_size = cfg->obs_type == TTRL_OBS_GRID ? cfg->n_features * cfg->grid_w * cfg->grid_h : cfg->obs_vehicles * cfg->n_features;
    CK(cudaMalloc(&s->d_scene, sizeof(SceneDev)));
    CK(cudaMemcpy(s->d_scene, &s->scene_host, sizeof(SceneDev), cudaMemcpyHostToDevice));
    if (alloc_state(s->g, num_envs, vcap)) return 1;
    CK(cudaMalloc(&s->d_stats, sizeof(double) * kStatFields * num_envs));
    CK(cudaMemset(s->d_stats, 0, sizeof(double) * kStatFields * num_envs));
    CK(cudaMalloc(&s->d_stats_out, sizeof(double) * kStatFields));
    CK(cudaMalloc(&s->d_draws, sizeof(ttrl_spawn_draw) * num_envs));
    CK(cudaMalloc(&s->d_perm, sizeof(int32_t) * num_envs * (cfg->obs_vehicles > 1 ? cfg->obs_vehicles - 1 : 1)));
    CK(cudaMalloc(&s->d_accepted, sizeof(int32_t) * num_envs));
    CK(cudaMemset(s->d_accepted, 0, sizeof(int32_t) * num_envs));
    DISPATCH_V(s, if (configure_kernels<V>(s)) return 1);
    *out = s;
    return 0;
}

int ttrl_sim_destroy(ttrl_sim* s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    cudaFree(s->d_scene);
    free_state(s->g);
    if (s->pool.vd) free_state(s->pool);
    cudaFree(s->d_stats); cudaFree(s->d_stats_out); cudaFree(s->d_draws); cudaFree(s->d_perm); cudaFree(s->d_accepted);
    if (s->hstream) cudaStreamDestroy(s->hstream);
    cudaFreeHost(s->h_actions); cudaFreeHost(s->h_obs); cudaFreeHost(s->h_reward); cudaFreeHost(s->h_flags);
    cudaFree(s->d_actions); cudaFree(s->d_obs); cudaFree(s->d_reward); cudaFree(s->d_flags);
    delete s;
    return 0;
}

int ttrl_sim_num_envs(const ttrl_sim* s) { return s->E; }
int ttrl_sim_vcap(const ttrl_sim* s) { return s->vcap; }
int ttrl_sim_obs_size(const ttrl_sim* s) { return s->obs_size; }
int64_t ttrl_sim_launch_count(const ttrl_sim* s) { return s->launches; }

int ttrl_sim_set_spawn_routes(ttrl_sim* s, const int32_t* spawn_lane, const int32_t* route_len, const int32_t* route_road) {
    CK(cudaSetDevice(s->device));
    memcpy(s->scene_host.spawn_lane, spawn_lane, sizeof(int32_t) * 4);
    memcpy(s->scene_host.spawn_route_len, route_len, sizeof(int32_t) * 16);
    memcpy(s->scene_host.spawn_route_road, route_road, sizeof(int32_t) * 16 * TTRL_ROUTE_CAP);
    CK(cudaMemcpy(s->d_scene, &s->scene_host, sizeof(SceneDev), cudaMemcpyHostToDevice));
    return 0;
}

int ttrl_sim_set_state(ttrl_sim* s, const double* vd, const int32_t* vi, const int32_t* ei, const double* ed) {
    CK(cudaSetDevice(s->device));
    return upload_state(s->g, vd, vi, ei, ed);
}
int ttrl_sim_get_state(ttrl_sim* s, double* vd, int32_t* vi, int32_t* ei, double* ed) {
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    const GlobalState& g = s->g;
    CK(cudaMemcpy(vd, g.vd, sizeof(double) * TTRL_ND * g.E * g.V, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(vi, g.vi, sizeof(int32_t) * TTRL_NI * g.E * g.V, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ei, g.ei, sizeof(int32_t) * TTRL_NEI * g.E, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ed, g.ed, sizeof(double) * TTRL_NED * g.E, cudaMemcpyDeviceToHost));
    return 0;
}
int ttrl_sim_set_reset_pool(ttrl_sim* s, int pool_size, const double* vd, const int32_t* vi, const int32_t* ei, const double* ed) {
    CK(cudaSetDevice(s->device));
    if (pool_size <= 0) return fail("pool_size must be > 0");
    if (s->pool.vd) free_state(s->pool);
    if (alloc_state(s->pool, pool_size, s->vcap)) return 1;
    return upload_state(s->pool, vd, vi, ei, ed);
}
int ttrl_sim_set_autoreset(ttrl_sim* s, int enabled) { s->autoreset = enabled; return 0; }
int ttrl_sim_seed(ttrl_sim* s, uint64_t seed, int64_t first_global_env) { s->seed = seed; s->first_global_env = first_global_env; return 0; }

int ttrl_sim_inject_spawn(ttrl_sim* s, const ttrl_spawn_draw* draws) {
    CK(cudaSetDevice(s->device));
    s->have_draws = draws != nullptr;
    if (draws) CK(cudaMemcpy(s->d_draws, draws, sizeof(ttrl_spawn_draw) * s->E, cudaMemcpyHostToDevice));
    return 0;
}
int ttrl_sim_inject_shuffle(ttrl_sim* s, const int32_t* inv_perm) {
    CK(cudaSetDevice(s->device));
    s->have_perm = inv_perm != nullptr;
    const int m = s->scene_host.cfg.obs_vehicles - 1;
    if (inv_perm && m > 0) CK(cudaMemcpy(s->d_perm, inv_perm, sizeof(int32_t) * s->E * m, cudaMemcpyHostToDevice));
    return 0;
}
int ttrl_sim_spawn_accepted(ttrl_sim* s, int32_t* accepted_host) {
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(accepted_host, s->d_accepted, sizeof(int32_t) * s->E, cudaMemcpyDeviceToHost));
    return 0;
}

int ttrl_sim_substep(ttrl_sim* s, const int32_t* actions_dev, void* stream) {
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_V(s, (k_substep<V><<<s->E, TeamOf<V>::T, s->lay.total, st>>>(s->d_scene, s->g, actions_dev, s->lay)));
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int launch_step(ttrl_sim* s, const int32_t* actions_dev, float* obs_dev, float* reward_dev, uint8_t* term_dev, uint8_t* trunc_dev, cudaStream_t st) {
    StepIO io{};
    io.actions = actions_dev; io.obs = obs_dev; io.reward = reward_dev; io.terminated = term_dev; io.truncated = trunc_dev;
    io.draws = s->have_draws ? s->d_draws : nullptr;
    io.spawn_accepted = s->d_accepted;
    io.inv_perm = s->have_perm ? s->d_perm : nullptr;
    io.stats = s->d_stats;
    io.pool = s->pool;
    io.autoreset = s->autoreset;
    io.seed = s->seed; io.first_global_env = s->first_global_env;
    io.obs_size = s->obs_size;
    DISPATCH_V(s, (k_step<V><<<s->E, TeamOf<V>::T, s->lay.total, st>>>(s->d_scene, s->g, io, s->lay)));
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

int ttrl_sim_step(ttrl_sim* s, const int32_t* actions_dev, float* obs_dev, float* reward_dev, uint8_t* terminated_dev,
                  uint8_t* truncated_dev, void* stream) {
    CK(cudaSetDevice(s->device));
    return launch_step(s, actions_dev, obs_dev, reward_dev, terminated_dev, truncated_dev, (cudaStream_t)stream);
}

int ttrl_sim_step_host(ttrl_sim* s, const int32_t* actions_host, float* obs_host, float* reward_host, uint8_t* terminated_host,
                       uint8_t* truncated_host) {
    CK(cudaSetDevice(s->device));
    const size_t E = s->E, osz = s->obs_size;
    if (!s->hstream) {
        CK(cudaStreamCreateWithFlags(&s->hstream, cudaStreamNonBlocking));
        CK(cudaMallocHost(&s->h_actions, sizeof(int32_t) * E));
        CK(cudaMallocHost(&s->h_obs, sizeof(float) * E * osz));
        CK(cudaMallocHost(&s->h_reward, sizeof(float) * E));
        CK(cudaMallocHost(&s->h_flags, 2 * E));
        CK(cudaMalloc(&s->d_actions, sizeof(int32_t) * E));
        CK(cudaMalloc(&s->d_obs, sizeof(float) * E * osz));
        CK(cudaMalloc(&s->d_reward, sizeof(float) * E));
        CK(cudaMalloc(&s->d_flags, 2 * E));
    }
    if (actions_host) {
        memcpy(s->h_actions, actions_host, sizeof(int32_t) * E);
        CK(cudaMemcpyAsync(s->d_actions, s->h_actions, sizeof(int32_t) * E, cudaMemcpyHostToDevice, s->hstream));
    }
    if (launch_step(s, actions_host ? s->d_actions : nullptr, s->d_obs, s->d_reward, s->d_flags, s->d_flags + E, s->hstream)) return 1;
    CK(cudaMemcpyAsync(s->h_obs, s->d_obs, sizeof(float) * E * osz, cudaMemcpyDeviceToHost, s->hstream));
    CK(cudaMemcpyAsync(s->h_reward, s->d_reward, sizeof(float) * E, cudaMemcpyDeviceToHost, s->hstream));
    CK(cudaMemcpyAsync(s->h_flags, s->d_flags, 2 * E, cudaMemcpyDeviceToHost, s->hstream));
    CK(cudaStreamSynchronize(s->hstream));
    if (obs_host) memcpy(obs_host, s->h_obs, sizeof(float) * E * osz);
    if (reward_host) memcpy(reward_host, s->h_reward, sizeof(float) * E);
    if (terminated_host) memcpy(terminated_host, s->h_flags, E);
    if (truncated_host) memcpy(truncated_host, s->h_flags + E, E);
    return 0;
}

int ttrl_sim_observe(ttrl_sim* s, float* obs_dev, void* stream) {
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t* perm = s->have_perm ? s->d_perm : nullptr;
    DISPATCH_V(s, (k_observe<V><<<s->E, TeamOf<V>::T, s->lay.total, st>>>(s->d_scene, s->g, obs_dev, s->obs_size, perm, s->lay)));
    s->launches++;
    CK(cudaGetLastError());
    return 0;
}

/* Host-driven reset primitive: one _spawn_vehicle attempt with explicit arguments (intersection_env.py:265-283). */
int ttrl_sim_spawn(ttrl_sim* s, const ttrl_spawn_draw* draws_host, double longitudinal, double position_deviation,
                   double speed_deviation, double spawn_probability, int go_straight, int32_t* accepted_host) {
    CK(cudaSetDevice(s->device));
    CK(cudaMemcpy(s->d_draws, draws_host, sizeof(ttrl_spawn_draw) * s->E, cudaMemcpyHostToDevice));
    SpawnParams sp{longitudinal, position_deviation, speed_deviation, spawn_probability, go_straight};
    DISPATCH_V(s, (k_spawn<V><<<s->E, TeamOf<V>::T, s->lay.total, 0>>>(s->d_scene, s->g, s->d_draws, sp, s->d_accepted, s->lay)));
    s->launches++;
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (accepted_host) CK(cudaMemcpy(accepted_host, s->d_accepted, sizeof(int32_t) * s->E, cudaMemcpyDeviceToHost));
    return 0;
}

int ttrl_sim_read_stats(ttrl_sim* s, ttrl_episode_stats* out, int reset_after_read) {
    CK(cudaSetDevice(s->device));
    CK(cudaDeviceSynchronize());
    k_reduce_stats<<<kStatFields, 256>>>(s->d_stats, s->E, s->d_stats_out);
    s->launches++;
    CK(cudaGetLastError());
    double h[kStatFields];
    CK(cudaMemcpy(h, s->d_stats_out, sizeof(h), cudaMemcpyDeviceToHost));
    out->episodes = h[0]; out->total_return = h[1]; out->total_length = h[2]; out->crashes = h[3];
    out->arrivals = h[4]; out->total_speed = h[5]; out->vehicle_steps = h[6]; out->env_steps = h[7];
    if (reset_after_read) CK(cudaMemset(s->d_stats, 0, sizeof(double) * kStatFields * s->E));
    return 0;
}

}  // extern "C"
