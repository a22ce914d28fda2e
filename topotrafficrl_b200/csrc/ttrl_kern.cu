// ttrl_kern.cu -- the sm_100a kernels of the batched simulator for ONE slot capacity V (-DTT_V=<V>).
//
// Launch shape: one CTA ("team" of T threads) per env instance; slots and task lists are strided over the team.
// k_step keeps the env's state in shared memory across the F sub-steps of an env-step: one coalesced HBM read
// and one write of the state per env-step.  Shared-memory arrays are sized by V, so the
// library carries one instantiation per common capacity (Makefile: TT_VS) and picks the smallest V >= vcap.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "ttrl_kernels.cuh"

using namespace ttrl;

// ------------------------------------------------------------------------------------------------
// device execution policy: a "team" of T threads per env; slots / tasks are strided over the team.
//  * T > 32: one team per CTA, barriers are __syncthreads().
//  * T == 32 (up to 32 vehicle slots): the team is one warp (barriers are __syncwarp()) and a CTA holds G teams.
//    `align()` is a CTA-wide barrier at the phase boundaries of a sub-step: it keeps the warps of the CTA in the
//    same code region, so an instruction-cache line fetched by one warp serves all of them.  The step kernel is
//    bound by instruction-cache miss traffic (the GPC-level cache serves instruction requests at ~90 % of its peak
//    rate and kernel time tracks the request count, profiles/r1f_icache.txt); for the intersection scene (few
//    vehicles per env, long scalar phases) aligned multi-env CTAs are 23 % faster, for 2-warp and larger teams the
//    barrier skew costs more than the fetch sharing saves (profiles/r1c_variants.txt).
// ------------------------------------------------------------------------------------------------
// CTA-wide alignment points per sub-step of the multi-env CTAs (bit k = point k of Exec::align_at).  Measured on B200
// (profiles/r2_alignment.txt): start only 1.88 ms/step (intersection, 8192 envs), start + before the integration 1.67 - 1.78, start + before the controls + before the integration 1.68 - 1.70, all four 1.76.
// Two-warp highway teams (10 per CTA, named team barriers): the start-of-sub-step point only -- 1.048 ms/step against 1.14 with
// none (0x0) or two (0x5), 1.16 with three, and 1.175 for one env per CTA (profiles/r2_step_scheduling.txt item 12).
#ifndef TT_ALIGN_MASK
#define TT_ALIGN_MASK (TT_V <= 32 ? 0x7 : 0x1)
#endif
// MULTI: the CTA holds several teams (compile-time, so that single-env CTAs carry none of the multi-team paths)
template <int V, int T, bool MULTI>
struct DevExec {
    int tid;        // thread within the team
    int G;          // teams in this CTA (1 unless T <= 32)
    unsigned mask;  // T <= 32: the team's lanes of its warp (a sub-warp team shares the warp with 32 / T - 1 other envs)
    int lane;       // lane within the warp
    int team;       // team within the CTA (teams of several warps in a multi-env CTA meet at named barrier 1 + team)
    __device__ __forceinline__ DevExec(int tid_, int G_) : tid(tid_), G(MULTI ? G_ : 1) {
        team = (int)threadIdx.x / T;
        lane = (int)(threadIdx.x & 31u);
        mask = T >= 32 ? 0xffffffffu : (((1u << (T & 31)) - 1u) << (lane & ~(T - 1)));
    }
    __device__ __forceinline__ bool first() const { return tid == 0; }
    __device__ __forceinline__ void sync() {
        if (T <= 32) __syncwarp(mask);
        else if (MULTI) asm volatile("bar.sync %0, %1;" :: "r"(1 + team), "r"(T) : "memory");  // this team's warps only
        else __syncthreads();
    }
    // CTA-wide phase alignment; every team of the CTA calls it the same number of times
    __device__ __forceinline__ void align() {
#if defined(TT_NO_ALIGN)
        return;  // experiment: free-running teams
#endif
        if (MULTI && G > 1) __syncthreads();
    }
    // alignment point k of a sub-step (0 its start, 1 before the controls, 2 before the integration, 3 before the collisions)
    template <int K> __device__ __forceinline__ void align_at() {
        if ((TT_ALIGN_MASK >> K) & 1) align();
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
        if (T <= 32) return __ballot_sync(mask, p) != 0;
        if (MULTI) {  // OR-reduction over this team's warps at its named barrier
            int r;
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %3, 0;\n\tbar.red.or.pred p, %1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(r) : "r"(1 + team), "r"(T), "r"(p) : "memory");
            return r != 0;
        }
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
        static_assert(T % 8 == 0, "a pair's 8 axis lanes must lie inside one team");
        const unsigned wmask = T <= 32 ? mask : 0xffffffffu;
        const int total = np * 8;
#pragma unroll 1
        for (int base = 0; base < total; base += T) {
            const int idx = base + tid;
            const bool act = idx < total;
            const int q = idx >> 3, axis = idx & 7;
            AxisRes r;
            r.absd = INFINITY; r.nx = 0; r.ny = 0; r.fl = 0;
            if (act) r = fa(q, axis);
            const unsigned gbase = (unsigned)lane & ~7u;
            const unsigned sep_now = __ballot_sync(wmask, act && (r.fl & 1));
            const unsigned sep_after = __ballot_sync(wmask, act && (r.fl & 2));
            const bool inter = ((sep_now >> gbase) & 0xFFu) == 0, will = ((sep_after >> gbase) & 0xFFu) == 0;
            double bd = r.absd;
            int bi = axis;
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) {
                const double od = __shfl_xor_sync(wmask, bd, off);
                const int oi = __shfl_xor_sync(wmask, bi, off);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            const double nx = __shfl_sync(wmask, r.nx, (int)gbase + bi), ny = __shfl_sync(wmask, r.ny, (int)gbase + bi);
            if (act && axis == 0) fp(q, inter, will, bd, nx, ny);
        }
        sync();
    }
    __device__ __forceinline__ void atomic_min(int32_t* a, int32_t v) { atomicMin(a, v); }
    __device__ __forceinline__ void atomic_max(int32_t* a, int32_t v) { atomicMax(a, v); }
    __device__ __forceinline__ void atomic_or(uint32_t* a, uint32_t v) { atomicOr(a, v); }
    __device__ __forceinline__ int atomic_add(int32_t* a, int32_t v) { return atomicAdd(a, v); }
    __device__ __forceinline__ int atomic_add_global(int32_t* a, int32_t v) { return atomicAdd(a, v); }
    // acquire load (gpu scope): pairs with the release in k_regen_list
    __device__ __forceinline__ int load_acquire(const int32_t* a) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
        return v;
    }
};

// threads per env for a slot capacity
// Threads per env ("team") and minimum resident CTAs per SM (register cap) for a slot capacity.  Tuned on B200
// (profiles/r1c_variants.txt): the kernel is latency-bound (fp64 dependency chains, instruction fetch), so
// resident warps matter more than registers per thread.
#ifndef TT_V
#error "compile with -DTT_V=<slot capacity>"
#endif
#ifndef TT_T
#define TT_T (TT_V <= 16 ? 16 : TT_V <= 32 ? 32 : TT_V <= 64 ? 64 : TT_V <= 128 ? 128 : 256)
#endif
#ifndef TT_MINB
#define TT_MINB (TT_V <= 32 ? 1 : TT_V <= 64 ? 10 : TT_V <= 128 ? 5 : 3)
#endif
// envs per CTA of k_step (1 = one env per CTA, no phase alignment)
#ifndef TT_G
/* upper bound; the launch uses as many as fit in shared memory.  V <= 32: 512 threads.  V <= 64 (two-warp teams): 10 envs =
 * 640 threads at ~100 registers, the same 20 warps per SM as one env per CTA had, but fetching in step (G = 8: 1.22, 10: 1.05, 12: 1.11 ms) */
#define TT_G (TT_V <= 32 ? 512 / TT_T : TT_V <= 64 ? 10 : 1)
#endif
template <int V> struct TeamOf {
    static constexpr int T = TT_T, MINB = TT_MINB, G = TT_G;
    static_assert(G == 1 || T <= 32 || G <= 15, "multi-warp teams of a multi-env CTA use named barriers 1..15");
};

template <int V, int P>
__device__ __forceinline__ void make_ctx(EnvCtx<V, P>& c, unsigned char* smem_cta, int team, const SceneDev* sc, const SmemLayout& lay, int vcap) {
    // the lane table (n_lanes * 160 B) is shared by the teams of the CTA: copied once with 16-byte vector loads
    ttrl_lane* lanes_s = reinterpret_cast<ttrl_lane*>(smem_cta);
    {
        const int4* src = reinterpret_cast<const int4*>(sc->lanes);
        int4* dst = reinterpret_cast<int4*>(lanes_s);
        const int n16 = sc->cfg.n_lanes * (int)(sizeof(ttrl_lane) / 16);
        for (int k = threadIdx.x; k < n16; k += blockDim.x) dst[k] = __ldg(src + k);
    }
    unsigned char* smem = smem_cta + lay.lanes_bytes + (size_t)team * lay.per_env;
    c.st = reinterpret_cast<EnvState<V>*>(smem);
    c.sc = sc;
    c.L = sc->cfg.n_lanes;
    c.vcap = vcap < V ? vcap : V;  // a buffer with more slots than this kernel's capacity: the small size class
    c.lanes = lanes_s;
    c.SR = reinterpret_cast<d2*>(smem + lay.off_SR);
    c.NC = P == 1 ? sc->cfg.n_lanes : sc->n_curved;
    static_assert(P >= 0 && P <= 4, "scene profile");
    c.lmask = reinterpret_cast<uint32_t*>(smem + lay.off_lmask);
    c.pred = lay.off_pred >= 0 ? reinterpret_cast<double*>(smem + lay.off_pred) : nullptr;
    c.pbits = lay.off_pred >= 0 ? reinterpret_cast<uint32_t*>(smem + lay.off_pred + sizeof(double) * 4 * V) : nullptr;
    c.obs_s = reinterpret_cast<float*>(smem + lay.off_obs);
    c.perm_s = reinterpret_cast<uint32_t*>(smem + lay.off_perm);
    c.lin = lay.off_lin >= 0 ? reinterpret_cast<double*>(smem + lay.off_lin) : nullptr;
    c.cell = reinterpret_cast<int32_t*>(smem + lay.off_cell);
    c.gap_den = 2 * sqrt(-sc->cfg.comfort_acc_max * sc->cfg.comfort_acc_min);  // behavior.py:214-216
    c.tan_max_steer = tan(kPi / 3);
    __syncthreads();
}

// single-team CTA (k_substep, k_observe, k_spawn, k_reset): the general profile (any number of controlled vehicles)
#define TT_KERNEL_PROLOGUE                                     \
    extern __shared__ __align__(16) unsigned char smem[];      \
    constexpr int T = TeamOf<V>::T;                            \
    EnvCtx<V, 4> c;                                            \
    make_ctx<V, 4>(c, smem, 0, sc, lay, g.V);                  \
    DevExec<V, T, false> ex((int)threadIdx.x, 1);

template <int V, int P>
__global__ void __launch_bounds__(TeamOf<V>::T * TeamOf<V>::G, TeamOf<V>::G > 1 ? 1 : TeamOf<V>::MINB)
k_step(const SceneDev* __restrict__ sc, GlobalState g, StepIO io, SmemLayout lay) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int T = TeamOf<V>::T;
    const int G = (int)blockDim.x / T;
    const int team = threadIdx.x / T;
    EnvCtx<V, P> c;
    DevExec<V, T, (TeamOf<V>::G > 1)> ex((int)threadIdx.x % T, G);
    int e = io.env_first + (int)blockIdx.x * G + team;
    if (io.env_count > 0 && e >= io.env_first + io.env_count) e = g.E;  // past the slice of this launch
    if (io.cls_list) {  // binned launch: this CTA takes G envs of one bin
        int cta = (int)blockIdx.x, b = io.cls_first, cnt = 0;
        for (; b < io.cls_first + io.cls_n; ++b) {
            cnt = io.cls_count[b];
            const int nct = (cnt + G - 1) / G;
            if (cta < nct) break;
            cta -= nct;
        }
        if (b == io.cls_first + io.cls_n) return;  // uniform per CTA (before any barrier)
        const int idx = cta * G + team;
        e = idx < cnt ? io.cls_list[(size_t)b * g.E + idx] : g.E;
    }
    make_ctx<V, P>(c, smem, team, sc, lay, g.V);
    // The teams of the CTA run their sub-steps in lockstep: one alignment barrier per sub-step keeps the warps in the same
    // code region (shared instruction fetch).  The barrier is ONE instruction outside the per-team branches, reached by teams
    // with and without an env alike (a team without an env: last CTA of a bin).
    const bool active = e < g.E;
    if (TeamOf<V>::G == 1) {  // one env per CTA: nothing to align
        if (active) env_step(c, ex, g, io, e);
        return;
    }
    if (active) env_step_load(c, ex, g, io, e);
    const int32_t* actions = (active && io.actions) ? io.actions + (size_t)e * n_agents(c) : nullptr;
    double veh_steps = 0;
    const int F = sc->F;
#pragma unroll 1
    for (int f = 0; f < F; ++f) {
        env_substep_lockstep(c, ex, active, actions);
        if (active) veh_steps += c.st->n;
    }
    if (active) env_step_finish(c, ex, g, io, e, veh_steps);
}

template <int V>
__global__ void __launch_bounds__(TeamOf<V>::T) k_substep(const SceneDev* __restrict__ sc, GlobalState g, const int32_t* __restrict__ actions, SmemLayout lay) {
    TT_KERNEL_PROLOGUE
    const int e = blockIdx.x;
    load_env(c, ex, g, e);
    env_substep(c, ex, actions ? actions + (size_t)e * n_agents(c) : nullptr);
    store_env(c, ex, g, e);
}

template <int V>
__global__ void __launch_bounds__(TeamOf<V>::T) k_observe(const SceneDev* __restrict__ sc, GlobalState g, float* __restrict__ obs, int obs_size,
                                                          const int32_t* __restrict__ inv_perm, uint64_t seed, int64_t first_global_env,
                                                          double* __restrict__ info, SmemLayout lay) {
    TT_KERNEL_PROLOGUE
    const int e = blockIdx.x;
    load_env(c, ex, g, e);
    observe(c, ex, obs + (size_t)e * obs_size, inv_perm ? inv_perm + (size_t)e * n_agents(c) * (sc->cfg.obs_vehicles - 1) : nullptr,
            seed, first_global_env + e);
    if (info && ex.first()) write_info(c, info, g.E, e, nullptr);
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


// AbstractEnv.reset for every (masked) env: a fresh episode generated on the device (ttrl_core.cuh: env_reset)
template <int V>
__global__ void __launch_bounds__(TeamOf<V>::T) k_reset(const SceneDev* __restrict__ sc, GlobalState g, const uint8_t* __restrict__ mask, uint64_t seed,
                                                        int64_t first_global_env, SmemLayout lay) {
    TT_KERNEL_PROLOGUE
    const int e = blockIdx.x;
    if (mask && !mask[e]) return;  // uniform per CTA
    const int episode = mask ? g.ei[TTRL_EI_EPISODE * g.E + e] + 1 : 0;
    env_reset(c, ex, seed, first_global_env + e, episode);
    store_env(c, ex, g, e);
}

// Device autoreset of scenes whose reset has warm-up sub-steps (intersection): the envs k_step found finished are
// listed in `done_list`; they are reset here PACKED (G per CTA, like k_step) and phase-aligned, and their first
// observation replaces the terminal one (gymnasium autoreset semantics).  A reset inside k_step would leave one warp
// running 45 sub-steps while the other teams of its CTA -- and the SM -- wait.
template <int V, int P>
__global__ void __launch_bounds__(TeamOf<V>::T * TeamOf<V>::G, TeamOf<V>::G > 1 ? 1 : TeamOf<V>::MINB)
k_reset_list(const SceneDev* __restrict__ sc, GlobalState g, StepIO io, SmemLayout lay) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int T = TeamOf<V>::T;
    const int G = (int)blockDim.x / T;
    const int team = threadIdx.x / T;
    const int n_done = *io.done_count;
    const int nb = (int)gridDim.x, b = (int)blockIdx.x;
    if (b >= n_done) return;  // uniform per CTA
    EnvCtx<V, P> c;
    make_ctx<V, P>(c, smem, team, sc, lay, g.V);
    DevExec<V, T, (TeamOf<V>::G > 1)> ex((int)threadIdx.x % T, G);
    // The finished envs are dealt ROUND-ROBIN over the CTAs (one CTA per SM), team by team: a typical step finishes ~1/13 of
    // the envs, and a few teams on every SM run faster than 14 teams on a third of the SMs (a team's speed is set by
    // dependency latency and by its share of the SM's instruction fetch).  With more finished envs than teams the CTA
    // makes several passes; a team without an env in a pass only takes part in the phase alignment.
    const int passes = (n_done + G * nb - 1) / (G * nb);
    for (int pass = 0; pass < passes; ++pass) {
        const int k = (pass * G + team) * nb + b;
        const bool active = k < n_done;
        const int e = active ? io.done_list[k] : 0;
        const int episode = active ? g.ei[TTRL_EI_EPISODE * g.E + e] + 1 : 0;
        env_reset_lockstep(c, ex, active, io.seed, io.first_global_env + e, episode);
        if (active) {
            if (io.obs) observe(c, ex, io.obs + (size_t)e * io.obs_size,
                                io.inv_perm ? io.inv_perm + (size_t)e * n_agents(c) * (sc->cfg.obs_vehicles - 1) : nullptr,
                                io.seed, io.first_global_env + e);
            store_env(c, ex, g, e);
        }
    }
}

// Asynchronous device reset: generate the episodes queued in io.regen_list = (env, episode) pairs into the SHADOW buffers,
// on a side stream, while the next env-steps run.  Same dealing of the work as k_reset_list.  The record is published with
// a release store of its episode number after every thread's writes have been fenced.
template <int V, int P>
__global__ void __launch_bounds__(TeamOf<V>::T * TeamOf<V>::G, TeamOf<V>::G > 1 ? 1 : TeamOf<V>::MINB)
k_regen_list(const SceneDev* __restrict__ sc, StepIO io, SmemLayout lay, int part, int parts) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int T = TeamOf<V>::T;
    const int G = (int)blockDim.x / T;
    const int team = threadIdx.x / T;
    const int n_regen = *io.regen_count;
    const int nb = (int)gridDim.x, b = (int)blockIdx.x;
    if (b >= n_regen) return;  // uniform per CTA
    EnvCtx<V, P> c;
    make_ctx<V, P>(c, smem, team, sc, lay, io.shadow.V);
    DevExec<V, T, (TeamOf<V>::G > 1)> ex((int)threadIdx.x % T, G);
    const int passes = (n_regen + G * nb - 1) / (G * nb);
    for (int pass = 0; pass < passes; ++pass) {
        const int k = (pass * G + team) * nb + b;
        const bool active = k < n_regen;
        const int e = active ? io.regen_list[2 * k] : 0, episode = active ? io.regen_list[2 * k + 1] : 0;
        const int D = io.shadow_depth > 0 ? io.shadow_depth : 1;
        const int rec = (episode % D) * (io.shadow.E / D) + e;  // the ring record of this episode
        env_reset_lockstep(c, ex, active, io.seed, io.first_global_env + e, episode, part, parts, &io.shadow, rec);
        if (active) {
            store_env(c, ex, io.shadow, rec);  // (an unfinished reset parks its state in the record: ready stays -1)
            if (part == parts - 1) {
                __threadfence();
                ex.sync();
                if (ex.first()) asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(io.shadow_ready + rec), "r"(episode) : "memory");
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launch table
// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// the LinearVehicle profile is instantiated for the small capacities only (its scenes: roundabout, intersection)
template <int V> constexpr bool kHasLinearProfile = V <= 32;

template <int V>
static int configure(const ttrl_config& cfg, const ttrl_lane* lanes, int vcap, SmemLayout* out) {
    SmemLayout l{};
    l.lanes_bytes = (int)align_up(sizeof(ttrl_lane) * cfg.n_lanes, 128);
    size_t off = align_up(sizeof(EnvState<V>), 16);
    int n_curved = 0;
    for (int k = 0; k < cfg.n_lanes; ++k) n_curved += lanes[k].kind != TTRL_LANE_STRAIGHT ? 1 : 0;
    // coordinate cache: the curved lanes; scenes without any (they run the plain profile or never read it): every lane
    l.off_SR = (int)off; off += sizeof(d2) * (size_t)vcap * (n_curved > 0 ? n_curved : cfg.n_lanes);
    l.off_lmask = (int)off; off += align_up(sizeof(uint32_t) * ((V + 31) / 32) * cfg.n_lanes, 16);
    if (cfg.regulated) { l.off_pred = (int)off; off += sizeof(double) * 4 * V + align_up(sizeof(uint32_t) * ((V * (V - 1) / 2 + 31) / 32), 16); } else l.off_pred = -1;
    l.off_obs = (int)off; off += align_up(sizeof(float) * (cfg.obs_type == TTRL_OBS_KINEMATICS ? cfg.obs_vehicles * cfg.n_features : 4), 16);
    l.off_perm = (int)off; off += align_up(sizeof(uint32_t) * 2 * (cfg.obs_type == TTRL_OBS_KINEMATICS && cfg.order == TTRL_ORDER_SHUFFLED && cfg.obs_vehicles > 1 ? cfg.obs_vehicles - 1 : 0), 16);
    if (cfg.vehicle_model == TTRL_VEHICLE_LINEAR) { l.off_lin = (int)off; off += sizeof(double) * TTRL_NLIN * V; } else l.off_lin = -1;
    l.off_cell = (int)off; off += align_up(sizeof(int32_t) * (cfg.obs_type == TTRL_OBS_GRID ? cfg.grid_w * cfg.grid_h :
                                                                cfg.obs_type == TTRL_OBS_TTC ? TTRL_MAX_TTC_CELLS : 4), 16);
    l.per_env = (int)align_up(off, 128);
    l.total = l.lanes_bytes + l.per_env;                       // single-team kernels
    int G = TeamOf<V>::G;                                      // k_step: G envs per CTA, as many as fit
    if (const char* v = getenv("TTRL_G")) { const int cap = atoi(v); if (cap >= 1 && cap < G) G = cap; }  // tuning experiments
    while (G > 1 && l.lanes_bytes + G * l.per_env > 227 * 1024) --G;
    if (TeamOf<V>::T < 32 && G > 1) G -= G % (32 / TeamOf<V>::T);  // whole warps
    if (G < 1) G = 1;
    l.G = G;
    l.total_step = l.lanes_bytes + G * l.per_env;
    *out = l;
    if (l.total_step > 227 * 1024) return (int)cudaErrorInvalidValue;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_step<V, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_step<V, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_step<V, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
    if constexpr (kHasLinearProfile<V>) {
        if ((e = cudaFuncSetAttribute(k_step<V, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
        if ((e = cudaFuncSetAttribute(k_reset_list<V, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
        if ((e = cudaFuncSetAttribute(k_regen_list<V, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
        cudaFuncSetAttribute(k_step<V, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
    if ((e = cudaFuncSetAttribute(k_substep<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_observe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_spawn<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_reset<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_reset_list<V, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_reset_list<V, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_regen_list<V, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_regen_list<V, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, l.total_step)) != cudaSuccess) return (int)e;
    // all of the SM's unified L1/shared storage as shared memory: resident CTAs are what hides latency here
    cudaFuncSetAttribute(k_step<V, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_step<V, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_step<V, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    // "plain" scene profile (EnvCtx::kPlain): straight lanes only, no regulation, no spawn / clear, highway reward
    bool plain = !cfg.regulated && !cfg.spawn_enabled && cfg.reward_type == TTRL_REWARD_HIGHWAY && cfg.controlled_vehicles <= 1 &&
                 cfg.obs_type != TTRL_OBS_TTC;
    for (int k = 0; k < cfg.n_lanes; ++k) plain = plain && lanes[k].kind == TTRL_LANE_STRAIGHT;
    out->plain = plain ? 1 : (cfg.controlled_vehicles > 1 ? 2 : 0);  // scene profile: 0 general single-agent, 1 plain, 2 multi-agent
    if (cfg.vehicle_model == TTRL_VEHICLE_LINEAR) {  // 3: general single-agent with LinearVehicle traffic (capacities <= 32)
        if (cfg.controlled_vehicles > 1 || !kHasLinearProfile<V>) return (int)cudaErrorNotSupported;
        out->plain = 3;
    }
    int dev = 0, n_sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
    out->n_sms = n_sms > 0 ? n_sms : 148;
    return 0;
}
template <int V>
static void launch_step(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const StepIO& io) {
    const int G = lay.G;
    // binned launch: every bin may leave one partly filled CTA; CTAs beyond the bins' envs exit at once
    const int n_envs = (!io.cls_list && io.env_count > 0) ? io.env_count : E;
    const int nb = (n_envs + G - 1) / G + (io.cls_list ? io.cls_n : 0);
    if (lay.plain == 3) { if constexpr (kHasLinearProfile<V>) k_step<V, 3><<<nb, TeamOf<V>::T * G, lay.total_step, st>>>(sc, g, io, lay); }
    else if (lay.plain == 1) k_step<V, 1><<<nb, TeamOf<V>::T * G, lay.total_step, st>>>(sc, g, io, lay);
    else if (lay.plain == 2) k_step<V, 2><<<nb, TeamOf<V>::T * G, lay.total_step, st>>>(sc, g, io, lay);
    else k_step<V, 0><<<nb, TeamOf<V>::T * G, lay.total_step, st>>>(sc, g, io, lay);
}
// the packed reset of the envs the step found finished (io.done_list)
template <int V>
static void launch_reset_list(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const StepIO& io) {
    const int G = lay.G;
    // one CTA per SM (fewer for tiny batches); CTAs beyond the number of finished envs exit at once
    const int nb_reset = (E + G - 1) / G < lay.n_sms ? (E + G - 1) / G : lay.n_sms;
    if (io.done_list && lay.plain == 3) { if constexpr (kHasLinearProfile<V>) k_reset_list<V, 3><<<nb_reset, TeamOf<V>::T * G, lay.total_step, st>>>(sc, g, io, lay); }
    else if (io.done_list && lay.plain == 2) k_reset_list<V, 2><<<nb_reset, TeamOf<V>::T * G, lay.total_step, st>>>(sc, g, io, lay);
    else if (io.done_list) k_reset_list<V, 0><<<nb_reset, TeamOf<V>::T * G, lay.total_step, st>>>(sc, g, io, lay);
}
// the regeneration of the episodes queued by the last step, on the side stream
template <int V>
static void launch_regen(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const StepIO& io, int parts) {
    const int G = lay.G;
    // One CTA per SM, default stream priority: the regeneration CTAs (4-5 busy teams each) slot into the ragged tail of the
    // step kernel's last wave and the next step's CTAs follow as they drain.  Measured alternatives (profiles/r1k_async_reset.txt):
    // a third of the SMs at the lowest priority starves the regeneration (5.02 vs 3.45 ms/step).
    static const int div = [] { const char* v = getenv("TTRL_REGEN_SM_DIV"); const int d = v ? atoi(v) : 1; return d > 0 ? d : 1; }();  // tuning experiments
    const int cap = lay.n_sms / div > 0 ? lay.n_sms / div : 1;
    const int nb = (E + G - 1) / G < cap ? (E + G - 1) / G : cap;
    for (int part = 0; part < parts; ++part) {  // consecutive launches on the same stream: part k + 1 continues from the records of part k
        if (lay.plain == 3) { if constexpr (kHasLinearProfile<V>) k_regen_list<V, 3><<<nb, TeamOf<V>::T * G, lay.total_step, st>>>(sc, io, lay, part, parts); }
        else if (lay.plain == 2) k_regen_list<V, 2><<<nb, TeamOf<V>::T * G, lay.total_step, st>>>(sc, io, lay, part, parts);
        else k_regen_list<V, 0><<<nb, TeamOf<V>::T * G, lay.total_step, st>>>(sc, io, lay, part, parts);
    }
}
template <int V>
static void launch_substep(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const int32_t* actions) {
    k_substep<V><<<E, TeamOf<V>::T, lay.total, st>>>(sc, g, actions, lay);
}
template <int V>
static void launch_observe(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, float* obs, int obs_size,
                           const int32_t* inv_perm, uint64_t seed, int64_t first_global_env, double* info) {
    k_observe<V><<<E, TeamOf<V>::T, lay.total, st>>>(sc, g, obs, obs_size, inv_perm, seed, first_global_env, info, lay);
}
template <int V>
static void launch_spawn(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const ttrl_spawn_draw* draws,
                         SpawnParams sp, int32_t* accepted) {
    k_spawn<V><<<E, TeamOf<V>::T, lay.total, st>>>(sc, g, draws, sp, accepted, lay);
}

template <int V>
static void launch_reset(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const uint8_t* mask, uint64_t seed,
                         int64_t first_global_env) {
    k_reset<V><<<E, TeamOf<V>::T, lay.total, st>>>(sc, g, mask, seed, first_global_env, lay);
}

#define TT_CAT_(a, b) a##b
#define TT_CAT(a, b) TT_CAT_(a, b)
extern "C" const KernelSet* TT_CAT(ttrl_kernel_set_, TT_V)(void) {
    static const KernelSet ks = {TT_V, TeamOf<TT_V>::T, configure<TT_V>, launch_step<TT_V>, launch_reset_list<TT_V>, launch_substep<TT_V>, launch_observe<TT_V>, launch_spawn<TT_V>, launch_reset<TT_V>,
                                 launch_regen<TT_V>};
    return &ks;
}
