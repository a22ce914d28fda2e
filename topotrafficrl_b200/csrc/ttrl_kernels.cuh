// ttrl_kernels.cuh -- interface between the host side of the C ABI (ttrl_sim.cu) and the per-capacity kernel
// translation units (ttrl_kern.cu compiled once per slot capacity V with -DTT_V=<V>).
#pragma once
#include <cuda_runtime.h>
#include "ttrl_core.cuh"

namespace ttrl {

struct SmemLayout {
    // offsets inside one env's region; a CTA's dynamic shared memory = lane table (lanes_bytes) + G env regions (per_env)
    int off_SR, off_lmask, off_pred, off_obs, off_perm, off_lin, off_cell, lanes_bytes, per_env, total, total_step, G, plain, n_sms;
};

// One set of launchers per compiled slot capacity.
struct KernelSet {
    int V;        // slot capacity the kernels were compiled for (arrays in shared memory are sized by it)
    int T;        // threads per env
    int (*configure)(const ttrl_config& cfg, const ttrl_lane* lanes, int vcap, SmemLayout* lay);  // layout + shared-memory opt-in; cudaError_t as int
    void (*step)(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const StepIO& io);
    void (*reset_list)(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const StepIO& io);
    void (*substep)(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const int32_t* actions);
    void (*observe)(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, float* obs, int obs_size,
                    const int32_t* inv_perm, uint64_t seed, int64_t first_global_env, double* info);
    void (*spawn)(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const ttrl_spawn_draw* draws,
                  SpawnParams sp, int32_t* accepted);
    void (*reset)(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const GlobalState& g, const uint8_t* mask, uint64_t seed,
                  int64_t first_global_env);
    void (*regen)(int E, const SmemLayout& lay, cudaStream_t st, const SceneDev* sc, const StepIO& io, int parts);
};

}  // namespace ttrl

// capacities compiled into the library (Makefile: TT_VS); ttrl_kernel_set_<V>() is defined by ttrl_kern.cu -DTT_V=<V>
#define TT_DECLARE_KERNEL_SET(N_) extern "C" const ttrl::KernelSet* ttrl_kernel_set_##N_(void);
