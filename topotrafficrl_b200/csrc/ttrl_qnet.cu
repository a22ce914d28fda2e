// ttrl_qnet.cu -- fused DQN Q-network rollout: obs -> Q-values -> epsilon-greedy action, one launch.
//
// Replaces AbstractDQNAgent.act (ttrl_agent/agents/deep_q_network/abstract.py:65-83) ->
// DQNAgent.get_batch_state_action_values (pytorch.py:79-80) -> model forward (agents/common/models.py)
// -> EpsilonGreedy.update / DiscreteDistribution.sample (exploration/epsilon_greedy.py:32-48,
// exploration/abstract.py:20-25) for a whole batch of observations that already live in HBM (the step
// kernel's output buffer): no host round trip per decision, batch = E instead of 1.
//
// This file holds the fp32 SIMT path (bit-for-bit fp32 semantics of the torch modules up to summation
// order; it is the parity path).  Persistent CTAs keep ALL weights in shared memory (<= ~137 KB for the
// shipped configs); each warp forwards one env (ego-attention) or a group of 8 envs (MLP / dueling).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ttrl_b200.h"

extern "C" void ttrl_set_error(const char* msg);  // ttrl_sim.cu

namespace {

int qfail(const std::string& m) { ttrl_set_error(m.c_str()); return 1; }

constexpr int kWarps = 6;
constexpr int kMaxRows = 16;

struct DenseDesc { int K, N, w_off, b_off, relu; };  // Wt[K][N] at w_off, bias at b_off (-1: none)

struct QnetDev {
    ttrl_qnet_desc d;
    int n_weights;
    // layer tables (offsets into the weight blob)
    int n_ego, n_oth, n_out, n_base, n_val, n_adv;
    DenseDesc ego[5], oth[5], out[5], base[5], val[5], adv[5];
    DenseDesc wk, wv, wq, wc;
    int width;        // widest activation
    int scratch_per_warp;  // floats
};

__device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
__device__ void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// out[r][c] = act(b[c] + sum_k in[r][k] * Wt[k][c]) for r < rows; lane owns columns lane + 32 j.
// Weights are read conflict-free (consecutive lanes -> consecutive c), inputs are warp broadcasts.
template <int MAXR, int NCOL>
__device__ __forceinline__ void warp_dense(const float* __restrict__ w, const DenseDesc& L, const float* in, int ldin, int rows,
                                           float* out, int ldout, int lane) {
    float acc[MAXR][NCOL];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        const int c = lane + 32 * j;
        const float b = (L.b_off >= 0 && c < L.N) ? w[L.b_off + c] : 0.f;
#pragma unroll
        for (int r = 0; r < MAXR; ++r) acc[r][j] = b;
    }
    const float* Wt = w + L.w_off;
    for (int k = 0; k < L.K; ++k) {
        float wv[NCOL];
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const int c = lane + 32 * j;
            wv[j] = c < L.N ? Wt[k * L.N + c] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
            if (r < rows) {
                const float x = in[r * ldin + k];
#pragma unroll
                for (int j = 0; j < NCOL; ++j) acc[r][j] = fmaf(x, wv[j], acc[r][j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        const int c = lane + 32 * j;
        if (c < L.N) {
#pragma unroll
            for (int r = 0; r < MAXR; ++r)
                if (r < rows) out[r * ldout + c] = L.relu ? fmaxf(acc[r][j], 0.f) : acc[r][j];
        }
    }
    __syncwarp();
}
template <int MAXR>
__device__ __forceinline__ void warp_dense_n(const float* w, const DenseDesc& L, const float* in, int ldin, int rows, float* out, int ldout, int lane) {
    if (L.N <= 32) warp_dense<MAXR, 1>(w, L, in, ldin, rows, out, ldout, lane);
    else if (L.N <= 64) warp_dense<MAXR, 2>(w, L, in, ldin, rows, out, ldout, lane);
    else warp_dense<(MAXR > 8 ? 8 : MAXR), 4>(w, L, in, ldin, rows > 8 ? 8 : rows, out, ldout, lane);
}

// epsilon-greedy / greedy selection.  optimal = np.argmax (first maximum).  With exploration the action is
// np_random.choice(actions, p=dist) = searchsorted(cumsum(p)/sum(p), u, side='right') for u ~ U[0,1).
__device__ int select_action(const float* q, int n, double eps, double u) {
    int best = 0;
    for (int a = 1; a < n; ++a) if (q[a] > q[best]) best = a;
    if (!(eps > 0.0)) return best;
    double cdf[16], run = 0;
    for (int a = 0; a < n; ++a) { run += eps / n + (a == best ? 1 - eps : 0.0); cdf[a] = run; }
    int idx = 0;
    for (int a = 0; a < n; ++a) if (cdf[a] / run <= u) ++idx;
    return idx < n ? idx : n - 1;
}

__device__ double uniform_for(uint64_t seed, uint64_t step, int env) {
    uint32_t c[4] = {(uint32_t)env, (uint32_t)step, (uint32_t)(step >> 32), 0x51u};
    philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t x = (((uint64_t)c[0] << 32) | c[1]) >> 11;
    return (double)x * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(kWarps * 32) k_qnet_fp32(QnetDev net, const float* __restrict__ weights, const float* __restrict__ obs, int E,
                                                           double eps, uint64_t seed, uint64_t step, const double* __restrict__ u_inj,
                                                           int32_t* __restrict__ actions, float* __restrict__ qout) {
    extern __shared__ __align__(16) float smem[];
    float* w = smem;
    for (int k = threadIdx.x; k < net.n_weights; k += blockDim.x) w[k] = __ldg(weights + k);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* scr = smem + ((net.n_weights + 3) & ~3) + warp * net.scratch_per_warp;
    const ttrl_qnet_desc& d = net.d;
    const int nin = d.n_entities * d.n_features;
    const int A = d.n_actions;

    if (d.type == TTRL_QNET_EGO_ATTENTION) {
        const int Fs = d.feature_size, H = d.heads, dk = Fs / H, NE = d.n_entities, Fe = d.n_features;
        float* x = scr;                       // [NE][Fe]
        float* bufA = x + ((NE * Fe + 3) & ~3);  // [NE][width]
        float* bufB = bufA + kMaxRows * net.width;
        float* bufC = bufB + kMaxRows * net.width;
        for (int e = blockIdx.x * kWarps + warp; e < E; e += gridDim.x * kWarps) {
            for (int k = lane; k < nin; k += 32) x[k] = obs[(size_t)e * nin + k];
            __syncwarp();
            // embeddings: row 0 through ego_embedding, rows 1.. through others_embedding (models.py:302-304)
            const float* cur = x; int ld = Fe;
            float* pong[2] = {bufA, bufB};
            for (int l = 0; l < net.n_ego; ++l) {
                float* o = pong[(net.n_ego - 1 - l) & 1];  // last layer lands in bufA
                warp_dense_n<1>(w, net.ego[l], cur, ld, 1, o, net.width, lane);
                warp_dense_n<kMaxRows>(w, net.oth[l], cur + ld, ld, NE - 1, o + net.width, net.width, lane);
                cur = o; ld = net.width;
            }
            // bufA = input_all [NE][Fs]; K -> bufB, V -> bufC, Q(ego) -> x region is too small, use tail of bufC row NE
            warp_dense_n<kMaxRows>(w, net.wk, bufA, net.width, NE, bufB, net.width, lane);
            warp_dense_n<kMaxRows>(w, net.wv, bufA, net.width, NE, bufC, net.width, lane);
            float* qe = bufB + (kMaxRows - 1) * net.width;  // row 15 of bufB is free when NE <= 15
            warp_dense_n<1>(w, net.wq, bufA, net.width, 1, qe, net.width, lane);
            // attention (models.py:370-388): scores / sqrt(dk), masked_fill(-1e9), softmax over entities
            float* p = bufC + (kMaxRows - 1) * net.width;   // [H][NE] probabilities in the free row of bufC
            const float inv = 1.0f / sqrtf((float)dk);
            for (int idx = lane; idx < H * NE; idx += 32) {
                const int h = idx / NE, n = idx % NE;
                float s = 0.f;
                for (int t = 0; t < dk; ++t) s = fmaf(qe[h * dk + t], bufB[n * net.width + h * dk + t], s);
                s *= inv;
                if (x[n * Fe + d.presence_feature_idx] < 0.5f) s = -1e9f;
                p[idx] = s;
            }
            __syncwarp();
            if (lane < H) {
                float m = -INFINITY;
                for (int n = 0; n < NE; ++n) m = fmaxf(m, p[lane * NE + n]);
                float sum = 0.f;
                for (int n = 0; n < NE; ++n) { const float ev = expf(p[lane * NE + n] - m); p[lane * NE + n] = ev; sum += ev; }
                for (int n = 0; n < NE; ++n) p[lane * NE + n] /= sum;
            }
            __syncwarp();
            // value = p @ V  -> row 0 of bufB region reused: write into qe
            for (int c = lane; c < Fs; c += 32) {
                const int h = c / dk;
                float a = 0.f;
                for (int n = 0; n < NE; ++n) a = fmaf(p[h * NE + n], bufC[n * net.width + c], a);
                qe[c] = a;
            }
            __syncwarp();
            // result = (attention_combine(value) + ego) / 2 (models.py:193)
            float* res = bufB;  // row 0 of bufB (K no longer needed)
            warp_dense_n<1>(w, net.wc, qe, net.width, 1, res, net.width, lane);
            for (int c = lane; c < Fs; c += 32) res[c] = (res[c] + bufA[c]) / 2.f;
            __syncwarp();
            // output MLP (models.py:69-76)
            const float* oc = res;
            float* opong[2] = {bufC, bufA};
            for (int l = 0; l < net.n_out; ++l) {
                float* o = opong[l & 1];
                warp_dense_n<1>(w, net.out[l], oc, net.width, 1, o, net.width, lane);
                oc = o;
            }
            if (lane == 0) {
                if (qout) for (int a = 0; a < A; ++a) qout[(size_t)e * A + a] = oc[a];
                actions[e] = select_action(oc, A, eps, u_inj ? u_inj[e] : uniform_for(seed, step, e));
            }
            __syncwarp();
        }
    } else {
        // MLP (models.py:50-76) or DuelingNetwork (models.py:79-104): 8 envs per warp iteration
        constexpr int G = 8;
        const int ldx = (nin + 3) & ~3;
        float* x = scr;  // [G][ldx]
        float* bufA = x + G * ldx;
        float* bufB = bufA + G * net.width;
        float* bufC = bufB + G * net.width;
        const int groups = (E + G - 1) / G;
        for (int gi = blockIdx.x * kWarps + warp; gi < groups; gi += gridDim.x * kWarps) {
            const int e0 = gi * G, rows = min(G, E - e0);
            for (int k = lane; k < rows * nin; k += 32) x[(k / nin) * ldx + (k % nin)] = obs[(size_t)e0 * nin + k];
            __syncwarp();
            const float* cur = x; int ld = ldx;
            float* pong[2] = {bufA, bufB};
            const int nl = d.type == TTRL_QNET_MLP ? net.n_out : net.n_base;
            const DenseDesc* Ls = d.type == TTRL_QNET_MLP ? net.out : net.base;
            for (int l = 0; l < nl; ++l) {
                float* o = pong[l & 1];
                warp_dense_n<G>(w, Ls[l], cur, ld, rows, o, net.width, lane);
                cur = o; ld = net.width;
            }
            if (d.type == TTRL_QNET_DUELING) {
                // value head -> bufC[.][0], advantage head -> other pong buffer
                const float* vcur = cur; int vld = ld;
                float* vp[2] = {bufC, bufC + G * net.width / 2};
                for (int l = 0; l < net.n_val; ++l) { float* o = vp[l & 1]; warp_dense_n<G>(w, net.val[l], vcur, vld, rows, o, net.width / 2, lane); vcur = o; vld = net.width / 2; }
                float* other = (cur == bufA) ? bufB : bufA;
                const float* acur = cur; int ald = ld;
                for (int l = 0; l < net.n_adv; ++l) {
                    // intermediate advantage layers would overwrite `cur`; only the shipped shape (no hidden layers) is supported
                    warp_dense_n<G>(w, net.adv[l], acur, ald, rows, other, net.width, lane);
                    acur = other; ald = net.width;
                }
                if (lane < rows) {
                    float mean = 0.f;
                    for (int a = 0; a < A; ++a) mean += acur[lane * ald + a];
                    mean /= (float)A;
                    for (int a = 0; a < A; ++a) other[lane * net.width + a] = vcur[lane * vld] + acur[lane * ald + a] - mean;
                }
                __syncwarp();
                cur = other; ld = net.width;
            }
            if (lane < rows) {
                const int e = e0 + lane;
                const float* qv = cur + lane * ld;
                if (qout) for (int a = 0; a < A; ++a) qout[(size_t)e * A + a] = qv[a];
                actions[e] = select_action(qv, A, eps, u_inj ? u_inj[e] : uniform_for(seed, step, e));
            }
            __syncwarp();
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Tensor-core path (throughput mode) for the MultiLayerPerceptron Q-network (models.py:50-76, the reference's
// DQN baseline.json: 105 -> 128 -> 128 -> 3): the two hidden GEMMs run on the 5th-generation tensor cores.
//
//  * one CTA = one tile of 128 observations (UMMA M = 128), 128 threads; persistent over tiles;
//  * operands live in shared memory in the canonical K-major, no-swizzle UMMA layout (8 x 16-byte core
//    matrices): element (row, k) at (row/8)*SBO + (k/8)*128 + (row%8)*16 + (k%8)*2 bytes, SBO = K/8*128;
//  * tcgen05.mma (kind::f16, BF16 inputs, FP32 accumulate) is issued by ONE thread, accumulators live in TMEM
//    (128 lanes x N columns), completion is signalled with tcgen05.commit on an mbarrier, and the epilogue reads
//    the accumulators back with tcgen05.ld (32x32b: thread t <-> TMEM lane t <-> observation t of the tile);
//  * fp32 accuracy from BF16 tensor cores: every operand is split x = hi + lo (two BF16 values) and each product
//    is accumulated as hi*hi + hi*lo + lo*hi (the dropped lo*lo term is ~2^-16 relative), so Q-values agree with
//    the fp32 path to ~1e-5 and the greedy action only differs on near-ties (tests/test_gpu_qnet.py);
//  * bias + ReLU + re-split of the activations happen in the epilogue, straight into the next layer's A operand;
//    the 3-wide head and the epsilon-greedy selection stay on the CUDA cores.
// ------------------------------------------------------------------------------------------------
struct TcMlp {
    int nin, K1, H1, H2, A;                              // K1 = nin rounded up to 16
    int w1, b1, w2, b2, w3, b3;                          // offsets into the fp32 blob ([in][out] matrices)
    // shared-memory bytes.  [0, img_bytes) is the WEIGHT IMAGE: split BF16 canonical B operands + the fp32 tail, prepared once
    // per weight update in HBM (k_qnet_mlp_prep) and brought in by every CTA with bulk (TMA) copies
    int off_w1hi, off_w1lo, off_w2hi, off_w2lo, off_f32, img_bytes, off_ahi, off_alo, off_raw, slot_bytes, n_slots, off_bar, total;
    int tmem_cols;
};
constexpr int kMlpSlotRows = 32;   // observations per slot of the raw-observation ring (a quarter of a 128-row tile)
constexpr int kMlpMaxSlots = 4;
constexpr int kMlpThreads = 512, kMlpColGroups = kMlpThreads / 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    // SmemDescriptor: start address [0,14), leading byte offset [16,30), stride byte offset [32,46) (all >> 4),
    // version 1 (Blackwell) at [46,48), layout type SWIZZLE_NONE = 0 at [61,64)
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy HBM -> shared memory on the TMA unit (SASS UBLKCP); completion counts `bytes` on the mbarrier.  16-byte aligned.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void split_bf16(float x, unsigned short& hi, unsigned short& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
}
// two values -> packed BF16x2 (hi pair, lo pair), the first value in the low half: one packed conversion per pair
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// canonical K-major no-swizzle offset (bytes) of element (row, k) in an operand with K columns
__device__ __forceinline__ uint32_t canon_off(int row, int k, int K) {
    return (uint32_t)((row >> 3) * (K >> 3) * 128 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

// issue the 3-term split product D (+)= A * B^T for K columns; one thread
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int K, uint32_t idesc) {
    const uint32_t sbo = (uint32_t)(K >> 3) * 128u;
    uint32_t acc = 0;
    for (int j = 0; j < K / 16; ++j) {
        const uint32_t o = (uint32_t)j * 256u;  // two 128-byte core matrices per K = 16 step
        umma_bf16(tmem_d, umma_desc(a_hi + o, 128, sbo), umma_desc(b_hi + o, 128, sbo), idesc, acc);
        acc = 1;
        umma_bf16(tmem_d, umma_desc(a_hi + o, 128, sbo), umma_desc(b_lo + o, 128, sbo), idesc, 1);
        umma_bf16(tmem_d, umma_desc(a_lo + o, 128, sbo), umma_desc(b_hi + o, 128, sbo), idesc, 1);
    }
}

#define TT_TMEM_LD32(v, addr)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                    \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),  \
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),  \
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                          \
                 : "r"(addr) : "memory");                                                                                    \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// The weight image of the tensor-core MLP: what k_qnet_mlp_tc keeps in shared memory bytes [0, img_bytes), laid out in HBM.
__global__ void k_qnet_mlp_prep(TcMlp d, const float* __restrict__ weights, unsigned char* __restrict__ img) {
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = t0; i < d.K1 * d.H1; i += stride) {   // B1[n][k] = W1t[k][n]; coalesced reads along n
        const int k = i / d.H1, n = i - k * d.H1;
        unsigned short hi, lo;
        split_bf16(k < d.nin ? weights[d.w1 + k * d.H1 + n] : 0.f, hi, lo);
        const uint32_t o = canon_off(n, k, d.K1);
        *reinterpret_cast<unsigned short*>(img + d.off_w1hi + o) = hi;
        *reinterpret_cast<unsigned short*>(img + d.off_w1lo + o) = lo;
    }
    for (int i = t0; i < d.H1 * d.H2; i += stride) {   // B2[n][k] = W2t[k][n]
        const int k = i / d.H2, n = i - k * d.H2;
        unsigned short hi, lo;
        split_bf16(weights[d.w2 + k * d.H2 + n], hi, lo);
        const uint32_t o = canon_off(n, k, d.H1);
        *reinterpret_cast<unsigned short*>(img + d.off_w2hi + o) = hi;
        *reinterpret_cast<unsigned short*>(img + d.off_w2lo + o) = lo;
    }
    float* f32 = reinterpret_cast<float*>(img + d.off_f32);   // b1[H1] b2[H2] W3[A][H2] b3[A]
    for (int i = t0; i < d.H1; i += stride) f32[i] = weights[d.b1 + i];
    for (int i = t0; i < d.H2; i += stride) f32[d.H1 + i] = weights[d.b2 + i];
    for (int i = t0; i < d.A * d.H2; i += stride) { const int a = i / d.H2, k = i - a * d.H2; f32[d.H1 + d.H2 + i] = weights[d.w3 + k * d.A + a]; }
    for (int i = t0; i < d.A; i += stride) f32[d.H1 + d.H2 + d.A * d.H2 + i] = weights[d.b3 + i];
}

// Staging is a TMA pipeline: thread 0 streams the raw fp32 observations of the CTA's tiles, a quarter tile (32 rows, contiguous in
// HBM) per bulk copy, through a ring of n_slots shared-memory slots, one mbarrier each; the 128 threads turn a landed slot into the
// split-BF16 canonical A operand from shared memory (conflict-free reads: a warp reads 32 rows at one column, row stride nin is odd for
// the shipped 105) and the slot is refilled with the quarter n_slots ahead -- so the loads of the next tile run under the MMAs and
// epilogues of this one.  A ragged last quarter (or an unaligned observation pointer) takes guarded global loads instead.
__global__ void __launch_bounds__(kMlpThreads, 1) k_qnet_mlp_tc(TcMlp d, const unsigned char* __restrict__ img, const float* __restrict__ obs, int E,
                                                         double eps, uint64_t seed, uint64_t step, const double* __restrict__ u_inj,
                                                         int32_t* __restrict__ actions, float* __restrict__ qout) {
    extern __shared__ __align__(1024) unsigned char sm[];
    // 16 warps: warp w owns TMEM lanes (= tile rows) 32 (w % 4) .. + 32 and, in the epilogues, the 32-column chunks cg, cg + 4, ..
    // (cg = w / 4): the four warps of a lane quadrant split the columns of a row, four warps per scheduler hide the TMEM / shared
    // memory latencies one warp per scheduler could not (profiles/r2_qnet_mlp_raw.txt: 10 % issue slots with 4 warps)
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, cg = tid >> 7;
    float* f32 = reinterpret_cast<float*>(sm + d.off_f32);   // b1[H1] b2[H2] W3[A][H2] b3[A]
    float* sb1 = f32; float* sb2 = sb1 + d.H1; float* sw3 = sb2 + d.H2; float* sb3 = sw3 + d.A * d.H2;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + d.off_bar);               // MMA completion
    const uint32_t bar_w = smem_u32(bar) + 8, bar_full = smem_u32(bar) + 16;   // weight image; ring slots
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + d.off_bar + 16 + 8 * kMlpMaxSlots);
    const int NS = d.n_slots, tiles = (E + 127) / 128;
    const bool tma = NS >= 2 && (reinterpret_cast<uintptr_t>(obs) & 15) == 0;
    // quarter g of this CTA: tile blockIdx.x + (g / 4) gridDim.x, rows 32 (g % 4) .. + 32; only whole quarters go through the ring
    auto quarter_full = [&](int g) { const int tl = blockIdx.x + (g >> 2) * gridDim.x; return tl < tiles && tl * 128 + (g & 3) * kMlpSlotRows + kMlpSlotRows <= E; };
    auto issue_quarter = [&](int g) {
        const int slot = g % NS, tl = blockIdx.x + (g >> 2) * gridDim.x;
        mbar_expect_tx(bar_full + 8 * slot, (uint32_t)d.slot_bytes);
        bulk_g2s(smem_u32(sm + d.off_raw + slot * d.slot_bytes), obs + (size_t)(tl * 128 + (g & 3) * kMlpSlotRows) * d.nin, (uint32_t)d.slot_bytes, bar_full + 8 * slot);
    };

    if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        mbar_init(bar_w, 1);
        for (int k = 0; k < kMlpMaxSlots; ++k) mbar_init(bar_full + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // ---- one-time setup: the weight image (split BF16 canonical B operands, biases, head) by bulk copies ----
        mbar_expect_tx(bar_w, (uint32_t)d.img_bytes);
        for (int o = 0; o < d.img_bytes; o += 32768) bulk_g2s(smem_u32(sm + o), img + o, (uint32_t)min(32768, d.img_bytes - o), bar_w);
        if (tma) for (int g = 0; g < NS; ++g) if (quarter_full(g)) issue_quarter(g);
    }
    if (warp == 0) {  // TMEM allocation by one warp; the address lands in shared memory
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)d.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(bar_w, 0);
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 TMEM lanes
    // instruction descriptor: D = F32 [4,6), A = B = BF16 [7,10) [10,13), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d.H1 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(d.H2 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hi = smem_u32(sm + d.off_ahi), a_lo = smem_u32(sm + d.off_alo);
    uint32_t parity = 0, slot_phase = 0;
    int gq = 0;   // running quarter index of this CTA

    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int e0 = tile * 128, rows = min(128, E - e0);
        // ---- A1 = split(obs tile) in the K1 layout, a quarter at a time; task = (8 columns, row), row fastest ----
        for (int q = 0; q < 128 / kMlpSlotRows; ++q, ++gq) {
            const int slot = tma ? gq % NS : 0;
            const bool landed = tma && e0 + q * kMlpSlotRows + kMlpSlotRows <= E;
            const float* raw = reinterpret_cast<const float*>(sm + d.off_raw + slot * d.slot_bytes);
            if (landed) { mbar_wait(bar_full + 8 * slot, (slot_phase >> slot) & 1u); slot_phase ^= 1u << slot; }
            for (int i = tid; i < kMlpSlotRows * (d.K1 >> 3); i += kMlpThreads) {
                const int r = i & (kMlpSlotRows - 1), k0 = (i / kMlpSlotRows) * 8, arow = q * kMlpSlotRows + r;
                float x[8];
                if (landed) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = k0 + j < d.nin ? raw[r * d.nin + k0 + j] : 0.f;
                } else {
                    const float* src = obs + (size_t)(e0 + arow) * d.nin;
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = (arow < rows && k0 + j < d.nin) ? __ldg(src + k0 + j) : 0.f;
                }
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    split_bf16x2(x[2 * j], x[2 * j + 1], ph[j], pl[j]);
                }
                const uint32_t o = canon_off(arow, k0, d.K1);
                *reinterpret_cast<uint4*>(sm + d.off_ahi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4*>(sm + d.off_alo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
            if (tma) {
                __syncthreads();   // every thread has read the slot: refill it with the quarter n_slots ahead
                if (tid == 0 && quarter_full(gq + NS)) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue_quarter(gq + NS);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {   // layer 1: [128 x K1] x [K1 x H1]
            issue_layer(tmem, a_hi, a_lo, smem_u32(sm + d.off_w1hi), smem_u32(sm + d.off_w1lo), d.K1, idesc1);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
        }
        mbar_wait(smem_u32(bar), parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 1: h1 = relu(acc + b1) -> split -> A2 (K = H1 layout); thread = (row, column group) ----
        for (int c0 = cg * 32; c0 < d.H1; c0 += 32 * kMlpColGroups) {
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                           "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                           "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(tmem_row + (uint32_t)c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int g = 0; g < 4; ++g) {  // 8 columns = one 16-byte row of a core matrix
                if (c0 + g * 8 >= d.H1) break;   // widths are multiples of 16, chunks of 32
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    split_bf16x2(fmaxf(__uint_as_float(v[g * 8 + 2 * q]) + sb1[c0 + g * 8 + 2 * q], 0.f), fmaxf(__uint_as_float(v[g * 8 + 2 * q + 1]) + sb1[c0 + g * 8 + 2 * q + 1], 0.f), ph[q], pl[q]);
                }
                const uint32_t o = canon_off(row, c0 + g * 8, d.H1);
                *reinterpret_cast<uint4*>(sm + d.off_ahi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4*>(sm + d.off_alo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {   // layer 2: [128 x H1] x [H1 x H2]
            issue_layer(tmem, a_hi, a_lo, smem_u32(sm + d.off_w2hi), smem_u32(sm + d.off_w2lo), d.H1, idesc2);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
        }
        mbar_wait(smem_u32(bar), parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 2: h2 = relu(acc + b2); head (A outputs, 4 at a time in registers) per column group, summed through the
        // free A operand; action selection on the CUDA cores ----
        float* part = reinterpret_cast<float*>(sm + d.off_ahi);   // [column group][row][A rounded up to 4]
        const int A4 = (d.A + 3) & ~3;
        for (int a0 = 0; a0 < d.A; a0 += 4) {
            float q4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c0 = cg * 32; c0 < d.H2; c0 += 32 * kMlpColGroups) {
                uint32_t v[32];
                TT_TMEM_LD32(v, tmem_row + (uint32_t)c0);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (c0 + j >= d.H2) break;
                    const float h = fmaxf(__uint_as_float(v[j]) + sb2[c0 + j], 0.f);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (a0 + u < d.A) q4[u] = fmaf(h, sw3[(a0 + u) * d.H2 + c0 + j], q4[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (a0 + u < d.A) part[(cg * 128 + row) * A4 + a0 + u] = q4[u];
        }
        __syncthreads();
        if (cg == 0 && row < rows) {
            const int e = e0 + row;
            float q[16];
            for (int a = 0; a < d.A; ++a) {
                float sum = sb3[a];
                for (int g = 0; g < kMlpColGroups; ++g) sum += part[(g * 128 + row) * A4 + a];
                q[a] = sum;
            }
            if (qout) for (int a = 0; a < d.A; ++a) qout[(size_t)e * d.A + a] = q[a];
            actions[e] = select_action(q, d.A, eps, u_inj ? u_inj[e] : uniform_for(seed, step, e));
        }
        // all TMEM reads of this tile are done (wait::ld above) before the next tile's MMAs overwrite the accumulators
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)d.tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Tensor-core path for the EgoAttentionNetwork (models.py:237-312; the reference's default DQN model,
// ego_attention_2h.json): the embedding MLPs and the key / value / query projections -- 89 % of its FLOPs -- run
// on tcgen05; attention, the combine layer and the 64-wide output MLP stay on the CUDA cores.
//
//  * tile = 8 observations = 128 rows: row r = 16 * env + entity (15 entities + 1 zero pad row), UMMA M = 128;
//  * ego and others embeddings share each MMA: B = [W_others ; W_ego] stacked along N (N = 128), the epilogue of
//    row r takes columns 0..63 (others) or 64..127 (ego row, entity 0);
//  * K | V | Q come from ONE MMA with N = 192 ([W_k ; W_v ; W_q]); they are never written to shared memory:
//    thread r keeps its K and V rows in registers (tcgen05.ld), scores are softmaxed across the 16 threads of an
//    observation with shuffles, the p-weighted V rows are summed through a shared scratch;
//  * BF16x3 split operands, FP32 accumulation, like the MLP path.
// ------------------------------------------------------------------------------------------------
struct TcEgo {
    int NE, Fe, Fs, H, A, pidx;
    int ego_w1, ego_b1, ego_w2, ego_b2, oth_w1, oth_b1, oth_w2, oth_b2, wk, wv, wq, wc, o_w1, o_b1, o_w2, o_b2, p_w, p_b;  // blob offsets
    int off_b1hi, off_b1lo, off_b2hi, off_b2lo, off_b3hi, off_b3lo, off_ahi, off_f32, off_x, x_stride, off_small, small_stride, off_bar, total;
};


__device__ __forceinline__ void tc_sync_before_mma() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// The ego kernel runs TWO tiles per CTA, one per group of 8 warps (threads 0..255 / 256..511), each with its own A operand,
// accumulators (256 TMEM columns), mbarrier and named barrier (ids 1 and 2): while one group waits for its MMA or runs an
// epilogue, the other one stages, issues or reads back -- the tensor pipe and the CUDA cores of the SM overlap across the two tiles.
// Inside a group, thread (row, cg) owns tile row `row` (= TMEM lane) and the 32-column half `cg` of every 64-wide quantity, so
// 16 warps (four per scheduler) share the epilogues (profiles/r2_qnet_ego_raw.txt: 28 % issue slots with 8 warps).
constexpr int kEgoGT = 256;
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, %1;" :: "r"(grp + 1), "n"(kEgoGT) : "memory"); }
__device__ __forceinline__ void tc_group_sync_before_mma(int grp) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    group_sync(grp);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// out[env][n] = act(bias[n] + sum_k in[env][k] * Wt[k][n]) for the 8 observations of a tile; K % 4 == 0.  Wide layers: thread =
// (pair of observations, one column) -- a warp reads 32 consecutive weights (one shared-memory wavefront per k) and two broadcast
// input vectors per four k, 6 wavefronts per 8 FMAs instead of 9 for (observation, column pair)
__device__ __forceinline__ void tile_dense8(const float* in, const float* Wt, const float* bias, int K, int N, bool relu, float* out, int t) {
    const int env = t >> 5, j = t & 31;
    if (N >= 64) {
        const int col = t & 63, e0 = (t >> 6) * 2;
        float a0 = bias ? bias[col] : 0.f, a1 = a0;
        const float* in0 = in + e0 * 64;
        const float* in1 = in0 + 64;
        for (int k = 0; k < K; k += 4) {
            const float4 x0 = *reinterpret_cast<const float4*>(in0 + k), x1 = *reinterpret_cast<const float4*>(in1 + k);
            const float w0 = Wt[k * N + col], w1 = Wt[(k + 1) * N + col], w2 = Wt[(k + 2) * N + col], w3 = Wt[(k + 3) * N + col];
            a0 = fmaf(x0.x, w0, a0); a1 = fmaf(x1.x, w0, a1);
            a0 = fmaf(x0.y, w1, a0); a1 = fmaf(x1.y, w1, a1);
            a0 = fmaf(x0.z, w2, a0); a1 = fmaf(x1.z, w2, a1);
            a0 = fmaf(x0.w, w3, a0); a1 = fmaf(x1.w, w3, a1);
        }
        if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
        out[e0 * 64 + col] = a0;
        out[(e0 + 1) * 64 + col] = a1;
    } else {   // N <= 4 (the action head): lane = slice of k, butterfly sum over the warp
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = j; k < K; k += 32) {
            const float x = in[env * 64 + k];
#pragma unroll
            for (int n = 0; n < 4; ++n)
                if (n < N) acc[n] = fmaf(x, Wt[k * N + n], acc[n]);
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], off);
        }
        if (j < N) {
            const float a = (j == 0 ? acc[0] : j == 1 ? acc[1] : j == 2 ? acc[2] : acc[3]) + (bias ? bias[j] : 0.f);
            out[env * 64 + j] = relu ? fmaxf(a, 0.f) : a;
        }
    }
}

__global__ void __launch_bounds__(2 * kEgoGT, 1) k_qnet_ego_tc(TcEgo d, const float* __restrict__ weights, const float* __restrict__ obs, int E,
                                                                double eps, uint64_t seed, uint64_t step, const double* __restrict__ u_inj,
                                                                int32_t* __restrict__ actions, float* __restrict__ qout) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const int tid = threadIdx.x, grp = tid / kEgoGT, t = tid % kEgoGT;
    const int row = t & 127, cg = t >> 7, warp = t >> 5;   // warp w of the group reads TMEM lanes 32 (w % 4) ..
    const int env = row >> 4, ent = row & 15;              // row of the group's tile = entity `ent` of observation `env`
    const int denv = t >> 5, dj = t & 31;                  // (observation, column pair) mapping of the per-observation tail
    const int Fs = d.Fs, nin = d.NE * d.Fe;
    float* f32 = reinterpret_cast<float*>(sm + d.off_f32);
    float* s_wc = f32; float* s_ow1 = s_wc + Fs * Fs; float* s_ow2 = s_ow1 + Fs * Fs; float* s_pw = s_ow2 + Fs * Fs;
    float* s_b1o = s_pw + Fs * 4; float* s_b1e = s_b1o + Fs; float* s_b2o = s_b1e + Fs; float* s_b2e = s_b2o + Fs;
    float* s_ob1 = s_b2e + Fs; float* s_ob2 = s_ob1 + Fs; float* s_pb = s_ob2 + Fs;
    float* s_x = reinterpret_cast<float*>(sm + d.off_x + grp * d.x_stride);          // [8][nin]
    float* s_q = reinterpret_cast<float*>(sm + d.off_small + grp * d.small_stride);  // [8][64] query of the ego
    // s_val and s_t1 reuse the query's storage: the query is dead after the scores, s_val after the combine layer
    float* s_ego = s_q + 8 * 64; float* s_val = s_q; float* s_t0 = s_ego + 8 * 64; float* s_t1 = s_q;
    unsigned char* a_hi_p = sm + d.off_ahi + grp * (2 * 128 * 64 * 2);               // this group's A operand: hi then lo
    unsigned char* a_lo_p = a_hi_p + 128 * 64 * 2;
    float* scratch = reinterpret_cast<float*>(a_hi_p);            // [128][64] fp32, aliases the A operand (hi + lo)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + d.off_bar + grp * 8);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + d.off_bar + 16);

    // ---- one-time setup: stacked B operands (split BF16, canonical layout) and the fp32 tail weights ----
    for (int i = tid; i < 128 * 16; i += 2 * kEgoGT) {            // B1[n][k], K = 16: n < 64 others layer 1, n >= 64 ego layer 1
        const int n = i >> 4, k = i & 15;
        float w = 0.f;
        if (k < d.Fe) w = __ldg(weights + (n < 64 ? d.oth_w1 : d.ego_w1) + k * Fs + (n & 63));
        unsigned short hi, lo; split_bf16(w, hi, lo);
        const uint32_t o = canon_off(n, k, 16);
        *reinterpret_cast<unsigned short*>(sm + d.off_b1hi + o) = hi; *reinterpret_cast<unsigned short*>(sm + d.off_b1lo + o) = lo;
    }
    for (int i = tid; i < 128 * 64; i += 2 * kEgoGT) {            // B2[n][k], K = 64
        const int n = i >> 6, k = i & 63;
        unsigned short hi, lo; split_bf16(__ldg(weights + (n < 64 ? d.oth_w2 : d.ego_w2) + k * Fs + (n & 63)), hi, lo);
        const uint32_t o = canon_off(n, k, 64);
        *reinterpret_cast<unsigned short*>(sm + d.off_b2hi + o) = hi; *reinterpret_cast<unsigned short*>(sm + d.off_b2lo + o) = lo;
    }
    for (int i = tid; i < 192 * 64; i += 2 * kEgoGT) {            // B3[n][k] = [W_k ; W_v ; W_q]
        const int n = i >> 6, k = i & 63;
        const int base = n < 64 ? d.wk : n < 128 ? d.wv : d.wq;
        unsigned short hi, lo; split_bf16(__ldg(weights + base + k * Fs + (n & 63)), hi, lo);
        const uint32_t o = canon_off(n, k, 64);
        *reinterpret_cast<unsigned short*>(sm + d.off_b3hi + o) = hi; *reinterpret_cast<unsigned short*>(sm + d.off_b3lo + o) = lo;
    }
    for (int i = tid; i < Fs * Fs; i += 2 * kEgoGT) { s_wc[i] = __ldg(weights + d.wc + i); s_ow1[i] = __ldg(weights + d.o_w1 + i); s_ow2[i] = __ldg(weights + d.o_w2 + i); }
    for (int i = tid; i < Fs * d.A; i += 2 * kEgoGT) s_pw[i] = __ldg(weights + d.p_w + i);
    for (int i = tid; i < Fs; i += 2 * kEgoGT) {
        s_b1o[i] = __ldg(weights + d.oth_b1 + i); s_b1e[i] = __ldg(weights + d.ego_b1 + i);
        s_b2o[i] = __ldg(weights + d.oth_b2 + i); s_b2e[i] = __ldg(weights + d.ego_b2 + i);
        s_ob1[i] = __ldg(weights + d.o_b1 + i); s_ob2[i] = __ldg(weights + d.o_b2 + i);
    }
    for (int i = tid; i < d.A; i += 2 * kEgoGT) s_pb[i] = __ldg(weights + d.p_b + i);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {   // 512 columns: 256 per group (one CTA per SM: the shared-memory footprint excludes a second one)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_sync_before_mma();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem = tmem_base + (uint32_t)grp * 256u;
    const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc192 = (1u << 4) | (1u << 7) | (1u << 10) | ((192u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hi = smem_u32(a_hi_p), a_lo = smem_u32(a_lo_p);
    uint32_t parity = 0;
    const int dk = Fs / d.H;
    const float inv = 1.0f / sqrtf((float)dk);
    const uint32_t c0 = (uint32_t)cg * 32u;   // this thread's half of every 64-wide quantity

    const int tiles = (E + 7) / 8;
    constexpr int kEgoPre = 8;   // 8 observations x (15 entities x 16 features) / 256 threads, rounded up
    float pre[kEgoPre];
    auto fetch_obs = [&](int tl) {
        const int valid = tl < tiles ? min(8, E - tl * 8) * nin : 0;
#pragma unroll
        for (int u = 0; u < kEgoPre; ++u) {
            const int i = t + kEgoGT * u;
            pre[u] = i < valid ? __ldg(obs + (size_t)tl * 8 * nin + i) : 0.f;
        }
    };
    fetch_obs(2 * blockIdx.x + grp);
    for (int tile = 2 * blockIdx.x + grp; tile < tiles; tile += 2 * gridDim.x) {
        const int e0 = tile * 8, nenv = min(8, E - e0);
        // ---- observations of the tile (contiguous in HBM): fetched into registers one tile ahead, so that the load latency
        // runs under the previous tile's work ----
#pragma unroll
        for (int u = 0; u < kEgoPre; ++u)
            if (t + kEgoGT * u < 8 * nin) s_x[t + kEgoGT * u] = pre[u];
        group_sync(grp);
        fetch_obs(tile + 2 * gridDim.x);
        // ---- A1[r][k] (K = 16): features 8 cg .. 8 cg + 7 of entity `ent` of observation `env`, zero padded ----
        {
            uint32_t ph[4], pl[4];
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) {
                const int k0 = 8 * cg + 2 * q2, k1 = k0 + 1;
                split_bf16x2((ent < d.NE && k0 < d.Fe) ? s_x[env * nin + ent * d.Fe + k0] : 0.f,
                             (ent < d.NE && k1 < d.Fe) ? s_x[env * nin + ent * d.Fe + k1] : 0.f, ph[q2], pl[q2]);
            }
            const uint32_t o = canon_off(row, 8 * cg, 16);
            *reinterpret_cast<uint4*>(a_hi_p + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            *reinterpret_cast<uint4*>(a_lo_p + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
        // ---- two embedding layers: MMA (N = 128: others | ego) + epilogue selecting the row's half ----
        for (int layer = 0; layer < 2; ++layer) {
            tc_group_sync_before_mma(grp);
            if (t == 0) {
                if (layer == 0) issue_layer(tmem, a_hi, a_lo, smem_u32(sm + d.off_b1hi), smem_u32(sm + d.off_b1lo), 16, idesc128);
                else issue_layer(tmem, a_hi, a_lo, smem_u32(sm + d.off_b2hi), smem_u32(sm + d.off_b2lo), 64, idesc128);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
            }
            mbar_wait(smem_u32(bar), parity);
            parity ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float* bb = ent == 0 ? (layer == 0 ? s_b1e : s_b2e) : (layer == 0 ? s_b1o : s_b2o);   // ego | others bias
            {
                uint32_t vo[32], ve[32];
                TT_TMEM_LD32(vo, tmem_row + c0);
                TT_TMEM_LD32(ve, tmem_row + 64u + c0);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t ph[4], pl[4];
                    float hv[8];
#pragma unroll
                    for (int q2 = 0; q2 < 8; ++q2) {
                        const int c = (int)c0 + g * 8 + q2;
                        hv[q2] = fmaxf(__uint_as_float(ent == 0 ? ve[g * 8 + q2] : vo[g * 8 + q2]) + bb[c], 0.f);
                    }
#pragma unroll
                    for (int q2 = 0; q2 < 4; ++q2) split_bf16x2(hv[2 * q2], hv[2 * q2 + 1], ph[q2], pl[q2]);
                    const uint32_t o = canon_off(row, (int)c0 + g * 8, 64);
                    *reinterpret_cast<uint4*>(a_hi_p + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    *reinterpret_cast<uint4*>(a_lo_p + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    if (layer == 1 && ent == 0) {   // the ego's embedding: residual of the attention block
                        float4* eg = reinterpret_cast<float4*>(s_ego + env * 64 + (int)c0 + g * 8);
                        eg[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                        eg[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
                    }
                }
            }
        }
        // ---- K | V | Q = input_all x [W_k ; W_v ; W_q]^T ----
        tc_group_sync_before_mma(grp);
        if (t == 0) {
            issue_layer(tmem, a_hi, a_lo, smem_u32(sm + d.off_b3hi), smem_u32(sm + d.off_b3lo), 64, idesc192);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
        }
        mbar_wait(smem_u32(bar), parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {   // the ego rows publish their query
            uint32_t v0[32];
            TT_TMEM_LD32(v0, tmem_row + 128u + c0);
            if (ent == 0) {
                float4* qd = reinterpret_cast<float4*>(s_q + env * 64 + (int)c0);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    qd[j] = make_float4(__uint_as_float(v0[4 * j]), __uint_as_float(v0[4 * j + 1]), __uint_as_float(v0[4 * j + 2]), __uint_as_float(v0[4 * j + 3]));
            }
        }
        group_sync(grp);
        // ---- attention (models.py:370-388): scores / sqrt(d_k), masked_fill(-1e9), softmax over the entities.  The thread's 32
        // key columns hold two heads (H = 4: heads 2 cg, 2 cg + 1), one head (H = 2: head cg) or half a head (H = 1) ----
        float p_lo, p_hi;   // probability of this row for the head of its columns 0..15 / 16..31
        {
            uint32_t kk[32];
            TT_TMEM_LD32(kk, tmem_row + c0);
            const bool masked = ent < d.NE ? s_x[env * nin + ent * d.Fe + d.pidx] < 0.5f : true;
            const float4* qrow = reinterpret_cast<const float4*>(s_q + env * 64 + c0);
            float sc_lo = 0.f, sc_hi = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 ql = qrow[c], qh = qrow[4 + c];
                sc_lo = fmaf(ql.x, __uint_as_float(kk[4 * c]), sc_lo); sc_lo = fmaf(ql.y, __uint_as_float(kk[4 * c + 1]), sc_lo);
                sc_lo = fmaf(ql.z, __uint_as_float(kk[4 * c + 2]), sc_lo); sc_lo = fmaf(ql.w, __uint_as_float(kk[4 * c + 3]), sc_lo);
                sc_hi = fmaf(qh.x, __uint_as_float(kk[16 + 4 * c]), sc_hi); sc_hi = fmaf(qh.y, __uint_as_float(kk[16 + 4 * c + 1]), sc_hi);
                sc_hi = fmaf(qh.z, __uint_as_float(kk[16 + 4 * c + 2]), sc_hi); sc_hi = fmaf(qh.w, __uint_as_float(kk[16 + 4 * c + 3]), sc_hi);
            }
            if (d.H != 4) { sc_lo += sc_hi; sc_hi = sc_lo; }
            if (d.H == 1) {   // the two halves of the single head meet through shared memory (s_t0 is free here)
                s_t0[cg * 128 + row] = sc_lo;
                group_sync(grp);
                sc_lo = s_t0[row] + s_t0[128 + row];
                sc_hi = sc_lo;
            }
            auto softmax16 = [&](float sc) {
                float v = sc * inv;
                if (masked) v = -1e9f;
                if (ent >= d.NE) v = -INFINITY;  // pad row: not an entity
                float m = v;
#pragma unroll
                for (int off = 1; off < 16; off <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                const float ev = ent < d.NE ? expf(v - m) : 0.f;
                float sum = ev;
#pragma unroll
                for (int off = 1; off < 16; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                return ev / sum;
            };
            p_lo = softmax16(sc_lo);
            p_hi = d.H == 4 ? softmax16(sc_hi) : p_lo;
        }
        {   // p-weighted value rows -> scratch (the A operand is free: its MMA has completed)
            uint32_t vv[32];
            TT_TMEM_LD32(vv, tmem_row + 64u + c0);
            float* orow = scratch + row * 64 + c0;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float pw = c < 16 ? p_lo : p_hi;
                *reinterpret_cast<float4*>(orow + c) = make_float4(pw * __uint_as_float(vv[c]), pw * __uint_as_float(vv[c + 1]),
                                                                   pw * __uint_as_float(vv[c + 2]), pw * __uint_as_float(vv[c + 3]));
            }
        }
        group_sync(grp);
        {   // value[env][c] = sum over the entities; thread = (env, 2 columns)
            float2 a = make_float2(0.f, 0.f);
            for (int n = 0; n < d.NE; ++n) {
                const float2 v = *reinterpret_cast<const float2*>(scratch + (denv * 16 + n) * 64 + 2 * dj);
                a.x += v.x; a.y += v.y;
            }
            *reinterpret_cast<float2*>(s_val + denv * 64 + 2 * dj) = a;
        }
        group_sync(grp);
        // ---- (attention_combine(value) + ego) / 2 (models.py:193), output MLP (models.py:69-76), action ----
        tile_dense8(s_val, s_wc, nullptr, Fs, Fs, false, s_t0, t);
        group_sync(grp);
        for (int c = 2 * dj; c < 2 * dj + 2; ++c) s_t0[denv * 64 + c] = (s_t0[denv * 64 + c] + s_ego[denv * 64 + c]) / 2.f;
        group_sync(grp);
        tile_dense8(s_t0, s_ow1, s_ob1, Fs, Fs, true, s_t1, t);
        group_sync(grp);
        tile_dense8(s_t1, s_ow2, s_ob2, Fs, Fs, true, s_t0, t);
        group_sync(grp);
        tile_dense8(s_t0, s_pw, s_pb, Fs, d.A, false, s_t1, t);
        group_sync(grp);
        if (cg == 0 && ent == 0 && env < nenv) {
            const int e = e0 + env;
            const float* qv = s_t1 + env * 64;
            if (qout) for (int a = 0; a < d.A; ++a) qout[(size_t)e * d.A + a] = qv[a];
            actions[e] = select_action(qv, d.A, eps, u_inj ? u_inj[e] : uniform_for(seed, step, e));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync(grp);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // both groups are done with their accumulators
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

struct ttrl_qnet {
    QnetDev net;
    float* d_weights = nullptr;
    unsigned char* d_img = nullptr;   // tensor-core MLP: the weight image (TcMlp), rebuilt when the weights change
    bool img_dirty = true;            // set by set_weights; stays set once the blob pointer has been handed out (ttrl_qnet_weights_dev)
    bool external_writer = false;
    int device = 0;
    int smem_bytes = 0;
    int n_sms = 0;
    int64_t launches = 0;
    int mode = 0;       // TTRL_QNET_MODE_FP32 (parity) | TTRL_QNET_MODE_TENSOR (tcgen05 BF16x3, MLP only)
    bool tc_ok = false;
    TcMlp tc{};
    TcEgo tce{};
};

#define QCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return qfail(std::string(#call) + ": " + cudaGetErrorString(e_));   \
    } while (0)

extern "C" {

// Weight blob order (all matrices TRANSPOSED to [in][out], float32):
//   MLP:      hidden layers (W, b)..., predict (W, b)
//   EgoAttn:  ego_embedding layers (W, b)..., others_embedding layers (W, b)..., key_all W, value_all W, query_ego W,
//             attention_combine W, output_layer hidden (W, b)..., output_layer.predict (W, b)
//   Dueling:  base_module layers (W, b)..., value layers (W, b)... + value.predict (W, b), advantage layers... + predict
int ttrl_qnet_create(const ttrl_qnet_desc* desc, const float* weights_host, int64_t n_weights, int device, ttrl_qnet** out) {
    if (!desc || !weights_host || !out) return qfail("null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return qfail("no CUDA device: the Q-network rollout has no CPU fallback");
    QCK(cudaSetDevice(device));
    ttrl_qnet* q = new ttrl_qnet();
    q->device = device;
    QnetDev& n = q->net;
    memset(&n, 0, sizeof n);
    n.d = *desc;
    int off = 0, width = 0;
    auto dense = [&](int K, int N, bool bias, bool relu) {
        DenseDesc L{K, N, off, -1, relu ? 1 : 0};
        off += K * N;
        if (bias) { L.b_off = off; off += N; }
        if (N > width) width = N;
        return L;
    };
    const int nin = desc->n_entities * desc->n_features;
    if (desc->n_actions > 16 || desc->n_entities > kMaxRows - 1) { delete q; return qfail("unsupported Q-network shape"); }
    if (desc->type == TTRL_QNET_EGO_ATTENTION) {
        if (desc->feature_size > 64 || desc->feature_size % desc->heads) { delete q; return qfail("EgoAttention feature_size must be <= 64 and divisible by heads"); }
        int k = desc->n_features;
        for (int l = 0; l < desc->embed_layers; ++l) { n.ego[n.n_ego++] = dense(k, desc->embed[l], true, true); k = desc->embed[l]; }
        k = desc->n_features;
        for (int l = 0; l < desc->embed_layers; ++l) { n.oth[n.n_oth++] = dense(k, desc->embed[l], true, true); k = desc->embed[l]; }
        if (k != desc->feature_size) { delete q; return qfail("embedding width must equal attention feature_size"); }
        const int Fs = desc->feature_size;
        n.wk = dense(Fs, Fs, false, false); n.wv = dense(Fs, Fs, false, false); n.wq = dense(Fs, Fs, false, false); n.wc = dense(Fs, Fs, false, false);
        k = Fs;
        for (int l = 0; l < desc->out_layers; ++l) { n.out[n.n_out++] = dense(k, desc->out_hidden[l], true, true); k = desc->out_hidden[l]; }
        n.out[n.n_out++] = dense(k, desc->n_actions, true, false);
        if (width > 64) { delete q; return qfail("EgoAttention layer widths must be <= 64"); }
        if (desc->heads * desc->n_entities > width) { delete q; return qfail("heads * n_entities must be <= layer width"); }
        n.width = width;
        n.scratch_per_warp = ((nin + 3) & ~3) + 3 * kMaxRows * width;
    } else if (desc->type == TTRL_QNET_MLP) {
        int k = nin;
        for (int l = 0; l < desc->n_hidden; ++l) { n.out[n.n_out++] = dense(k, desc->hidden[l], true, true); k = desc->hidden[l]; }
        n.out[n.n_out++] = dense(k, desc->n_actions, true, false);
        n.width = width;
        n.scratch_per_warp = 8 * ((nin + 3) & ~3) + 3 * 8 * width;
    } else if (desc->type == TTRL_QNET_DUELING) {
        int k = nin;
        for (int l = 0; l < desc->n_hidden; ++l) { n.base[n.n_base++] = dense(k, desc->hidden[l], true, true); k = desc->hidden[l]; }
        n.val[n.n_val++] = dense(k, 1, true, false);
        n.adv[n.n_adv++] = dense(k, desc->n_actions, true, false);
        n.width = width;
        n.scratch_per_warp = 8 * ((nin + 3) & ~3) + 3 * 8 * width;
    } else { delete q; return qfail("Unknown model type"); }
    if (width > 128) { delete q; return qfail("layer widths must be <= 128"); }
    if (off != n_weights) { delete q; return qfail("weight blob size does not match the network description"); }
    n.n_weights = off;
    q->smem_bytes = (int)(sizeof(float) * (((off + 3) & ~3) + kWarps * n.scratch_per_warp));
    if (q->smem_bytes > 227 * 1024) { delete q; return qfail("Q-network does not fit in shared memory"); }
    QCK(cudaMalloc(&q->d_weights, sizeof(float) * off));
    QCK(cudaMemcpy(q->d_weights, weights_host, sizeof(float) * off, cudaMemcpyHostToDevice));
    QCK(cudaFuncSetAttribute(k_qnet_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, q->smem_bytes));
    cudaDeviceProp prop;
    QCK(cudaGetDeviceProperties(&prop, device));
    q->n_sms = prop.multiProcessorCount;
    // tensor-core path: MLP with two hidden layers whose widths are UMMA N sizes (multiples of 16, <= 256)
    if (desc->type == TTRL_QNET_MLP && desc->n_hidden == 2 && desc->hidden[0] % 16 == 0 && desc->hidden[1] % 16 == 0 &&
        desc->hidden[0] >= 16 && desc->hidden[1] >= 16 && desc->hidden[0] <= 256 && desc->hidden[1] <= 256) {
        TcMlp& t = q->tc;
        t.nin = nin; t.K1 = (nin + 15) / 16 * 16; t.H1 = desc->hidden[0]; t.H2 = desc->hidden[1]; t.A = desc->n_actions;
        t.w1 = n.out[0].w_off; t.b1 = n.out[0].b_off; t.w2 = n.out[1].w_off; t.b2 = n.out[1].b_off; t.w3 = n.out[2].w_off; t.b3 = n.out[2].b_off;
        auto up = [](int x) { return (x + 127) / 128 * 128; };
        int o = 0;
        t.off_w1hi = o; o += up(t.H1 * t.K1 * 2);
        t.off_w1lo = o; o += up(t.H1 * t.K1 * 2);
        t.off_w2hi = o; o += up(t.H2 * t.H1 * 2);
        t.off_w2lo = o; o += up(t.H2 * t.H1 * 2);
        t.off_f32 = o; o += up((int)sizeof(float) * (t.H1 + t.H2 + t.A * t.H2 + t.A));
        t.img_bytes = o;
        const int ka = t.K1 > t.H1 ? t.K1 : t.H1;
        t.off_ahi = o; o += up(128 * ka * 2);
        t.off_alo = o; o += up(128 * ka * 2);
        const int part_bytes = kMlpColGroups * 128 * ((t.A + 3) & ~3) * (int)sizeof(float);   // epilogue 2 reuses the A operand
        if (o - t.off_ahi < part_bytes) o = t.off_ahi + up(part_bytes);
        t.slot_bytes = kMlpSlotRows * nin * (int)sizeof(float);   // a multiple of 128
        t.off_raw = o;
        const int room = 227 * 1024 - 64 - o;
        t.n_slots = room < 2 * t.slot_bytes ? 0 : room / t.slot_bytes > kMlpMaxSlots ? kMlpMaxSlots : room / t.slot_bytes;
        o += t.n_slots * t.slot_bytes;
        t.off_bar = o; o += 64;                                   // MMA, weight-image and slot mbarriers, the TMEM address
        t.total = o;
        const int mx = t.H1 > t.H2 ? t.H1 : t.H2;
        t.tmem_cols = mx <= 32 ? 32 : mx <= 64 ? 64 : mx <= 128 ? 128 : 256;
        if (t.total <= 227 * 1024 &&
            cudaFuncSetAttribute(k_qnet_mlp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, t.total) == cudaSuccess &&
            cudaMalloc(&q->d_img, t.img_bytes) == cudaSuccess && cudaMemset(q->d_img, 0, t.img_bytes) == cudaSuccess)
            q->tc_ok = true;
        (void)cudaGetLastError();
    }
    // tensor-core path for the EgoAttentionNetwork: the shipped shape family (two 64-wide embedding layers,
    // feature_size 64, <= 4 heads, two 64-wide output layers, <= 15 entities x <= 16 features)
    if (desc->type == TTRL_QNET_EGO_ATTENTION && desc->embed_layers == 2 && desc->embed[0] == 64 && desc->embed[1] == 64 &&
        desc->feature_size == 64 && desc->heads >= 1 && desc->heads <= 4 && 64 % desc->heads == 0 && desc->out_layers == 2 &&
        desc->out_hidden[0] == 64 && desc->out_hidden[1] == 64 && desc->n_entities <= 15 && desc->n_features <= 16 && desc->n_actions <= 4) {
        TcEgo& t = q->tce;
        t.NE = desc->n_entities; t.Fe = desc->n_features; t.Fs = 64; t.H = desc->heads; t.A = desc->n_actions; t.pidx = desc->presence_feature_idx;
        t.ego_w1 = n.ego[0].w_off; t.ego_b1 = n.ego[0].b_off; t.ego_w2 = n.ego[1].w_off; t.ego_b2 = n.ego[1].b_off;
        t.oth_w1 = n.oth[0].w_off; t.oth_b1 = n.oth[0].b_off; t.oth_w2 = n.oth[1].w_off; t.oth_b2 = n.oth[1].b_off;
        t.wk = n.wk.w_off; t.wv = n.wv.w_off; t.wq = n.wq.w_off; t.wc = n.wc.w_off;
        t.o_w1 = n.out[0].w_off; t.o_b1 = n.out[0].b_off; t.o_w2 = n.out[1].w_off; t.o_b2 = n.out[1].b_off;
        t.p_w = n.out[2].w_off; t.p_b = n.out[2].b_off;
        auto up = [](int x) { return (x + 1023) / 1024 * 1024; };
        int o = 0;
        t.off_b1hi = o; o += up(128 * 16 * 2);
        t.off_b1lo = o; o += up(128 * 16 * 2);
        t.off_b2hi = o; o += up(128 * 64 * 2);
        t.off_b2lo = o; o += up(128 * 64 * 2);
        t.off_b3hi = o; o += up(192 * 64 * 2);
        t.off_b3lo = o; o += up(192 * 64 * 2);
        auto up16 = [](int x) { return (x + 15) / 16 * 16; };
        t.off_ahi = o; o += 2 * (2 * 128 * 64 * 2);   // per group: hi and lo contiguous, together the [128][64] fp32 scratch
        t.off_f32 = o; o += up16((int)sizeof(float) * (3 * 64 * 64 + 64 * 4 + 6 * 64 + 16));
        t.x_stride = up16((int)sizeof(float) * 8 * nin);
        t.off_x = o; o += 2 * t.x_stride;
        t.small_stride = (int)sizeof(float) * 3 * 8 * 64;
        t.off_small = o; o += 2 * t.small_stride;
        t.off_bar = o; o += 32;                        // two mbarriers + the TMEM address
        t.total = o;
        if (t.total <= 227 * 1024 &&
            cudaFuncSetAttribute(k_qnet_ego_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, t.total) == cudaSuccess)
            q->tc_ok = true;
        (void)cudaGetLastError();
    }
    *out = q;
    return 0;
}

int ttrl_qnet_set_mode(ttrl_qnet* q, int mode) {
    if (!q) return qfail("null argument");
    if (mode == TTRL_QNET_MODE_FP32) { q->mode = mode; return 0; }
    if (mode != TTRL_QNET_MODE_TENSOR) return qfail("unknown Q-network mode");
    if (!q->tc_ok) return qfail("the tensor-core path supports MultiLayerPerceptron with two hidden layers of width 16..256 (multiples of 16) and "
                                "EgoAttentionNetwork with 64-wide embedding / attention / output layers (the shipped ego_attention*.json shapes)");
    q->mode = mode;
    return 0;
}

int ttrl_qnet_set_weights(ttrl_qnet* q, const float* weights, int64_t n_weights, int on_device, void* stream) {
    if (!q || !weights) return qfail("null argument");
    if (n_weights != q->net.n_weights) return qfail("weight blob size does not match the network description");
    QCK(cudaSetDevice(q->device));
    QCK(cudaMemcpyAsync(q->d_weights, weights, sizeof(float) * n_weights, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                        (cudaStream_t)stream));
    if (!on_device) QCK(cudaStreamSynchronize((cudaStream_t)stream));  // the host buffer may be reused by the caller
    q->img_dirty = true;
    return 0;
}

float* ttrl_qnet_weights_dev(ttrl_qnet* q) {
    if (q) q->external_writer = true;   // the caller's kernels update the blob in place: derived images are rebuilt on every launch
    return q ? q->d_weights : nullptr;
}

int ttrl_qnet_destroy(ttrl_qnet* q) {
    if (!q) return 0;
    cudaSetDevice(q->device);
    cudaFree(q->d_weights);
    cudaFree(q->d_img);
    delete q;
    return 0;
}

static int qnet_launch(ttrl_qnet* q, const float* obs_dev, int E, double eps, uint64_t seed, uint64_t step, const double* u_dev,
                       int32_t* actions_dev, float* q_dev, void* stream) {
    QCK(cudaSetDevice(q->device));
    if (q->mode == TTRL_QNET_MODE_TENSOR && q->net.d.type == TTRL_QNET_EGO_ATTENTION) {
        int grid = (E + 15) / 16;   // two tiles of 8 observations in flight per CTA
        if (grid > q->n_sms) grid = q->n_sms;
        if (grid < 1) grid = 1;
        k_qnet_ego_tc<<<grid, 2 * kEgoGT, q->tce.total, (cudaStream_t)stream>>>(q->tce, q->d_weights, obs_dev, E, eps, seed, step, u_dev, actions_dev, q_dev);
        q->launches++;
        QCK(cudaGetLastError());
        return 0;
    }
    if (q->mode == TTRL_QNET_MODE_TENSOR) {
        int grid = (E + 127) / 128;
        if (grid > q->n_sms) grid = q->n_sms;
        if (grid < 1) grid = 1;
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        QCK(cudaStreamIsCapturing((cudaStream_t)stream, &cap));   // a captured launch is replayed after later weight updates
        if (q->img_dirty || q->external_writer || cap != cudaStreamCaptureStatusNone) {
            k_qnet_mlp_prep<<<32, 256, 0, (cudaStream_t)stream>>>(q->tc, q->d_weights, q->d_img);
            q->launches++;
            if (cap == cudaStreamCaptureStatusNone) q->img_dirty = false;
        }
        k_qnet_mlp_tc<<<grid, kMlpThreads, q->tc.total, (cudaStream_t)stream>>>(q->tc, q->d_img, obs_dev, E, eps, seed, step, u_dev, actions_dev, q_dev);
        q->launches++;
        QCK(cudaGetLastError());
        return 0;
    }
    const int per = q->net.d.type == TTRL_QNET_EGO_ATTENTION ? 1 : 8;
    const int items = (E + per - 1) / per;
    int grid = (items + kWarps - 1) / kWarps;
    if (grid > q->n_sms) grid = q->n_sms;
    if (grid < 1) grid = 1;
    k_qnet_fp32<<<grid, kWarps * 32, q->smem_bytes, (cudaStream_t)stream>>>(q->net, q->d_weights, obs_dev, E, eps, seed, step, u_dev, actions_dev, q_dev);
    q->launches++;
    QCK(cudaGetLastError());
    return 0;
}

int ttrl_qnet_act(ttrl_qnet* q, const float* obs_dev, int num_envs, double epsilon, uint64_t seed, uint64_t step,
                  int32_t* actions_dev, float* q_dev, void* stream) {
    return qnet_launch(q, obs_dev, num_envs, epsilon, seed, step, nullptr, actions_dev, q_dev, stream);
}
/* Same, with the exploration uniforms supplied by the caller (parity with np_random.choice: one U[0,1) per env). */
int ttrl_qnet_act_injected(ttrl_qnet* q, const float* obs_dev, int num_envs, double epsilon, const double* u_dev,
                           int32_t* actions_dev, float* q_dev, void* stream) {
    return qnet_launch(q, obs_dev, num_envs, epsilon, 0, 0, u_dev, actions_dev, q_dev, stream);
}
int64_t ttrl_qnet_launch_count(const ttrl_qnet* q) { return q->launches; }

}  // extern "C"
