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

}  // namespace

struct ttrl_qnet {
    QnetDev net;
    float* d_weights = nullptr;
    int device = 0;
    int smem_bytes = 0;
    int n_sms = 0;
    int64_t launches = 0;
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
    *out = q;
    return 0;
}

int ttrl_qnet_destroy(ttrl_qnet* q) {
    if (!q) return 0;
    cudaSetDevice(q->device);
    cudaFree(q->d_weights);
    delete q;
    return 0;
}

static int qnet_launch(ttrl_qnet* q, const float* obs_dev, int E, double eps, uint64_t seed, uint64_t step, const double* u_dev,
                       int32_t* actions_dev, float* q_dev, void* stream) {
    QCK(cudaSetDevice(q->device));
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
