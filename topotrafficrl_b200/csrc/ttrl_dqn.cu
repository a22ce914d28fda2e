// ttrl_dqn.cu -- the DQN update of the batched training driver as two kernels.
//
// Replaces, for the MultiLayerPerceptron model family (scripts/configs/*/agents/DQNAgent/baseline.json: [128, 128]):
//   DQNAgent.compute_bellman_residual (ttrl_agent/agents/deep_q_network/pytorch.py:41-73): the three forward passes
//     (value_net(s), value_net(s'), target_net(s')), the double-DQN target, the loss;
//   DQNAgent.step_optimizer (pytorch.py:32-39): backward, gradient clamp to [-1, 1], Adam step.
//
//   k_dqn_grad   ONE CTA of 512 threads: the minibatch is gathered from the replay memory into shared memory, every
//                activation of the batch stays there (X, H1, H2, dH2, dH1 as [64][129] fp32 tiles + one staged weight
//                matrix: 198 KB), each GEMM is a 4 x 4 (forward, dH1) or 8 x 4 (weight gradients) register tile per thread
//                over operands read from shared memory.  Output: the flat gradient (torch parameter order) and the loss.
//                The whole update is ~9 M multiply-adds: a single SM does it in tens of microseconds, and one CTA makes the
//                summation order -- hence the loss curve -- deterministic.
//   k_dqn_adam   elementwise over the 30 K parameters: [mean over ranks,] clamp, Adam (torch.optim.Adam's arithmetic), and the
//                refreshed weight blob of the rollout kernels (ttrl_qnet.cu layout: W^T then bias per layer) in the same pass.
// Between the two the host may all-reduce the flat gradient (NCCL) -- the only data-path collective of the system.
//
// fp32 throughout (the reference trains in fp32); the batch GEMMs have M = 64 rows, below the 128-row tcgen05 tile, and
// are latency- not throughput-bound: CUDA cores.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string>

#include "../../include/ttrl_b200.h"

namespace {

constexpr int DB = 64;    // batch rows per pass
constexpr int DK = 128;   // widest layer
constexpr int DS = 129;   // shared-memory row stride (floats): conflict-free columns
constexpr int NT = 512;

struct Net { const float* W[3]; const float* b[3]; };
struct GradArgs {
    Net val, tgt;
    const float* state; const float* next_state; const int64_t* action; const float* reward; const uint8_t* terminal; const int64_t* idx;
    float* grad; float* loss;
    int n_in, h1, h2, na, batch, loss_type, double_q;
    float gamma;
};

// out[b][j] = act(sum_k in[b][k] W[j][k] + bias[j]) for the DB rows; W [J][K] row-major in global memory (torch layout)
template <bool RELU>
__device__ void dense(const float* __restrict__ W, const float* __restrict__ bias, int J, int K, const float* in, float* out, float* Ws) {
    const int tid = threadIdx.x;
    for (int i = tid; i < J * K; i += NT) { const int j = i / K, k = i - j * K; Ws[k * DS + j] = W[i]; }  // staged transposed: Ws[k][j]
    __syncthreads();
    const int bt = tid >> 5, jt = tid & 31;
    float acc[4][4] = {};
    for (int k = 0; k < K; ++k) {
        float x[4], w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = in[(bt + 16 * i) * DS + k];
#pragma unroll
        for (int m = 0; m < 4; ++m) w[m] = (jt + 32 * m < J) ? Ws[k * DS + jt + 32 * m] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[i][m] = fmaf(x[i], w[m], acc[i][m]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int j = jt + 32 * m;
            float v = 0.f;
            if (j < J) { v = acc[i][m] + bias[j]; if (RELU) v = fmaxf(v, 0.f); }
            out[(bt + 16 * i) * DS + j] = v;
        }
    __syncthreads();
}

// q[b][a] = H2[b] . W3[a] + b3[a]
__device__ void head(const float* __restrict__ W3, const float* __restrict__ b3, int na, int K, const float* H2, float* q) {
    for (int t = threadIdx.x; t < DB * na; t += NT) {
        const int b = t / na, a = t - b * na;
        float s = 0.f;
        for (int k = 0; k < K; ++k) s = fmaf(H2[b * DS + k], W3[a * K + k], s);
        q[b * 8 + a] = s + b3[a];
    }
    __syncthreads();
}

__device__ void gather(const float* __restrict__ src, const int64_t* __restrict__ idx, int row0, int rows, int n_in, float* X) {
    for (int t = threadIdx.x; t < DB * DS; t += NT) {
        const int b = t / DS, k = t - b * DS;
        X[t] = (b < rows && k < n_in) ? src[(size_t)idx[row0 + b] * n_in + k] : 0.f;
    }
    __syncthreads();
}

// gW[j][k] (+)= sum_b G[b][j] * A[b][k];  gb[j] (+)= sum_b G[b][j]
__device__ void weight_grad(const float* G, const float* A, int J, int K, float* gW, float* gb, bool accumulate) {
    const int tid = threadIdx.x, jt = tid >> 5, kt = tid & 31;
    float acc[8][4] = {};
    for (int b = 0; b < DB; ++b) {
        float g[8], h[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = G[b * DS + jt + 16 * i];
#pragma unroll
        for (int m = 0; m < 4; ++m) h[m] = A[b * DS + kt + 32 * m];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[i][m] = fmaf(g[i], h[m], acc[i][m]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int j = jt + 16 * i, k = kt + 32 * m;
            if (j < J && k < K) gW[j * K + k] = accumulate ? gW[j * K + k] + acc[i][m] : acc[i][m];
        }
    for (int j = tid; j < J; j += NT) {
        float s = 0.f;
        for (int b = 0; b < DB; ++b) s += G[b * DS + j];
        gb[j] = accumulate ? gb[j] + s : s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT, 1) k_dqn_grad(GradArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* X = sm;                 // [DB][DS] states (then next states, then states again)
    float* H1 = X + DB * DS;
    float* H2 = H1 + DB * DS;      // later: dH1
    float* G2 = H2 + DB * DS;      // dH2
    float* Ws = G2 + DB * DS;      // [DK][DS] staged weights
    float* q = Ws + DK * DS;       // [DB][8] value_net(s)
    float* qn = q + DB * 8;        // value_net(s')
    float* qt = qn + DB * 8;       // target_net(s')
    float* delta = qt + DB * 8;    // dLoss/dq[b][a_b]
    int* act = reinterpret_cast<int*>(delta + DB);
    const int tid = threadIdx.x;
    const int n_in = a.n_in, h1 = a.h1, h2 = a.h2, na = a.na;
    const int oW1 = 0, ob1 = oW1 + h1 * n_in, oW2 = ob1 + h1, ob2 = oW2 + h2 * h1, oW3 = ob2 + h2, ob3 = oW3 + na * h2;
    float loss_sum = 0.f;  // thread 0 only
    for (int i = tid; i < DK * DS; i += NT) Ws[i] = 0.f;
    __syncthreads();
    for (int row0 = 0; row0 < a.batch; row0 += DB) {
        const int rows = a.batch - row0 < DB ? a.batch - row0 : DB;
        const bool accumulate = row0 > 0;
        // ---- forward passes on the next states: target_net and (double DQN) value_net
        gather(a.next_state, a.idx, row0, rows, n_in, X);
        dense<true>(a.tgt.W[0], a.tgt.b[0], h1, n_in, X, H1, Ws);
        dense<true>(a.tgt.W[1], a.tgt.b[1], h2, h1, H1, H2, Ws);
        head(a.tgt.W[2], a.tgt.b[2], na, h2, H2, qt);
        if (a.double_q) {
            dense<true>(a.val.W[0], a.val.b[0], h1, n_in, X, H1, Ws);
            dense<true>(a.val.W[1], a.val.b[1], h2, h1, H1, H2, Ws);
            head(a.val.W[2], a.val.b[2], na, h2, H2, qn);
        }
        // ---- forward pass on the states (activations kept for the backward pass)
        gather(a.state, a.idx, row0, rows, n_in, X);
        dense<true>(a.val.W[0], a.val.b[0], h1, n_in, X, H1, Ws);
        dense<true>(a.val.W[1], a.val.b[1], h2, h1, H1, H2, Ws);
        head(a.val.W[2], a.val.b[2], na, h2, H2, q);
        // ---- Bellman residual per row (pytorch.py:41-73): target = r + gamma * [not terminal] * Q_target(s', a*)
        if (tid < DB) {
            const int b = tid;
            float d = 0.f;
            int ab = 0;
            if (b < rows) {
                const int64_t m = a.idx[row0 + b];
                ab = (int)a.action[m];
                const float* sel = a.double_q ? qn + b * 8 : qt + b * 8;
                int best = 0;
                for (int k = 1; k < na; ++k) if (sel[k] > sel[best]) best = k;  // torch.max: the first maximum
                const float next_v = a.terminal[m] ? 0.f : qt[b * 8 + best];
                const float target = a.reward[m] + a.gamma * next_v;
                const float e = q[b * 8 + ab] - target;
                float l;
                if (a.loss_type == 0) { l = e * e; d = 2.f * e; }                                 // F.mse_loss
                else if (a.loss_type == 1) { l = fabsf(e); d = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f); }  // F.l1_loss
                else { const float ae = fabsf(e); l = ae < 1.f ? 0.5f * e * e : ae - 0.5f; d = ae < 1.f ? e : (e > 0.f ? 1.f : -1.f); }  // F.smooth_l1_loss
                d /= (float)a.batch;  // reduction = mean
                delta[b] = d;
                q[b * 8 + 7] = l;
            } else { delta[b] = 0.f; q[b * 8 + 7] = 0.f; }
            act[b] = ab;
        }
        __syncthreads();
        if (tid == 0) for (int b = 0; b < rows; ++b) loss_sum += q[b * 8 + 7];
        // ---- backward: predict layer
        for (int t = tid; t < na * h2; t += NT) {  // dW3[a][k] = sum_b [a_b == a] delta_b H2[b][k]
            const int aa = t / h2, k = t - aa * h2;
            float s = 0.f;
            for (int b = 0; b < DB; ++b) if (act[b] == aa) s = fmaf(delta[b], H2[b * DS + k], s);
            a.grad[oW3 + t] = accumulate ? a.grad[oW3 + t] + s : s;
        }
        if (tid < na) {
            float s = 0.f;
            for (int b = 0; b < DB; ++b) if (act[b] == tid) s += delta[b];
            a.grad[ob3 + tid] = accumulate ? a.grad[ob3 + tid] + s : s;
        }
        for (int t = tid; t < DB * DS; t += NT) {  // dH2 = delta_b W3[a_b] masked by relu'
            const int b = t / DS, k = t - b * DS;
            G2[t] = (k < h2 && H2[t] > 0.f) ? delta[b] * a.val.W[2][act[b] * h2 + k] : 0.f;
        }
        __syncthreads();
        // ---- dH1[b][k] = relu'(H1) sum_j dH2[b][j] W2[j][k]  (into the H2 tile, dead from here on)
        for (int i = tid; i < h2 * h1; i += NT) { const int j = i / h1, k = i - j * h1; Ws[j * DS + k] = a.val.W[1][i]; }  // natural layout
        __syncthreads();
        {
            const int bt = tid >> 5, kt = tid & 31;
            float acc[4][4] = {};
            for (int j = 0; j < h2; ++j) {
                float g[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) g[i] = G2[(bt + 16 * i) * DS + j];
#pragma unroll
                for (int m = 0; m < 4; ++m) w[m] = Ws[j * DS + kt + 32 * m];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int m = 0; m < 4; ++m) acc[i][m] = fmaf(g[i], w[m], acc[i][m]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int o = (bt + 16 * i) * DS + kt + 32 * m;
                    H2[o] = (kt + 32 * m < h1 && H1[o] > 0.f) ? acc[i][m] : 0.f;
                }
        }
        __syncthreads();
        weight_grad(G2, H1, h2, h1, a.grad + oW2, a.grad + ob2, accumulate);
        weight_grad(H2, X, h1, n_in, a.grad + oW1, a.grad + ob1, accumulate);
    }
    if (tid == 0) *a.loss = loss_sum / (float)a.batch;
}

struct AdamTensor { float* p; float* m; float* v; int n, goff, rows, cols, boff; };
struct AdamArgs {
    AdamTensor t[6];
    const float* grad;
    float* blob;  // rollout weight blob (or null)
    int total;
    float grad_scale, clamp, step_size, inv_sqrt_bc2, beta1, beta2, eps, weight_decay;
};

// torch.optim.Adam (single tensor form): exp_avg.lerp_(g, 1 - b1); exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2);
// denom = sqrt(exp_avg_sq) / sqrt(bias_correction2) + eps; param.addcdiv_(exp_avg, denom, -lr / bias_correction1)
__global__ void k_dqn_adam(AdamArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.total) return;
    int k = 0;
    while (k < 5 && i >= a.t[k + 1].goff) ++k;
    const AdamTensor& t = a.t[k];
    const int e = i - t.goff;
    float g = a.grad[i] * a.grad_scale;
    g = fminf(fmaxf(g, -a.clamp), a.clamp);  // p.grad.data.clamp_(-1, 1) (pytorch.py:36-37)
    float p = t.p[e];
    if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
    float m = t.m[e], v = t.v[e];
    m = m + (g - m) * (1.f - a.beta1);
    v = v * a.beta2 + (1.f - a.beta2) * g * g;
    const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
    p = p - a.step_size * (m / denom);
    t.p[e] = p; t.m[e] = m; t.v[e] = v;
    if (a.blob) {  // rollout blob: weights transposed ([in][out]), then the bias
        if (t.cols > 1) { const int j = e / t.cols, c = e - j * t.cols; a.blob[t.boff + c * t.rows + j] = p; }
        else a.blob[t.boff + e] = p;
    }
}

thread_local std::string g_dqn_err;

}  // namespace

struct ttrl_dqn {
    ttrl_dqn_desc d;
    int device;
    int n_params;
    size_t smem;
    int64_t launches;
};

extern "C" {

#define DQN_FAIL(msg) do { ttrl_set_error(msg); return 1; } while (0)
#define DQN_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ttrl_set_error((std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); return 1; } } while (0)

int ttrl_dqn_create(const ttrl_dqn_desc* desc, int device, ttrl_dqn** out) {
    if (!desc || !out) DQN_FAIL("null argument");
    if (desc->n_in < 1 || desc->n_in > DK || desc->h1 < 1 || desc->h1 > DK || desc->h2 < 1 || desc->h2 > DK || desc->n_actions < 1 || desc->n_actions > 7)
        DQN_FAIL("ttrl_dqn: layer widths must be in 1..128 and n_actions in 1..7 (MultiLayerPerceptron with two hidden layers)");
    if (desc->batch < 1 || desc->loss < 0 || desc->loss > 2) DQN_FAIL("ttrl_dqn: bad batch size or loss function");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) DQN_FAIL("no CUDA device: the DQN update kernels have no CPU fallback");
    if (device < 0 || device >= ndev) DQN_FAIL("bad device index");
    DQN_CK(cudaSetDevice(device));
    ttrl_dqn* q = new ttrl_dqn();
    q->d = *desc;
    q->device = device;
    q->n_params = desc->h1 * desc->n_in + desc->h1 + desc->h2 * desc->h1 + desc->h2 + desc->n_actions * desc->h2 + desc->n_actions;
    q->smem = sizeof(float) * (4 * DB * DS + DK * DS + 3 * DB * 8 + DB) + sizeof(int) * DB;
    q->launches = 0;
    cudaError_t e = cudaFuncSetAttribute(k_dqn_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q->smem);
    if (e != cudaSuccess) { delete q; ttrl_set_error((std::string("k_dqn_grad shared memory opt-in: ") + cudaGetErrorString(e)).c_str()); return 1; }
    *out = q;
    return 0;
}
int ttrl_dqn_destroy(ttrl_dqn* q) { delete q; return 0; }
int64_t ttrl_dqn_num_params(const ttrl_dqn* q) { return q->n_params; }
int64_t ttrl_dqn_launch_count(const ttrl_dqn* q) { return q->launches; }

int ttrl_dqn_grad(ttrl_dqn* q, const float* const* value_params, const float* const* target_params, const float* state_dev,
                  const float* next_state_dev, const int64_t* action_dev, const float* reward_dev, const uint8_t* terminal_dev,
                  const int64_t* idx_dev, float* grad_dev, float* loss_dev, void* stream) {
    if (!q || !value_params || !target_params || !grad_dev || !loss_dev) DQN_FAIL("null argument");
    DQN_CK(cudaSetDevice(q->device));
    GradArgs a{};
    for (int k = 0; k < 3; ++k) {
        a.val.W[k] = value_params[2 * k]; a.val.b[k] = value_params[2 * k + 1];
        a.tgt.W[k] = target_params[2 * k]; a.tgt.b[k] = target_params[2 * k + 1];
    }
    a.state = state_dev; a.next_state = next_state_dev; a.action = action_dev; a.reward = reward_dev; a.terminal = terminal_dev; a.idx = idx_dev;
    a.grad = grad_dev; a.loss = loss_dev;
    a.n_in = q->d.n_in; a.h1 = q->d.h1; a.h2 = q->d.h2; a.na = q->d.n_actions; a.batch = q->d.batch;
    a.loss_type = q->d.loss; a.double_q = q->d.double_q; a.gamma = q->d.gamma;
    k_dqn_grad<<<1, NT, q->smem, (cudaStream_t)stream>>>(a);
    q->launches++;
    DQN_CK(cudaGetLastError());
    return 0;
}

int ttrl_dqn_adam(ttrl_dqn* q, float* const* params, float* const* exp_avg, float* const* exp_avg_sq, const float* grad_dev, int64_t step,
                  double lr, double beta1, double beta2, double eps, double weight_decay, double grad_clamp, double grad_scale,
                  float* rollout_blob_dev, void* stream) {
    if (!q || !params || !exp_avg || !exp_avg_sq || !grad_dev) DQN_FAIL("null argument");
    if (step < 1) DQN_FAIL("ttrl_dqn_adam: step counts from 1");
    DQN_CK(cudaSetDevice(q->device));
    const ttrl_dqn_desc& d = q->d;
    const int rows[6] = {d.h1, d.h1, d.h2, d.h2, d.n_actions, d.n_actions};
    const int cols[6] = {d.n_in, 1, d.h1, 1, d.h2, 1};
    AdamArgs a{};
    int off = 0;
    for (int k = 0; k < 6; ++k) {
        a.t[k] = AdamTensor{params[k], exp_avg[k], exp_avg_sq[k], rows[k] * cols[k], off, rows[k], cols[k], off};  // blob order == parameter order
        off += rows[k] * cols[k];
    }
    a.grad = grad_dev; a.blob = rollout_blob_dev; a.total = off;
    a.grad_scale = (float)grad_scale; a.clamp = (float)grad_clamp;
    a.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)step)));
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.weight_decay = (float)weight_decay;
    k_dqn_adam<<<(off + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    q->launches++;
    DQN_CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
