// Probe: CTA barriers reached by the two 16-lane halves of a warp (a) at the same instruction at different times,
// (b) from the two branches of an if / else (different instructions), (c) like (b) with a full-warp __syncwarp first.
// Prints "ok" or hangs (run under `timeout`).   nvcc -gencode arch=compute_100a,code=sm_100a -o bar_divergence bar_divergence.cu
#include <cstdio>
#include <cstdlib>
__device__ __noinline__ void spin(int n, volatile int* sink) { for (int k = 0; k < n; ++k) *sink += k; }
__global__ void k(int mode, int iters, int* out) {
    __shared__ int sink[512];
    const int half = (threadIdx.x >> 4) & 1;
    const unsigned hmask = 0xFFFFu << (16 * half);
    int acc = 0;
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
            if (half == (it & 1)) spin(200, sink + threadIdx.x);  // team-divergent work
            __syncwarp(hmask);
            __syncthreads();                                      // same instruction, the halves arrive apart
            acc += it;
        }
    } else {
        if (half == 0) {
            for (int it = 0; it < iters; ++it) { spin(100, sink + threadIdx.x); __syncwarp(hmask); if (mode == 2) __syncwarp(); __syncthreads(); acc += it; }
        } else {
            for (int it = 0; it < iters; ++it) { if (mode == 2) __syncwarp(); __syncthreads(); acc -= it; }  // "idle team": barriers only
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    int* out;
    cudaMalloc(&out, 4 * 148 * 512);
    k<<<148, 512>>>(mode, 1000, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
    return 0;
}
