// Register-only rate of the radix-32 butterfly + 31 twiddle products (what every pass of the 1024-point plan
// executes between two exchanges), as a function of the warps per SM: tells whether the propagation passes are
// bound by the FP32 pipe (rate independent of the warp count) or by latency (rate grows with the warp count).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../binary_hologram_reinforcement_learning_b200/csrc/bh_fft.cuh"
using namespace bh;

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(float2* out, int iters, float2 seed) {
    float2 v[32], w[4];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        v[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f + seed.x);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = make_float2(1.f - 1e-6f * (i + threadIdx.x), 1e-3f * i + seed.y);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 1; r < 32; ++r) v[r] = cmul(v[r], w[r & 3]);
        dft<32, false>(v);
    }
    float2 r = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 32; ++i) { r.x += v[i].x; r.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int THREADS, int MINB> void run(float2* d, int ctas_per_sm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, grid = 148 * ctas_per_sm;
    k<THREADS, MINB><<<grid, THREADS>>>(d, 16, make_float2(0.f, 0.f));
    cudaEventRecord(e0);
    k<THREADS, MINB><<<grid, THREADS>>>(d, iters, make_float2(0.f, 0.f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int warps_per_sm = THREADS / 32 * ctas_per_sm;
    const double dfts_per_sm = double(warps_per_sm) * iters;          // warp-level butterflies
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("threads %4d x %d CTAs/SM = %2d warps/SM: %.3f ms, %.0f cycles per warp-butterfly and scheduler (at %d MHz nominal), "
           "%.1f G butterflies(32 pt, per thread)/s\n", THREADS, ctas_per_sm, warps_per_sm, ms,
           cycles / (dfts_per_sm / 4.0), clk / 1000, 148.0 * dfts_per_sm * 32 / ms / 1e6);
}
int main() {
    float2* d; cudaMalloc(&d, 148 * 16 * 1024 * sizeof(float2));
    run<128, 1>(d, 1); run<256, 1>(d, 1); run<256, 2>(d, 2); run<512, 1>(d, 1); run<256, 3>(d, 3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
