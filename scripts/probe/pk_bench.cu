// Throughput of scalar FFMA vs packed FFMA2 / FADD2 on sm_100a (one warp-instruction = 32 or 64 fp32 FMAs).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, int iters, float2 s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, s.x, s.y); a[i].y = fmaf(a[i].y, s.x, s.y); }
            if (MODE == 1) a[i] = __ffma2_rn(a[i], s, make_float2(s.y, s.y));
            if (MODE == 2) { a[i].x += s.x; a[i].y += s.y; }
            if (MODE == 3) a[i] = __fadd2_rn(a[i], s);
        }
    }
    float2 r = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { r.x += a[i].x; r.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float2* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, grid = 148 * 8;
    k<MODE><<<grid, 256>>>(d, 64, make_float2(1.0001f, 1e-6f));
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, iters, make_float2(1.0001f, 1e-6f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fp = double(grid) * 256 * iters * 16.0;       // fp32 element-operations
    printf("%-8s %.3f ms  %.1f T element-ops/s (x2 flops for FMA)\n", name, ms, fp / ms / 1e9);
}
int main() {
    float2* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float2));
    run<0>("FFMA", d); run<1>("FFMA2", d); run<2>("FADD", d); run<3>("FADD2", d);
    return 0;
}
