// Read-bandwidth probe for the delta-eval access pattern (diagnostic, not part of the library).
//   A  one stream of float4 loads over a 768 MB buffer
//   B  the k_eval streams without the impulse response: per candidate U (8 B/px) + I + T (4 B/px)
//      of 8 different environments, 1024-pixel units, balanced partition over a one-wave grid
//   C  B plus four 8-byte taps per quad from a 25 MB table kept in L2 (evict_last)
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/membw_probe scripts/membw_probe.cu && /tmp/membw_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pol_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t pol_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ float4 ld4(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ld2(const float2* p, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}

template <int UF>
__global__ void __launch_bounds__(256, 2) k_stream(const float4* __restrict__ a, size_t n4, float* out) {
    const uint64_t pf = pol_first();
    const size_t per = n4 / gridDim.x;
    const float4* p = a + per * blockIdx.x;
    float s = 0.f;
    for (size_t i = threadIdx.x; i + (UF - 1) * 256 < per; i += UF * 256) {
        float4 v[UF];
#pragma unroll
        for (int k = 0; k < UF; ++k) v[k] = ld4(p + i + k * 256, pf);
#pragma unroll
        for (int k = 0; k < UF; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    if (s == 12345.678f) *out = s;
}

// candidates: U_k, I_k, T_k of 1024x1024; units of 1024 px; TAPS adds the L2-resident table reads
template <int UF, bool TAPS>
__global__ void __launch_bounds__(256, 2)
k_pattern(const float2* __restrict__ U, const float* __restrict__ I, const float* __restrict__ T,
          const float2* __restrict__ h, int n_cand, float* out) {
    const int N = 1024, HP = 1028;
    const uint64_t pf = pol_first(), pl = pol_last();
    const long long total = (long long)n_cand * 1024;
    const long long beg = blockIdx.x * total / gridDim.x, end = (blockIdx.x + 1) * total / gridDim.x;
    float s = 0.f;
    long long u = beg;
    while (u < end) {
        const int k = int(u >> 10);
        const long long seg_end = ((long long)(k + 1) << 10) < end ? ((long long)(k + 1) << 10) : end;
        const size_t base = (size_t)k * N * N;
        const int r = (k * 397 + 11) & 1023, c = (k * 613 + 5) & 1023;
        int y = int(u & 1023);
        const int x = threadIdx.x * 4;
        int hx = x - c; if (hx < 0) hx += N;
        int hy = y - r; if (hy < 0) hy += N;
        int ho = hy * HP + hx;
        size_t p = base + (size_t)y * N + x;
        long long v = u;
        for (; v + UF <= seg_end; v += UF) {
            float4 ua[UF], ub[UF], iv[UF], tv[UF]; float2 hq[UF][4];
#pragma unroll
            for (int q = 0; q < UF; ++q) {
                const float4* Up = reinterpret_cast<const float4*>(U + p);
                ua[q] = ld4(Up, pf); ub[q] = ld4(Up + 1, pf);
                iv[q] = ld4(reinterpret_cast<const float4*>(I + p), pf);
                tv[q] = ld4(reinterpret_cast<const float4*>(T + p), pf);
                p += N;
                if (TAPS) {
                    const float2* hp = h + ho;
                    hq[q][0] = ld2(hp, pl); hq[q][1] = ld2(hp + 1, pl); hq[q][2] = ld2(hp + 2, pl); hq[q][3] = ld2(hp + 3, pl);
                    ho += HP; if (ho >= N * HP) ho -= N * HP;
                }
            }
#pragma unroll
            for (int q = 0; q < UF; ++q) {
                s += ua[q].x * iv[q].x + ua[q].y * tv[q].x + ub[q].z * iv[q].w + ub[q].w * tv[q].w;
                if (TAPS) s += hq[q][0].x * hq[q][1].y + hq[q][2].x * hq[q][3].y;
            }
        }
        u = seg_end;
    }
    if (s == 12345.678f) *out = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch(i);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch(i);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const size_t N2 = 1024 * 1024;
    const int n_env = 8, F = 24, G = 3;
    float2 *U, *h; float *I, *T, *out;
    cudaMalloc(&U, n_env * F * N2 * sizeof(float2));         // 1.6 GB: consecutive launches stream new frames
    const int planes = 2 * n_env * G;                          // 48 planes each: no L2 reuse between launches
    cudaMalloc(&I, planes * N2 * sizeof(float));
    cudaMalloc(&T, planes * N2 * sizeof(float));
    cudaMalloc(&h, 3 * 1024 * 1028 * sizeof(float2));
    cudaMalloc(&out, 4);
    cudaMemset(U, 0, n_env * F * N2 * sizeof(float2)); cudaMemset(I, 0, planes * N2 * sizeof(float));
    cudaMemset(T, 0, planes * N2 * sizeof(float)); cudaMemset(h, 0, 3 * 1024 * 1028 * sizeof(float2));
    const int grid = 296;
    {
        const size_t bytes = 768ull << 20;
        float ms = time_ms([&](int) { k_stream<4><<<grid, 256>>>(reinterpret_cast<const float4*>(U), bytes / 16, out); }, 20);
        printf("A  one float4 stream, 768 MB, UF=4, grid 296:      %.1f GB/s\n", bytes / ms / 1e6);
        ms = time_ms([&](int) { k_stream<8><<<grid, 256>>>(reinterpret_cast<const float4*>(U), bytes / 16, out); }, 20);
        printf("A  one float4 stream, 768 MB, UF=8, grid 296:      %.1f GB/s\n", bytes / ms / 1e6);
        ms = time_ms([&](int) { k_stream<4><<<grid * 8, 256>>>(reinterpret_cast<const float4*>(U), bytes / 16, out); }, 20);
        printf("A  one float4 stream, 768 MB, UF=4, grid 2368:     %.1f GB/s\n", bytes / ms / 1e6);
    }
    // B / C: candidate k of launch i reads frame (i*8 + k) % 184 of U and plane (i*8 + k) % 40 of I / T
    auto pat = [&](auto kern, const char* name) {
        float ms = time_ms([&](int i) {
            const int f0 = (i * 8) % (n_env * F - 8), g0 = (i * 8) % (planes - 8);
            kern<<<grid, 256>>>(U + (size_t)f0 * N2, I + (size_t)g0 * N2, T + (size_t)g0 * N2, h, 8, out);
        }, 200);
        printf("%s %.2f us / 8 candidates = %.1f GB/s of U+I+T\n", name, ms * 1e3, 8 * 16.0 * N2 / ms / 1e6);
    };
    pat(k_pattern<3, false>, "B  U+I+T streams, UF=3:               ");
    pat(k_pattern<4, false>, "B  U+I+T streams, UF=4:               ");
    pat(k_pattern<3, true>,  "C  U+I+T streams + L2 taps, UF=3:     ");
    pat(k_pattern<4, true>,  "C  U+I+T streams + L2 taps, UF=4:     ");
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
