// Stand-alone timing probe of the column pass (pass B) of the propagation: runs k2_cols on one colour group
// (Fg frames of P x P) the way launch_prop2_state does, with a per-phase clock64() breakdown of one CTA, and
// candidate variants next to it (checked against k2_cols).  Build + run: scripts/probe/run_probe.sh (GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../binary_hologram_reinforcement_learning_b200/csrc/bh_fft2.cuh"
#include "../../binary_hologram_reinforcement_learning_b200/csrc/bh_tables.hpp"
#include "cols_variants.cuh"

using namespace bh;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn encoder() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    return reinterpret_cast<encode_tiled_fn>(p);
}
static void make_map(CUtensorMap* map, const float2* base, int P, int planes, int colw, int rows_per_box) {
    const cuuint64_t dims[3] = {cuuint64_t(2) * P, cuuint64_t(P), cuuint64_t(planes)};
    const cuuint64_t strides[2] = {cuuint64_t(P) * sizeof(float2), cuuint64_t(P) * P * sizeof(float2)};
    const cuuint32_t box[3] = {cuuint32_t(2 * colw), cuuint32_t(rows_per_box), 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = encoder()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float2*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", int(r)); exit(1); }
}

int main(int argc, char** argv) {
    constexpr int P = 1024, Fg = 8;
    const int reps = argc > 1 ? atoi(argv[1]) : 20;
    int sms = 148; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t n2 = size_t(P) * P;
    std::vector<float2> h_in(n2 * Fg), h_H(n2);
    srand(1);
    for (auto& v : h_in) v = make_float2(rand() / float(RAND_MAX) - 0.5f, rand() / float(RAND_MAX) - 0.5f);
    auto tab = build_tables(P, 515e-9, 7.56e-6, 2e-3, 0);
    memcpy(h_H.data(), tab->H.data(), n2 * sizeof(float2));
    std::vector<float> tw = build_twiddles(P);
    float2 *d_src, *d_buf, *d_ref, *d_H, *d_tw;
    CK(cudaMalloc(&d_src, n2 * Fg * sizeof(float2))); CK(cudaMalloc(&d_buf, n2 * Fg * sizeof(float2)));
    CK(cudaMalloc(&d_ref, n2 * Fg * sizeof(float2))); CK(cudaMalloc(&d_H, n2 * sizeof(float2)));
    CK(cudaMalloc(&d_tw, tw.size() * sizeof(float) + 16));
    CK(cudaMemcpy(d_src, h_in.data(), n2 * Fg * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_H, h_H.data(), n2 * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice));
    long long* d_clk; CK(cudaMalloc(&d_clk, 64 * sizeof(long long)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    cudaStream_t st; CK(cudaStreamCreate(&st));

    auto run = [&](const char* name, auto launch, bool check) {
        float best = 1e9f, sum = 0.f;
        for (int i = 0; i < reps + 2; ++i) {
            CK(cudaMemcpyAsync(d_buf, d_src, n2 * Fg * sizeof(float2), cudaMemcpyDeviceToDevice, st));
            CK(cudaMemsetAsync(d_clk, 0, 64 * sizeof(long long), st));
            CK(cudaEventRecord(e0, st));
            launch();
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (i >= 2) { best = std::min(best, ms); sum += ms; }
        }
        double err = -1.0;
        if (check) {
            std::vector<float2> a(n2 * Fg), b(n2 * Fg);
            CK(cudaMemcpy(a.data(), d_buf, n2 * Fg * sizeof(float2), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(b.data(), d_ref, n2 * Fg * sizeof(float2), cudaMemcpyDeviceToHost));
            double mx = 0, ref = 0;
            for (size_t i = 0; i < a.size(); ++i) {
                mx = std::max(mx, double(std::max(fabsf(a[i].x - b[i].x), fabsf(a[i].y - b[i].y))));
                ref = std::max(ref, double(std::max(fabsf(b[i].x), fabsf(b[i].y))));
            }
            err = mx / ref;
        } else {
            CK(cudaMemcpy(d_ref, d_buf, n2 * Fg * sizeof(float2), cudaMemcpyDeviceToDevice));
        }
        long long clk[64]; CK(cudaMemcpy(clk, d_clk, sizeof clk, cudaMemcpyDeviceToHost));
        printf("%-34s best %.2f us  mean %.2f us  rel.err %.2e", name, best * 1e3, sum / reps * 1e3, err);
        if (clk[0]) { printf("  | CTA0 cycles/tile:"); for (int i = 1; i <= int(clk[0]) && i < 20; ++i) printf(" %lld", clk[i] / std::max(1ll, clk[32])); printf("  (tiles %lld)", clk[32]); }
        printf("\n");
    };

    // baseline: product kernel, W = 8, TMA input
    {
        CUtensorMap mb, mh; make_map(&mb, d_buf, P, Fg, 8, P / 4); make_map(&mh, d_H, P, 1, 8, P / 4);
        auto k = k2_cols<P, true, true, 8>;
        const size_t sm = cols2_smem_bytes<P, 8>();
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
        const int tiles = (P / 16 + 1) * Fg;
        run("k2_cols<W=8,TMA> (product)", [&]() { k<<<std::min(tiles, sms), 256, sm, st>>>(mb, mh, d_buf, d_tw, 1, Fg, 0); }, false);
    }
    run_variants(P, Fg, sms, d_buf, d_H, d_tw, d_clk, st, run, make_map);
    return 0;
}
