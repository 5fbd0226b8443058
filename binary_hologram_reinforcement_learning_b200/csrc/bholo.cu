// C ABI of the B200 hologram reward / DBS engine (see include/bholo.h).
// Host logic only: context, table upload, launch sequencing.  Kernels live in
// bh_kernels.cuh / bh_fft.cuh.  Built with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/bholo.h"
#include "bh_kernels.cuh"
#include "bh_fft2.cuh"
#include "bh_tables.hpp"

using namespace bh;

static_assert(sizeof(Result) == sizeof(bh_result), "Result must mirror bh_result");
static_assert(sizeof(bh_result) == 40, "bh_result layout");

static thread_local std::string g_err;

struct bh_ctx {
    int device = 0, E = 0, N = 0, F = 0, G = 0, Fg = 0, P = 0, pad = 1, relative = 1, method = 0;
    double dx = 0, z = 0;
    std::vector<double> wl;
    cudaStream_t stream = nullptr;
    size_t n2 = 0;
    float2 *dH = nullptr, *dh = nullptr, *dtw = nullptr, *dU = nullptr, *dscratch = nullptr;
    float *dI = nullptr, *dT = nullptr, *drecon = nullptr;
    int8_t* dstate = nullptr;
    double* dsums = nullptr;
    double* dloss_partial = nullptr;
    int max_tasks = 0, units_per_task = 0, grid_cap = 0, grid_cap_commit = 0;
    int grid_cap_bundle = 0, bundle = 0;     // bundle = slots per bundle of k_eval_bundle_t, 0 = off
    int32_t* d_envs = nullptr;
    long long* d_actions = nullptr;
    Result* d_results = nullptr;
    unsigned long long* d_acc = nullptr;
    unsigned* d_tickets = nullptr;
    long long* d_scalars = nullptr;      // [0] dbs cursor, [1] accepted count
    double* d_dbs_s0 = nullptr;          // [4] sums + PSNR a greedy-DBS window was scored against
    // correlation sweep (allocated on first use)
    float2 *dK3 = nullptr, *dK4 = nullptr, *dK5 = nullptr, *dK6 = nullptr, *dsw_in = nullptr, *dsw_out = nullptr;
    float2* dsw_buf = nullptr;           // pad = 2: P x P working planes of the correlation passes
    float *dsw_it = nullptr, *dsw_ii = nullptr;
    double* dsw_psnr = nullptr;
    std::vector<double> m4;
    int32_t* h_envs = nullptr;           // pinned staging
    long long* h_actions = nullptr;
    Result* h_results = nullptr;
    long long* h_scalars = nullptr;
    double* h_sums = nullptr;
    Result* h_results_dev = nullptr;     // device alias of the mapped h_results
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_eval = nullptr;
    cudaStream_t own_stream = nullptr;   // capture needs a real stream when the caller gave none
    void (*k_eval)(const DeltaArgs) = nullptr;
    void (*k_eval_bundle)(const DeltaArgs) = nullptr;
    void (*k_commit)(const DeltaArgs) = nullptr;
    bool use_pdl = true;
    int l2_prefetch = 4;                 // k_eval: units requested into L2 before the PDL wait (BHOLO_EVAL_L2PF; measured 0..8: 4 is best)
    bool fp64_eval = false;              // per-quad arithmetic of the delta evaluation in double (BHOLO_EVAL_FP64=1)
    bool fft2 = true;                    // register-resident passes (bh_fft2.cuh) where the side allows
    int fft2_mode = 1;                   // launch shape of those passes, see propagate_env
    int sms = 148;
    // observation path (bh_recon_batch)
    uint8_t* d_recon_stale = nullptr;    // [E][RECON_MAX_BUFFERS]
    ReconPlan* d_recon_plan = nullptr;   // [max_tasks]
    unsigned long long* d_ring = nullptr; // [E][ROLLOUT_RING][2] + error flag: barriers of the persistent rollout kernel
    int rollout_cap = 0;                 // co-resident CTAs of k_rollout_t (0: cooperative launch unavailable)
    float* d_recon_obs[RECON_MAX_BUFFERS] = {nullptr, nullptr, nullptr, nullptr};   // device observation blocks
    cudaEvent_t ev_recon = nullptr;
    // scratch of bh_sweep_stats, kept between calls
    double* d_stat_map = nullptr; float* d_stat_pre = nullptr; unsigned long long* d_stat_out = nullptr;
    int64_t launches = 0;
    std::string err;
};

#define BH_FAIL(ctx, code, ...)                                   \
    do {                                                          \
        char _b[512];                                             \
        snprintf(_b, sizeof _b, __VA_ARGS__);                     \
        if (ctx) (ctx)->err = _b;                                 \
        g_err = _b;                                               \
        return (code);                                            \
    } while (0)

#define BH_CUDA(ctx, expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            BH_FAIL(ctx, -2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define BH_CHECK_CTX(ctx)                       \
    do {                                        \
        if (!(ctx)) BH_FAIL((bh_ctx*)nullptr, -1, "null context"); \
        cudaError_t _e = cudaSetDevice((ctx)->device);             \
        if (_e != cudaSuccess) BH_FAIL(ctx, -2, "cudaSetDevice: %s", cudaGetErrorString(_e)); \
    } while (0)

#define BH_CHECK_ENV(ctx, env) \
    do { if ((env) < 0 || (env) >= (ctx)->E) BH_FAIL(ctx, -3, "env %d out of range [0,%d)", (env), (ctx)->E); } while (0)

extern "C" int bh_abi_version(void) { return 1; }

extern "C" const char* bh_last_error(const bh_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_err.c_str();
}

// ---------------------------------------------------------------------------
// register-resident passes (bh_fft2.cuh) for P = 896 / 1024, pad = 1
// ---------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn tensor_map_encoder() {
    static encode_tiled_fn fn = []() -> encode_tiled_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<encode_tiled_fn>(p);
    }();
    return fn;
}
// [planes][P][P] complex64 seen as float32 [planes][P][2 P]; box = colw columns x P/4 rows of one plane
static int make_tile_map(CUtensorMap* map, const float2* base, int P, int planes, int colw) {
    encode_tiled_fn enc = tensor_map_encoder();
    if (!enc) return -1;
    const cuuint64_t dims[3] = {cuuint64_t(2) * P, cuuint64_t(P), cuuint64_t(planes)};
    const cuuint64_t strides[2] = {cuuint64_t(P) * sizeof(float2), cuuint64_t(P) * P * sizeof(float2)};
    const cuuint32_t box[3] = {cuuint32_t(2 * colw), cuuint32_t(P / 4), 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float2*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

// launch with programmatic stream serialization (the kernels call griddepcontrol.wait before their first
// global access); BHOLO_NO_PDL / BHOLO_FFT_NO_PDL fall back to plain stream order
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    static const bool pdl = !std::getenv("BHOLO_NO_PDL") && !std::getenv("BHOLO_FFT_NO_PDL");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// the three passes over `groups` colour groups of Fg frames each (state / U / I / T point at the first group);
// H: table of colour group h_group0 + g.  fused: pass C also produces I and the loss partials.
template <int P>
static cudaError_t launch_prop2_state(const int8_t* state, float2* U, float* I, const float* T, const float2* H_all,
                                      int G_total, int h_group0, const float2* tw, int groups, int Fg, double* partial,
                                      int sms, cudaStream_t st, cudaEvent_t* ev, int phases = 3) {
    constexpr int RA = Plan2<P, false>::RA;
    constexpr size_t row_smem = size_t(RowLaySize<P, RA>::value) * sizeof(float2);
    static const int colw = std::getenv("BHOLO_FFT_W") ? std::atoi(std::getenv("BHOLO_FFT_W")) : 8;
    const size_t smA = ROWS_WARPS * row_smem, smB = colw == 8 ? cols2_smem_bytes<P, 8>() : cols2_smem_bytes<P, 4>();
    const size_t smC = INVG_WARPS * row_smem + size_t(INVG_WARPS) * P * sizeof(float);
    // experiment switches (profiles/r2_notes.md): BHOLO_FFT_A / _B / _C = 1 selects the shared-memory kernel of
    // round 1 for that pass (launched per colour group), BHOLO_FFT_B = 3 the TMA-staged input tile
    // BHOLO_FFT_A = 4 (default): k4_rows_fwd_real (state rows prefetched by cp.async, twiddles in shared memory)
    static const int selA = std::getenv("BHOLO_FFT_A") ? std::atoi(std::getenv("BHOLO_FFT_A")) : 4;
    static const int selB = std::getenv("BHOLO_FFT_B") ? std::atoi(std::getenv("BHOLO_FFT_B")) : 3;
    static const int selC = std::getenv("BHOLO_FFT_C") ? std::atoi(std::getenv("BHOLO_FFT_C")) : 2;
    static const bool late_ok = !std::getenv("BHOLO_FFT_NO_LATE_WAIT");
    auto kA = k2_rows_fwd_real<P>;
    auto kA4 = k4_rows_fwd_real<P>;
    auto kB = colw == 8 ? (selB == 3 ? k2_cols<P, true, true, 8> : k2_cols<P, true, false, 8>)
                        : (selB == 3 ? k2_cols<P, true, true, 4> : k2_cols<P, true, false, 4>);
    auto kC = k2_rows_inv_group<P>;
    auto kC3 = k3_rows_inv_group<P>;
    auto kA1 = k_rows_fwd<P, 1, int8_t, false>;
    auto kB1 = k_cols_herm<P, 1>;
    auto kC1 = k_rows_inv_group<P, 1>;
    const size_t smr1 = FftCfg<P>::smem_row, smc1 = FftCfg<P>::smem_col;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smA)))) return e;
    if ((e = cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smB)))) return e;
    if ((e = cudaFuncSetAttribute(kC, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smC)))) return e;
    if ((e = cudaFuncSetAttribute(kC3, cudaFuncAttributeMaxDynamicSharedMemorySize, int(invg3_smem_bytes<P>())))) return e;
    if ((e = cudaFuncSetAttribute(kA4, cudaFuncAttributeMaxDynamicSharedMemorySize, int(rows4_fwd_smem_bytes<P>())))) return e;
    if ((e = cudaFuncSetAttribute(kA1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smr1)))) return e;
    if ((e = cudaFuncSetAttribute(kB1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * smc1)))) return e;
    if ((e = cudaFuncSetAttribute(kC1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * smr1)))) return e;
    CUtensorMap map_buf, map_h;
    const int frames = groups * Fg;
    if (make_tile_map(&map_buf, U, P, frames, colw) || make_tile_map(&map_h, H_all, P, G_total, colw)) return cudaErrorNotSupported;
    const size_t n2 = size_t(P) * P;
    if (ev && (phases & 1)) cudaEventRecord(ev[0], st);
    const int n_pairs = frames * (P / 2);
    if (!(phases & 1)) {
    } else if (selA == 1) {
        kA1<<<dim3(P / (2 * TILE_W), frames), FftCfg<P>::T, smr1, st>>>(state, U, tw);
    } else if (selA == 4) {
        // colour groups g >= 1 of a propagation do not depend on their predecessor (pass C of group g - 1)
        const int late = (late_ok && h_group0 > 0 && phases == 3) ? 1 : 0;
        launch_pdl(kA4, std::min((n_pairs + ROWS_WARPS - 1) / ROWS_WARPS, sms * 4), 32 * ROWS_WARPS,
                   rows4_fwd_smem_bytes<P>(), st, state, U, tw, n_pairs, late);
    } else {
        launch_pdl(kA, std::min((n_pairs + ROWS_WARPS - 1) / ROWS_WARPS, sms * 4 * 4), 32 * ROWS_WARPS, smA, st, state, U, tw, n_pairs);
    }
    if (!(phases & 2)) {
        if (ev) cudaEventRecord(ev[1], st);
        return cudaGetLastError();
    }
    if (ev) cudaEventRecord(ev[1], st);
    static const bool rev = std::getenv("BHOLO_FFT_REV") != nullptr;     // most recently written group first (L2)
    if (selB == 1 && selC == 1 && rev) {
        // pass B and pass C of a group back to back, groups in reverse order of pass A's writes
        for (int g = groups - 1; g >= 0; --g) {
            kB1<<<dim3(P / (2 * FftCfg<P>::WC) + 1, Fg), FftCfg<P>::TC, 2 * smc1, st>>>(
                U + size_t(g) * Fg * n2, H_all + size_t(h_group0 + g) * n2, tw);
            kC1<<<dim3(P / TILE_W, 1), FftCfg<P>::T, 2 * smr1, st>>>(
                U + size_t(g) * Fg * n2, U + size_t(g) * Fg * n2, I + size_t(g) * n2, T + size_t(g) * n2, tw, Fg,
                partial + size_t(g) * (P / TILE_W) * 3);
        }
        if (ev) { cudaEventRecord(ev[2], st); cudaEventRecord(ev[3], st); }
        return cudaGetLastError();
    }
    if (selB == 1) {
        for (int g = 0; g < groups; ++g)
            kB1<<<dim3(P / (2 * FftCfg<P>::WC) + 1, Fg), FftCfg<P>::TC, 2 * smc1, st>>>(
                U + size_t(g) * Fg * n2, H_all + size_t(h_group0 + g) * n2, tw);
    } else {
        const int tiles = groups * (P / (2 * colw) + 1) * Fg;
        launch_pdl(kB, std::min(tiles, sms * (colw == 8 ? 1 : 2)), 32 * colw, smB, st, map_buf, map_h, U, tw, groups, Fg, h_group0);
    }
    if (ev) cudaEventRecord(ev[2], st);
    if (selC == 1) {
        for (int g = 0; g < groups; ++g)
            kC1<<<dim3(P / TILE_W, 1), FftCfg<P>::T, 2 * smr1, st>>>(
                U + size_t(g) * Fg * n2, U + size_t(g) * Fg * n2, I + size_t(g) * n2, T + size_t(g) * n2, tw, Fg,
                partial + size_t(g) * (P / TILE_W) * 3);
    } else if (selC == 3) {
        kC3<<<std::min(groups * P, sms), 32 * INVG_WARPS, invg3_smem_bytes<P>(), st>>>(U, I, T, tw, groups, Fg, partial);
    } else {
        launch_pdl(kC, std::min(groups * P, sms * 2 * 4), 32 * INVG_WARPS, smC, st, U, I, T, tw, groups, Fg, partial);
    }
    if (ev) cudaEventRecord(ev[3], st);
    return cudaGetLastError();
}

// complex input planes (stand-alone operator, sweep correlations): in -> U, one spectrum K for all planes
template <int P>
static cudaError_t launch_prop2_cplx(const float2* in, float2* U, const float2* K, const float2* tw, int planes,
                                     int sms, cudaStream_t st) {
    constexpr int RA = Plan2<P, false>::RA;
    constexpr size_t row_smem = size_t(RowLaySize<P, RA>::value) * sizeof(float2);
    constexpr int COLW = 8;
    const size_t smA = ROWS_WARPS * row_smem, smB = cols2_smem_bytes<P, COLW>();
    auto kA = k2_rows_fwd_cplx<P>;
    auto kB = k2_cols<P, false, true, COLW>;
    auto kC = k2_rows_inv<P>;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smA)))) return e;
    if ((e = cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smB)))) return e;
    if ((e = cudaFuncSetAttribute(kC, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smA)))) return e;
    CUtensorMap map_buf, map_h;
    if (make_tile_map(&map_buf, U, P, planes, COLW) || make_tile_map(&map_h, K, P, 1, COLW)) return cudaErrorNotSupported;
    const int rows = planes * P;
    const int grid_rows = std::min((rows + ROWS_WARPS - 1) / ROWS_WARPS, sms * 4 * 4);
    kA<<<grid_rows, 32 * ROWS_WARPS, smA, st>>>(in, U, tw, rows);
    // one "group" of `planes` frames: every tile uses the same spectrum
    kB<<<std::min(planes * (P / COLW), sms), 32 * COLW, smB, st>>>(map_buf, map_h, U, tw, 1, planes, 0);
    kC<<<grid_rows, 32 * ROWS_WARPS, smA, st>>>(U, U, tw, rows);
    return cudaGetLastError();
}

static bool use_fft2(const bh_ctx* c) { return c->fft2 && c->pad == 1 && fft2_side(c->P); }

static int current_sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
static bool fft2_enabled() { static const bool on = !std::getenv("BHOLO_FFT_V1"); return on; }

// ---------------------------------------------------------------------------
// propagation dispatch over the supported FFT sides
// ---------------------------------------------------------------------------
// The three FFT passes over `frames` consecutive planes (one colour group, or the C planes
// of the stand-alone operator).  Launching group by group keeps the 8 MB-per-frame
// intermediate of pass A / pass B resident in the 126 MB L2 between passes.
struct FusedC {            // non-null: pass C also produces I and the loss partials of group g
    float* I; const float* T; double* partial;
};

template <int P, int PAD, typename InT, bool CPLX>
static cudaError_t launch_prop(const InT* in, float2* buf, float2* U, const float2* H,
                               const float2* tw, int frames, int Fg, cudaStream_t st,
                               cudaEvent_t* ev = nullptr, const FusedC* fused = nullptr) {
    constexpr int N = P / PAD;
    constexpr int T = FftCfg<P>::T;
    const size_t smr = FftCfg<P>::smem_row, smc = FftCfg<P>::smem_col;
    auto kA = k_rows_fwd<P, PAD, InT, CPLX>;
    auto kB = k_cols<P, PAD>;
    auto kC = k_rows_inv<P, PAD>;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smr)))) return e;
    if ((e = cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smc)))) return e;
    if ((e = cudaFuncSetAttribute(kC, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smr)))) return e;
    (void)Fg;
    if (ev) cudaEventRecord(ev[0], st);
    // real int8 input: two rows per sequence, so a tile covers 2 * TILE_W rows
    constexpr int rows_per_tile = !CPLX ? 2 * TILE_W : TILE_W;
    kA<<<dim3(P / rows_per_tile, frames), T, smr, st>>>(in, buf, tw);
    if (ev) cudaEventRecord(ev[1], st);
    if constexpr (!CPLX) {
        // real input: pass A stored kx <= P/2 only; one forward FFT feeds columns kx and P - kx
        auto kBh = k_cols_herm<P, PAD>;
        if ((e = cudaFuncSetAttribute(kBh, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * smc)))) return e;
        kBh<<<dim3(P / (2 * FftCfg<P>::WC) + 1, frames), FftCfg<P>::TC, 2 * smc, st>>>(buf, H, tw);
    } else {
        kB<<<dim3(P / FftCfg<P>::WC, frames), FftCfg<P>::TC, smc, st>>>(buf, H, tw);
    }
    if (ev) cudaEventRecord(ev[2], st);
    if (fused) {
        auto kCg = k_rows_inv_group<P, PAD>;
        if ((e = cudaFuncSetAttribute(kCg, cudaFuncAttributeMaxDynamicSharedMemorySize, int(2 * smr)))) return e;
        kCg<<<dim3(N / TILE_W, 1), T, 2 * smr, st>>>(buf, U, fused->I, fused->T, tw, Fg, fused->partial);
    } else {
        kC<<<dim3(N / TILE_W, frames), T, smr, st>>>(buf, U, tw);
    }
    if (ev) cudaEventRecord(ev[3], st);
    return cudaGetLastError();
}

template <typename InT, bool CPLX>
static cudaError_t dispatch_prop(int P, int pad, const InT* in, float2* buf, float2* U,
                                 const float2* H, const float2* tw, int frames, int Fg,
                                 cudaStream_t st, bool* supported, cudaEvent_t* ev = nullptr,
                                 const FusedC* fused = nullptr) {
    *supported = true;
    if constexpr (CPLX) {
        if (pad == 1 && fft2_side(P) && fft2_enabled() && !fused && !ev) {
            // in -> U (buf aliases U when pad = 1); one spectrum for all planes
            if (P == 1024) return launch_prop2_cplx<1024>(in, U, H, tw, frames, current_sm_count(), st);
            if (P == 896) return launch_prop2_cplx<896>(in, U, H, tw, frames, current_sm_count(), st);
        }
    }
#define BH_CASE(PP, PD) \
    if (P == PP && pad == PD) return launch_prop<PP, PD, InT, CPLX>(in, buf, U, H, tw, frames, Fg, st, ev, fused);
    BH_CASE(32, 1) BH_CASE(64, 1) BH_CASE(128, 1) BH_CASE(256, 1) BH_CASE(512, 1)
    BH_CASE(896, 1) BH_CASE(1024, 1)
    BH_CASE(64, 2) BH_CASE(128, 2) BH_CASE(256, 2) BH_CASE(512, 2) BH_CASE(1792, 2) BH_CASE(2048, 2)
#undef BH_CASE
    *supported = false;
    return cudaSuccess;
}

static bool fft_side_supported(int P, int pad) {
    static const int p1[] = {32, 64, 128, 256, 512, 896, 1024};
    static const int p2[] = {64, 128, 256, 512, 1792, 2048};
    if (pad == 1) { for (int v : p1) if (v == P) return true; }
    if (pad == 2) { for (int v : p2) if (v == P) return true; }
    return false;
}

static int propagate_env(bh_ctx* c, int env, float* pass_ms = nullptr) {
    const size_t n2 = c->n2, p2 = size_t(c->P) * c->P;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (pass_ms)
        for (auto& e : ev) BH_CUDA(c, cudaEventCreate(&e));
    if (use_fft2(c)) {
        // launch shapes (measured, profiles/r2_notes.md): mode 0 all groups per pass; 1 (default) colour group by
        // colour group, so a group's 100 MB of intermediates stay in the 126 MB L2 between its passes;
        // 2: pass A for all groups in one launch, passes B and C group by group
        const int mode = c->fft2_mode;
        const int per_call = mode == 0 ? c->G : 1;
        if (std::getenv("BHOLO_FFT_C"))     // experiment: the round-1 pass C fills only half of the partial slots
            BH_CUDA(c, cudaMemsetAsync(c->dloss_partial, 0, size_t(c->G) * c->N * 3 * sizeof(double), c->stream));
        if (mode == 2) {
            cudaError_t e = cudaErrorNotSupported;
            if (c->P == 1024)
                e = launch_prop2_state<1024>(c->dstate + size_t(env) * c->F * n2, c->dU + size_t(env) * c->F * n2, nullptr, nullptr,
                                             c->dH, c->G, 0, c->dtw, c->G, c->Fg, nullptr, c->sms, c->stream, pass_ms ? ev : nullptr, 1);
            else if (c->P == 896)
                e = launch_prop2_state<896>(c->dstate + size_t(env) * c->F * n2, c->dU + size_t(env) * c->F * n2, nullptr, nullptr,
                                            c->dH, c->G, 0, c->dtw, c->G, c->Fg, nullptr, c->sms, c->stream, pass_ms ? ev : nullptr, 1);
            BH_CUDA(c, e);
            c->launches += 1;
            if (pass_ms) {
                float ms = 0.f;
                BH_CUDA(c, cudaEventSynchronize(ev[1]));
                BH_CUDA(c, cudaEventElapsedTime(&ms, ev[0], ev[1]));
                pass_ms[0] += ms;
            }
        }
        const int phases = mode == 2 ? 2 : 3;
        for (int g0 = 0; g0 < c->G; g0 += per_call) {
            const int8_t* st = c->dstate + (size_t(env) * c->F + size_t(g0) * c->Fg) * n2;
            float2* U = c->dU + (size_t(env) * c->F + size_t(g0) * c->Fg) * n2;
            float* I = c->dI + (size_t(env) * c->G + g0) * n2;
            const float* T = c->dT + (size_t(env) * c->G + g0) * n2;
            double* part = c->dloss_partial + size_t(g0) * c->N * 3;
            cudaError_t e = cudaErrorNotSupported;
            if (c->P == 1024)
                e = launch_prop2_state<1024>(st, U, I, T, c->dH, c->G, g0, c->dtw, per_call, c->Fg, part, c->sms, c->stream, pass_ms ? ev : nullptr, phases);
            else if (c->P == 896)
                e = launch_prop2_state<896>(st, U, I, T, c->dH, c->G, g0, c->dtw, per_call, c->Fg, part, c->sms, c->stream, pass_ms ? ev : nullptr, phases);
            BH_CUDA(c, e);
            c->launches += (phases == 3 ? 3 : 2);
            if (pass_ms) {
                BH_CUDA(c, cudaEventRecord(ev[4], c->stream));
                BH_CUDA(c, cudaEventSynchronize(ev[4]));
                for (int i = (phases == 2 ? 1 : 0); i < 4; ++i) {
                    float ms = 0.f;
                    BH_CUDA(c, cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
                    pass_ms[i] += ms;
                }
            }
        }
        if (pass_ms)
            for (auto& e : ev) cudaEventDestroy(e);
        k_loss_final<<<1, 256, 0, c->stream>>>(c->dloss_partial, c->G * c->N, double(c->G) * double(n2),
                                              c->dsums + size_t(env) * 4, c->relative);
        BH_CUDA(c, cudaGetLastError());
        c->launches += 1;
        BH_CUDA(c, cudaMemsetAsync(c->d_recon_stale + size_t(env) * RECON_MAX_BUFFERS, 0xff, RECON_MAX_BUFFERS, c->stream));
        return 0;
    }
    for (int g = 0; g < c->G; ++g) {
        const int f0 = g * c->Fg;
        const int8_t* st = c->dstate + (size_t(env) * c->F + f0) * n2;
        float2* U = c->dU + (size_t(env) * c->F + f0) * n2;
        float* I = c->dI + (size_t(env) * c->G + g) * n2;
        const float* T = c->dT + (size_t(env) * c->G + g) * n2;
        float2* buf = (c->pad == 1) ? U : c->dscratch;
        bool ok = false;
        const int tiles = c->N / TILE_W;            // partials of the fused pass C, per group
        FusedC fused{I, T, c->dloss_partial + size_t(g) * tiles * 3};
        BH_CUDA(c, (dispatch_prop<int8_t, false>(c->P, c->pad, st, buf, U, c->dH + size_t(g) * p2, c->dtw,
                                                 c->Fg, c->Fg, c->stream, &ok, pass_ms ? ev : nullptr, &fused)));
        if (!ok) BH_FAIL(c, -4, "unsupported FFT side P=%d pad=%d", c->P, c->pad);
        c->launches += 3;
        if (pass_ms) {
            BH_CUDA(c, cudaEventRecord(ev[4], c->stream));
            BH_CUDA(c, cudaEventSynchronize(ev[4]));
            for (int i = 0; i < 4; ++i) {
                float ms = 0.f;
                BH_CUDA(c, cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
                pass_ms[i] += ms;
            }
        }
    }
    if (pass_ms)
        for (auto& e : ev) cudaEventDestroy(e);
    k_loss_final<<<1, 256, 0, c->stream>>>(c->dloss_partial, c->G * (c->N / TILE_W), double(c->G) * double(n2),
                                          c->dsums + size_t(env) * 4, c->relative);
    BH_CUDA(c, cudaGetLastError());
    c->launches += 1;
    // every observation buffer has to re-read all planes of this environment
    BH_CUDA(c, cudaMemsetAsync(c->d_recon_stale + size_t(env) * RECON_MAX_BUFFERS, 0xff, RECON_MAX_BUFFERS, c->stream));
    return 0;
}

// k_eval variants: units in flight per thread x CTAs per SM x arithmetic type of the per-quad sums.
// Measured at 1024^2 x 24, 8 candidates per launch (scripts/tune_eval.py, profiles/r2_notes.md).
// BHOLO_EVAL_VARIANT selects another shape.  BHOLO_EVAL_FP64=1 forms the per-quad sums in double: measured
// (profiles/r2_parity_error_dist.json) it changes the error of dPSNR by ~10 % -- the error is set by the fp32
// fields U themselves -- and costs 13 % of the launch, so float is the default.
typedef void (*eval_fn)(const DeltaArgs);
static eval_fn commit_variant(int v) {
    switch (v) {
        case 1: return k_commit_t<1, 6>;
        case 2: return k_commit_t<2, 4>;
        case 3: return k_commit_t<3, 2>;
        case 4: return k_commit_t<4, 2>;
        case 5: return k_commit_t<2, 3>;
        case 6: return k_commit_t<3, 3>;
        default: return k_commit_t<3, 2>;
    }
}
// k_eval_bundle variants: slots per bundle x CTAs per SM; BHOLO_BUNDLE_VARIANT=9 disables bundling
// (candidate lists then go through k_eval_t)
template <typename R>
static eval_fn bundle_variant(int v, int* slots) {
    switch (v) {
        case 1: *slots = 8; return k_eval_bundle_t<8, 1, R>;
        case 2: *slots = 6; return k_eval_bundle_t<6, 1, R>;
        case 3: *slots = 3; return k_eval_bundle_t<3, 2, R>;
        case 4: *slots = 2; return k_eval_bundle_t<2, 3, R>;
        case 9: *slots = 0; return nullptr;
        default: *slots = 4; return k_eval_bundle_t<4, 2, R>;
    }
}
template <typename R>
static eval_fn eval_variant(int v) {
    switch (v) {
        case 1: return k_eval_t<2, 4, R>;
        case 2: return k_eval_t<4, 2, R>;
        case 3: return k_eval_t<2, 3, R>;
        case 4: return k_eval_t<3, 3, R>;
        case 5: return k_eval_t<6, 1, R>;
        case 6: return k_eval_t<5, 2, R>;
        case 7: return k_eval_t<2, 2, R>;
        default: return k_eval_t<3, 2, R>;
    }
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" int bh_destroy(bh_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream); else cudaDeviceSynchronize();
    cudaFree(c->dH); cudaFree(c->dh); cudaFree(c->dtw); cudaFree(c->dU); cudaFree(c->dscratch);
    cudaFree(c->dI); cudaFree(c->dT); cudaFree(c->drecon); cudaFree(c->dstate); cudaFree(c->dsums);
    cudaFree(c->dloss_partial);
    cudaFree(c->d_envs); cudaFree(c->d_actions); cudaFree(c->d_results); cudaFree(c->d_acc);
    cudaFree(c->d_tickets); cudaFree(c->d_scalars); cudaFree(c->d_dbs_s0);
    cudaFree(c->dK3); cudaFree(c->dK4); cudaFree(c->dK5); cudaFree(c->dK6); cudaFree(c->dsw_in); cudaFree(c->dsw_out);
    cudaFree(c->dsw_buf);
    cudaFree(c->dsw_it); cudaFree(c->dsw_ii); cudaFree(c->dsw_psnr);
    cudaFree(c->d_recon_stale); cudaFree(c->d_recon_plan); cudaFree(c->d_ring);
    for (auto& b : c->d_recon_obs) cudaFree(b);
    cudaFree(c->d_stat_map); cudaFree(c->d_stat_pre); cudaFree(c->d_stat_out);
    if (c->ev_recon) cudaEventDestroy(c->ev_recon);
    cudaFreeHost(c->h_envs); cudaFreeHost(c->h_actions); cudaFreeHost(c->h_results);
    cudaFreeHost(c->h_scalars); cudaFreeHost(c->h_sums);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_eval) cudaEventDestroy(c->ev_eval);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

extern "C" int bh_create(bh_ctx** out, int device, int n_env, int N, int F, int G, const double* wl,
                         double dx, double z, int pad, int relative, int method) {
    bh_ctx* nul = nullptr;
    if (!out) BH_FAIL(nul, -1, "out is null");
    *out = nullptr;
    if (n_env < 1 || N < 8 || F < 1 || G < 1 || !wl) BH_FAIL(nul, -1, "bad shape arguments");
    if (F % G) BH_FAIL(nul, -1, "F=%d is not a multiple of G=%d", F, G);
    if (N % 32) BH_FAIL(nul, -1, "N=%d must be a multiple of 32", N);
    if (pad != 1 && pad != 2) BH_FAIL(nul, -1, "pad must be 1 or 2");
    if (method != BH_METHOD_ASM && method != BH_METHOD_FRESNEL) BH_FAIL(nul, -1, "bad method");
    if (!fft_side_supported(N * pad, pad))
        BH_FAIL(nul, -4, "unsupported FFT side P=%d (pad=%d); supported: pad 1 {32,64,128,256,512,896,1024}, pad 2 {64,128,256,512,1792,2048}", N * pad, pad);
    int ndev = 0;
    BH_CUDA(nul, cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) BH_FAIL(nul, -1, "device %d not present (%d devices)", device, ndev);
    BH_CUDA(nul, cudaSetDevice(device));
    cudaDeviceProp prop;
    BH_CUDA(nul, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) BH_FAIL(nul, -5, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

    bh_ctx* c = new bh_ctx;
    c->device = device; c->E = n_env; c->N = N; c->F = F; c->G = G; c->Fg = F / G;
    c->P = N * pad; c->pad = pad; c->relative = relative ? 1 : 0; c->method = method;
    c->dx = dx; c->z = z; c->wl.assign(wl, wl + G);
    c->n2 = size_t(N) * N;
    const size_t n2 = c->n2, p2 = size_t(c->P) * c->P;
    int rc = 0;
#define BH_TRY(expr)                                                                   \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess && rc == 0) {                                            \
            char _b[256];                                                              \
            snprintf(_b, sizeof _b, "%s failed: %s", #expr, cudaGetErrorString(_e));   \
            g_err = _b; rc = -2;                                                       \
        }                                                                              \
    } while (0)
    BH_TRY(cudaMalloc(&c->dH, size_t(G) * p2 * sizeof(float2)));
    BH_TRY(cudaMalloc(&c->dh, size_t(G) * c->P * h_stride(c->P) * sizeof(float2)));
    BH_TRY(cudaMalloc(&c->dtw, (build_twiddles(c->P).size() / 2 + 1) * sizeof(float2)));
    BH_TRY(cudaMalloc(&c->dU, size_t(n_env) * F * n2 * sizeof(float2)));
    if (pad == 2) BH_TRY(cudaMalloc(&c->dscratch, size_t(c->Fg) * p2 * sizeof(float2)));
    BH_TRY(cudaMalloc(&c->dI, size_t(n_env) * G * n2 * sizeof(float)));
    BH_TRY(cudaMalloc(&c->dT, size_t(n_env) * G * n2 * sizeof(float)));
    BH_TRY(cudaMalloc(&c->drecon, size_t(G) * n2 * sizeof(float)));
    BH_TRY(cudaMalloc(&c->dstate, size_t(n_env) * F * n2));
    BH_TRY(cudaMalloc(&c->dsums, size_t(n_env) * 4 * sizeof(double)));
    BH_TRY(cudaMalloc(&c->dloss_partial, size_t(G) * N * 3 * sizeof(double)));
    c->units_per_task = int(n2 / UNIT_PX);
    c->max_tasks = std::max(4096, n_env);
    BH_TRY(cudaMalloc(&c->d_envs, size_t(c->max_tasks) * sizeof(int32_t)));
    BH_TRY(cudaMalloc(&c->d_actions, size_t(c->max_tasks) * sizeof(long long)));
    BH_TRY(cudaMalloc(&c->d_results, size_t(c->max_tasks) * sizeof(Result)));
    BH_TRY(cudaMalloc(&c->d_acc, size_t(c->max_tasks) * 2 * sizeof(unsigned long long)));
    BH_TRY(cudaMalloc(&c->d_tickets, size_t(c->max_tasks) * sizeof(unsigned)));
    BH_TRY(cudaMalloc(&c->d_scalars, 4 * sizeof(long long)));
    BH_TRY(cudaMalloc(&c->d_dbs_s0, 4 * sizeof(double)));
    BH_TRY(cudaMalloc(&c->d_recon_stale, size_t(n_env) * RECON_MAX_BUFFERS));
    BH_TRY(cudaMalloc(&c->d_recon_plan, size_t(c->max_tasks) * sizeof(ReconPlan)));
    BH_TRY(cudaEventCreateWithFlags(&c->ev_recon, cudaEventDisableTiming));
    BH_TRY(cudaMallocHost(&c->h_envs, size_t(c->max_tasks) * sizeof(int32_t)));
    BH_TRY(cudaMallocHost(&c->h_actions, size_t(c->max_tasks) * sizeof(long long)));
    BH_TRY(cudaHostAlloc(&c->h_results, size_t(c->max_tasks) * sizeof(Result), cudaHostAllocMapped));
    if (rc == 0) BH_TRY(cudaHostGetDevicePointer(&c->h_results_dev, c->h_results, 0));
    BH_TRY(cudaMallocHost(&c->h_scalars, 4 * sizeof(long long)));
    BH_TRY(cudaMallocHost(&c->h_sums, size_t(n_env) * 4 * sizeof(double)));
    BH_TRY(cudaEventCreate(&c->ev0));
    BH_TRY(cudaEventCreate(&c->ev1));
    BH_TRY(cudaEventCreateWithFlags(&c->ev_eval, cudaEventDisableTiming));
    if (rc == 0) {
        BH_TRY(cudaMemset(c->d_scalars, 0, 4 * sizeof(long long)));
        BH_TRY(cudaMemset(c->d_tickets, 0, size_t(c->max_tasks) * sizeof(unsigned)));
        BH_TRY(cudaMemset(c->d_acc, 0, size_t(c->max_tasks) * 2 * sizeof(unsigned long long)));
        BH_TRY(cudaMemset(c->dsums, 0, size_t(n_env) * 4 * sizeof(double)));
        BH_TRY(cudaMemset(c->dT, 0, size_t(n_env) * G * n2 * sizeof(float)));
        BH_TRY(cudaMemset(c->dstate, 0, size_t(n_env) * F * n2));
        BH_TRY(cudaMemset(c->dU, 0, size_t(n_env) * F * n2 * sizeof(float2)));
        BH_TRY(cudaMemset(c->dI, 0, size_t(n_env) * G * n2 * sizeof(float)));
        BH_TRY(cudaMemset(c->d_recon_stale, 0xff, size_t(n_env) * RECON_MAX_BUFFERS));
        const size_t HP = size_t(h_stride(c->P));
        for (int g = 0; g < G && rc == 0; ++g) {
            auto t = get_tables(c->P, wl[g], dx, z, method);
            BH_TRY(cudaMemcpy(c->dH + size_t(g) * p2, t->H.data(), p2 * sizeof(float2), cudaMemcpyHostToDevice));
            // rows of h carry H_PAD wrapped columns (bh_kernels.cuh: h_stride)
            BH_TRY(cudaMemcpy2D(c->dh + size_t(g) * c->P * HP, HP * sizeof(float2), t->h.data(), c->P * sizeof(float2),
                                c->P * sizeof(float2), c->P, cudaMemcpyHostToDevice));
            BH_TRY(cudaMemcpy2D(c->dh + size_t(g) * c->P * HP + c->P, HP * sizeof(float2), t->h.data(), c->P * sizeof(float2),
                                H_PAD * sizeof(float2), c->P, cudaMemcpyHostToDevice));
        }
        auto tw = build_twiddles(c->P);
        BH_TRY(cudaMemcpy(c->dtw, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice));
        int nb = 0;
        c->use_pdl = !std::getenv("BHOLO_NO_PDL");
        c->fft2 = fft2_enabled();
        c->fft2_mode = std::getenv("BHOLO_FFT_MODE") ? std::atoi(std::getenv("BHOLO_FFT_MODE")) : 1;
        c->sms = prop.multiProcessorCount;
        const char* ev = std::getenv("BHOLO_EVAL_VARIANT");
        c->fp64_eval = std::getenv("BHOLO_EVAL_FP64") != nullptr;
        if (const char* pf = std::getenv("BHOLO_EVAL_L2PF")) c->l2_prefetch = std::max(0, std::atoi(pf));
        c->k_eval = c->fp64_eval ? eval_variant<double>(ev ? std::atoi(ev) : 0) : eval_variant<float>(ev ? std::atoi(ev) : 0);
        BH_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, c->k_eval, 256, 0));
        c->grid_cap = std::min(MAX_DELTA_GRID, std::max(1, nb) * prop.multiProcessorCount);
        const char* cv = std::getenv("BHOLO_COMMIT_VARIANT");
        c->k_commit = commit_variant(cv ? std::atoi(cv) : 0);
        BH_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, c->k_commit, 256, 0));
        c->grid_cap_commit = std::max(1, nb) * prop.multiProcessorCount;
        const char* bv = std::getenv("BHOLO_BUNDLE_VARIANT");
        c->k_eval_bundle = c->fp64_eval ? bundle_variant<double>(bv ? std::atoi(bv) : 0, &c->bundle)
                                        : bundle_variant<float>(bv ? std::atoi(bv) : 0, &c->bundle);
        if (c->k_eval_bundle) {
            BH_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, c->k_eval_bundle, 256, 0));
            c->grid_cap_bundle = std::min(MAX_DELTA_GRID, std::max(1, nb) * prop.multiProcessorCount);
        }
    }
#undef BH_TRY
    if (rc) { std::string keep = g_err; bh_destroy(c); g_err = keep; return rc; }
    *out = c;
    return 0;
}

extern "C" int bh_set_stream(bh_ctx* c, void* s) {
    BH_CHECK_CTX(c);
    c->stream = reinterpret_cast<cudaStream_t>(s);
    return 0;
}

extern "C" int bh_max_tasks(const bh_ctx* c) { return c ? c->max_tasks : 0; }
extern "C" int64_t bh_launch_count(const bh_ctx* c) { return c ? c->launches : 0; }

extern "C" void* bh_device_ptr(bh_ctx* c, int which) {
    if (!c) return nullptr;
    switch (which) {
        case 0: return c->dU; case 1: return c->dI; case 2: return c->dT; case 3: return c->dstate;
        case 4: return c->dsums; case 5: return c->dh; case 6: return c->dH;
    }
    return nullptr;
}

extern "C" int bh_set_target(bh_ctx* c, int env, const float* T, int on_host) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (!T) BH_FAIL(c, -1, "target is null");
    const size_t bytes = size_t(c->G) * c->n2 * sizeof(float);
    BH_CUDA(c, cudaMemcpyAsync(c->dT + size_t(env) * c->G * c->n2, T, bytes,
                               on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
    if (on_host) BH_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int bh_load_state(bh_ctx* c, int env, const int8_t* state, int on_host) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (!state) BH_FAIL(c, -1, "state is null");
    const size_t bytes = size_t(c->F) * c->n2;
    BH_CUDA(c, cudaMemcpyAsync(c->dstate + size_t(env) * bytes, state, bytes,
                               on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
    int rc = propagate_env(c, env);
    if (rc) return rc;
    if (on_host) BH_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int bh_clone_env(bh_ctx* c, int src, int dst) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, src); BH_CHECK_ENV(c, dst);
    if (src == dst) return 0;
    const size_t n2 = c->n2;
    const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
    BH_CUDA(c, cudaMemcpyAsync(c->dU + size_t(dst) * c->F * n2, c->dU + size_t(src) * c->F * n2, size_t(c->F) * n2 * sizeof(float2), k, c->stream));
    BH_CUDA(c, cudaMemcpyAsync(c->dI + size_t(dst) * c->G * n2, c->dI + size_t(src) * c->G * n2, size_t(c->G) * n2 * sizeof(float), k, c->stream));
    BH_CUDA(c, cudaMemcpyAsync(c->dT + size_t(dst) * c->G * n2, c->dT + size_t(src) * c->G * n2, size_t(c->G) * n2 * sizeof(float), k, c->stream));
    BH_CUDA(c, cudaMemcpyAsync(c->dstate + size_t(dst) * c->F * n2, c->dstate + size_t(src) * c->F * n2, size_t(c->F) * n2, k, c->stream));
    BH_CUDA(c, cudaMemcpyAsync(c->dsums + size_t(dst) * 4, c->dsums + size_t(src) * 4, 4 * sizeof(double), k, c->stream));
    BH_CUDA(c, cudaMemsetAsync(c->d_recon_stale + size_t(dst) * RECON_MAX_BUFFERS, 0xff, RECON_MAX_BUFFERS, c->stream));
    return 0;
}

extern "C" int bh_resync(bh_ctx* c, int env) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    return propagate_env(c, env);
}

extern "C" int bh_get_metrics(bh_ctx* c, int env, double* psnr, double* mse, double* sums3) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    double* hs = c->h_sums + size_t(env) * 4;
    BH_CUDA(c, cudaMemcpyAsync(hs, c->dsums + size_t(env) * 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA(c, cudaStreamSynchronize(c->stream));
    const double n = double(c->G) * double(c->n2);
    const double m = c->relative ? (hs[2] - hs[1] * hs[1] / hs[0]) / n : (hs[0] - 2.0 * hs[1] + hs[2]) / n;
    if (psnr) *psnr = hs[3];
    if (mse) *mse = m;
    if (sums3) { sums3[0] = hs[0]; sums3[1] = hs[1]; sums3[2] = hs[2]; }
    return 0;
}

// ---------------------------------------------------------------------------
// incremental path
// ---------------------------------------------------------------------------
static DeltaArgs make_args(bh_ctx* c, int n, int env_fixed, const int32_t* d_envs,
                           const long long* d_actions, int rule, Result* d_results) {
    DeltaArgs a;
    a.U = c->dU; a.I = c->dI; a.T = c->dT; a.state = c->dstate; a.h = c->dh; a.sums = c->dsums;
    a.envs = d_envs; a.actions = d_actions; a.offset_ptr = nullptr; a.n_total = n;
    a.env_fixed = env_fixed;
    a.n_tasks = n; a.N = c->N; a.P = c->P; a.HP = h_stride(c->P); a.F = c->F; a.G = c->G; a.Fg = c->Fg;
    a.relative = c->relative; a.rule = rule;
    a.units_per_task = c->units_per_task;
    a.unit_dy = UNIT_PX / c->N; a.unit_dx = UNIT_PX % c->N;
    a.acc = c->d_acc; a.tickets = c->d_tickets; a.results = d_results;
    a.results_host = nullptr; a.n_inline = 0; a.sort_window = 0;
    a.recon_stale = c->d_recon_stale;
    a.l2_prefetch_units = c->use_pdl ? c->l2_prefetch : 0;
    a.log_accept = nullptr; a.log_psnr = nullptr;
    a.dbs_accepted = nullptr; a.dbs_trace = nullptr; a.dbs_count = nullptr; a.dbs_cursor = nullptr;
    a.dbs_s0 = nullptr;
    return a;
}

static inline int delta_grid(const bh_ctx* c, int n) {
    const long long total = (long long)n * c->units_per_task;
    return int(std::min<long long>(total, c->grid_cap));
}

// delta kernels are launched with programmatic stream serialization (see pdl_wait_then_release)
static cudaError_t launch_delta(bh_ctx* c, void (*kern)(const DeltaArgs), int grid, const DeltaArgs& a) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = c->use_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

static int launch_eval(bh_ctx* c, const DeltaArgs& a) {
    launch_delta(c, c->k_eval, delta_grid(c, a.n_tasks), a);
    c->launches += 1;
    return 0;
}
// Candidate lists that revisit frames of an environment: bundled evaluation (shared U / I / T
// loads, k_eval_bundle_t).  It pays once runs of same-frame candidates are likely -- host lists are
// sorted by frame, windows of one environment are sorted inside the kernel and need about four
// candidates per frame.  Short windows (the greedy DBS while its accept rate is high) are faster
// through k_eval_t, which keeps more independent units in flight (scripts/tune_bundle.py).
static int launch_eval_list(bh_ctx* c, const DeltaArgs& a, bool sorted_by_caller) {
    // images that are not row-regular (896^2 crop) run the bundled kernel's generic loop, one unit in
    // flight: it only wins on long lists (896^2 x 24: 128 per call 580 k -> 530 k/s, 65 536 per call
    // 616 k -> 863 k/s)
    const bool regular = UNIT_PX % c->N == 0;
    const int need = !regular ? 512 : (sorted_by_caller ? 2 * c->bundle : 4 * c->F);
    if (!c->k_eval_bundle || a.n_tasks < need || (!sorted_by_caller && (!regular || a.n_tasks > SORT_WINDOW_MAX)))
        return launch_eval(c, a);
    const long long total = (long long)((a.n_tasks + c->bundle - 1) / c->bundle) * c->units_per_task;
    launch_delta(c, c->k_eval_bundle, int(std::min<long long>(total, c->grid_cap_bundle)), a);
    c->launches += 1;
    return 0;
}
// commit launches carry at most COMMIT_MAX_TASKS tasks; larger batches are split
static int launch_commit(bh_ctx* c, const DeltaArgs& a0) {
    for (int base = 0; base < a0.n_tasks; base += COMMIT_MAX_TASKS) {
        DeltaArgs a = a0;
        a.n_tasks = std::min(COMMIT_MAX_TASKS, a0.n_tasks - base);
        if (a.envs) a.envs += base;
        a.results += base;
        // one accepted task already fills the chip; more only add units
        const int grid = int(std::min<long long>((long long)a.units_per_task, c->grid_cap_commit));
        launch_delta(c, c->k_commit, grid, a);
        c->launches += 1;
    }
    return 0;
}

static int check_actions(bh_ctx* c, const int64_t* actions, int64_t n) {
    const int64_t lim = int64_t(c->F) * int64_t(c->n2);
    for (int64_t i = 0; i < n; ++i)
        if (actions[i] < 0 || actions[i] >= lim)
            BH_FAIL(c, -3, "action[%lld]=%lld out of range [0,%lld)", (long long)i, (long long)actions[i], (long long)lim);
    return 0;
}

extern "C" int bh_eval_flips_device(bh_ctx* c, int env, int n, const int32_t* d_env_ids,
                                    const int64_t* d_actions, bh_result* d_results) {
    BH_CHECK_CTX(c);
    if (n < 0 || n > c->max_tasks) BH_FAIL(c, -3, "n=%d exceeds max_tasks=%d", n, c->max_tasks);
    if (n == 0) return 0;
    if (!d_env_ids) BH_CHECK_ENV(c, env);
    DeltaArgs a = make_args(c, n, env, d_env_ids, reinterpret_cast<const long long*>(d_actions),
                            RULE_NEVER, reinterpret_cast<Result*>(d_results));
    if (d_env_ids) {
        launch_eval(c, a);
    } else {
        a.sort_window = 1;              // a window of one environment: frame order inside the CTA
        launch_eval_list(c, a, false);
    }
    BH_CUDA(c, cudaGetLastError());
    return 0;
}

extern "C" int bh_eval_flips(bh_ctx* c, int env, int64_t n, const int32_t* env_ids,
                             const int64_t* actions, double* psnr_after) {
    BH_CHECK_CTX(c);
    if (n < 0 || !actions || !psnr_after) BH_FAIL(c, -1, "bad arguments");
    if (!env_ids) BH_CHECK_ENV(c, env);
    if (int rc = check_actions(c, actions, n)) return rc;
    if (env_ids)
        for (int64_t i = 0; i < n; ++i) BH_CHECK_ENV(c, env_ids[i]);
    std::vector<int> perm;
    for (int64_t base = 0; base < n; base += c->max_tasks) {
        const int m = int(std::min<int64_t>(c->max_tasks, n - base));
        // order the chunk by (env, action): consecutive tasks then share a frame,
        // so its U / I / T stream from L2 instead of HBM
        perm.resize(m);
        std::iota(perm.begin(), perm.end(), 0);
        std::sort(perm.begin(), perm.end(), [&](int x, int y) {
            const int ex = env_ids ? env_ids[base + x] : 0, ey = env_ids ? env_ids[base + y] : 0;
            if (ex != ey) return ex < ey;
            if (actions[base + x] != actions[base + y]) return actions[base + x] < actions[base + y];
            return x < y;
        });
        for (int i = 0; i < m; ++i) {
            c->h_actions[i] = actions[base + perm[i]];
            if (env_ids) c->h_envs[i] = env_ids[base + perm[i]];
        }
        BH_CUDA(c, cudaMemcpyAsync(c->d_actions, c->h_actions, size_t(m) * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
        if (env_ids)
            BH_CUDA(c, cudaMemcpyAsync(c->d_envs, c->h_envs, size_t(m) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        DeltaArgs a = make_args(c, m, env, env_ids ? c->d_envs : nullptr, c->d_actions, RULE_NEVER, c->d_results);
        launch_eval_list(c, a, true);
        BH_CUDA(c, cudaGetLastError());
        BH_CUDA(c, cudaMemcpyAsync(c->h_results, c->d_results, size_t(m) * sizeof(Result), cudaMemcpyDeviceToHost, c->stream));
        BH_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int i = 0; i < m; ++i) psnr_after[base + perm[i]] = c->h_results[i].psnr_after;
    }
    return 0;
}

extern "C" int bh_step_batch_device(bh_ctx* c, int n, const int32_t* d_env_ids, const int64_t* d_actions,
                                    int rule, bh_result* d_results) {
    BH_CHECK_CTX(c);
    if (n < 0 || n > c->max_tasks) BH_FAIL(c, -3, "n=%d exceeds max_tasks=%d", n, c->max_tasks);
    if (rule < 0 || rule > 3) BH_FAIL(c, -1, "bad rule %d", rule);
    if (n == 0) return 0;
    if (!d_env_ids && n > 1) BH_FAIL(c, -1, "a batch step needs one distinct env per task");
    DeltaArgs a = make_args(c, n, 0, d_env_ids, reinterpret_cast<const long long*>(d_actions), rule,
                            reinterpret_cast<Result*>(d_results));
    launch_eval(c, a);
    if (rule != RULE_NEVER) launch_commit(c, a);
    BH_CUDA(c, cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// open-loop rollouts in one persistent launch (k_rollout_t)
// ---------------------------------------------------------------------------
static int rollout_setup(bh_ctx* c) {
    if (c->d_ring) return 0;
    int coop = 0, nb = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
    auto kern = c->fp64_eval ? k_rollout_t<3, double, true> : k_rollout_t<3, float, true>;
    BH_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, 0));
    c->rollout_cap = (coop && !std::getenv("BHOLO_NO_ROLLOUT_KERNEL")) ? nb * c->sms : 0;
    BH_CUDA(c, cudaMalloc(&c->d_ring, (size_t(c->E) * ROLLOUT_RING * 2 + 2) * sizeof(unsigned long long)));
    BH_CUDA(c, cudaMemsetAsync(c->d_ring, 0, (size_t(c->E) * ROLLOUT_RING * 2 + 2) * sizeof(unsigned long long), c->stream));
    return 0;
}

// 1: launched; 0: this shape / rule needs the two-kernel chain; < 0: error
static int launch_rollout(bh_ctx* c, int n_env, const int32_t* d_env_ids, const long long* d_actions,
                          long long act_step, long long act_env, int steps, int rule, Result* d_results,
                          long long res_step, long long res_env, uint8_t* log_accept, double* log_psnr) {
    if (int rc = rollout_setup(c)) return rc;
    const bool regular = UNIT_PX % c->N == 0;
    int cpe = std::min(c->rollout_cap / std::max(1, n_env), std::min(c->units_per_task, MAX_DELTA_GRID));
    static const int cpe_cap = std::getenv("BHOLO_ROLLOUT_CPE") ? std::atoi(std::getenv("BHOLO_ROLLOUT_CPE")) : 0;
    static const int backoff = std::getenv("BHOLO_ROLLOUT_BACKOFF") ? std::atoi(std::getenv("BHOLO_ROLLOUT_BACKOFF")) : 0;
    if (cpe_cap > 0) cpe = std::min(cpe, cpe_cap);
    if (!regular || cpe < 1 || rule == RULE_NEVER) return 0;
    RolloutArgs ra;
    ra.a = make_args(c, n_env, 0, nullptr, nullptr, rule, nullptr);
    ra.envs = d_env_ids;
    ra.actions = d_actions;
    ra.act_step = act_step; ra.act_env = act_env;
    ra.results = d_results;
    ra.res_step = res_step; ra.res_env = res_env;
    ra.log_accept = log_accept; ra.log_psnr = log_psnr;
    ra.ring = c->d_ring;
    ra.error = reinterpret_cast<int*>(c->d_ring + size_t(c->E) * ROLLOUT_RING * 2);
    ra.n_env = n_env; ra.steps = steps; ra.cpe = cpe;
    ra.backoff_ns = unsigned(std::max(0, backoff));
    // the ring words are zeroed, the error flag (the word behind them) is sticky until bh_rollout_status
    BH_CUDA(c, cudaMemsetAsync(c->d_ring, 0, size_t(c->E) * ROLLOUT_RING * 2 * sizeof(unsigned long long), c->stream));
    void* params[] = {&ra};
    // look-ahead pays while an environment's barrier latency is exposed (few environments per GPU); with many, the
    // other environments fill the gap and a discarded look-ahead would only cost bandwidth
    static const int look_max = std::getenv("BHOLO_ROLLOUT_LOOK_MAX_ENVS") ? std::atoi(std::getenv("BHOLO_ROLLOUT_LOOK_MAX_ENVS")) : 2;
    const bool look = n_env <= look_max;
    const void* kern = c->fp64_eval
        ? (look ? reinterpret_cast<const void*>(k_rollout_t<3, double, true>) : reinterpret_cast<const void*>(k_rollout_t<3, double, false>))
        : (look ? reinterpret_cast<const void*>(k_rollout_t<3, float, true>) : reinterpret_cast<const void*>(k_rollout_t<3, float, false>));
    const cudaError_t le = cudaLaunchCooperativeKernel(kern, dim3(cpe * n_env), dim3(256), params, 0, c->stream);
    if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorNotSupported || le == cudaErrorLaunchOutOfResources) {
        // co-residency refused (a shared or partitioned GPU): the two-kernel chain computes the same thing
        cudaGetLastError();
        c->rollout_cap = 0;
        return 0;
    }
    BH_CUDA(c, le);
    c->launches += 1;
    return 1;
}

extern "C" int bh_rollout_device(bh_ctx* c, int n_env, const int32_t* d_env_ids, const int64_t* d_actions,
                                 int64_t act_step_stride, int64_t act_env_stride, int steps, int rule,
                                 bh_result* d_results, int64_t res_step_stride, int64_t res_env_stride) {
    BH_CHECK_CTX(c);
    if (n_env < 1 || n_env > c->E || n_env > c->max_tasks || steps < 0 || !d_actions) BH_FAIL(c, -1, "bad arguments");
    if (rule < 0 || rule > 3) BH_FAIL(c, -1, "bad rule %d", rule);
    if (steps == 0) return 0;
    const int rc = launch_rollout(c, n_env, d_env_ids, reinterpret_cast<const long long*>(d_actions), act_step_stride,
                                  act_env_stride, steps, rule, reinterpret_cast<Result*>(d_results), res_step_stride,
                                  res_env_stride, nullptr, nullptr);
    if (rc != 0) return rc < 0 ? rc : 0;
    // the two-kernel chain, step by step (every image size, any number of environments)
    if (act_env_stride != 1 || (d_results && res_env_stride != 1))
        BH_FAIL(c, -4, "this shape needs step-major action / result lists (env stride 1)");
    if (!d_env_ids && n_env > 1) BH_FAIL(c, -1, "a batch step needs one distinct env per task");
    for (int t = 0; t < steps; ++t) {
        DeltaArgs a = make_args(c, n_env, 0, d_env_ids, reinterpret_cast<const long long*>(d_actions) + t * act_step_stride,
                                rule, d_results ? reinterpret_cast<Result*>(d_results) + t * res_step_stride : c->d_results);
        launch_eval(c, a);
        if (rule != RULE_NEVER) launch_commit(c, a);
    }
    BH_CUDA(c, cudaGetLastError());
    return 0;
}

// 1 if the last rollout aborted at a barrier (a CTA waited longer than ROLLOUT_SPIN_LIMIT polls); synchronises
extern "C" int bh_rollout_status(bh_ctx* c) {
    BH_CHECK_CTX(c);
    if (!c->d_ring) return 0;
    int flag = 0;
    BH_CUDA(c, cudaMemcpyAsync(&flag, c->d_ring + size_t(c->E) * ROLLOUT_RING * 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    BH_CUDA(c, cudaStreamSynchronize(c->stream));
    if (flag) {
        cudaMemsetAsync(c->d_ring + size_t(c->E) * ROLLOUT_RING * 2, 0, sizeof(unsigned long long), c->stream);
        BH_FAIL(c, -5, "rollout kernel aborted at a barrier (co-residency lost?)");
    }
    return 0;
}

extern "C" int bh_step_batch(bh_ctx* c, int n, const int32_t* env_ids, const int64_t* actions, int rule,
                             bh_result* results) {
    BH_CHECK_CTX(c);
    if (n < 0 || n > c->max_tasks || !actions || !results) BH_FAIL(c, -1, "bad arguments");
    if (rule < 0 || rule > 3) BH_FAIL(c, -1, "bad rule %d", rule);
    if (n == 0) return 0;
    if (int rc = check_actions(c, actions, n)) return rc;
    if (n > c->E) BH_FAIL(c, -3, "n=%d tasks but only %d environments", n, c->E);
    // distinct environments: a launch commits at most one flip per env
    std::vector<char> seen(c->E, 0);
    for (int i = 0; i < n; ++i) {
        const int e = env_ids ? env_ids[i] : i;
        BH_CHECK_ENV(c, e);
        if (seen[e]) BH_FAIL(c, -3, "environment %d appears twice in one batch step", e);
        seen[e] = 1;
        c->h_envs[i] = e;
        c->h_actions[i] = actions[i];
    }
    DeltaArgs a = make_args(c, n, 0, c->d_envs, c->d_actions, rule, c->d_results);
    if (n <= INLINE_MAX) {
        // step path: tasks ride in the kernel parameters, the finaliser mirrors the results
        // into mapped pinned memory, and the host only waits for the evaluation -- the
        // commit of the accepted flips overlaps the caller's bookkeeping (stream order
        // keeps every later call behind it)
        a.n_inline = n;
        for (int i = 0; i < n; ++i) { a.inl_actions[i] = c->h_actions[i]; a.inl_envs[i] = c->h_envs[i]; }
        a.results_host = c->h_results_dev;
        launch_eval(c, a);
        BH_CUDA(c, cudaEventRecord(c->ev_eval, c->stream));
        if (rule != RULE_NEVER) launch_commit(c, a);
        BH_CUDA(c, cudaGetLastError());
        BH_CUDA(c, cudaEventSynchronize(c->ev_eval));
    } else {
        BH_CUDA(c, cudaMemcpyAsync(c->d_actions, c->h_actions, size_t(n) * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
        BH_CUDA(c, cudaMemcpyAsync(c->d_envs, c->h_envs, size_t(n) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        launch_eval(c, a);
        if (rule != RULE_NEVER) launch_commit(c, a);
        BH_CUDA(c, cudaGetLastError());
        BH_CUDA(c, cudaMemcpyAsync(c->h_results, c->d_results, size_t(n) * sizeof(Result), cudaMemcpyDeviceToHost, c->stream));
        BH_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    std::memcpy(results, c->h_results, size_t(n) * sizeof(Result));
    return 0;
}

extern "C" int bh_vec_book_update(int n, const int32_t* env_ids, const int64_t* actions,
                                  const bh_result* results, const bh_vec_book* b) {
    if (!b || !b->prev_psnr || !b->init_psnr || !b->steps || !b->flips || !b->rewards)
        BH_FAIL((bh_ctx*)nullptr, -1, "incomplete bh_vec_book");
    if (n < 0 || (n > 0 && (!actions || !results))) BH_FAIL((bh_ctx*)nullptr, -1, "bad arguments");
    for (int i = 0; i < n; ++i) {
        const int e = env_ids ? env_ids[i] : i;
        const bh_result& r = results[i];
        const int64_t a = actions[i];
        b->steps[e] += 1;
        if (b->state_record) b->state_record[size_t(e) * b->stride + a] += 1;      // env.py:165
        const double change = r.psnr_after - b->prev_psnr[e];                      // env.py:184
        const double diff = r.psnr_after - b->init_psnr[e];
        b->rewards[i] = change * b->reward_scale;                                   // env.py:188
        if (b->psnr_change) b->psnr_change[i] = change;
        if (b->psnr_diff) b->psnr_diff[i] = diff;
        uint8_t ev = 0;
        if (r.accept) {
            if (b->state) b->state[size_t(e) * b->stride + a] ^= 1;                 // env.py:164
            b->flips[e] += 1;
            b->prev_psnr[e] = r.psnr_after;                                         // env.py:214
            const bool success = b->t_psnr_diff && b->t_psnr &&
                (diff >= b->t_psnr_diff[e] || (r.psnr_after >= b->t_psnr[e] && diff < 0.1));
            const bool over = b->max_steps && b->steps[e] >= b->max_steps[e];
            ev = (success || over) ? 1 : 0;                                          // env.py:216,237
        }
        if (b->last_candidate) b->last_candidate[e] = r.accept ? -1 : a;
        if (b->event) b->event[i] = ev;
    }
    return 0;
}

extern "C" int bh_vec_step(bh_ctx* c, int n, const int32_t* env_ids, const int64_t* actions, int rule,
                           bh_result* results, const bh_vec_book* b) {
    if (!b || !b->prev_psnr || !b->init_psnr || !b->steps || !b->flips || !b->rewards)
        BH_FAIL(c, -1, "incomplete bh_vec_book");
    if (int rc = bh_step_batch(c, n, env_ids, actions, rule, results)) return rc;
    return bh_vec_book_update(n, env_ids, actions, results, b);
}

extern "C" int bh_commit_flip(bh_ctx* c, int env, int64_t action) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (int rc = check_actions(c, &action, 1)) return rc;
    // score with the "always keep" rule, then apply: no host round trip in between
    DeltaArgs a = make_args(c, 1, env, nullptr, c->d_actions, RULE_ALWAYS, c->d_results);
    a.n_inline = 1; a.inl_actions[0] = action; a.inl_envs[0] = env;
    launch_eval(c, a);
    launch_commit(c, a);
    BH_CUDA(c, cudaGetLastError());
    BH_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int bh_dbs_run(bh_ctx* c, int env, const int64_t* order, int64_t n, int k_spec,
                          int64_t resync_every, uint8_t* accepted, double* psnr_trace,
                          int64_t* n_accepted, double* final_psnr) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (n < 0 || (n > 0 && (!order || !accepted))) BH_FAIL(c, -1, "bad arguments");
    if (int rc = check_actions(c, order, n)) return rc;
    long long* d_order = nullptr; uint8_t* d_acc = nullptr; double* d_trace = nullptr;
    int rc = 0;
    // The iterations of a chunk are identical launches (the cursor lives on the device), so a
    // chunk is captured once per speculation depth into a CUDA graph and replayed: the loop is
    // launch bound at small N (256^2: ~4 us of GPU work per iteration).
    std::map<int, cudaGraphExec_t> graphs;
    cudaStream_t caller_stream = c->stream;
    const bool use_graph = n >= 4096 && !std::getenv("BHOLO_NO_GRAPH");
    if (use_graph && !c->stream) {
        if (!c->own_stream) BH_CUDA(c, cudaStreamCreate(&c->own_stream));   // blocking: ordered with stream 0
        c->stream = c->own_stream;
    }
    auto cleanup = [&]() {
        for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
        cudaFree(d_order); cudaFree(d_acc); cudaFree(d_trace);
        c->stream = caller_stream;
    };
#define BH_DBS(expr)                                                                       \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            cleanup();                                                                     \
            BH_FAIL(c, -2, "%s failed: %s", #expr, cudaGetErrorString(_e));                \
        }                                                                                  \
    } while (0)
    if (n > 0) {
        BH_DBS(cudaMalloc(&d_order, size_t(n) * sizeof(long long)));
        BH_DBS(cudaMalloc(&d_acc, size_t(n)));
        if (psnr_trace) BH_DBS(cudaMalloc(&d_trace, size_t(n) * sizeof(double)));
        BH_DBS(cudaMemcpyAsync(d_order, order, size_t(n) * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
        BH_DBS(cudaMemsetAsync(d_acc, 0, size_t(n), c->stream));
        BH_DBS(cudaMemsetAsync(c->d_scalars, 0, 2 * sizeof(long long), c->stream));
        const int kmax = std::min(SORT_WINDOW_MAX, c->max_tasks);
        int K = k_spec > 0 ? std::min(k_spec, c->max_tasks) : 2;
        const int iters_per_sync = 32;
        long long cursor = 0, nacc = 0, last_resync = 0;
        // High accept rates: the strictly sequential loop in the persistent rollout kernel with one step of
        // look-ahead (k_rollout_t<LOOK>: no launch per candidate, the next candidate is scored while the barrier
        // of the current one resolves) beats the speculation windows, whose evaluations are discarded behind every
        // kept flip of the same colour group: measured 114 k -> 125 k candidates/s at 50 % kept flips, 130 k -> 137 k
        // at 36 % (profiles/r2_notes.md 8.7).  Low accept rates: the windows (many candidates per pass over the
        // image, 320-340 k/s at 1 %) win.  Chunks of up to DBS_ROLL_CHUNK candidates; the accept rate of a chunk
        // picks the mode of the next (BHOLO_DBS_ROLLOUT_MIN_ACCEPT, default 0.25; > 1 disables the rollout mode).
        static const double roll_thr = std::getenv("BHOLO_DBS_ROLLOUT_MIN_ACCEPT")
                                           ? std::atof(std::getenv("BHOLO_DBS_ROLLOUT_MIN_ACCEPT")) : 0.25;
        const long long DBS_ROLL_CHUNK = 512;
        bool roll_ok = k_spec <= 0 && UNIT_PX % c->N == 0 && roll_thr <= 1.0;
        double recent_accept = 1.0;
        bool scalars_dirty = false;
        uint8_t* h_chunk = nullptr;
        if (roll_ok) {
            c->h_envs[0] = env;
            BH_DBS(cudaMemcpyAsync(c->d_envs, c->h_envs, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            BH_DBS(cudaMallocHost(&h_chunk, size_t(DBS_ROLL_CHUNK)));
        }
        while (cursor < n) {
            if (roll_ok && recent_accept >= roll_thr) {
                long long m = std::min<long long>(DBS_ROLL_CHUNK, n - cursor);
                if (resync_every > 0) {                  // end the chunk about where the next re-synchronisation is due
                    const double need = double(resync_every - (nacc - last_resync));
                    m = std::min<long long>(m, std::max<long long>(16, (long long)(need / std::max(recent_accept, 0.02)) + 1));
                }
                const int launched = launch_rollout(c, 1, c->d_envs, d_order + cursor, 1, 0, int(m), RULE_DBS, nullptr, 1, 0,
                                                    d_acc + cursor, d_trace ? d_trace + cursor : nullptr);
                if (launched < 0) { cudaFreeHost(h_chunk); cleanup(); return launched; }
                if (launched == 0) { roll_ok = false; continue; }
                BH_DBS(cudaMemcpyAsync(h_chunk, d_acc + cursor, size_t(m), cudaMemcpyDeviceToHost, c->stream));
                BH_DBS(cudaStreamSynchronize(c->stream));
                long long kept = 0;
                for (long long i = 0; i < m; ++i) kept += h_chunk[i];
                cursor += m; nacc += kept;
                recent_accept = double(kept) / double(m);
                scalars_dirty = true;                    // the windows read cursor and count from the device
                if (resync_every > 0 && nacc - last_resync >= resync_every && cursor < n) {
                    rc = propagate_env(c, env);
                    if (rc) { cudaFreeHost(h_chunk); cleanup(); return rc; }
                    last_resync = nacc;
                }
                continue;
            }
            if (scalars_dirty) {
                c->h_scalars[0] = cursor; c->h_scalars[1] = nacc;
                BH_DBS(cudaMemcpyAsync(c->d_scalars, c->h_scalars, 2 * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
                BH_DBS(cudaStreamSynchronize(c->stream));      // h_scalars is read back below
                scalars_dirty = false;
            }
            DeltaArgs a = make_args(c, K, env, nullptr, d_order, RULE_DBS, c->d_results);
            a.offset_ptr = c->d_scalars;
            a.n_total = n;
            a.sort_window = 1;
            a.dbs_s0 = c->d_dbs_s0;             // the finalisers publish the sums the window was scored against
            DeltaArgs ac = a;                   // the commit kernel also selects and logs
            ac.dbs_cursor = c->d_scalars; ac.dbs_count = c->d_scalars + 1;
            ac.dbs_accepted = d_acc; ac.dbs_trace = d_trace;
            if (use_graph) {
                auto it = graphs.find(K);
                if (it == graphs.end()) {
                    cudaGraph_t g = nullptr;
                    cudaGraphExec_t ge = nullptr;
                    const int64_t before = c->launches;
                    BH_DBS(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                    for (int i = 0; i < iters_per_sync; ++i) { launch_eval_list(c, a, false); launch_commit(c, ac); }
                    // always leave capture mode, also when a captured launch failed
                    const cudaError_t cap_err = cudaStreamEndCapture(c->stream, &g);
                    if (cap_err != cudaSuccess && g) { cudaGraphDestroy(g); g = nullptr; }
                    BH_DBS(cap_err);
                    c->launches = before;       // captured, not launched
                    BH_DBS(cudaGraphInstantiate(&ge, g, 0));
                    cudaGraphDestroy(g);
                    it = graphs.emplace(K, ge).first;
                }
                BH_DBS(cudaGraphLaunch(it->second, c->stream));
                c->launches += 2 * iters_per_sync;
            } else {
                for (int i = 0; i < iters_per_sync; ++i) { launch_eval_list(c, a, false); launch_commit(c, ac); }
            }
            BH_DBS(cudaGetLastError());
            BH_DBS(cudaMemcpyAsync(c->h_scalars, c->d_scalars, 2 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
            BH_DBS(cudaStreamSynchronize(c->stream));
            const long long consumed = c->h_scalars[0] - cursor;
            if (consumed > 0) recent_accept = double(c->h_scalars[1] - nacc) / double(consumed);
            cursor = c->h_scalars[0];
            nacc = c->h_scalars[1];
            if (k_spec <= 0) {
                // adapt the speculation depth to the accept rate: an iteration costs ~7.5 us + ~4 us per scored
                // candidate, a window is used up to the first candidate whose colour group was touched
                // (profiles/r2_notes.md 7: at 50 % kept flips 2.8 of 4 candidates are used, 3.2 of 16)
                const double per_iter = double(consumed) / iters_per_sync;
                if (per_iter > 0.75 * K && K < kmax) K = std::min(kmax, K * 2);
                else if (per_iter < 0.50 * K && K > 1) K = std::max(1, K / 2);
            }
            if (resync_every > 0 && nacc - last_resync >= resync_every && cursor < n) {
                rc = propagate_env(c, env);
                if (rc) { cleanup(); return rc; }
                last_resync = nacc;
            }
        }
        cudaFreeHost(h_chunk);
        BH_DBS(cudaMemcpyAsync(accepted, d_acc, size_t(n), cudaMemcpyDeviceToHost, c->stream));
        if (psnr_trace)
            BH_DBS(cudaMemcpyAsync(psnr_trace, d_trace, size_t(n) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        BH_DBS(cudaStreamSynchronize(c->stream));
        if (n_accepted) *n_accepted = nacc;
    } else if (n_accepted) {
        *n_accepted = 0;
    }
#undef BH_DBS
    cleanup();
    if (final_psnr) return bh_get_metrics(c, env, final_psnr, nullptr, nullptr);
    return 0;
}

// Greedy DBS of several images at once (the dataset loop of DBS.py:208 / DBS_1024_24.py:211 is independent
// per image): iteration i scores candidate orders[e][i] of every environment e in ONE k_eval launch and keeps
// the improving ones in ONE k_commit launch.  No speculation: every evaluation is used, the decisions of each
// image are exactly those of its sequential loop, and the launch overhead is shared by n_env candidates.
extern "C" int bh_dbs_run_batch(bh_ctx* c, int n_env, const int32_t* env_ids, const int64_t* orders, int64_t n,
                                int64_t resync_every, uint8_t* accepted, double* psnr_trace,
                                int64_t* n_accepted, double* final_psnr) {
    BH_CHECK_CTX(c);
    if (n_env < 1 || n_env > c->E || n_env > COMMIT_MAX_TASKS || n < 0 || (n > 0 && (!orders || !accepted)))
        BH_FAIL(c, -1, "bad arguments");
    std::vector<int32_t> ids(n_env);
    std::vector<char> seen(c->E, 0);
    for (int e = 0; e < n_env; ++e) {
        ids[e] = env_ids ? env_ids[e] : e;
        BH_CHECK_ENV(c, ids[e]);
        if (seen[ids[e]]) BH_FAIL(c, -3, "environment %d appears twice", ids[e]);
        seen[ids[e]] = 1;
    }
    for (int e = 0; e < n_env; ++e)
        if (int rc = check_actions(c, orders + size_t(e) * n, n)) return rc;
    const int64_t chunk = std::min<int64_t>(n, 1 << 18);
    long long* d_ord = nullptr; uint8_t* d_acc = nullptr; double* d_tr = nullptr; int32_t* d_ids = nullptr;
    long long* h_ord = nullptr; uint8_t* h_acc = nullptr; double* h_tr = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_ord); cudaFree(d_acc); cudaFree(d_tr); cudaFree(d_ids);
        cudaFreeHost(h_ord); cudaFreeHost(h_acc); cudaFreeHost(h_tr);
    };
#define BH_DB(expr)                                                                        \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            cleanup();                                                                     \
            BH_FAIL(c, -2, "%s failed: %s", #expr, cudaGetErrorString(_e));                \
        }                                                                                  \
    } while (0)
    if (n > 0) {
        const size_t cells = size_t(chunk) * n_env;
        BH_DB(cudaMalloc(&d_ord, cells * sizeof(long long)));
        BH_DB(cudaMalloc(&d_acc, cells));
        if (psnr_trace) BH_DB(cudaMalloc(&d_tr, cells * sizeof(double)));
        BH_DB(cudaMalloc(&d_ids, size_t(n_env) * sizeof(int32_t)));
        BH_DB(cudaMallocHost(&h_ord, cells * sizeof(long long)));
        BH_DB(cudaMallocHost(&h_acc, cells));
        if (psnr_trace) BH_DB(cudaMallocHost(&h_tr, cells * sizeof(double)));
        BH_DB(cudaMemcpyAsync(d_ids, ids.data(), size_t(n_env) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        int64_t since_resync = 0;
        for (int64_t base = 0; base < n; base += chunk) {
            const int64_t m = std::min<int64_t>(chunk, n - base);
            for (int64_t i = 0; i < m; ++i)                     // [iteration][env]
                for (int e = 0; e < n_env; ++e) h_ord[size_t(i) * n_env + e] = orders[size_t(e) * n + base + i];
            BH_DB(cudaMemcpyAsync(d_ord, h_ord, size_t(m) * n_env * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
            for (int64_t i = 0; i < m;) {
                // iterations up to the next re-synchronisation: ONE persistent launch (k_rollout_t: identical
                // decisions and fields), or -- image sizes that are not row regular -- one k_eval + one k_commit
                // launch per iteration
                int64_t seg = m - i;
                if (resync_every > 0) seg = std::min<int64_t>(seg, std::max<int64_t>(1, resync_every - since_resync));
                seg = std::min<int64_t>(seg, 1 << 20);
                const int launched = launch_rollout(c, n_env, d_ids, d_ord + size_t(i) * n_env, n_env, 1, int(seg), RULE_DBS,
                                                    nullptr, n_env, 1, d_acc + size_t(i) * n_env,
                                                    d_tr ? d_tr + size_t(i) * n_env : nullptr);
                if (launched < 0) { cleanup(); return launched; }
                if (!launched) {
                    for (int64_t j = i; j < i + seg; ++j) {
                        DeltaArgs a = make_args(c, n_env, 0, d_ids, d_ord + size_t(j) * n_env, RULE_DBS, c->d_results);
                        a.log_accept = d_acc + size_t(j) * n_env;
                        a.log_psnr = d_tr ? d_tr + size_t(j) * n_env : nullptr;
                        launch_eval(c, a);
                        launch_commit(c, a);
                    }
                }
                i += seg;
                since_resync += seg;
                if (resync_every > 0 && since_resync >= resync_every && base + i < n) {
                    for (int e = 0; e < n_env; ++e)
                        if (int rc = propagate_env(c, ids[e])) { cleanup(); return rc; }
                    since_resync = 0;
                }
            }
            if (c->d_ring) {                                   // a barrier of the persistent kernel gave up?
                int flag = 0;
                BH_DB(cudaMemcpyAsync(&flag, c->d_ring + size_t(c->E) * ROLLOUT_RING * 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                BH_DB(cudaStreamSynchronize(c->stream));
                if (flag) {
                    cudaMemsetAsync(c->d_ring + size_t(c->E) * ROLLOUT_RING * 2, 0, sizeof(unsigned long long), c->stream);
                    cleanup();
                    BH_FAIL(c, -5, "rollout kernel aborted at a barrier");
                }
            }
            BH_DB(cudaGetLastError());
            BH_DB(cudaMemcpyAsync(h_acc, d_acc, size_t(m) * n_env, cudaMemcpyDeviceToHost, c->stream));
            if (d_tr) BH_DB(cudaMemcpyAsync(h_tr, d_tr, size_t(m) * n_env * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            BH_DB(cudaStreamSynchronize(c->stream));
            for (int64_t i = 0; i < m; ++i)
                for (int e = 0; e < n_env; ++e) {
                    accepted[size_t(e) * n + base + i] = h_acc[size_t(i) * n_env + e];
                    if (psnr_trace) psnr_trace[size_t(e) * n + base + i] = h_tr[size_t(i) * n_env + e];
                }
        }
    }
#undef BH_DB
    cleanup();
    for (int e = 0; e < n_env; ++e) {
        if (n_accepted) {
            int64_t k = 0;
            for (int64_t i = 0; i < n; ++i) k += accepted[size_t(e) * n + i];
            n_accepted[e] = k;
        }
        if (final_psnr)
            if (int rc = bh_get_metrics(c, ids[e], final_psnr + e, nullptr, nullptr)) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// exhaustive sweep by correlation
// ---------------------------------------------------------------------------
static int sweep_setup(bh_ctx* c) {
    if (c->dK3) return 0;
    const size_t p2 = size_t(c->P) * c->P, n2 = c->n2;
    BH_CUDA(c, cudaMalloc(&c->dK3, size_t(c->G) * p2 * sizeof(float2)));
    BH_CUDA(c, cudaMalloc(&c->dK4, size_t(c->G) * p2 * sizeof(float2)));
    BH_CUDA(c, cudaMalloc(&c->dK5, size_t(c->G) * p2 * sizeof(float2)));
    if (c->pad == 2) {
        BH_CUDA(c, cudaMalloc(&c->dK6, size_t(c->G) * p2 * sizeof(float2)));
        BH_CUDA(c, cudaMalloc(&c->dsw_buf, size_t(c->Fg) * p2 * sizeof(float2)));
    }
    BH_CUDA(c, cudaMalloc(&c->dsw_in, size_t(c->Fg) * n2 * sizeof(float2)));
    BH_CUDA(c, cudaMalloc(&c->dsw_out, size_t(c->Fg) * n2 * sizeof(float2)));
    BH_CUDA(c, cudaMalloc(&c->dsw_it, size_t(c->Fg) * n2 * sizeof(float)));
    BH_CUDA(c, cudaMalloc(&c->dsw_ii, size_t(c->Fg) * n2 * sizeof(float)));
    BH_CUDA(c, cudaMalloc(&c->dsw_psnr, size_t(c->Fg) * n2 * sizeof(double)));
    c->m4.assign(c->G, 0.0);
    for (int g = 0; g < c->G; ++g) {
        auto t = get_tables(c->P, c->wl[g], c->dx, c->z, c->method);
        auto sw = build_sweep_tables(*t);
        c->m4[g] = sw->m4;
        BH_CUDA(c, cudaMemcpy(c->dK3 + size_t(g) * p2, sw->K3.data(), p2 * sizeof(float2), cudaMemcpyHostToDevice));
        BH_CUDA(c, cudaMemcpy(c->dK4 + size_t(g) * p2, sw->K4.data(), p2 * sizeof(float2), cudaMemcpyHostToDevice));
        BH_CUDA(c, cudaMemcpy(c->dK5 + size_t(g) * p2, sw->K5.data(), p2 * sizeof(float2), cudaMemcpyHostToDevice));
        if (c->pad == 2)
            BH_CUDA(c, cudaMemcpy(c->dK6 + size_t(g) * p2, sw->K6.data(), p2 * sizeof(float2), cudaMemcpyHostToDevice));
    }
    return 0;
}

extern "C" int bh_sweep_all(bh_ctx* c, int env, double* psnr_after, int on_host) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (!psnr_after) BH_FAIL(c, -1, "psnr_after is null");
    if (c->Fg % 2) BH_FAIL(c, -4, "bh_sweep_all needs an even number of frames per colour group");
    if (int rc = sweep_setup(c)) return rc;
    const size_t n2 = c->n2, p2 = size_t(c->P) * c->P;
    const int Fg = c->Fg;
    const float iF = 1.f / float(Fg), iF2 = iF * iF;
    const dim3 blk(256);
    const int gx = int(std::min<size_t>((n2 + 255) / 256, 148 * 8));
    cudaStream_t st = c->stream;
    for (int g = 0; g < c->G; ++g) {
        const float2* U = c->dU + (size_t(env) * c->F + size_t(g) * Fg) * n2;
        const float* I = c->dI + (size_t(env) * c->G + g) * n2;
        const float* T = c->dT + (size_t(env) * c->G + g) * n2;
        const int8_t* stt = c->dstate + (size_t(env) * c->F + size_t(g) * Fg) * n2;
        const float2 *Hh = c->dH + size_t(g) * p2, *K3 = c->dK3 + size_t(g) * p2,
                     *K4 = c->dK4 + size_t(g) * p2, *K5 = c->dK5 + size_t(g) * p2;
        BH_CUDA(c, cudaMemsetAsync(c->dsw_it, 0, size_t(Fg) * n2 * sizeof(float), st));
        BH_CUDA(c, cudaMemsetAsync(c->dsw_ii, 0, size_t(Fg) * n2 * sizeof(float), st));
        auto corr = [&](int mode, const float* A, const float* B, int planes, const float2* K) -> int {
            k_sweep_prep<<<dim3(gx, planes), blk, 0, st>>>(U, A, B, c->dsw_in, n2, mode);
            bool ok = false;
            // pad = 2: the N x N plane is embedded in the P x P canvas by pass A and pass C returns the
            // window, i.e. exactly the window-restricted correlation the linear propagation needs
            BH_CUDA(c, (dispatch_prop<float2, true>(c->P, c->pad, c->dsw_in, c->pad == 1 ? c->dsw_out : c->dsw_buf,
                                                    c->dsw_out, K, c->dtw, planes, planes, st, &ok)));
            if (!ok) BH_FAIL(c, -4, "unsupported FFT side P=%d", c->P);
            c->launches += 4;
            return 0;
        };
        auto acc = [&](int mode, int planes, int to_ii, int use_sign, float c1, float c2, float cst) {
            k_sweep_acc<<<dim3(gx, planes), blk, 0, st>>>(c->dsw_out, c->dsw_it, c->dsw_ii, stt, n2, Fg, mode,
                                                         to_ii, use_sign, c1, c2, cst);
            c->launches += 1;
        };
        int rc;
        if ((rc = corr(PREP_UA, T, nullptr, Fg, Hh))) return rc;          // C1
        acc(ACC_RE, Fg, 0, 1, 2.f * iF, 0.f, 0.f);
        if ((rc = corr(PREP_UA, I, nullptr, Fg, Hh))) return rc;          // C2
        acc(ACC_RE, Fg, 1, 1, 4.f * iF, 0.f, 0.f);
        if ((rc = corr(PREP_U, nullptr, nullptr, Fg, K3))) return rc;     // C3
        acc(ACC_RE, Fg, 1, 1, 4.f * iF2, 0.f, 0.f);
        if ((rc = corr(PREP_U2, nullptr, nullptr, Fg, K5))) return rc;    // C5
        acc(ACC_RE, Fg, 1, 0, 2.f * iF2, 0.f, 0.f);
        if ((rc = corr(PREP_ABS2_PAIR, nullptr, nullptr, Fg / 2, K4))) return rc;   // C4, two frames per plane
        acc(ACC_PAIR, Fg / 2, 1, 0, 2.f * iF2, 0.f, 0.f);
        if ((rc = corr(PREP_TI, T, I, 1, K4))) return rc;                 // BT + i BI
        if (c->pad == 1) {
            acc(ACC_GROUP, Fg, 1, 0, iF, 2.f * iF, float(c->m4[g] * double(iF2)));    // sum |h|^4 is a constant
        } else {
            acc(ACC_GROUP, Fg, 1, 0, iF, 2.f * iF, 0.f);
            if ((rc = corr(PREP_ONES, nullptr, nullptr, 1, c->dK6 + size_t(g) * p2))) return rc;   // window * |h|^4
            acc(ACC_GROUP_RE, Fg, 1, 0, iF2, 0.f, 0.f);
        }
        const size_t count = size_t(Fg) * n2;
        double* dst = on_host ? c->dsw_psnr : psnr_after + size_t(g) * count;
        k_sweep_final<<<int(std::min<size_t>((count + 255) / 256, 148 * 16)), blk, 0, st>>>(
            c->dsw_it, c->dsw_ii, c->dsums + size_t(env) * 4, dst, count, double(c->G) * double(n2), c->relative);
        c->launches += 1;
        BH_CUDA(c, cudaGetLastError());
        if (on_host) {
            BH_CUDA(c, cudaMemcpyAsync(psnr_after + size_t(g) * count, c->dsw_psnr, count * sizeof(double),
                                       cudaMemcpyDeviceToHost, st));
            BH_CUDA(c, cudaStreamSynchronize(st));
        }
    }
    return 0;
}

extern "C" int bh_sweep_stats(bh_ctx* c, int env, const float* pre_model, const double* edges,
                              int64_t* attempted, int64_t* improved, double* gains, double* psnr_after) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (!pre_model || !edges || !attempted || !improved || !gains) BH_FAIL(c, -1, "bad arguments");
    const size_t count = size_t(c->F) * c->n2;
    // scratch (~450 MB at 1024^2 x 24) is allocated on first use and kept for the life of the context
    double*& d_map = c->d_stat_map; float*& d_pre = c->d_stat_pre; unsigned long long*& d_out = c->d_stat_out;
    auto cleanup = [&]() {};
#define BH_ST(expr)                                                          \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) {                                             \
            cleanup();                                                       \
            BH_FAIL(c, -2, "%s failed: %s", #expr, cudaGetErrorString(_e));  \
        }                                                                    \
    } while (0)
    if (!d_map) BH_ST(cudaMalloc(&d_map, count * sizeof(double)));
    if (!d_pre) BH_ST(cudaMalloc(&d_pre, count * sizeof(float)));
    if (!d_out) BH_ST(cudaMalloc(&d_out, 3 * N_BINS * sizeof(unsigned long long)));
    BH_ST(cudaMemcpyAsync(d_pre, pre_model, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    BH_ST(cudaMemsetAsync(d_out, 0, 3 * N_BINS * sizeof(unsigned long long), c->stream));
    double previous = 0.0;
    if (int rc = bh_get_metrics(c, env, &previous, nullptr, nullptr)) { cleanup(); return rc; }
    if (int rc = bh_sweep_all(c, env, d_map, 0)) { cleanup(); return rc; }
    BinEdges be;
    for (int i = 0; i <= N_BINS; ++i) be.e[i] = edges[i];
    k_sweep_stats<<<148 * 8, 256, 0, c->stream>>>(d_map, d_pre, count, previous, be, d_out);
    c->launches += 1;
    BH_ST(cudaGetLastError());
    unsigned long long h_out[3 * N_BINS];
    BH_ST(cudaMemcpyAsync(h_out, d_out, sizeof h_out, cudaMemcpyDeviceToHost, c->stream));
    if (psnr_after) BH_ST(cudaMemcpyAsync(psnr_after, d_map, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BH_ST(cudaStreamSynchronize(c->stream));
#undef BH_ST
    for (int i = 0; i < N_BINS; ++i) {
        attempted[i] = (int64_t)h_out[i];
        improved[i] = (int64_t)h_out[N_BINS + i];
        gains[i] = double((long long)h_out[2 * N_BINS + i]) * FIX_INV;
    }
    cleanup();
    return 0;
}

extern "C" void* bh_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        g_err = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}

extern "C" int bh_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) { g_err = "cudaFreeHost failed"; return -2; }
    return 0;
}

extern "C" int bh_get_recon(bh_ctx* c, int env, float* out, int on_host, int64_t cand) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (!out) BH_FAIL(c, -1, "out is null");
    const size_t n2 = c->n2, bytes = size_t(c->G) * n2 * sizeof(float);
    const float* I = c->dI + size_t(env) * c->G * n2;
    if (cand < 0) {                       // committed reconstruction: one copy, no staging
        BH_CUDA(c, cudaMemcpyAsync(out, I, bytes, on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
        if (on_host) BH_CUDA(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
    if (int rc = check_actions(c, &cand, 1)) return rc;
    float* dst = on_host ? c->drecon : out;
    BH_CUDA(c, cudaMemcpyAsync(dst, I, bytes, cudaMemcpyDeviceToDevice, c->stream));
    const int f = int(cand / (long long)n2);
    const int pix = int(cand - (long long)f * (long long)n2);
    const int r = pix / c->N, col = pix % c->N, g = f / c->Fg;
    // the sign of the flip is read from the resident state by the kernel (no host round trip)
    k_recon_candidate<<<std::min<size_t>((n2 + 255) / 256, 148 * 8), 256, 0, c->stream>>>(
        c->dU + (size_t(env) * c->F + f) * n2, c->dh + size_t(g) * c->P * h_stride(c->P),
        c->dstate + (size_t(env) * c->F + f) * n2 + pix, dst + size_t(g) * n2, c->N, c->P, r, col, c->Fg);
    BH_CUDA(c, cudaGetLastError());
    c->launches += 1;
    if (on_host) {
        BH_CUDA(c, cudaMemcpyAsync(out, c->drecon, bytes, cudaMemcpyDeviceToHost, c->stream));
        BH_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// ---------------------------------------------------------------------------
// batched observation path
// ---------------------------------------------------------------------------
extern "C" void* bh_recon_device_block(bh_ctx* c, int buffer) {
    if (!c || buffer < 0 || buffer >= RECON_MAX_BUFFERS) return nullptr;
    if (cudaSetDevice(c->device) != cudaSuccess) return nullptr;
    if (!c->d_recon_obs[buffer]) {
        const size_t bytes = size_t(c->E) * c->G * c->n2 * sizeof(float);
        if (cudaMalloc(&c->d_recon_obs[buffer], bytes) != cudaSuccess) { c->err = "cudaMalloc of the observation block failed"; return nullptr; }
        cudaMemsetAsync(c->d_recon_obs[buffer], 0, bytes, c->stream);
        // a new block holds nothing: every plane of every environment is stale in it
        std::vector<uint8_t> ones(size_t(c->E) * RECON_MAX_BUFFERS);
        cudaMemcpyAsync(ones.data(), c->d_recon_stale, ones.size(), cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        for (int e = 0; e < c->E; ++e) ones[size_t(e) * RECON_MAX_BUFFERS + buffer] = 0xff;
        cudaMemcpyAsync(c->d_recon_stale, ones.data(), ones.size(), cudaMemcpyHostToDevice, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    return c->d_recon_obs[buffer];
}

extern "C" int bh_recon_batch(bh_ctx* c, int n, const int32_t* env_ids, const bh_result* d_results,
                              float* out, int out_kind, int buffer, int flags) {
    BH_CHECK_CTX(c);
    if (n < 0 || n > c->max_tasks) BH_FAIL(c, -3, "n=%d exceeds max_tasks=%d", n, c->max_tasks);
    if (buffer < 0 || buffer >= RECON_MAX_BUFFERS) BH_FAIL(c, -1, "buffer %d outside [0,%d)", buffer, RECON_MAX_BUFFERS);
    if (n == 0) return 0;
    float* dst = nullptr;
    if (out_kind == BH_OBS_DEVICE) {
        dst = out;
    } else if (out_kind == BH_OBS_PINNED_HOST) {
        if (!out) BH_FAIL(c, -1, "out is null");
        BH_CUDA(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&dst), out, 0));
    } else if (out_kind == BH_OBS_CONTEXT) {
        dst = static_cast<float*>(bh_recon_device_block(c, buffer));
    } else {
        BH_FAIL(c, -1, "bad out_kind %d", out_kind);
    }
    if (!dst) BH_FAIL(c, -1, "no observation block");
    ReconArgs a;
    a.U = c->dU; a.I = c->dI; a.h = c->dh; a.state = c->dstate;
    a.results = (flags & BH_OBS_COMMITTED_ONLY) ? nullptr
              : (d_results ? reinterpret_cast<const Result*>(d_results) : c->d_results);
    a.envs = nullptr; a.n_inline = 0;
    if (env_ids) {
        for (int i = 0; i < n; ++i) BH_CHECK_ENV(c, env_ids[i]);
        if (n <= INLINE_MAX) {
            a.n_inline = n;
            for (int i = 0; i < n; ++i) a.inl_envs[i] = env_ids[i];
        } else {
            // the staging arrays are free here: every host-array entry point synchronises before returning
            std::memcpy(c->h_envs, env_ids, size_t(n) * sizeof(int32_t));
            BH_CUDA(c, cudaMemcpyAsync(c->d_envs, c->h_envs, size_t(n) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            a.envs = c->d_envs;
        }
    } else if (n > c->E) {
        BH_FAIL(c, -3, "n=%d tasks but only %d environments", n, c->E);
    }
    a.stale = c->d_recon_stale; a.plan = c->d_recon_plan; a.out = dst;
    a.planes_written = reinterpret_cast<unsigned long long*>(c->d_scalars + 2);
    a.n_tasks = n; a.E = c->E; a.N = c->N; a.P = c->P; a.HP = h_stride(c->P); a.F = c->F; a.G = c->G; a.Fg = c->Fg;
    a.buffer = buffer; a.full = (flags & BH_OBS_FULL) ? 1 : 0;
    k_recon_plan<<<1, 256, 0, c->stream>>>(a);
    const int chunks = int(std::min<size_t>((c->n2 + 1023) / 1024, 96));
    k_recon_batch<<<dim3(chunks, c->G, n), 256, 0, c->stream>>>(a);
    BH_CUDA(c, cudaGetLastError());
    c->launches += 2;
    if (flags & BH_OBS_SYNC) {
        BH_CUDA(c, cudaEventRecord(c->ev_recon, c->stream));
        BH_CUDA(c, cudaEventSynchronize(c->ev_recon));
    }
    return 0;
}

extern "C" int64_t bh_recon_planes_written(bh_ctx* c) {
    if (!c || cudaSetDevice(c->device) != cudaSuccess) return -1;
    long long v = 0;
    if (cudaMemcpyAsync(&v, c->d_scalars + 2, sizeof v, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    return v;
}

extern "C" int bh_stream_sync(bh_ctx* c) {
    BH_CHECK_CTX(c);
    BH_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int bh_get_state(bh_ctx* c, int env, int8_t* out, int on_host) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (!out) BH_FAIL(c, -1, "out is null");
    const size_t bytes = size_t(c->F) * c->n2;
    BH_CUDA(c, cudaMemcpyAsync(out, c->dstate + size_t(env) * bytes, bytes,
                               on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    if (on_host) BH_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int bh_get_field(bh_ctx* c, int env, int frame, float* out, int on_host) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (!out || frame < 0 || frame >= c->F) BH_FAIL(c, -1, "bad arguments");
    const size_t bytes = c->n2 * sizeof(float2);
    BH_CUDA(c, cudaMemcpyAsync(out, c->dU + (size_t(env) * c->F + frame) * c->n2, bytes,
                               on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    if (on_host) BH_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// stand-alone tt.simulate operator
// ---------------------------------------------------------------------------
__global__ void k_real_to_complex(const float* __restrict__ in, float2* __restrict__ out, size_t n) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        out[i] = make_float2(in[i], 0.f);
}

// device tables of the stand-alone operator, kept for the life of the process: a script that
// calls tt.simulate every step (the reference's loops) must not rebuild and upload 8 MB per call
struct SimTables { float2* dH = nullptr; float2* dtw = nullptr; };
static int sim_tables(int device, int P, double wl, double dx, double z, int method, SimTables* out) {
    using Key = std::tuple<int, int, double, double, double, int>;
    static std::mutex mu;
    static std::map<Key, SimTables> cache;
    std::lock_guard<std::mutex> lk(mu);
    const Key key(device, P, wl, dx, z, method);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
    bh_ctx* nul = nullptr;
    const size_t p2 = size_t(P) * P;
    auto t = get_tables(P, wl, dx, z, method);
    auto tw = build_twiddles(P);
    SimTables st;
    BH_CUDA(nul, cudaMalloc(&st.dH, p2 * sizeof(float2)));
    BH_CUDA(nul, cudaMalloc(&st.dtw, (tw.size() / 2 + 1) * sizeof(float2)));
    BH_CUDA(nul, cudaMemcpy(st.dH, t->H.data(), p2 * sizeof(float2), cudaMemcpyHostToDevice));
    BH_CUDA(nul, cudaMemcpy(st.dtw, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice));
    cache[key] = st;
    *out = st;
    return 0;
}

// Work buffers of the stand-alone operator, per device, grown on demand and kept: a script that
// calls tt.simulate every step must not pay cudaMalloc / cudaFree (each a device synchronisation).
struct SimPool {
    void* p[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t cap[4] = {0, 0, 0, 0};
};
static cudaError_t sim_pool_get(int device, int slot, size_t bytes, void** out) {
    static std::map<int, SimPool> pools;          // guarded by the caller's lock
    SimPool& sp = pools[device];
    if (sp.cap[slot] < bytes) {
        if (sp.p[slot]) cudaFree(sp.p[slot]);
        sp.p[slot] = nullptr; sp.cap[slot] = 0;
        cudaError_t e = cudaMalloc(&sp.p[slot], bytes);
        if (e != cudaSuccess) return e;
        sp.cap[slot] = bytes;
    }
    *out = sp.p[slot];
    return cudaSuccess;
}

extern "C" int bh_simulate(int device, void* stream_, const float* in, int is_complex, int C, int N,
                           double wl, double dx, double z, int pad, int method, float* out, int on_host) {
    bh_ctx* nul = nullptr;
    if (!in || !out || C < 1 || N < 8) BH_FAIL(nul, -1, "bad arguments");
    if (pad != 1 && pad != 2) BH_FAIL(nul, -1, "pad must be 1 or 2");
    const int P = N * pad;
    if (!fft_side_supported(P, pad)) BH_FAIL(nul, -4, "unsupported FFT side P=%d (pad=%d)", P, pad);
    BH_CUDA(nul, cudaSetDevice(device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
    const size_t n2 = size_t(N) * N, p2 = size_t(P) * P, cnt = size_t(C) * n2;
    SimTables tabs;
    if (int rc = sim_tables(device, P, wl, dx, z, method, &tabs)) return rc;
    float2 *dH = tabs.dH, *dtw = tabs.dtw;
    static std::mutex pool_mu;                     // the call synchronises before it returns, so one
    std::lock_guard<std::mutex> lk(pool_mu);       // caller at a time owns the pooled buffers
    float2 *din = nullptr, *dbuf = nullptr, *dout = nullptr;
    float* draw = nullptr;
    BH_CUDA(nul, sim_pool_get(device, 0, cnt * sizeof(float2), reinterpret_cast<void**>(&din)));
    const cudaMemcpyKind kin = on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (is_complex) {
        BH_CUDA(nul, cudaMemcpyAsync(din, in, cnt * sizeof(float2), kin, st));
    } else {
        BH_CUDA(nul, sim_pool_get(device, 1, cnt * sizeof(float), reinterpret_cast<void**>(&draw)));
        BH_CUDA(nul, cudaMemcpyAsync(draw, in, cnt * sizeof(float), kin, st));
        k_real_to_complex<<<std::min<size_t>((cnt + 255) / 256, 148 * 16), 256, 0, st>>>(draw, din, cnt);
    }
    float2* U = reinterpret_cast<float2*>(out);
    if (on_host) { BH_CUDA(nul, sim_pool_get(device, 2, cnt * sizeof(float2), reinterpret_cast<void**>(&dout))); U = dout; }
    float2* buf = U;
    if (pad == 2) { BH_CUDA(nul, sim_pool_get(device, 3, size_t(C) * p2 * sizeof(float2), reinterpret_cast<void**>(&dbuf))); buf = dbuf; }
    bool ok = false;
    BH_CUDA(nul, (dispatch_prop<float2, true>(P, pad, din, buf, U, dH, dtw, C, C, st, &ok)));
    if (on_host) BH_CUDA(nul, cudaMemcpyAsync(out, dout, cnt * sizeof(float2), cudaMemcpyDeviceToHost, st));
    BH_CUDA(nul, cudaStreamSynchronize(st));
    return 0;
}

// ---------------------------------------------------------------------------
// timing hooks (CUDA events on the context stream)
// ---------------------------------------------------------------------------
extern "C" int bh_time_eval(bh_ctx* c, int n, const int32_t* d_env_ids, const int64_t* d_actions,
                            int n_sets, int reps, float* ms_per_launch) {
    BH_CHECK_CTX(c);
    if (n < 1 || n > c->max_tasks || reps < 1 || n_sets < 1 || !ms_per_launch) BH_FAIL(c, -1, "bad arguments");
    DeltaArgs a = make_args(c, n, 0, d_env_ids, reinterpret_cast<const long long*>(d_actions), RULE_NEVER, c->d_results);
    // no env ids: a window of candidates of environment 0, scored like a DBS speculation window
    a.sort_window = d_env_ids ? 0 : 1;
    auto launch = [&]() { return d_env_ids ? launch_eval(c, a) : launch_eval_list(c, a, false); };
    for (int i = 0; i < 3; ++i) launch();                    // warm
    BH_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i) {
        a.actions = reinterpret_cast<const long long*>(d_actions) + size_t(i % n_sets) * n;
        launch();
    }
    BH_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    BH_CUDA(c, cudaEventSynchronize(c->ev1));
    BH_CUDA(c, cudaGetLastError());
    float ms = 0.f;
    BH_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_per_launch = ms / float(reps);
    return 0;
}

// eval (+ commit) chains as the step path launches them: `reps` vectorised steps over n_sets action
// sets, device-resident inputs; with_commit = 0 times the evaluation alone under the same rule.
extern "C" int bh_time_step(bh_ctx* c, int n, const int32_t* d_env_ids, const int64_t* d_actions,
                            int n_sets, int reps, int rule, int with_commit, float* ms_per_step) {
    BH_CHECK_CTX(c);
    if (n < 1 || n > c->max_tasks || reps < 1 || n_sets < 1 || !ms_per_step || !d_env_ids) BH_FAIL(c, -1, "bad arguments");
    if (rule < 0 || rule > 3) BH_FAIL(c, -1, "bad rule %d", rule);
    DeltaArgs a = make_args(c, n, 0, d_env_ids, reinterpret_cast<const long long*>(d_actions), rule, c->d_results);
    auto step = [&](int i) {
        a.actions = reinterpret_cast<const long long*>(d_actions) + size_t(i % n_sets) * n;
        launch_eval(c, a);
        if (with_commit) launch_commit(c, a);
    };
    for (int i = 0; i < 3; ++i) step(i);
    BH_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i) step(i);
    BH_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    BH_CUDA(c, cudaEventSynchronize(c->ev1));
    BH_CUDA(c, cudaGetLastError());
    float ms = 0.f;
    BH_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_per_step = ms / float(reps);
    return 0;
}

// k_commit alone: n_sets result sets are prepared once (every flip marked "keep"), then `reps`
// commit launches run back to back, launch i applying set i % n_sets (n accepted flips, 24 N^2 B
// each).  The fields drift by the repeated additions; the environments are re-propagated from
// their (consistent) state afterwards.
extern "C" int bh_time_commit(bh_ctx* c, int n, const int32_t* d_env_ids, const int64_t* d_actions,
                              int n_sets, int reps, float* ms_per_launch) {
    BH_CHECK_CTX(c);
    if (n < 1 || n > COMMIT_MAX_TASKS || reps < 1 || n_sets < 1 || !ms_per_launch || !d_env_ids) BH_FAIL(c, -1, "bad arguments");
    if (size_t(n) * n_sets > size_t(c->max_tasks)) BH_FAIL(c, -1, "n * n_sets exceeds max_tasks");
    DeltaArgs a = make_args(c, n, 0, d_env_ids, reinterpret_cast<const long long*>(d_actions), RULE_ALWAYS, c->d_results);
    for (int i = 0; i < n_sets; ++i) {
        DeltaArgs e = a;
        e.actions = reinterpret_cast<const long long*>(d_actions) + size_t(i) * n;
        e.results = c->d_results + size_t(i) * n;
        e.acc = c->d_acc + size_t(i) * n * 2; e.tickets = c->d_tickets + size_t(i) * n;
        launch_eval(c, e);
    }
    auto commit = [&](int i) {
        DeltaArgs k = a;
        k.results = c->d_results + size_t(i % n_sets) * n;
        launch_commit(c, k);
    };
    for (int i = 0; i < 3; ++i) commit(i);
    BH_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i) commit(i);
    BH_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    BH_CUDA(c, cudaEventSynchronize(c->ev1));
    BH_CUDA(c, cudaGetLastError());
    float ms = 0.f;
    BH_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_per_launch = ms / float(reps);
    for (int e = 0; e < c->E; ++e)
        if (int rc = propagate_env(c, e)) return rc;
    BH_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int bh_time_propagate_passes(bh_ctx* c, int env, int reps, float* ms4) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (reps < 1 || !ms4) BH_FAIL(c, -1, "bad arguments");
    if (int rc = propagate_env(c, env)) return rc;
    for (int i = 0; i < 4; ++i) ms4[i] = 0.f;
    for (int i = 0; i < reps; ++i)
        if (int rc = propagate_env(c, env, ms4)) return rc;
    for (int i = 0; i < 4; ++i) ms4[i] /= float(reps);
    return 0;
}

extern "C" int bh_time_propagate(bh_ctx* c, int env, int reps, float* ms_per_launch) {
    BH_CHECK_CTX(c); BH_CHECK_ENV(c, env);
    if (reps < 1 || !ms_per_launch) BH_FAIL(c, -1, "bad arguments");
    if (int rc = propagate_env(c, env)) return rc;
    BH_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i)
        if (int rc = propagate_env(c, env)) return rc;
    BH_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    BH_CUDA(c, cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    BH_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_per_launch = ms / float(reps);
    return 0;
}
