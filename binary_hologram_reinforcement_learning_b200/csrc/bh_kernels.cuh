// CUDA kernels of the hologram reward / DBS hot path (sm_100a).
//
//  propagation (reset / re-sync), per colour group:
//      k_rows_fwd -> k_cols -> k_rows_inv_group (inverse rows + intensity + loss partials) ; then k_loss_final
//      (k_rows_inv alone is the tail of the stand-alone tt.simulate operator)
//      restates tt.simulate + .abs()**2 + mean(dim=1) + tt.relativeLoss
//      (reference env.py:123-132, env_1024_24.py:149-166)
//  incremental path (every step / candidate):  k_eval -> k_commit
//      replaces the per-step full re-simulation of env.py:170-174,
//      DBS.py:259-270, DBS_1024_24.py:324-352, env_group.py:96-119 by the
//      delta identity U' = U + s * shift(h)   (SURVEY.md 8c).
#pragma once
#include <cstdint>
#include "bh_fft.cuh"

namespace bh {

// ---------------------------------------------------------------------------
// data layout in HBM (per context; E environments)
//   U     float2 [E][F][N][N]   propagated field of every frame
//   I     float  [E][G][N][N]   frame-averaged intensity per colour group
//   T     float  [E][G][N][N]   target
//   state int8   [E][F][N][N]   binary hologram
//   sums  double [E][4]         sum I^2, sum I*T, sum T^2, current PSNR
//   H     float2 [G][P][P]      transfer function, pre-scaled by 1/P^2
//   h     float2 [G][P][P]      impulse response ifft2(H)
// ---------------------------------------------------------------------------

struct Result {            // mirrored by bh_result in include/bholo.h (40 bytes)
    double psnr_after;
    double d_sii;
    double d_sit;
    long long action;
    int32_t accept;
    int32_t sgn;
};

enum { RULE_ENV = 0, RULE_DBS = 1, RULE_NEVER = 2 };

// Tile shape knobs (sequences per tile, CTAs per SM the compiler must allow).  Tuned on B200 at
// P = 1024 (profiles/r1_notes.md): narrow tiles with a large register budget beat wide tiles at
// the 64-register cap -- the passes are latency bound, and independent small CTAs overlap their
// load / barrier phases.
#ifndef BH_ROWS_W
#define BH_ROWS_W 2
#endif
#ifndef BH_ROWS_MINB
#define BH_ROWS_MINB 4
#endif
#ifndef BH_COLS_W
#define BH_COLS_W 8
#endif
#ifndef BH_COLS_MINB
#define BH_COLS_MINB 1
#endif
constexpr int TILE_W = BH_ROWS_W;  // sequences per FFT tile of the row passes

constexpr int ilog2_c(int v) { return v <= 1 ? 0 : 1 + ilog2_c(v >> 1); }

// threads: P/16 per sequence at P >= 896 (16 complex values in registers per thread and pass)
template <int P> struct FftCfg {
    // threads per sequence: one radix-32 butterfly per thread for the two-pass plans (896 = 32 x 28,
    // 1024 = 32 x 32), P/32 = 64 for the three-pass plans above them
    static constexpr int Q = (P > 1024) ? 64 : 32;
    static constexpr int T = TILE_W * Q;                // threads per CTA, row passes
    static constexpr int MINB = (P <= 1024) ? BH_ROWS_MINB : 1;
    static constexpr int SKR = ilog2_c(Plan<P>::r[0]) > 4 ? ilog2_c(Plan<P>::r[0]) : 4;   // row layout: pad after
                                                        // every max(16, R0) elements
    static constexpr int SKC = ilog2_c(Plan<P>::r[0]);  // column layout: pad after every R0 elements
    static constexpr int SEQ = SeqLen<P, SKR>::value;   // row layout: stride between sequences
    static constexpr size_t smem_row = size_t(SEQ) * TILE_W * sizeof(float2);
    static constexpr int WC = (P > 1024) ? 4 : BH_COLS_W;   // columns per tile of the column pass (32 values per
                                                            // thread at P = 2048: 4 columns x 64 threads, 255 regs)
    static constexpr int TC = WC * Q;
    static constexpr int MINBC = (P <= 1024) ? BH_COLS_MINB : 1;
    static constexpr size_t smem_col = size_t(SeqLen<P, SKC>::value) * WC * sizeof(float2);
};

// ---------------------------------------------------------------------------
// cp.async helpers (LDGSTS): global -> shared without staging registers
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(float2* smem, const float2* gmem) {
    const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------
// pass A: FFT along x of W canvas rows.  in: [frames][N][N] int8 state (or
// complex float2 for the stand-alone operator); out: buf [frames][P][P].
// Row layout: thread (w = tid / Q, q = tid % Q).
// ---------------------------------------------------------------------------
template <int P, int PAD, typename InT, bool IS_CPLX>
__global__ void __launch_bounds__(FftCfg<P>::T, FftCfg<P>::MINB)
k_rows_fwd(const InT* __restrict__ in, float2* __restrict__ buf, const float2* __restrict__ tw) {
    using C = FftCfg<P>;
    constexpr int N = P / PAD, O = (P - N) / 2, T = C::T, Q = C::Q, SK = C::SKR, SEQ = C::SEQ;
    extern __shared__ float2 s[];
    const int tid = threadIdx.x, f = blockIdx.y;
    float2* out = buf + size_t(f) * P * P;
    if constexpr (!IS_CPLX) {
        // Real input: canvas rows 2w and 2w+1 of the tile ride as real and imaginary part of ONE
        // complex sequence z = a + i b; one FFT gives both spectra, A[k] = (Z[k] + conj Z[-k]) / 2 and
        // B[k] = (Z[k] - conj Z[-k]) / (2i).  A tile of W sequences covers 2 W rows, so the forward
        // row pass of the binary state costs half the butterflies, and only kx <= P/2 is stored
        // (the rest follows from Hermitian symmetry, see k_cols_herm).  16 pixels per 128-bit load.
        const int Y0 = blockIdx.x * (2 * TILE_W);
        constexpr int NXS = (P + T - 1) / T;
        if (PAD > 1 && (Y0 + 2 * TILE_W <= O || Y0 >= O + N)) {      // all-zero canvas rows
            for (int i = tid; i < 2 * TILE_W * (P / 2 + 1); i += T) {
                const int r = i / (P / 2 + 1), X = i - r * (P / 2 + 1);
                out[size_t(Y0 + r) * P + X] = make_float2(0.f, 0.f);
            }
            return;
        }
        if (PAD > 1) {                                   // zero margins of the canvas rows
#pragma unroll
            for (int w = 0; w < TILE_W; ++w)
                for (int i = tid; i < P - N; i += T) {
                    const int X = (i < O) ? i : i + N;
                    s[w * SEQ + padded<SK>(X)] = make_float2(0.f, 0.f);
                }
        }
        constexpr int CH = N / 16;                       // 16-pixel chunks per row
        constexpr int NI = (TILE_W * CH + T - 1) / T;
        int4 ra[NI], rb[NI];
#pragma unroll
        for (int k = 0; k < NI; ++k) {
            const int i = tid + k * T;
            ra[k] = make_int4(0, 0, 0, 0); rb[k] = make_int4(0, 0, 0, 0);
            if ((TILE_W * CH) % T == 0 || i < TILE_W * CH) {
                const int w = i / CH, xc = i - w * CH;
                const int ya = Y0 + 2 * w - O, yb = ya + 1;
                const int8_t* base = reinterpret_cast<const int8_t*>(in) + size_t(f) * N * N + 16 * xc;
                if (PAD == 1 || (ya >= 0 && ya < N)) ra[k] = __ldg(reinterpret_cast<const int4*>(base + size_t(ya) * N));
                if (PAD == 1 || (yb >= 0 && yb < N)) rb[k] = __ldg(reinterpret_cast<const int4*>(base + size_t(yb) * N));
            }
        }
#pragma unroll
        for (int k = 0; k < NI; ++k) {
            const int i = tid + k * T;
            if ((TILE_W * CH) % T == 0 || i < TILE_W * CH) {
                const int w = i / CH, xc = i - w * CH;
                const int wa[4] = {ra[k].x, ra[k].y, ra[k].z, ra[k].w};
                const int wb[4] = {rb[k].x, rb[k].y, rb[k].z, rb[k].w};
                float2* d = s + w * SEQ + padded<SK>(O + 16 * xc);   // 16 | O + 16 xc: never straddles a pad slot
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    d[j] = make_float2(float(int8_t((wa[j >> 2] >> (8 * (j & 3))) & 0xff)),
                                       float(int8_t((wb[j >> 2] >> (8 * (j & 3))) & 0xff)));
            }
        }
        __syncthreads();
        tile_fft<P, Q, 1, SK, false>(s + (tid / Q) * SEQ, tid % Q, tw);
#pragma unroll
        for (int w = 0; w < TILE_W; ++w)
#pragma unroll
            for (int k = 0; k < NXS; ++k) {
                const int X = tid + k * T;
                if ((P % T == 0 || X < P) && X <= P / 2) {       // kx > P/2 follows from Hermitian symmetry
                    const float2 zk = s[w * SEQ + padded<SK>(X)];
                    const float2 zm = s[w * SEQ + padded<SK>(X == 0 ? 0 : P - X)];
                    out[size_t(Y0 + 2 * w) * P + X] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                    out[size_t(Y0 + 2 * w + 1) * P + X] = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
                }
            }
        return;
    }
    const int Y0 = blockIdx.x * TILE_W;
    if (PAD > 1 && (Y0 + TILE_W <= O || Y0 >= O + N)) {     // all-zero canvas rows
        for (int i = tid; i < TILE_W * P; i += T) out[size_t(Y0) * P + i] = make_float2(0.f, 0.f);
        return;
    }
    {
        constexpr int NX = (P + T - 1) / T;
        float2 v[TILE_W][NX];
#pragma unroll
        for (int w = 0; w < TILE_W; ++w) {
            const int y = Y0 + w - O;
            const bool rowok = (y >= 0 && y < N);
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                const int X = tid + k * T, x = X - O;
                v[w][k] = make_float2(0.f, 0.f);
                if ((P % T == 0 || X < P) && rowok && x >= 0 && x < N) {
                    const size_t idx = (size_t(f) * N + y) * N + x;
                    v[w][k] = reinterpret_cast<const float2*>(in)[idx];
                }
            }
        }
#pragma unroll
        for (int w = 0; w < TILE_W; ++w)
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                const int X = tid + k * T;
                if (P % T == 0 || X < P) s[w * SEQ + padded<SK>(X)] = v[w][k];
            }
    }
    __syncthreads();
    tile_fft<P, Q, 1, SK, false>(s + (tid / Q) * SEQ, tid % Q, tw);
    constexpr int NXS = (P + T - 1) / T;
#pragma unroll
    for (int w = 0; w < TILE_W; ++w)
#pragma unroll
        for (int k = 0; k < NXS; ++k) {
            const int X = tid + k * T;
            if (P % T == 0 || X < P) out[size_t(Y0 + w) * P + X] = s[w * SEQ + padded<SK>(X)];
        }
}

// ---------------------------------------------------------------------------
// pass B: for W canvas columns: FFT along y, multiply by H, inverse FFT along y.
// In place on buf.  H is pre-scaled by 1/P^2 so no later normalisation.
// Column layout: thread (w = tid % WC, q = tid / WC), WC = 16 or 8 columns per tile; H
// points at the colour group of the frames of this launch.
// ---------------------------------------------------------------------------
template <int P, int PAD>
__global__ void __launch_bounds__(FftCfg<P>::TC, FftCfg<P>::MINBC)
k_cols(float2* __restrict__ buf, const float2* __restrict__ H, const float2* __restrict__ tw) {
    using C = FftCfg<P>;
    constexpr int N = P / PAD, O = (P - N) / 2, Q = C::Q, SK = C::SKC, TILE_W = C::WC;
    extern __shared__ float2 s[];
    const int tid = threadIdx.x, f = blockIdx.y, X0 = blockIdx.x * TILE_W;
    float2* b = buf + size_t(f) * P * P + X0;
    const float2* Hg = H + X0;
    const int w = tid % TILE_W, q = tid / TILE_W;
    constexpr int NY = (P + Q - 1) / Q;
    {
        float2 v[NY];
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int y = q + k * Q;
            v[k] = make_float2(0.f, 0.f);
            if ((P % Q == 0 || y < P) && (PAD == 1 || (y >= O && y < O + N))) v[k] = b[size_t(y) * P + w];
        }
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int y = q + k * Q;
            if (P % Q == 0 || y < P) s[padded<SK>(y) * TILE_W + w] = v[k];
        }
    }
    __syncthreads();
    tile_fft<P, Q, TILE_W, SK, false>(s + w, q, tw);
    {
        float2 hv[NY];
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int y = q + k * Q;
            if (P % Q == 0 || y < P) hv[k] = __ldg(Hg + size_t(y) * P + w);
        }
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int y = q + k * Q;
            if (P % Q == 0 || y < P) {
                float2* e = s + padded<SK>(y) * TILE_W + w;
                *e = cmul(*e, hv[k]);
            }
        }
    }
    __syncthreads();
    tile_fft<P, Q, TILE_W, SK, true>(s + w, q, tw);
#pragma unroll
    for (int k = 0; k < NY; ++k) {
        const int y = q + k * Q;
        if ((P % Q == 0 || y < P) && (PAD == 1 || (y >= O && y < O + N)))
            b[size_t(y) * P + w] = s[padded<SK>(y) * TILE_W + w];
    }
}

// ---------------------------------------------------------------------------
// pass B for REAL input (the binary state).  Every row spectrum is Hermitian along
// kx, so column P-kx of the row-transformed frame is the complex conjugate of column kx and
// pass A stored only kx in [0, P/2].  One forward column FFT F = FFT_y(col_kx) serves two
// output columns:
//     out[:, kx]   = IFFT_y(F * H[:, kx])
//     out[:, P-kx] = conj( IFFT_y(F * conj(H[:, kx])) )        (H is even in kx and ky)
// i.e. 3 tile FFTs per 16 output columns instead of 4, and half the tile reads.
// grid (P / (2 W) + 1, frames): the last tile holds the Nyquist column kx = P/2 alone.
// ---------------------------------------------------------------------------
template <int P, int PAD>
__global__ void __launch_bounds__(FftCfg<P>::TC, FftCfg<P>::MINBC)
k_cols_herm(float2* __restrict__ buf, const float2* __restrict__ H, const float2* __restrict__ tw) {
    using C = FftCfg<P>;
    constexpr int N = P / PAD, O = (P - N) / 2;         // only window rows carry data / are needed
    constexpr int Q = C::Q, SK = C::SKC, W = C::WC;
    constexpr int NY = (P + Q - 1) / Q;
    constexpr int TILE = SeqLen<P, SK>::value * W;
    extern __shared__ float2 s[];                   // [2][TILE]
    float2* sA = s;
    float2* sB = s + TILE;
    const int tid = threadIdx.x, f = blockIdx.y, X0 = blockIdx.x * W;
    const int w = tid % W, q = tid / W;
    const int kx = X0 + w;
    const bool col_ok = kx <= P / 2;                // only the last tile has invalid columns
    float2* b = buf + size_t(f) * P * P;
    {
        float2 v[NY];
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int y = q + k * Q;
            v[k] = make_float2(0.f, 0.f);
            if ((P % Q == 0 || y < P) && col_ok && (PAD == 1 || (y >= O && y < O + N))) v[k] = b[size_t(y) * P + kx];
        }
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int y = q + k * Q;
            if (P % Q == 0 || y < P) sA[padded<SK>(y) * W + w] = v[k];
        }
    }
    __syncthreads();
    tile_fft<P, Q, W, SK, false>(sA + w, q, tw);
    {
        float2 hv[NY];
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int y = q + k * Q;
            hv[k] = make_float2(0.f, 0.f);
            if ((P % Q == 0 || y < P) && col_ok) hv[k] = __ldg(H + size_t(y) * P + kx);
        }
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const int y = q + k * Q;
            if (P % Q == 0 || y < P) {
                const int a = padded<SK>(y) * W + w;
                const float2 F = sA[a];
                sA[a] = cmul(F, hv[k]);
                sB[a] = cmul(F, make_float2(hv[k].x, -hv[k].y));
            }
        }
    }
    __syncthreads();
    tile_fft<P, Q, W, SK, true>(sA + w, q, tw);
    tile_fft<P, Q, W, SK, true>(sB + w, q, tw);
    const bool mirror_ok = col_ok && kx != 0 && kx != P / 2;
#pragma unroll
    for (int k = 0; k < NY; ++k) {
        const int y = q + k * Q;
        if ((P % Q == 0 || y < P) && (PAD == 1 || (y >= O && y < O + N))) {
            const int a = padded<SK>(y) * W + w;
            if (col_ok) b[size_t(y) * P + kx] = sA[a];
            if (mirror_ok) {
                const float2 m = sB[a];
                b[size_t(y) * P + (P - kx)] = make_float2(m.x, -m.y);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// pass C: inverse FFT along x of W window rows of one frame; writes the field U
// (cropped to the N x N window).  grid (N / W, frames).  buf may alias U when
// PAD == 1 (in place).  This alone is the tail of the tt.simulate operator.
// ---------------------------------------------------------------------------
template <int P, int PAD>
__global__ void __launch_bounds__(FftCfg<P>::T, FftCfg<P>::MINB)
k_rows_inv(const float2* buf, float2* U, const float2* __restrict__ tw) {
    using C = FftCfg<P>;
    constexpr int N = P / PAD, O = (P - N) / 2, T = C::T, Q = C::Q, SK = C::SKR, SEQ = C::SEQ;
    extern __shared__ float2 s[];
    const int tid = threadIdx.x, f = blockIdx.y, y0 = blockIdx.x * TILE_W;
    const float2* src = buf + size_t(f) * P * P + size_t(y0 + O) * P;
    float2* dst = U + size_t(f) * N * N + size_t(y0) * N;
    constexpr int NX = (P + T - 1) / T;
    {
        float2 v[TILE_W][NX];
#pragma unroll
        for (int w = 0; w < TILE_W; ++w)
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                const int X = tid + k * T;
                if (P % T == 0 || X < P) v[w][k] = src[size_t(w) * P + X];
            }
#pragma unroll
        for (int w = 0; w < TILE_W; ++w)
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                const int X = tid + k * T;
                if (P % T == 0 || X < P) s[w * SEQ + padded<SK>(X)] = v[w][k];
            }
    }
    __syncthreads();
    tile_fft<P, Q, 1, SK, true>(s + (tid / Q) * SEQ, tid % Q, tw);
    constexpr int NXO = (N + T - 1) / T;
#pragma unroll
    for (int w = 0; w < TILE_W; ++w)
#pragma unroll
        for (int k = 0; k < NXO; ++k) {
            const int x = tid + k * T;
            if (N % T == 0 || x < N) dst[size_t(w) * N + x] = s[w * SEQ + padded<SK>(x + O)];
        }
}

// ---------------------------------------------------------------------------
// deterministic block reduction helpers (fixed shuffle/tree order)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------
// pass C fused with the reconstruction and the loss sums (reset / re-sync path):
// one CTA owns W window rows of one colour group and walks the group's Fg frames.  While
// frame f is transformed in one shared-memory buffer, the rows of frame f+1 stream into the
// other with cp.async, so the tile load (37 % of the stand-alone pass) hides behind the FFT.
// |U|^2 accumulates in registers across the frames (.abs()**2 + torch.mean(dim=1),
// env.py:172-173), the epilogue writes I and this tile's float64 partial sums of
// tt.relativeLoss (sum I^2, sum I*T, sum T^2); k_loss_final folds them in index order.
// grid (N / W, groups).  buf may alias U when PAD == 1 (frame f+1 is only read before
// frame f+1 is written).
// ---------------------------------------------------------------------------
template <int P, int PAD>
__global__ void __launch_bounds__(FftCfg<P>::T, FftCfg<P>::MINB)
k_rows_inv_group(const float2* buf, float2* U, float* __restrict__ I, const float* __restrict__ Tg,
                 const float2* __restrict__ tw, int Fg, double* __restrict__ partial) {
    using C = FftCfg<P>;
    constexpr int N = P / PAD, O = (P - N) / 2, T = C::T, Q = C::Q, SK = C::SKR, SEQ = C::SEQ;
    constexpr int NX = (P + T - 1) / T, NXO = (N + T - 1) / T;
    constexpr int TILE = SEQ * TILE_W;
    extern __shared__ float2 s[];                   // two tiles
    __shared__ double sh[3][T / 32];
    const int tid = threadIdx.x, g = blockIdx.y, y0 = blockIdx.x * TILE_W;
    const size_t n2 = size_t(N) * N;
    float acc[TILE_W][NXO];
#pragma unroll
    for (int w = 0; w < TILE_W; ++w)
#pragma unroll
        for (int k = 0; k < NXO; ++k) acc[w][k] = 0.f;
    auto prefetch = [&](int fi, float2* dst) {
        const float2* src = buf + size_t(g * Fg + fi) * P * P + size_t(y0 + O) * P;
#pragma unroll
        for (int w = 0; w < TILE_W; ++w)
#pragma unroll
            for (int k = 0; k < NX; ++k) {
                const int X = tid + k * T;
                if (P % T == 0 || X < P) cp_async8(dst + w * SEQ + padded<SK>(X), src + size_t(w) * P + X);
            }
        cp_async_commit();
    };
    prefetch(0, s);
    for (int fi = 0; fi < Fg; ++fi) {
        float2* cur = s + (fi & 1) * TILE;
        if (fi + 1 < Fg) { prefetch(fi + 1, s + ((fi + 1) & 1) * TILE); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        tile_fft<P, Q, 1, SK, true>(cur + (tid / Q) * SEQ, tid % Q, tw);
        float2* dst = U + size_t(g * Fg + fi) * n2 + size_t(y0) * N;
#pragma unroll
        for (int w = 0; w < TILE_W; ++w)
#pragma unroll
            for (int k = 0; k < NXO; ++k) {
                const int x = tid + k * T;
                if (N % T == 0 || x < N) {
                    const float2 v = cur[w * SEQ + padded<SK>(x + O)];
                    dst[size_t(w) * N + x] = v;
                    acc[w][k] = fmaf(v.x, v.x, fmaf(v.y, v.y, acc[w][k]));
                }
            }
        __syncthreads();                            // the buffer is free for the prefetch of frame fi + 2
    }
    const float inv = 1.f / float(Fg);
    double a = 0, b = 0, c = 0;
#pragma unroll
    for (int w = 0; w < TILE_W; ++w)
#pragma unroll
        for (int k = 0; k < NXO; ++k) {
            const int x = tid + k * T;
            if (N % T == 0 || x < N) {
                const size_t p = size_t(g) * n2 + size_t(y0 + w) * N + x;
                const float iv = acc[w][k] * inv, tv = __ldg(Tg + p - size_t(g) * n2);
                I[p - size_t(g) * n2] = iv;
                a += double(iv) * iv; b += double(iv) * tv; c += double(tv) * tv;
            }
        }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; sh[2][warp] = c; }
    __syncthreads();
    if (tid == 0) {
        double x = 0, y = 0, z = 0;
        for (int i = 0; i < T / 32; ++i) { x += sh[0][i]; y += sh[1][i]; z += sh[2][i]; }
        double* out = partial + (size_t(g) * gridDim.x + blockIdx.x) * 3;
        out[0] = x; out[1] = y; out[2] = z;
    }
}

// ---------------------------------------------------------------------------
// fold the per-tile partials of k_rows_inv_group (all colour groups) -> sums[0..2], PSNR -> sums[3].  One CTA; thread t adds
// partials t, t+256, ... in index order and the 256 thread sums meet in a fixed shuffle/tree
// order, so the result is independent of scheduling.
__global__ void __launch_bounds__(256)
k_loss_final(const double* __restrict__ partial, int n_partial, double n_elems,
             double* __restrict__ sums, int relative) {
    __shared__ double sh[3][8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double x = 0, y = 0, z = 0;
    for (int i = tid; i < n_partial; i += 256) {
        x += partial[i * 3 + 0]; y += partial[i * 3 + 1]; z += partial[i * 3 + 2];
    }
    x = warp_sum(x); y = warp_sum(y); z = warp_sum(z);
    if (lane == 0) { sh[0][warp] = x; sh[1][warp] = y; sh[2][warp] = z; }
    __syncthreads();
    if (tid == 0) {
        x = y = z = 0;
        for (int i = 0; i < 8; ++i) { x += sh[0][i]; y += sh[1][i]; z += sh[2][i]; }
        const double mse = relative ? (z - y * y / x) / n_elems : (x - 2.0 * y + z) / n_elems;
        sums[0] = x; sums[1] = y; sums[2] = z;
        sums[3] = -10.0 * log10(mse);
    }
}

// ---------------------------------------------------------------------------
// incremental path
// ---------------------------------------------------------------------------
// Work decomposition.  A "unit" is UNIT_PX = 1024 consecutive pixels of one
// candidate's N x N image (one pass of a 256-thread CTA at 4 px per thread).
// The n_tasks * units_per_task units of a launch are split into gridDim.x
// contiguous, balanced ranges (CTA b owns [b*total/grid, (b+1)*total/grid)), so
// every resident CTA streams the same number of bytes (+-1 unit) and a CTA
// crosses at most a few task boundaries.
//
// Reduction.  Per-thread fp32 partials of one quad are converted to 2^-40
// fixed point and summed as 64-bit integers (registers -> warp shuffles ->
// shared -> one global atomic per CTA and task).  Integer addition is
// associative, so sum(dI*(2I+dI)) and sum(dI*T) are bit-identical for every
// grid size, batch composition, speculation depth and GPU count.
constexpr int UNIT_PX = 1024;
// Rows of the impulse-response table carry H_PAD wrapped columns (h[y][P + j] = h[y][j]), so the
// four taps of a pixel quad are always contiguous: one address, no per-tap wrap.
constexpr int H_PAD = 4;
__host__ __device__ constexpr int h_stride(int P) { return P + H_PAD; }
constexpr float FIX_SCALE = 1099511627776.0f;          // 2^40
constexpr double FIX_INV = 1.0 / 1099511627776.0;

struct DeltaArgs {
    float2* U; float* I; const float* T; int8_t* state; const float2* h;
    double* sums;                  // [E][4]
    const int32_t* envs;           // [n_tasks] or nullptr (then env_fixed)
    const long long* actions;      // device
    const long long* offset_ptr;   // speculative DBS: actions[*offset_ptr + k]; nullptr otherwise
    long long n_total;             // valid entries of actions
    int env_fixed;
    int n_tasks, N, P, F, G, Fg, relative, rule;
    int HP;                        // row stride of h: P + H_PAD (h_stride)
    int units_per_task;            // N*N / UNIT_PX
    int unit_dy, unit_dx;          // UNIT_PX / N, UNIT_PX % N
    unsigned long long* acc;       // [n_tasks][2] fixed-point accumulators (zero between launches)
    unsigned* tickets;             // [n_tasks]
    Result* results;               // [n_tasks]
    Result* results_host;          // optional mapped pinned mirror written by the finaliser
    // speculative greedy DBS (k_commit does the selection): decision log, PSNR trace, counter
    uint8_t* dbs_accepted; double* dbs_trace; long long* dbs_count; long long* dbs_cursor;
    // small batches (one env step of <= INLINE_MAX envs) carry their tasks in the
    // kernel parameters: no host-to-device copy on the step path
    int n_inline;
    long long inl_actions[32];
    int inl_envs[32];
    // bundled evaluation (k_eval_bundle_t): order a speculation window by frame inside the CTA
    int sort_window;
};
constexpr int INLINE_MAX = 32;

struct Decoded {
    int env, f, g, r, c; float sgn; bool active;
};

// decode (env.py:158-161) and read the sign of the flip from the resident state
__device__ __forceinline__ Decoded decode_action(const DeltaArgs& a, int k, long long act) {
    Decoded d;
    d.env = a.n_inline ? a.inl_envs[k] : (a.envs ? a.envs[k] : a.env_fixed);
    d.f = d.g = d.r = d.c = 0; d.sgn = 0.f;
    d.active = act >= 0;
    if (!d.active) return d;
    const int n2 = a.N * a.N;
    d.f = int(act / n2);
    const int pix = int(act - (long long)d.f * n2);
    d.r = pix / a.N;
    d.c = pix - d.r * a.N;
    d.g = d.f / a.Fg;
    d.sgn = 1.f - 2.f * float(a.state[(size_t(d.env) * a.F + d.f) * n2 + pix]);
    return d;
}

// per-pixel delta:  dI = (2 s Re(conj(U) h) + |h|^2) / Fg
__device__ __forceinline__ float delta_px(float ur, float ui, float hr, float hi, float s2,
                                          float invFg) {
    const float a = fmaf(ur, hr, ui * hi);
    const float m = fmaf(hr, hr, hi * hi);
    return fmaf(s2, a, m * invFg);
}

// L2 cache-policy loads: the streamed operands (U, I, T) are read once per
// candidate and marked evict-first; the impulse response is re-read by every
// candidate and marked evict-last so it stays L2 resident (25 MB for 3 colours).
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ld_keep2(const float2* p, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
                 : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}

// Programmatic dependent launch: the step path is a chain eval -> commit -> eval ... of grids
// that each fill the machine in one wave.  Every kernel waits for its predecessor's memory
// (griddepcontrol.wait) and immediately lets its successor be scheduled, so the successor's
// CTAs land on SMs as this grid's tail drains and the launch latency disappears from the chain.
__device__ __forceinline__ void pdl_wait_then_release() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

struct Quad {                      // operands of 4 consecutive pixels
    float4 ua, ub, iv, tv;
    float2 h0, h1, h2, h3;
};

// position of the thread's quad inside the image, advanced unit by unit
struct Cursor {
    int y, x;
    __device__ __forceinline__ void init(int unit, int tid, int N) {
        const int p = unit * UNIT_PX + tid * 4;
        y = p / N; x = p - y * N;
    }
    __device__ __forceinline__ void next(const DeltaArgs& a) {
        y += a.unit_dy; x += a.unit_dx;
        if (x >= a.N) { x -= a.N; ++y; }
    }
};

template <bool WITH_T>
__device__ __forceinline__ void load_quad(Quad& q, const float2* U, const float* I, const float* T,
                                          const float2* h, const Cursor& cu, int N, int P, int r, int c,
                                          uint64_t pf, uint64_t pl) {
    const size_t p = size_t(cu.y) * N + cu.x;
    const float4* Up = reinterpret_cast<const float4*>(U + p);
    q.ua = ld_stream4(Up, pf);
    q.ub = ld_stream4(Up + 1, pf);
    q.iv = ld_stream4(reinterpret_cast<const float4*>(I + p), pf);
    if (WITH_T) q.tv = ld_stream4(reinterpret_cast<const float4*>(T + p), pf);
    int hy = cu.y - r; if (hy < 0) hy += P;
    int hx = cu.x - c; if (hx < 0) hx += P;
    const float2* hq = h + size_t(hy) * h_stride(P) + hx;
    q.h0 = ld_keep2(hq, pl); q.h1 = ld_keep2(hq + 1, pl);
    q.h2 = ld_keep2(hq + 2, pl); q.h3 = ld_keep2(hq + 3, pl);
}

__device__ __forceinline__ void eval_quad(const Quad& q, float s2, float invFg, long long& aII,
                                          long long& aIT) {
    const float d0 = delta_px(q.ua.x, q.ua.y, q.h0.x, q.h0.y, s2, invFg);
    const float d1 = delta_px(q.ua.z, q.ua.w, q.h1.x, q.h1.y, s2, invFg);
    const float d2 = delta_px(q.ub.x, q.ub.y, q.h2.x, q.h2.y, s2, invFg);
    const float d3 = delta_px(q.ub.z, q.ub.w, q.h3.x, q.h3.y, s2, invFg);
    float ii = d0 * fmaf(2.f, q.iv.x, d0);
    ii = fmaf(d1, fmaf(2.f, q.iv.y, d1), ii);
    ii = fmaf(d2, fmaf(2.f, q.iv.z, d2), ii);
    ii = fmaf(d3, fmaf(2.f, q.iv.w, d3), ii);
    float it = d0 * q.tv.x;
    it = fmaf(d1, q.tv.y, it);
    it = fmaf(d2, q.tv.z, it);
    it = fmaf(d3, q.tv.w, it);
    aII += __float2ll_rn(ii * FIX_SCALE);
    aIT += __float2ll_rn(it * FIX_SCALE);
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA that owns unit u under the balanced partition of `total` units over `grid` CTAs
__device__ __forceinline__ int cta_of_unit(long long u, long long total, int grid) {
    return int(((u + 1) * grid - 1) / total);
}

// One CTA's exact partial sums of task k go to the task's accumulators; the last of the n_ctas
// contributors turns the totals into the PSNR and the accept decision and re-arms the slot.
__device__ __forceinline__ void contribute_and_finalise(const DeltaArgs& a, int k, const Decoded& d,
                                                        long long act, long long x, long long y,
                                                        unsigned n_ctas, size_t n2) {
    atomicAdd(a.acc + 2 * k, (unsigned long long)x);
    atomicAdd(a.acc + 2 * k + 1, (unsigned long long)y);
    __threadfence();
    const unsigned ticket = atomicAdd(a.tickets + k, 1u);
    if (ticket != n_ctas - 1u) return;
    __threadfence();                                         // every contribution has landed
    const long long sII = (long long)__ldcg(a.acc + 2 * k);
    const long long sIT = (long long)__ldcg(a.acc + 2 * k + 1);
    const double dII = double(sII) * FIX_INV, dIT = double(sIT) * FIX_INV;
    const double* S = a.sums + size_t(d.env) * 4;
    const double sii = S[0] + dII, sit = S[1] + dIT, stt = S[2];
    const double n = double(a.G) * double(n2);
    const double mse = a.relative ? (stt - sit * sit / sii) / n
                                  : (sii - 2.0 * sit + stt) / n;
    const double psnr = -10.0 * log10(mse);
    const double prev = S[3];
    int acc = 0;
    if (a.rule == RULE_ENV) acc = !(psnr - prev < 0.0);     // env.py:191
    else if (a.rule == RULE_DBS) acc = (psnr > prev);       // DBS.py:273
    Result r; r.psnr_after = psnr; r.d_sii = dII; r.d_sit = dIT;
    r.action = act; r.accept = acc; r.sgn = int(d.sgn);
    a.results[k] = r;
    if (a.results_host) a.results_host[k] = r;
    a.acc[2 * k] = 0ull; a.acc[2 * k + 1] = 0ull;
    a.tickets[k] = 0;
}

__device__ __forceinline__ void write_idle_result(const DeltaArgs& a, int k) {
    Result r; r.psnr_after = 0.0; r.d_sii = 0.0; r.d_sit = 0.0;
    r.action = -1; r.accept = 0; r.sgn = 0;
    a.results[k] = r;
}

// the action of task k: kernel parameters, or the (cursor-relative) device list; -1 = idle slot
__device__ __forceinline__ long long task_action(const DeltaArgs& a, int k) {
    if (a.n_inline) return a.inl_actions[k];
    long long idx = k;
    if (a.offset_ptr) idx += *a.offset_ptr;
    return (idx < a.n_total) ? a.actions[idx] : -1;
}

// Row-regular images (N divides UNIT_PX, so a unit is UNIT_PX / N whole rows): a thread keeps its
// column for the whole pass, the pixel offset advances by UNIT_PX per unit, and the impulse
// response of a candidate is ONE table offset that advances by unit_dy rows (modulo P).  That
// removes the per-quad decode of (y, x) and the tap addressing of the generic path -- about half
// of its instructions.  Scores L candidates of one frame per pass over units [w0, w1) of the task,
// UF units in flight; pixels, per-quad arithmetic and fixed-point sums are those of the generic
// path, so the results are bit-identical.
template <int L, int UF>
__device__ __forceinline__ void eval_run_rows(const DeltaArgs& a, const float2* U, const float* I,
                                              const float* T, const float2* h, const int* rr,
                                              const int* cc, const float* s2, long long* aII,
                                              long long* aIT, int w0, int w1, int tid, float invFg,
                                              uint64_t pf, uint64_t pl) {
    const int N = a.N, P = a.P, HP = a.HP, dy = a.unit_dy;
    const int y0 = (tid * 4) / N, x = tid * 4 - y0 * N;
    const int y = w0 * dy + y0;
    size_t p = size_t(y) * N + x;
    const int hstep = dy * HP, hwrap = P * HP;
    int ho[L];                                     // offset of the quad's first tap in the table
#pragma unroll
    for (int i = 0; i < L; ++i) {
        int hx = x - cc[i]; if (hx < 0) hx += P;
        int hy = y - rr[i]; if (hy < 0) hy += P;
        ho[i] = hy * HP + hx;
    }
    int w = w0;
#pragma unroll 1
    for (; w + UF <= w1; w += UF) {
        float4 ua[UF], ub[UF], iv[UF], tv[UF];
        float2 hq[UF][L][4];
#pragma unroll
        for (int k = 0; k < UF; ++k) {
            const float4* Up = reinterpret_cast<const float4*>(U + p);
            ua[k] = ld_stream4(Up, pf);
            ub[k] = ld_stream4(Up + 1, pf);
            iv[k] = ld_stream4(reinterpret_cast<const float4*>(I + p), pf);
            tv[k] = ld_stream4(reinterpret_cast<const float4*>(T + p), pf);
            p += UNIT_PX;
#pragma unroll
            for (int i = 0; i < L; ++i) {
                const float2* hp = h + ho[i];
                hq[k][i][0] = ld_keep2(hp, pl); hq[k][i][1] = ld_keep2(hp + 1, pl);
                hq[k][i][2] = ld_keep2(hp + 2, pl); hq[k][i][3] = ld_keep2(hp + 3, pl);
                ho[i] += hstep; if (ho[i] >= hwrap) ho[i] -= hwrap;
            }
        }
#pragma unroll
        for (int k = 0; k < UF; ++k) {
            Quad q; q.ua = ua[k]; q.ub = ub[k]; q.iv = iv[k]; q.tv = tv[k];
#pragma unroll
            for (int i = 0; i < L; ++i) {
                q.h0 = hq[k][i][0]; q.h1 = hq[k][i][1]; q.h2 = hq[k][i][2]; q.h3 = hq[k][i][3];
                eval_quad(q, s2[i], invFg, aII[i], aIT[i]);
            }
        }
    }
#pragma unroll 1
    for (; w < w1; ++w) {
        Quad q;
        const float4* Up = reinterpret_cast<const float4*>(U + p);
        q.ua = ld_stream4(Up, pf);
        q.ub = ld_stream4(Up + 1, pf);
        q.iv = ld_stream4(reinterpret_cast<const float4*>(I + p), pf);
        q.tv = ld_stream4(reinterpret_cast<const float4*>(T + p), pf);
        p += UNIT_PX;
        float2 hq[L][4];
#pragma unroll
        for (int i = 0; i < L; ++i) {
            const float2* hp = h + ho[i];
            hq[i][0] = ld_keep2(hp, pl); hq[i][1] = ld_keep2(hp + 1, pl);
            hq[i][2] = ld_keep2(hp + 2, pl); hq[i][3] = ld_keep2(hp + 3, pl);
            ho[i] += hstep; if (ho[i] >= hwrap) ho[i] -= hwrap;
        }
#pragma unroll
        for (int i = 0; i < L; ++i) {
            q.h0 = hq[i][0]; q.h1 = hq[i][1]; q.h2 = hq[i][2]; q.h3 = hq[i][3];
            eval_quad(q, s2[i], invFg, aII[i], aIT[i]);
        }
    }
}

// units in flight for a run of L candidates: about four quads of taps per thread
__host__ __device__ constexpr int run_uf(int L) { return L == 1 ? 3 : (L == 2 ? 2 : 1); }

template <int L, int B>
__device__ __forceinline__ void eval_run_dispatch(int len, const DeltaArgs& a, const float2* U,
                                                  const float* I, const float* T, const float2* h,
                                                  const int* rr, const int* cc, const float* s2,
                                                  long long* aII, long long* aIT, int w0, int w1,
                                                  int tid, float invFg, uint64_t pf, uint64_t pl) {
    if (len == L || L == B)
        eval_run_rows<L, run_uf(L)>(a, U, I, T, h, rr, cc, s2, aII, aIT, w0, w1, tid, invFg, pf, pl);
    else if constexpr (L < B)
        eval_run_dispatch<L + 1, B>(len, a, U, I, T, h, rr, cc, s2, aII, aIT, w0, w1, tid, invFg, pf, pl);
}

// k_eval: streams U (8 B/px), I and T (4 B/px each) once per candidate and the
// shifted impulse response from L2: 16 N^2 algorithmic HBM bytes per candidate.
// The last CTA to contribute to a task turns the exact sums into the PSNR and
// the accept decision.
template <int UF, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_eval_t(const DeltaArgs a) {
    __shared__ long long sh[2][8];
    pdl_wait_then_release();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, P = a.P, upt = a.units_per_task;
    const long long total = (long long)a.n_tasks * upt;
    const long long beg = (long long)blockIdx.x * total / gridDim.x;
    const long long end = (long long)(blockIdx.x + 1) * total / gridDim.x;
    const float invFg = 1.f / float(a.Fg);
    const uint64_t pf = policy_evict_first(), pl = policy_evict_last();
    const size_t n2 = size_t(N) * N;
    long long u = beg;
    while (u < end) {
        const int k = int(u / upt);
        const long long t_beg = (long long)k * upt;
        const long long seg_end = (t_beg + upt < end) ? t_beg + upt : end;
        const long long act = task_action(a, k);
        const Decoded d = decode_action(a, k, act);
        if (!d.active) {
            if (u == t_beg && tid == 0) write_idle_result(a, k);
            u = seg_end;
            continue;
        }
        const float2* U = a.U + (size_t(d.env) * a.F + d.f) * n2;
        const float* I = a.I + (size_t(d.env) * a.G + d.g) * n2;
        const float* T = a.T + (size_t(d.env) * a.G + d.g) * n2;
        const float2* h = a.h + size_t(d.g) * P * a.HP;
        const float s2 = 2.f * d.sgn * invFg;
        long long aII = 0, aIT = 0;
        if (a.unit_dx == 0) {
            eval_run_rows<1, UF>(a, U, I, T, h, &d.r, &d.c, &s2, &aII, &aIT, int(u - t_beg),
                                 int(seg_end - t_beg), tid, invFg, pf, pl);
        } else {
            Cursor cu; cu.init(int(u - t_beg), tid, N);
            long long v = u;
            for (; v + UF <= seg_end; v += UF) {         // UF units in flight per thread
                Quad q[UF];
#pragma unroll
                for (int i = 0; i < UF; ++i) {
                    load_quad<true>(q[i], U, I, T, h, cu, N, P, d.r, d.c, pf, pl); cu.next(a);
                }
#pragma unroll
                for (int i = 0; i < UF; ++i) eval_quad(q[i], s2, invFg, aII, aIT);
            }
            for (; v < seg_end; ++v) {
                Quad q0;
                load_quad<true>(q0, U, I, T, h, cu, N, P, d.r, d.c, pf, pl); cu.next(a);
                eval_quad(q0, s2, invFg, aII, aIT);
            }
        }
        aII = warp_sum_ll(aII); aIT = warp_sum_ll(aIT);
        if (lane == 0) { sh[0][warp] = aII; sh[1][warp] = aIT; }
        __syncthreads();
        if (tid == 0) {
            long long x = 0, y = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { x += sh[0][i]; y += sh[1][i]; }
            const int first = cta_of_unit(t_beg, total, gridDim.x);
            const int last = cta_of_unit(t_beg + upt - 1, total, gridDim.x);
            contribute_and_finalise(a, k, d, act, x, y, unsigned(last - first) + 1u, n2);
        }
        __syncthreads();
        u = seg_end;
    }
}

// k_eval_bundle: candidate lists that revisit an environment (speculation windows of the greedy
// DBS, DBS.py:247-294; the candidate tables of env_group.py:96-119 and the sweeps) are scored B
// slots at a time.  Inside a bundle every run of candidates of one frame is ONE pass over the
// image: the thread that owns a quad loads U, I and T once per run and, per candidate, only the
// shifted impulse response (L2 resident).  HBM traffic per candidate falls from 16 N^2 B towards
// 16 N^2 / B; the per-quad arithmetic and the 2^-40 fixed-point sums are those of k_eval_t, so
// both kernels return bit-identical results.  A speculation window (n_tasks <= SORT_WINDOW_MAX,
// sort_window set) is ordered by frame inside every CTA first; results stay indexed by the
// caller's task number.
constexpr int SORT_WINDOW_MAX = 128;

struct BundleTask { int task; long long act; Decoded d; };

__device__ __forceinline__ BundleTask bundle_task(const DeltaArgs& a, int slot, bool sorted,
                                                  const int* s_order) {
    BundleTask t;
    t.task = sorted ? s_order[slot] : slot;
    t.act = task_action(a, t.task);
    t.d = decode_action(a, t.task, t.act);
    return t;
}

template <int B, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_eval_bundle_t(const DeltaArgs a) {
    __shared__ long long sh[2 * B][8];
    __shared__ int s_key[SORT_WINDOW_MAX];
    __shared__ int s_order[SORT_WINDOW_MAX];
    pdl_wait_then_release();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, P = a.P, upt = a.units_per_task, n = a.n_tasks;
    const size_t n2 = size_t(N) * N;
    const bool sorted = a.sort_window && n <= SORT_WINDOW_MAX;
    if (sorted) {
        if (tid < n) {
            const long long act = task_action(a, tid);
            s_key[tid] = act < 0 ? 0x7fffffff : int(act / (long long)n2);     // idle slots last
        }
        __syncthreads();
        if (tid < n) {
            const int key = s_key[tid];
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const int kj = s_key[j];
                rank += (kj < key || (kj == key && j < tid)) ? 1 : 0;
            }
            s_order[rank] = tid;
        }
        __syncthreads();
    }
    const int n_bundles = (n + B - 1) / B;
    const long long total = (long long)n_bundles * upt;
    const long long beg = (long long)blockIdx.x * total / gridDim.x;
    const long long end = (long long)(blockIdx.x + 1) * total / gridDim.x;
    const float invFg = 1.f / float(a.Fg);
    const uint64_t pf = policy_evict_first(), pl = policy_evict_last();
    long long u = beg;
    while (u < end) {
        const int j = int(u / upt);
        const long long t_beg = (long long)j * upt;
        const long long seg_end = (t_beg + upt < end) ? t_beg + upt : end;
        const unsigned n_ctas = unsigned(cta_of_unit(t_beg + upt - 1, total, gridDim.x) -
                                         cta_of_unit(t_beg, total, gridDim.x)) + 1u;
        const int slots = (n - j * B < B) ? n - j * B : B;
        int rb = 0;
        while (rb < slots) {
            // the run: slots rb .. rb+len-1 of the bundle, all of one environment and frame
            const BundleTask head = bundle_task(a, j * B + rb, sorted, s_order);
            if (!head.d.active) {                           // idle slot
                if (u == t_beg && tid == 0) write_idle_result(a, head.task);
                rb += 1;
                continue;
            }
            int rr[B], cc[B]; float s2[B]; long long aII[B], aIT[B];
            int len = 1;
            rr[0] = head.d.r; cc[0] = head.d.c; s2[0] = 2.f * head.d.sgn * invFg;
            aII[0] = 0; aIT[0] = 0;
#pragma unroll
            for (int i = 1; i < B; ++i) {
                rr[i] = 0; cc[i] = 0; s2[i] = 0.f; aII[i] = 0; aIT[i] = 0;
                if (rb + i < slots && len == i) {
                    const BundleTask t = bundle_task(a, j * B + rb + i, sorted, s_order);
                    if (t.d.active && t.d.env == head.d.env && t.d.f == head.d.f) {
                        rr[i] = t.d.r; cc[i] = t.d.c; s2[i] = 2.f * t.d.sgn * invFg;
                        len = i + 1;
                    }
                }
            }
            const float2* U = a.U + (size_t(head.d.env) * a.F + head.d.f) * n2;
            const float* I = a.I + (size_t(head.d.env) * a.G + head.d.g) * n2;
            const float* T = a.T + (size_t(head.d.env) * a.G + head.d.g) * n2;
            const float2* h = a.h + size_t(head.d.g) * P * a.HP;
            if (a.unit_dx == 0) {
                eval_run_dispatch<1, B>(len, a, U, I, T, h, rr, cc, s2, aII, aIT, int(u - t_beg),
                                        int(seg_end - t_beg), tid, invFg, pf, pl);
            } else {
                Cursor cu; cu.init(int(u - t_beg), tid, N);
#pragma unroll 1
                for (long long v = u; v < seg_end; ++v) {
                    const size_t p = size_t(cu.y) * N + cu.x;
                    Quad q;
                    const float4* Up = reinterpret_cast<const float4*>(U + p);
                    q.ua = ld_stream4(Up, pf);
                    q.ub = ld_stream4(Up + 1, pf);
                    q.iv = ld_stream4(reinterpret_cast<const float4*>(I + p), pf);
                    q.tv = ld_stream4(reinterpret_cast<const float4*>(T + p), pf);
                    float2 hq[B][4];
#pragma unroll
                    for (int i = 0; i < B; ++i) {
                        if (i < len) {
                            int hy = cu.y - rr[i]; if (hy < 0) hy += P;
                            int hx = cu.x - cc[i]; if (hx < 0) hx += P;
                            const float2* hp = h + size_t(hy) * a.HP + hx;
                            hq[i][0] = ld_keep2(hp, pl); hq[i][1] = ld_keep2(hp + 1, pl);
                            hq[i][2] = ld_keep2(hp + 2, pl); hq[i][3] = ld_keep2(hp + 3, pl);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < B; ++i) {
                        if (i < len) {
                            q.h0 = hq[i][0]; q.h1 = hq[i][1]; q.h2 = hq[i][2]; q.h3 = hq[i][3];
                            eval_quad(q, s2[i], invFg, aII[i], aIT[i]);
                        }
                    }
                    cu.next(a);
                }
            }
#pragma unroll
            for (int i = 0; i < B; ++i) {
                if (i < len) {
                    const long long x = warp_sum_ll(aII[i]), y = warp_sum_ll(aIT[i]);
                    if (lane == 0) { sh[2 * i][warp] = x; sh[2 * i + 1][warp] = y; }
                }
            }
            __syncthreads();
            if (tid < len) {                                // one thread per candidate of the run
                const BundleTask t = bundle_task(a, j * B + rb + tid, sorted, s_order);
                long long x = 0, y = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { x += sh[2 * tid][w]; y += sh[2 * tid + 1][w]; }
                contribute_and_finalise(a, t.task, t.d, t.act, x, y, n_ctas, n2);
            }
            __syncthreads();
            rb += len;
        }
        u = seg_end;
    }
}

// k_commit: applies every accepted task: U_f += s*shift(h), I_g += dI, flips the
// state byte and advances the running sums.  24 N^2 algorithmic HBM bytes per
// accepted flip.  Tasks of one launch must target distinct environments
// (n_tasks <= COMMIT_MAX_TASKS).  Every CTA first compacts the accepted tasks,
// then the balanced unit partition runs over the accepted ones only, so a
// launch with one accepted flip out of K still uses the whole chip.
constexpr int COMMIT_MAX_TASKS = 256;

__device__ __forceinline__ void commit_quad(const Quad& q, float2* U, float* I, size_t p, float s2,
                                            float sg, float invFg) {
    float4 iv = q.iv, ua = q.ua, ub = q.ub;
    iv.x += delta_px(ua.x, ua.y, q.h0.x, q.h0.y, s2, invFg);
    iv.y += delta_px(ua.z, ua.w, q.h1.x, q.h1.y, s2, invFg);
    iv.z += delta_px(ub.x, ub.y, q.h2.x, q.h2.y, s2, invFg);
    iv.w += delta_px(ub.z, ub.w, q.h3.x, q.h3.y, s2, invFg);
    ua.x = fmaf(sg, q.h0.x, ua.x); ua.y = fmaf(sg, q.h0.y, ua.y);
    ua.z = fmaf(sg, q.h1.x, ua.z); ua.w = fmaf(sg, q.h1.y, ua.w);
    ub.x = fmaf(sg, q.h2.x, ub.x); ub.y = fmaf(sg, q.h2.y, ub.y);
    ub.z = fmaf(sg, q.h3.x, ub.z); ub.w = fmaf(sg, q.h3.y, ub.w);
    float4* Up = reinterpret_cast<float4*>(U + p);
    Up[0] = ua; Up[1] = ub;
    *reinterpret_cast<float4*>(I + p) = iv;
}

template <int UF, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_commit_t(const DeltaArgs a) {
    __shared__ int s_list[COMMIT_MAX_TASKS];
    __shared__ int s_cnt;
    const int tid = threadIdx.x;
    pdl_wait_then_release();
    if (tid == 0) {
        int n = 0;
        if (a.dbs_cursor) {
            // speculative greedy DBS (DBS.py:247-294 order): all K candidates were scored against
            // the same state; keep the first accepted one, the candidates after it are scored
            // again by the next batch.  Block 0 logs the decisions and advances the cursor
            // (idle slots carry accept = 0, so no other block needs the cursor).
            int first = -1;
            for (int k = 0; k < a.n_tasks; ++k)
                if (a.results[k].accept) { first = k; break; }
            if (first >= 0) s_list[n++] = first;
            if (blockIdx.x == 0) {
                const long long off = *a.dbs_cursor;
                long long cnt = a.n_total - off;
                if (cnt > a.n_tasks) cnt = a.n_tasks;
                if (cnt > 0) {
                    const int used = first >= 0 ? first + 1 : int(cnt);
                    for (int k = 0; k < used; ++k) {
                        a.dbs_accepted[off + k] = (k == first) ? 1 : 0;
                        if (a.dbs_trace) a.dbs_trace[off + k] = a.results[k].psnr_after;
                    }
                    if (first >= 0) *a.dbs_count += 1;
                    *a.dbs_cursor = off + used;
                }
            }
        } else {
            for (int k = 0; k < a.n_tasks; ++k)
                if (a.results[k].accept) s_list[n++] = k;
        }
        s_cnt = n;
    }
    __syncthreads();
    const int n_acc = s_cnt;
    if (n_acc == 0) return;
    const int N = a.N, P = a.P, upt = a.units_per_task;
    const long long total = (long long)n_acc * upt;
    const long long beg = (long long)blockIdx.x * total / gridDim.x;
    const long long end = (long long)(blockIdx.x + 1) * total / gridDim.x;
    const float invFg = 1.f / float(a.Fg);
    const uint64_t pf = policy_evict_first(), pl = policy_evict_last();
    const size_t n2 = size_t(N) * N;
    long long u = beg;
    while (u < end) {
        const int slot = int(u / upt);
        const int k = s_list[slot];
        const long long t_beg = (long long)slot * upt;
        const long long seg_end = (t_beg + upt < end) ? t_beg + upt : end;
        const Result res = a.results[k];
        Decoded d = decode_action(a, k, res.action);
        d.sgn = float(res.sgn);            // the state byte may already be flipped
        float2* U = a.U + (size_t(d.env) * a.F + d.f) * n2;
        float* I = a.I + (size_t(d.env) * a.G + d.g) * n2;
        const float2* h = a.h + size_t(d.g) * P * a.HP;
        const float s2 = 2.f * d.sgn * invFg, sg = d.sgn;
        Cursor cu; cu.init(int(u - t_beg), tid, N);
        long long v = u;
        for (; v + UF <= seg_end; v += UF) {
            Quad q[UF];
            size_t p[UF];
#pragma unroll
            for (int i = 0; i < UF; ++i) {
                load_quad<false>(q[i], U, I, nullptr, h, cu, N, P, d.r, d.c, pf, pl);
                p[i] = size_t(cu.y) * N + cu.x;
                cu.next(a);
            }
#pragma unroll
            for (int i = 0; i < UF; ++i) commit_quad(q[i], U, I, p[i], s2, sg, invFg);
        }
        for (; v < seg_end; ++v) {
            Quad q;
            load_quad<false>(q, U, I, nullptr, h, cu, N, P, d.r, d.c, pf, pl);
            commit_quad(q, U, I, size_t(cu.y) * N + cu.x, s2, sg, invFg);
            cu.next(a);
        }
        if (u == t_beg && tid == 0) {
            int8_t* st = a.state + (size_t(d.env) * a.F + d.f) * n2 + size_t(d.r) * N + d.c;
            *st = int8_t(1 - *st);
            double* S = a.sums + size_t(d.env) * 4;
            S[0] += res.d_sii;
            S[1] += res.d_sit;
            S[3] = res.psnr_after;
        }
        u = seg_end;
    }
}

// k_recon_candidate: out_g += dI of one (uncommitted) candidate flip; used to
// materialise obs["recon_image"] of a rejected step (env.py:176-181, appendix B-2).
__global__ void __launch_bounds__(256)
k_recon_candidate(const float2* __restrict__ U, const float2* __restrict__ h,
                  float* __restrict__ out_g, int N, int P, int r, int c, float sgn, int Fg) {
    const float invFg = 1.f / float(Fg), s2 = 2.f * sgn * invFg;
    const size_t n2 = size_t(N) * N;
    for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < n2;
         p += size_t(gridDim.x) * blockDim.x) {
        const int y = int(p / N), x = int(p - size_t(y) * N);
        int hy = y - r; if (hy < 0) hy += P;
        int hx = x - c; if (hx < 0) hx += P;
        const float2 u = U[p], hv = __ldg(h + size_t(hy) * h_stride(P) + hx);
        out_g[p] += delta_px(u.x, u.y, hv.x, hv.y, s2, invFg);
    }
}

// ---------------------------------------------------------------------------
// exhaustive sweep by correlation (score EVERY pixel of every frame against the
// fixed state: dbs-1024-1024-24-6464.py:330-395, range.py:294-335).
//
// For a flip of pixel (r,c) of frame f (sign s), with h_s = h shifted to (r,c),
//   a = Re(conj(U) h_s), m = |h_s|^2, dI = (2 s a + m)/Fg   and
//   d sum(I T) = (2s/Fg) Re C1 + (1/Fg) BT
//   d sum(I^2) = 2[(2s/Fg) Re C2 + (1/Fg) BI] + (1/Fg^2)[2 C4 + 2 Re C5 + 4 s Re C3 + sum|h|^4]
// where every term is a circular cross-correlation with an EVEN kernel,
//   C1 = (conj(U) T) * h      C2 = (conj(U) I) * h      C3 = conj(U) * (h|h|^2)
//   C4 = |U|^2 * |h|^2        C5 = conj(U)^2 * h^2      BT + i BI = (T + i I) * |h|^2
// i.e. ifft2(fft2(W) K) -- the propagation passes with a different spectrum K.
// All N^2 candidates of a frame cost a handful of FFTs instead of N^2 delta passes.
// ---------------------------------------------------------------------------
enum { PREP_UA = 0, PREP_U = 1, PREP_U2 = 2, PREP_ABS2_PAIR = 3, PREP_TI = 4, PREP_ONES = 5 };

// grid (blocks, planes).  U: group base [Fg][n2]; A, B: real planes of the group.
__global__ void __launch_bounds__(256)
k_sweep_prep(const float2* __restrict__ U, const float* __restrict__ A, const float* __restrict__ B,
             float2* __restrict__ out, size_t n2, int mode) {
    const int pl = blockIdx.y;
    float2* o = out + size_t(pl) * n2;
    for (size_t p = size_t(blockIdx.x) * 256 + threadIdx.x; p < n2; p += size_t(gridDim.x) * 256) {
        float2 v;
        if (mode == PREP_UA) {
            const float2 u = U[size_t(pl) * n2 + p]; const float a = A[p];
            v = make_float2(u.x * a, -u.y * a);
        } else if (mode == PREP_U) {
            const float2 u = U[size_t(pl) * n2 + p];
            v = make_float2(u.x, -u.y);
        } else if (mode == PREP_U2) {
            const float2 u = U[size_t(pl) * n2 + p];
            v = make_float2(u.x * u.x - u.y * u.y, -2.f * u.x * u.y);
        } else if (mode == PREP_ABS2_PAIR) {
            const float2 u0 = U[size_t(2 * pl) * n2 + p], u1 = U[size_t(2 * pl + 1) * n2 + p];
            v = make_float2(u0.x * u0.x + u0.y * u0.y, u1.x * u1.x + u1.y * u1.y);
        } else if (mode == PREP_TI) {
            v = make_float2(A[p], B[p]);
        } else {
            v = make_float2(1.f, 0.f);               // window indicator (pad = 2: sum |h_s|^4 over the window)
        }
        o[p] = v;
    }
}

enum { ACC_RE = 0, ACC_PAIR = 1, ACC_GROUP = 2, ACC_GROUP_RE = 3 };

// accumulate coef * part(S) into the per-candidate planes.
//   ACC_RE    plane i -> frame i:   dst[i] += coef * (use_sign ? sgn : 1) * Re S[i]
//   ACC_PAIR  plane i -> frames 2i (Re) and 2i+1 (Im) of dII
//   ACC_GROUP plane 0 -> every frame: dIT += cT * Re S, dII += cI * Im S + cst
__global__ void __launch_bounds__(256)
k_sweep_acc(const float2* __restrict__ S, float* __restrict__ dIT, float* __restrict__ dII,
            const int8_t* __restrict__ state /*group base*/, size_t n2, int Fg, int mode,
            int to_ii, int use_sign, float coef, float coef2, float cst) {
    const int pl = blockIdx.y;
    for (size_t p = size_t(blockIdx.x) * 256 + threadIdx.x; p < n2; p += size_t(gridDim.x) * 256) {
        if (mode == ACC_RE) {
            float v = coef * S[size_t(pl) * n2 + p].x;
            if (use_sign) v *= 1.f - 2.f * float(state[size_t(pl) * n2 + p]);
            float* d = (to_ii ? dII : dIT) + size_t(pl) * n2 + p;
            *d += v;
        } else if (mode == ACC_PAIR) {
            const float2 v = S[size_t(pl) * n2 + p];
            dII[size_t(2 * pl) * n2 + p] += coef * v.x;
            dII[size_t(2 * pl + 1) * n2 + p] += coef * v.y;
        } else if (mode == ACC_GROUP) {
            const float2 v = S[p];
            dIT[size_t(pl) * n2 + p] += coef * v.x;
            dII[size_t(pl) * n2 + p] += coef2 * v.y + cst;
        } else {                                     // ACC_GROUP_RE: plane 0 -> dII of every frame
            dII[size_t(pl) * n2 + p] += coef * S[p].x;
        }
    }
}

// PSNR of every candidate of the group from the accumulated deltas (float64 closed form)
__global__ void __launch_bounds__(256)
k_sweep_final(const float* __restrict__ dIT, const float* __restrict__ dII,
              const double* __restrict__ sums, double* __restrict__ out, size_t count,
              double n_elems, int relative) {
    const double sii0 = sums[0], sit0 = sums[1], stt = sums[2];
    for (size_t p = size_t(blockIdx.x) * 256 + threadIdx.x; p < count; p += size_t(gridDim.x) * 256) {
        const double sii = sii0 + double(dII[p]), sit = sit0 + double(dIT[p]);
        const double mse = relative ? (stt - sit * sit / sii) / n_elems : (sii - 2.0 * sit + stt) / n_elems;
        out[p] = -10.0 * log10(mse);
    }
}

// ---------------------------------------------------------------------------
// decile statistics of a sweep on the device (dbs-1024-1024-24-6464.py:377-395):
// per bin of the pre-model output: attempted, improved (psnr_after > previous) and the
// summed improvement.  Bins are half-open, the last one closed, compared in double like
// the reference's numpy code.  Gains are accumulated in 2^-40 fixed point so the sums do
// not depend on the order of the atomics.
// ---------------------------------------------------------------------------
constexpr int N_BINS = 10;
struct BinEdges { double e[N_BINS + 1]; };

__global__ void __launch_bounds__(256)
k_sweep_stats(const double* __restrict__ psnr, const float* __restrict__ pre, size_t count,
              double previous, BinEdges edges, unsigned long long* __restrict__ out /*[3][N_BINS]*/) {
    __shared__ unsigned long long sh[3][N_BINS];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 3 * N_BINS) (&sh[0][0])[tid] = 0ull;
    __syncthreads();
    // per-thread counters in registers (static indices), folded once at the end
    unsigned att[N_BINS], imp[N_BINS];
    long long gain[N_BINS];
#pragma unroll
    for (int i = 0; i < N_BINS; ++i) { att[i] = 0; imp[i] = 0; gain[i] = 0; }
    for (size_t p = size_t(blockIdx.x) * 256 + tid; p < count; p += size_t(gridDim.x) * 256) {
        const double v = double(pre[p]);
        const double d = psnr[p] - previous;
        const bool better = d > 0.0;
        const long long fx = better ? __double2ll_rn(d * 1099511627776.0) : 0ll;
        bool taken = false;
#pragma unroll
        for (int i = 0; i < N_BINS; ++i) {
            const bool in = !taken && ((i == N_BINS - 1) ? (v >= edges.e[i] && v <= edges.e[i + 1])
                                                         : (v >= edges.e[i] && v < edges.e[i + 1]));
            taken = taken || in;
            att[i] += in ? 1u : 0u;
            imp[i] += (in && better) ? 1u : 0u;
            gain[i] += in ? fx : 0ll;
        }
    }
#pragma unroll
    for (int i = 0; i < N_BINS; ++i) {
        unsigned a = att[i], m = imp[i];
        long long g = gain[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            m += __shfl_xor_sync(0xffffffffu, m, o);
            g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        if (lane == 0) {
            atomicAdd(&sh[0][i], (unsigned long long)a);
            atomicAdd(&sh[1][i], (unsigned long long)m);
            atomicAdd(&sh[2][i], (unsigned long long)g);
        }
    }
    __syncthreads();
    if (tid < 3 * N_BINS) atomicAdd(out + tid, (&sh[0][0])[tid]);
}

}  // namespace bh
