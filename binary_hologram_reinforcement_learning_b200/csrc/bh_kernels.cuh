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

}  // namespace bh

#include "bh_delta.cuh"      // incremental path: k_eval / k_eval_bundle / k_commit / k_recon_*

namespace bh {

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
