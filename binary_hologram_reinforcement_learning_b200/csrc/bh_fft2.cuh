// Register-resident propagation passes for the two-pass FFT sides (P = 896 = 32 x 28, P = 1024 = 32 x 32;
// pad = 1, i.e. the BASELINE shapes of env_1024_24.py and env_1024_24_128.py).  Restates
// tt.simulate + .abs()**2 + mean(dim=1) + the sums of tt.relativeLoss (env.py:123-132, env_1024_24.py:149-166).
//
// A length-P transform with radices (RA, RB), P = RA * RB, is two register butterflies per thread with ONE
// exchange through shared memory in between:
//     pass 1   thread j < RB holds x[j + r RB], r < RA        -> DFT_RA -> writes positions j RA + r
//     pass 2   thread j < RA reads  positions j + r RA, r < RB -> twiddle W_P^(r j) -> DFT_RB
//              and ends holding X[j + r RA], r < RB
// The entry pattern of pass 1 is a stride-RB gather and the exit pattern of pass 2 a stride-RA scatter -- for a
// ROW both are warp-wide contiguous runs in global memory, so a row transform goes global -> registers ->
// butterfly -> (shared) -> butterfly -> registers -> global: one shared-memory round trip instead of five, the
// sequence belongs to ONE warp and the exchange needs no CTA barrier.  The exit pattern of the plan (RA, RB)
// is the entry pattern of the reversed plan (RB, RA), so the COLUMN pass chains forward FFT -> multiply by H ->
// inverse FFT in registers: three exchanges for one forward and two inverse transforms (Hermitian pair).
//
//   k2_rows_fwd_real   binary state rows, two per complex sequence, byte gathers straight into the butterfly
//                      registers, Hermitian separation by warp shuffles; stores kx <= P/2
//   k2_cols            persistent CTAs; the [P x 4] column tile and the H tile arrive by TMA
//                      (cp.async.bulk.tensor + mbarrier) and are prefetched one tile ahead; HERM: one forward
//                      FFT feeds column kx (with H) and column P - kx (with conj H)
//   k2_rows_inv_group  one CTA per window row of a colour group, one warp per frame: inverse row FFT, U stored
//                      in place, |U|^2 of the frames reduced in shared memory -> I row + float64 loss partials
//   k2_rows_fwd_cplx / k2_rows_inv   complex-input head and plain tail (tt.simulate operator, sweep correlations)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "bh_fft.cuh"
#include "bh_async.cuh"

namespace bh {

constexpr bool fft2_side(int P) { return P == 896 || P == 1024; }

// Programmatic dependent launch between the passes (each pass needs ALL of its predecessor's output, so only the
// launch latency and the CTA ramp-up overlap): a kernel lets its successor be scheduled at once and waits for
// its predecessor's memory before its first global access.
__device__ __forceinline__ void fft_pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fft_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// radices of the (optionally reversed) two-pass plan and the offset of its twiddle block
template <int P, bool REV> struct Plan2 {
    static constexpr int RA = REV ? Plan<P>::r[1] : Plan<P>::r[0];
    static constexpr int RB = REV ? Plan<P>::r[0] : Plan<P>::r[1];
    // tw[off + (r - 1) RA + k] = exp(-2 pi i r k / P), r = 1..RB-1, k < RA.  Equal radices share one block.
    static constexpr int tw_off = (REV && Plan<P>::r[0] != Plan<P>::r[1]) ? TwLayout<P>::total : 0;
};

// shared-memory placement of sequence position p of an exchange whose FIRST pass has radix RA.
//   wr(j, r): position p = j RA + r   (pass 1, thread j, output r)
//   rd(j, r): position p = j + r RA   (pass 2, thread j, input r)
// Row layout: one padding slot after every RA elements (lanes j -> stride RA + 1: 2-way = ideal for 8-byte
// words).  Column layout: W sequences interleaved, dense (TMA tiles alias these buffers), the two low
// position bits XORed with the block index p / RA -- a bijection inside each block since 4 | RA (W = 4 or 8
// sequences interleaved: the 16 threads of a half-warp are 4 or 2 consecutive j, which land in distinct bank groups).
template <int RA> struct RowLay {
    static __device__ __forceinline__ int wr(int j, int r) { return j * (RA + 1) + r; }
    static __device__ __forceinline__ int rd(int j, int r) { return j + r * (RA + 1); }
};
template <int P, int RA> struct RowLaySize { static constexpr int value = P + P / RA; };
template <int W> struct ColLayW {
    template <int RA> struct L {
        static __device__ __forceinline__ int wr(int j, int r) { return ((j * RA + r) ^ (j & 3)) * W; }
        static __device__ __forceinline__ int rd(int j, int r) { return ((j + r * RA) ^ (r & 3)) * W; }
    };
};

template <int P, bool INV, bool REV, template <int> class Lay>
__device__ __forceinline__ void fft2_pass1(float2 (&v)[32], float2* s, int j) {
    constexpr int RA = Plan2<P, REV>::RA, RB = Plan2<P, REV>::RB;
    if (RB == 32 || j < RB) {
        dft<RA, INV>(v);
#pragma unroll
        for (int r = 0; r < RA; ++r) s[Lay<RA>::wr(j, r)] = v[r];
    }
}

template <int P, bool INV, bool REV, template <int> class Lay>
__device__ __forceinline__ void fft2_pass2(float2 (&v)[32], const float2* s, int j, const float2* __restrict__ tw) {
    constexpr int RA = Plan2<P, REV>::RA, RB = Plan2<P, REV>::RB;
    if (RA == 32 || j < RA) {
#pragma unroll
        for (int r = 0; r < RB; ++r) v[r] = s[Lay<RA>::rd(j, r)];
        const float2* twp = tw + Plan2<P, REV>::tw_off + j;
#pragma unroll
        for (int r = 1; r < RB; ++r) {
            float2 w = __ldg(twp + (r - 1) * RA);
            if (INV) w.y = -w.y;
            v[r] = cmul(v[r], w);
        }
        dft<RB, INV>(v);
    }
}

// pass 2 in two halves (an asynchronous copy can be issued between the reads of the exchange buffer and the
// butterfly), twiddles from a shared-memory copy of the plan's block
template <int P, bool REV, template <int> class Lay>
__device__ __forceinline__ void fft2_pass2_read(float2 (&v)[32], const float2* s, int j) {
    constexpr int RA = Plan2<P, REV>::RA, RB = Plan2<P, REV>::RB;
    if (RA == 32 || j < RA) {
#pragma unroll
        for (int r = 0; r < RB; ++r) v[r] = s[Lay<RA>::rd(j, r)];
    }
}
// stw: shared-memory copy of this plan's twiddle block, stw[(r - 1) RA + k] = exp(-2 pi i r k / P)
template <int P, bool INV, bool REV, bool TW_GLOBAL = false>
__device__ __forceinline__ void fft2_pass2_math(float2 (&v)[32], int j, const float2* stw) {
    constexpr int RA = Plan2<P, REV>::RA, RB = Plan2<P, REV>::RB;
    if (RA == 32 || j < RA) {
#pragma unroll
        for (int r = 1; r < RB; ++r) {
            float2 w = TW_GLOBAL ? __ldg(stw + (r - 1) * RA + j) : stw[(r - 1) * RA + j];
            if (INV) w.y = -w.y;
            v[r] = cmul(v[r], w);
        }
        dft<RB, INV>(v);
    }
}


// ---------------------------------------------------------------------------
// pass A, real input: int8 state [frames][P][P] -> row spectra kx <= P/2 of buf [frames][P][P].
// One warp per pair of rows (2 yp, 2 yp + 1): z = a + i b, one transform, A[k] = (Z[k] + conj Z[-k]) / 2,
// B[k] = (Z[k] - conj Z[-k]) / (2 i).  4 warps per CTA, grid-stride over the frames * P/2 row pairs.
// ---------------------------------------------------------------------------
constexpr int ROWS_WARPS = 4;

__device__ __forceinline__ float byte_to_float(int v) {          // exact for -128..127, no I2F
    return __int_as_float(0x4B400000 + v) - 12582912.0f;
}

template <int P>
__global__ void __launch_bounds__(32 * ROWS_WARPS, 4)
k2_rows_fwd_real(const int8_t* __restrict__ state, float2* __restrict__ buf, const float2* __restrict__ tw,
                 int n_pairs) {
    constexpr int RA = Plan2<P, false>::RA, RB = Plan2<P, false>::RB;
    constexpr int SZ = RowLaySize<P, RA>::value;
    extern __shared__ float2 s2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* sw = s2 + warp * SZ;
    fft_pdl_release();
    fft_pdl_wait();
    for (int task = blockIdx.x * ROWS_WARPS + warp; task < n_pairs; task += gridDim.x * ROWS_WARPS) {
        const int f = task / (P / 2), yp = task - f * (P / 2);
        const int8_t* pa = state + (size_t(f) * P + 2 * yp) * P + lane;
        const int8_t* pb = pa + P;
        float2 v[32];
        if (RB == 32 || lane < RB) {
            int ia[RA], ib[RA];
#pragma unroll
            for (int r = 0; r < RA; ++r) { ia[r] = __ldg(pa + r * RB); ib[r] = __ldg(pb + r * RB); }
#pragma unroll
            for (int r = 0; r < RA; ++r) v[r] = make_float2(byte_to_float(ia[r]), byte_to_float(ib[r]));
        }
        fft2_pass1<P, false, false, RowLay>(v, sw, lane);
        __syncwarp();
        fft2_pass2<P, false, false, RowLay>(v, sw, lane, tw);
        __syncwarp();                                  // the buffer is free for the next pair
        // lane j holds Z[j + r RA], r < RB.  Z[P - k] for k = j + r RA lives in lane (RA - j) % RA at
        // register RB - 1 - r (j >= 1) or in lane 0 at register (RB - r) % RB (j = 0).
        float2* oa = buf + (size_t(f) * P + 2 * yp) * P + lane;
        float2* ob = oa + P;
        const int src = (RA - lane) & (RA - 1);
#pragma unroll
        for (int r = 0; r <= RB / 2; ++r) {
            float2 zm;
            zm.x = __shfl_sync(0xffffffffu, v[RB - 1 - r].x, src);
            zm.y = __shfl_sync(0xffffffffu, v[RB - 1 - r].y, src);
            if (lane == 0) zm = v[(RB - r) % RB];
            const float2 zk = v[r];
            if (r < RB / 2 || lane == 0) {             // k <= P/2; the rest follows from Hermitian symmetry
                oa[r * RA] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                ob[r * RA] = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
            }
        }
    }
}

// pass A, complex input (stand-alone operator, sweep correlations): one warp per row
template <int P>
__global__ void __launch_bounds__(32 * ROWS_WARPS, 4)
k2_rows_fwd_cplx(const float2* __restrict__ in, float2* __restrict__ buf, const float2* __restrict__ tw, int n_rows) {
    constexpr int RA = Plan2<P, false>::RA, RB = Plan2<P, false>::RB;
    constexpr int SZ = RowLaySize<P, RA>::value;
    extern __shared__ float2 s2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* sw = s2 + warp * SZ;
    for (int row = blockIdx.x * ROWS_WARPS + warp; row < n_rows; row += gridDim.x * ROWS_WARPS) {
        const float2* src = in + size_t(row) * P + lane;
        float2 v[32];
        if (RB == 32 || lane < RB) {
#pragma unroll
            for (int r = 0; r < RA; ++r) v[r] = src[r * RB];
        }
        fft2_pass1<P, false, false, RowLay>(v, sw, lane);
        __syncwarp();
        fft2_pass2<P, false, false, RowLay>(v, sw, lane, tw);
        __syncwarp();
        float2* dst = buf + size_t(row) * P + lane;
#pragma unroll
        for (int r = 0; r < RB; ++r) dst[r * RA] = v[r];
    }
}

// pass C alone: inverse row transform, in place or into U (tail of tt.simulate and of the sweep correlations)
template <int P>
__global__ void __launch_bounds__(32 * ROWS_WARPS, 4)
k2_rows_inv(const float2* buf, float2* U, const float2* __restrict__ tw, int n_rows) {
    constexpr int RA = Plan2<P, false>::RA, RB = Plan2<P, false>::RB;
    constexpr int SZ = RowLaySize<P, RA>::value;
    extern __shared__ float2 s2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* sw = s2 + warp * SZ;
    for (int row = blockIdx.x * ROWS_WARPS + warp; row < n_rows; row += gridDim.x * ROWS_WARPS) {
        const float2* src = buf + size_t(row) * P + lane;
        float2 v[32];
        if (RB == 32 || lane < RB) {
#pragma unroll
            for (int r = 0; r < RA; ++r) v[r] = src[r * RB];
        }
        fft2_pass1<P, true, false, RowLay>(v, sw, lane);
        __syncwarp();
        fft2_pass2<P, true, false, RowLay>(v, sw, lane, tw);
        __syncwarp();
        float2* dst = U + size_t(row) * P + lane;
#pragma unroll
        for (int r = 0; r < RB; ++r) dst[r * RA] = v[r];
    }
}

// ---------------------------------------------------------------------------
// pass C fused with the reconstruction and the loss sums: CTA = INVG_WARPS warps = the frames of one colour
// group for ONE window row (more frames than warps: several rounds).  Warp w inverse-transforms the row of
// frame w, stores U in place and leaves |U|^2 in its slice of shared memory; after a barrier the CTA adds the
// slices in frame order (deterministic), writes the row of I = mean |U|^2 (.abs()**2 + torch.mean(dim=1),
// env.py:172-173) and this row's float64 partial sums of tt.relativeLoss (sum I^2, sum I T, sum T^2);
// k_loss_final folds the G * P partials in index order.  Grid-stride over the G * P (group, row) tasks.
// ---------------------------------------------------------------------------
constexpr int INVG_WARPS = 8;

template <int P>
__global__ void __launch_bounds__(32 * INVG_WARPS, 2)
k2_rows_inv_group(float2* U, float* __restrict__ I, const float* __restrict__ T, const float2* __restrict__ tw,
                  int G, int Fg, double* __restrict__ partial) {
    constexpr int RA = Plan2<P, false>::RA, RB = Plan2<P, false>::RB;
    constexpr int SZ = RowLaySize<P, RA>::value;
    extern __shared__ float2 s2[];                   // INVG_WARPS exchange buffers, then INVG_WARPS x P floats
    __shared__ double shd[3][INVG_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* sw = s2 + warp * SZ;
    float* sq_all = reinterpret_cast<float*>(s2 + INVG_WARPS * SZ);
    float* sq = sq_all + warp * P;
    const size_t n2 = size_t(P) * P;
    const float inv = 1.f / float(Fg);
    fft_pdl_release();
    fft_pdl_wait();
    for (int task = blockIdx.x; task < G * P; task += gridDim.x) {
        const int g = task / P, y = task - g * P;
        for (int f0 = 0; f0 < Fg; f0 += INVG_WARPS) {
            const int fi = f0 + warp;
            if (fi < Fg) {
                float2* row = U + (size_t(g) * Fg + fi) * n2 + size_t(y) * P + lane;
                float2 v[32];
                if (RB == 32 || lane < RB) {
#pragma unroll
                    for (int r = 0; r < RA; ++r) v[r] = row[r * RB];
                }
                fft2_pass1<P, true, false, RowLay>(v, sw, lane);
                __syncwarp();
                fft2_pass2<P, true, false, RowLay>(v, sw, lane, tw);
                __syncwarp();
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    row[r * RA] = v[r];
                    const float a = fmaf(v[r].x, v[r].x, v[r].y * v[r].y);
                    if (f0 == 0) sq[lane + r * RA] = a; else sq[lane + r * RA] += a;
                }
            }
        }
        __syncthreads();
        double a = 0, b = 0, c = 0;
        if (tid < P / 4) {
            const int nw = Fg < INVG_WARPS ? Fg : INVG_WARPS;
            float4 acc = *reinterpret_cast<const float4*>(sq_all + 4 * tid);
            for (int w = 1; w < nw; ++w) {
                const float4 q = *reinterpret_cast<const float4*>(sq_all + w * P + 4 * tid);
                acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
            }
            const size_t p = size_t(g) * n2 + size_t(y) * P + 4 * tid;
            const float4 tv = __ldg(reinterpret_cast<const float4*>(T + p));
            const float4 iv = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
            *reinterpret_cast<float4*>(I + p) = iv;
            a = double(iv.x) * iv.x + double(iv.y) * iv.y + double(iv.z) * iv.z + double(iv.w) * iv.w;
            b = double(iv.x) * tv.x + double(iv.y) * tv.y + double(iv.z) * tv.z + double(iv.w) * tv.w;
            c = double(tv.x) * tv.x + double(tv.y) * tv.y + double(tv.z) * tv.z + double(tv.w) * tv.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) { shd[0][warp] = a; shd[1][warp] = b; shd[2][warp] = c; }
        __syncthreads();
        if (tid == 0) {
            double x = 0, yv = 0, z = 0;
            for (int i = 0; i < INVG_WARPS; ++i) { x += shd[0][i]; yv += shd[1][i]; z += shd[2][i]; }
            double* out = partial + size_t(task) * 3;
            out[0] = x; out[1] = yv; out[2] = z;
        }
    }
}

// dense row buffer used in place for the exchange: the low 4 position bits XORed with the block index
// (p / RA) -- conflict-free for the 16 lanes of a half-warp in both directions, a bijection inside each
// block since 16 | RA
template <int RA> struct RowSwz {
    static __device__ __forceinline__ int wr(int j, int r) { return (j * RA + r) ^ (j & 15); }
    static __device__ __forceinline__ int rd(int j, int r) { return (j + r * RA) ^ (r & 15); }
};

// ---------------------------------------------------------------------------
// pass C (fused), TMA-fed: one persistent CTA per SM, 8 warps = 8 frames of one (colour group, window row)
// task.  Every warp owns two 8 KB row buffers and two mbarriers: while it transforms its row of task i out of
// one buffer (pass 1 reads the row as the bulk copy left it, the exchange happens in place, XOR-swizzled), the
// bulk copy of its row of task i + 1 lands in the other -- the rows are contiguous 8 KB runs, the shape a bulk
// copy is made for.  U is stored from registers, |U|^2 goes to a per-warp slice of shared memory, the CTA adds
// the slices in frame order -> I row + float64 loss partials (same arithmetic as k2_rows_inv_group).
// ---------------------------------------------------------------------------
template <int P> constexpr size_t invg3_smem_bytes() {
    return size_t(INVG_WARPS) * 2 * P * sizeof(float2) + size_t(INVG_WARPS) * P * sizeof(float) + 128 + INVG_WARPS * 2 * 8;
}

template <int P>
__global__ void __launch_bounds__(32 * INVG_WARPS, 1)
k3_rows_inv_group(float2* U, float* __restrict__ I, const float* __restrict__ T, const float2* __restrict__ tw,
                  int G, int Fg, double* __restrict__ partial) {
    constexpr int RA = Plan2<P, false>::RA, RB = Plan2<P, false>::RB;
    constexpr unsigned ROW_BYTES = unsigned(P) * sizeof(float2);
    extern __shared__ __align__(128) float2 s2a[];
    float2* rows = s2a;
    float* sq_all = reinterpret_cast<float*>(rows + size_t(INVG_WARPS) * 2 * P);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sq_all + size_t(INVG_WARPS) * P);
    __shared__ double shd[3][INVG_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* my_rows = rows + size_t(warp) * 2 * P;
    uint64_t* my_bar = bars + warp * 2;
    float* sq = sq_all + warp * P;
    const size_t n2 = size_t(P) * P;
    const float inv = 1.f / float(Fg);
    const int n_tasks = G * P;
    const int rounds = (Fg + INVG_WARPS - 1) / INVG_WARPS;
    // the sequence of (task, round) items of this CTA; item k uses buffer k & 1
    const int my_tasks = (n_tasks - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
    const int n_items = my_tasks * rounds;
    if (lane == 0) { mbar_init(&my_bar[0], 1); mbar_init(&my_bar[1], 1); fence_barrier_init(); }
    __syncwarp();
    auto row_of = [&](int item) -> float2* {
        const int task = int(blockIdx.x) + (item / rounds) * int(gridDim.x);
        const int fi = (item % rounds) * INVG_WARPS + warp;
        const int g = task / P, y = task - g * P;
        return fi < Fg ? U + (size_t(g) * Fg + fi) * n2 + size_t(y) * P : nullptr;
    };
    auto prefetch = [&](int item) {                    // lane 0
        const float2* src = row_of(item);
        if (src) {
            mbar_expect_tx(&my_bar[item & 1], ROW_BYTES);
            bulk_load(my_rows + size_t(item & 1) * P, src, ROW_BYTES, &my_bar[item & 1]);
        }
    };
    if (lane == 0 && n_items > 0) prefetch(0);
    unsigned phase[2] = {0u, 0u};
    for (int item = 0; item < n_items; ++item) {
        const int task = int(blockIdx.x) + (item / rounds) * int(gridDim.x);
        const int round = item % rounds;
        const int g = task / P, y = task - g * P;
        float2* row = row_of(item);
        float2* sb = my_rows + size_t(item & 1) * P;
        if (item + 1 < n_items) {
            // the other buffer was last written by this warp's exchange of item - 1: order those
            // generic-proxy accesses before the bulk copy that overwrites it
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) prefetch(item + 1);
        }
        if (row) {
            mbar_wait(&my_bar[item & 1], phase[item & 1]); phase[item & 1] ^= 1u;
            float2 v[32];
            if (RB == 32 || lane < RB) {
#pragma unroll
                for (int r = 0; r < RA; ++r) v[r] = sb[lane + r * RB];
            }
            __syncwarp();                              // every lane has its inputs before the in-place exchange
            fft2_pass1<P, true, false, RowSwz>(v, sb, lane);
            __syncwarp();
            fft2_pass2<P, true, false, RowSwz>(v, sb, lane, tw);
            float2* dst = row + lane;
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                dst[r * RA] = v[r];
                const float a = fmaf(v[r].x, v[r].x, v[r].y * v[r].y);
                if (round == 0) sq[lane + r * RA] = a; else sq[lane + r * RA] += a;
            }
        }
        if (round != rounds - 1) continue;
        __syncthreads();
        double a = 0, b = 0, c = 0;
        if (tid < P / 4) {
            const int nw = Fg < INVG_WARPS ? Fg : INVG_WARPS;
            float4 acc = *reinterpret_cast<const float4*>(sq_all + 4 * tid);
            for (int w = 1; w < nw; ++w) {
                const float4 q = *reinterpret_cast<const float4*>(sq_all + w * P + 4 * tid);
                acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
            }
            const size_t p = size_t(g) * n2 + size_t(y) * P + 4 * tid;
            const float4 tv = __ldg(reinterpret_cast<const float4*>(T + p));
            const float4 iv = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
            *reinterpret_cast<float4*>(I + p) = iv;
            a = double(iv.x) * iv.x + double(iv.y) * iv.y + double(iv.z) * iv.z + double(iv.w) * iv.w;
            b = double(iv.x) * tv.x + double(iv.y) * tv.y + double(iv.z) * tv.z + double(iv.w) * tv.w;
            c = double(tv.x) * tv.x + double(tv.y) * tv.y + double(tv.z) * tv.z + double(tv.w) * tv.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) { shd[0][warp] = a; shd[1][warp] = b; shd[2][warp] = c; }
        __syncthreads();
        if (tid == 0) {
            double x = 0, yv = 0, z = 0;
            for (int i = 0; i < INVG_WARPS; ++i) { x += shd[0][i]; yv += shd[1][i]; z += shd[2][i]; }
            double* out = partial + size_t(task) * 3;
            out[0] = x; out[1] = yv; out[2] = z;
        }
    }
}

// ---------------------------------------------------------------------------
// pass B: for tiles of COLW = 4 canvas columns: FFT along y, multiply by H, inverse FFT along y, in place on
// buf (H carries 1/P^2).  Persistent CTAs of 128 threads (thread = column w x butterfly slot j), 2 per SM,
// each owns a contiguous range of tiles ordered (group, column tile, frame) so that consecutive tiles share
// their H tile.  Per tile:
//     TMA   [P x 4] input tile  -> s_in   (4 boxes of P/4 rows, one mbarrier)      prefetched one tile ahead
//     TMA   [P x 4] H tile      -> s_h    (only when the column tile changes)
//     forward transform: s_in -> registers -> pass 1 -> s_wk -> pass 2 -> F (registers)
//     column kx:     F * H       -> inverse pass 1 -> s_in (free by now) -> pass 2 -> global column kx
//     column P - kx: F * conj H  -> inverse pass 1 -> s_wk               -> pass 2 -> conj -> global column P - kx
// HERM = false (complex input): every column is transformed, no mirror column.
// Four CTA barriers per tile; shared memory 3 x 8 P W bytes = 96 KB at P = 1024 (cols2_smem_bytes).
// ---------------------------------------------------------------------------
// + the twiddle blocks of both plans (LDS instead of 3 x 31 LDG per thread and tile)
template <int P, int W> constexpr size_t cols2_smem_bytes() {
    return size_t(3) * P * W * sizeof(float2) + 128 + size_t(TwLayout<P>::total_all) * sizeof(float2) + 16;
}

// IN_TMA = false: the input tile is gathered straight into the butterfly registers (32-byte row segments,
// 8 rows per warp instruction) instead of through s_in; only the H tile (shared by the Fg frames of a column
// tile) arrives by TMA.
template <int P, bool HERM, bool IN_TMA, int W>
__global__ void __launch_bounds__(32 * W, W == 4 ? 2 : 1)
k2_cols(const __grid_constant__ CUtensorMap map_buf, const __grid_constant__ CUtensorMap map_h,
        float2* buf, const float2* __restrict__ tw, int n_groups, int Fg, int h_group0) {
    constexpr int RA = Plan2<P, false>::RA, RB = Plan2<P, false>::RB;     // forward (RA, RB), inverse (RB, RA)
    constexpr int NT = HERM ? P / (2 * W) + 1 : P / W;                    // column tiles per frame
    constexpr unsigned TILE_BYTES = unsigned(P) * W * sizeof(float2);
    // 3 tiles + 2 mbarriers.  The array is declared 128-byte aligned (TMA destination) and indexed directly: an
    // address laundered through an integer makes the compiler emit generic LD / ST instead of LDS / STS.
    extern __shared__ __align__(128) float2 s2a[];
    float2* s_in = s2a;
    float2* s_h = s_in + P * W;
    float2* s_wk = s_h + P * W;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_wk + P * W);
    uint64_t& bar_in = bars[0];
    uint64_t& bar_h = bars[1];
    float2* stw = s_wk + P * W + 16;                   // 128 bytes after the barriers
    const float2* stw_fwd = stw + Plan2<P, false>::tw_off;
    const float2* stw_rev = stw + Plan2<P, true>::tw_off;
    const int tid = threadIdx.x, w = tid & (W - 1), j = tid / W;
    const long long total = (long long)n_groups * NT * Fg;
    const int beg = int((long long)blockIdx.x * total / gridDim.x);
    const int end = int((long long)(blockIdx.x + 1) * total / gridDim.x);
    fft_pdl_release();
    if (beg >= end) return;
    if (tid == 0) {
        mbar_init(&bar_in, 1); mbar_init(&bar_h, 1);
        fence_barrier_init();
    }
    for (int i = tid; i < TwLayout<P>::total_all; i += 32 * W) stw[i] = __ldg(tw + i);   // constant table
    __syncthreads();
    fft_pdl_wait();
    auto issue_in = [&](int idx) {                     // one thread
        const int gt = idx / Fg, fi = idx - gt * Fg, g = gt / NT, t = gt - g * NT;
        mbar_expect_tx(&bar_in, TILE_BYTES);
#pragma unroll
        for (int b = 0; b < 4; ++b)
            tma_load_3d(s_in + b * (P / 4) * W, &map_buf, &bar_in, 2 * W * t, b * (P / 4), g * Fg + fi);
    };
    auto issue_h = [&](int idx) {
        const int gt = idx / Fg, g = gt / NT, t = gt - g * NT;
        mbar_expect_tx(&bar_h, TILE_BYTES);
#pragma unroll
        for (int b = 0; b < 4; ++b)
            tma_load_3d(s_h + b * (P / 4) * W, &map_h, &bar_h, 2 * W * t, b * (P / 4), h_group0 + g);
    };
    if (tid == 0) { if (IN_TMA) issue_in(beg); issue_h(beg); }
    unsigned ph_in = 0, ph_h = 0;
    bool h_pending = true;
    for (int idx = beg; idx < end; ++idx) {
        const int gt = idx / Fg, fi = idx - gt * Fg, g = gt / NT, t = gt - g * NT;
        const int kx = W * t + w;
        const bool col_ok = HERM ? (kx <= P / 2) : true;
        const bool mirror_ok = HERM && col_ok && kx != 0 && kx != P / 2;
        float2* frame = buf + size_t(g * Fg + fi) * P * P;
        float2 v[32];
        if (IN_TMA) {
            __syncthreads();                           // B0: the previous tile's reads of s_wk are done
            mbar_wait(&bar_in, ph_in); ph_in ^= 1u;
            if (RB == 32 || j < RB) {
#pragma unroll
                for (int r = 0; r < RA; ++r) v[r] = s_in[(j + r * RB) * W + w];
            }
        } else {
            if (RB == 32 || j < RB) {
                const float2* src = frame + size_t(j) * P + kx;
#pragma unroll
                for (int r = 0; r < RA; ++r) v[r] = src[size_t(r) * RB * P];
            }
            __syncthreads();                           // B0 (the loads are in flight across it)
        }
        fft2_pass1<P, false, false, ColLayW<W>::template L>(v, s_wk + w, j);
        __syncthreads();                               // B1
        fft2_pass2_read<P, false, ColLayW<W>::template L>(v, s_wk + w, j);
        fft2_pass2_math<P, false, false>(v, j, stw_fwd);
        if (h_pending) { mbar_wait(&bar_h, ph_h); ph_h ^= 1u; h_pending = false; }
        // thread j < RA holds F[j + r RA], r < RB: the entry pattern of the reversed plan
        float2 a[32];
        if (RA == 32 || j < RA) {
#pragma unroll
            for (int r = 0; r < RB; ++r) a[r] = cmul(v[r], s_h[(j + r * RA) * W + w]);
        }
        fft2_pass1<P, true, true, ColLayW<W>::template L>(a, s_in + w, j);
        __syncthreads();                               // B2
        fft2_pass2_read<P, true, ColLayW<W>::template L>(a, s_in + w, j);
        fft2_pass2_math<P, true, true>(a, j, stw_rev);
        if (col_ok && (RB == 32 || j < RB)) {          // thread j < RB holds out[j + r RB], r < RA
#pragma unroll
            for (int r = 0; r < RA; ++r) frame[size_t(j + r * RB) * P + kx] = a[r];
        }
        if (HERM) {
            if (RA == 32 || j < RA) {
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const float2 hv = s_h[(j + r * RA) * W + w];
                    a[r] = cmul(v[r], make_float2(hv.x, -hv.y));
                }
            }
            fft2_pass1<P, true, true, ColLayW<W>::template L>(a, s_wk + w, j);
        }
        fence_proxy_async();                           // s_in was written by threads; the next TMA overwrites it
        __syncthreads();                               // B3: all reads of s_in and s_h are done
        if (idx + 1 < end) {
            const bool new_h = (idx + 1) / Fg != gt;
            if (tid == 0) { if (IN_TMA) issue_in(idx + 1); if (new_h) issue_h(idx + 1); }
            h_pending = new_h;
        }
        if (HERM) {
            fft2_pass2_read<P, true, ColLayW<W>::template L>(a, s_wk + w, j);
            fft2_pass2_math<P, true, true>(a, j, stw_rev);
            if (mirror_ok && (RB == 32 || j < RB)) {
#pragma unroll
                for (int r = 0; r < RA; ++r) frame[size_t(j + r * RB) * P + (P - kx)] = make_float2(a[r].x, -a[r].y);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Latency-hiding variants of pass A and pass C (round 2, session 4).  The passes are latency bound (issue slots
// 30-40 % busy, DRAM 25 %): a warp loads its row, waits a full L2 round trip, transforms, stores, and only then
// asks for the next row.  Here the NEXT row travels by cp.async (LDGSTS.128, no staging registers) while the
// current one is transformed, and the pass-2 twiddle block lives in shared memory (LDS instead of 31 LDG per
// transform competing with the row stores for the load/store queue).
// ---------------------------------------------------------------------------
template <int P> constexpr int tw_block_len() { return (Plan2<P, false>::RB - 1) * Plan2<P, false>::RA; }
template <int P> constexpr size_t rows4_fwd_smem_bytes() {
    return size_t(tw_block_len<P>()) * sizeof(float2)
         + size_t(ROWS_WARPS) * (RowLaySize<P, Plan2<P, false>::RA>::value * sizeof(float2) + 2 * P);
}

// pass A, real input.  The two state rows of a pair are 2 P contiguous bytes: they arrive in a per-warp staging
// buffer by cp.async one pair ahead; the stride-RB byte gather of the butterfly entry pattern reads shared memory
// (32 consecutive bytes per instruction: 8 banks, 4 lanes per word).
// late_wait: the launch does not depend on its predecessor's memory (colour groups g >= 1 of one propagation:
// the state was complete before group 0 started), so the CTAs may fill the SMs the predecessor's last wave
// leaves idle; the wait at the END keeps the completion order of the stream transitive.
template <int P>
__global__ void __launch_bounds__(32 * ROWS_WARPS, 4)
k4_rows_fwd_real(const int8_t* __restrict__ state, float2* __restrict__ buf, const float2* __restrict__ tw,
                 int n_pairs, int late_wait) {
    constexpr int RA = Plan2<P, false>::RA, RB = Plan2<P, false>::RB;
    constexpr int SZ = RowLaySize<P, RA>::value;
    constexpr int NTW = (RB - 1) * RA;
    extern __shared__ __align__(16) float2 s2[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* stw = s2;
    float2* sw = s2 + NTW + warp * SZ;
    int8_t* stg = reinterpret_cast<int8_t*>(s2 + NTW + ROWS_WARPS * SZ) + warp * 2 * P;
    fft_pdl_release();
    for (int i = threadIdx.x; i < NTW; i += 32 * ROWS_WARPS) stw[i] = __ldg(tw + i);   // constant table
    if (!late_wait) fft_pdl_wait();
    const int stride = gridDim.x * ROWS_WARPS;
    int task = blockIdx.x * ROWS_WARPS + warp;
    auto prefetch = [&](int t) {
        const int f = t / (P / 2), yp = t - f * (P / 2);
        const int8_t* src = state + (size_t(f) * P + 2 * yp) * P;
#pragma unroll
        for (int i = lane; i < 2 * P / 16; i += 32) cpa16(stg + 16 * i, src + 16 * i);
        cpa_commit();
    };
    if (task < n_pairs) prefetch(task);
    __syncthreads();
    for (; task < n_pairs; task += stride) {
        const int f = task / (P / 2), yp = task - f * (P / 2);
        cpa_wait_all();
        __syncwarp();
        float2 v[32];
        if (RB == 32 || lane < RB) {
#pragma unroll
            for (int r = 0; r < RA; ++r)
                v[r] = make_float2(byte_to_float(stg[lane + r * RB]), byte_to_float(stg[P + lane + r * RB]));
        }
        __syncwarp();                                  // the staging buffer is free
        if (task + stride < n_pairs) prefetch(task + stride);
        fft2_pass1<P, false, false, RowLay>(v, sw, lane);
        __syncwarp();
        fft2_pass2_read<P, false, RowLay>(v, sw, lane);
        fft2_pass2_math<P, false, false>(v, lane, stw);
        __syncwarp();
        float2* oa = buf + (size_t(f) * P + 2 * yp) * P + lane;
        float2* ob = oa + P;
        const int src = (RA - lane) & (RA - 1);
#pragma unroll
        for (int r = 0; r <= RB / 2; ++r) {
            float2 zm;
            zm.x = __shfl_sync(0xffffffffu, v[RB - 1 - r].x, src);
            zm.y = __shfl_sync(0xffffffffu, v[RB - 1 - r].y, src);
            if (lane == 0) zm = v[(RB - r) % RB];
            const float2 zk = v[r];
            if (r < RB / 2 || lane == 0) {
                oa[r * RA] = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                ob[r * RA] = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
            }
        }
    }
    if (late_wait) fft_pdl_wait();
}

}  // namespace bh
