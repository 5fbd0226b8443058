// Incremental path of the hologram reward / DBS engine (sm_100a): every env step and every DBS
// candidate is scored by k_eval / k_eval_bundle and kept by k_commit; k_recon_plan + k_recon_batch
// materialise obs["recon_image"] for all environments of a step.
//
//   replaces the per-step full re-simulation of env.py:170-174, DBS.py:259-270,
//   DBS_1024_24.py:324-352, env_group.py:96-119 by the delta identity
//       U'_f = U_f + s * shift(h_g)          
//       dI   = (2 s Re(conj(U_f) h) + |h|^2) / Fg
//
// Work decomposition.  A "unit" is UNIT_PX = 1024 consecutive pixels of one candidate's N x N
// image (one pass of a 256-thread CTA at 4 px per thread).  The n_tasks * units_per_task units of a
// launch are split into gridDim.x contiguous, balanced ranges (CTA b owns
// [b*total/grid, (b+1)*total/grid)), so every resident CTA streams the same number of bytes
// (+-1 unit) and a CTA crosses at most a few task boundaries.
//
// Arithmetic.  With A = (2/Fg) Re(conj(U) h) and M = |h|^2 / Fg  (dI = s A + M, s = +-1):
//     d sum(I T) = s * sum(A T)        + sum(M T)                    = s p1 + p2
//     d sum(I^2) = s * sum(2 A (I+M))  + sum(A^2 + M (2 I + M))      = s p3 + p4
// The four sums do not depend on the sign of the flip, so the streaming loop never waits for the
// state byte; the sign is applied to the exact integer totals afterwards.  Per pixel quad the four
// partial sums are formed in R = float (default) or R = double (BHOLO_EVAL_FP64=1; products of two fp32
// values are exact in double -- measured: the error of dPSNR is set by the fp32 fields, not by this
// arithmetic, profiles/r2_notes.md), converted to 2^-40 fixed point and from there on added as
// 64-bit integers (registers -> warp shuffles -> shared -> one global atomic per CTA, task and sum).
// Integer addition is associative, so the sums are bit-identical for every grid size, batch
// composition, speculation depth and GPU count.
//
// Fixed-point range (documented bound, tested in tests/test_gpu_parity.py):
//   one quad's partial sum is rounded to a multiple of 2^-40 (|error| <= 2^-41 per quad and sum);
//   an N x N task has N^2/4 quads, so |error of a total| <= N^2 * 2^-43  (1.2e-7 at N = 1024;
//   typical, random rounding: sqrt(N^2/4) * 2^-41 / sqrt(3) = 1.3e-10).  Totals are carried in the
//   upper 54 bits of a 64-bit word: |total| < 2^13 = 8192, while
//   |d sum(I^2)| <= (2 sqrt(sum|U|^2) + 1) / Fg * (2 max I + max dI) is below 10^3 for every
//   supported shape.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bh {

struct Result {            // mirrored by bh_result in include/bholo.h (40 bytes)
    double psnr_after;
    double d_sii;
    double d_sit;
    long long action;
    int32_t accept;
    int32_t sgn;
};

enum { RULE_ENV = 0, RULE_DBS = 1, RULE_NEVER = 2, RULE_ALWAYS = 3 };

constexpr int UNIT_PX = 1024;
// Rows of the impulse-response table carry H_PAD wrapped columns (h[y][P + j] = h[y][j]), so the
// four taps of a pixel quad are always contiguous: one address, no per-tap wrap.
constexpr int H_PAD = 4;
__host__ __device__ constexpr int h_stride(int P) { return P + H_PAD; }
constexpr double FIX_SCALE_D = 1099511627776.0;         // 2^40
constexpr float FIX_SCALE = 1099511627776.0f;
constexpr double FIX_INV = 1.0 / 1099511627776.0;
// accumulator word of one task and sum: (fixed-point total << CNT_BITS) + number of CTAs that
// have contributed.  One returning atomic per CTA adds its share and tells it whether it was last.
constexpr int CNT_BITS = 10;
constexpr unsigned long long CNT_MASK = (1ull << CNT_BITS) - 1ull;
constexpr int MAX_DELTA_GRID = int(CNT_MASK);           // CTAs per task must fit the counter

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
    unsigned long long* acc;       // [n_tasks][2] accumulator words (zero between launches)
    unsigned* tickets;             // [n_tasks] rendezvous of the (rare) split finalisation
    Result* results;               // [n_tasks]
    Result* results_host;          // optional mapped pinned mirror written by the finaliser
    // speculative greedy DBS (k_commit does the selection): decision log, PSNR trace, counter
    uint8_t* dbs_accepted; double* dbs_trace; long long* dbs_count; long long* dbs_cursor;
    double* dbs_s0;                // [4] running sums + PSNR the window was scored against (written by the finalisers)
    // small batches (one env step of <= INLINE_MAX envs) carry their tasks in the
    // kernel parameters: no host-to-device copy on the step path
    int n_inline;
    long long inl_actions[32];
    int inl_envs[32];
    // bundled evaluation (k_eval_bundle_t): order a speculation window by frame inside the CTA
    int sort_window;
    // batched greedy DBS (bh_dbs_run_batch): decision / PSNR log of this iteration, [n_tasks]
    uint8_t* log_accept; double* log_psnr;
    // observation bookkeeping: k_commit marks the plane of a kept flip stale in every
    // observation buffer (see k_recon_plan); [E][RECON_MAX_BUFFERS] or nullptr
    uint8_t* recon_stale;
    // k_eval: units of the CTA's stream requested into L2 before the wait for the predecessor (0 = off)
    int l2_prefetch_units;
};
constexpr int INLINE_MAX = 32;
constexpr int RECON_MAX_BUFFERS = 4;

struct Decoded {
    int env, f, g, r, c; bool active;
};

// decode (env.py:158-161); pure arithmetic, no memory access
__device__ __forceinline__ Decoded decode_action(const DeltaArgs& a, int k, long long act) {
    Decoded d;
    d.env = a.n_inline ? a.inl_envs[k] : (a.envs ? a.envs[k] : a.env_fixed);
    d.f = d.g = d.r = d.c = 0;
    d.active = act >= 0;
    if (!d.active) return d;
    const int n2 = a.N * a.N;
    d.f = int(act / n2);
    const int pix = int(act - (long long)d.f * n2);
    d.r = pix / a.N;
    d.c = pix - d.r * a.N;
    d.g = d.f / a.Fg;
    return d;
}

__device__ __forceinline__ const int8_t* state_byte(const DeltaArgs& a, const Decoded& d) {
    return a.state + (size_t(d.env) * a.F + d.f) * (size_t(a.N) * a.N) + size_t(d.r) * a.N + d.c;
}
// The sign of a flip comes from the resident state (+1: pixel 0 -> 1).  Issued as a volatile asm
// load so that it stays in front of the streaming loads in program order: its latency then hides
// behind them, the value is only used after the last pixel.
__device__ __forceinline__ int ld_state_issue(const int8_t* p) {
    int v;
    asm volatile("ld.global.s8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// the same through L2 (never served from a stale L1 line): k_rollout_t reads bytes other CTAs wrote earlier in
// the same launch
__device__ __forceinline__ int ld_state_relaxed(const int8_t* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s8 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// L2 cache-policy accesses: the streamed operands (U, I, T) are touched once per candidate and
// marked evict-first; the impulse response is re-read by every candidate and marked evict-last
// so it stays L2 resident (25 MB for 3 colours).
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p, uint64_t pol) {      // read-only data
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
// data the same kernel writes back (k_commit): the non-coherent path is only defined for data
// that stays read-only for the whole kernel, so these go through the ordinary global path
__device__ __forceinline__ float4 ld_rw4(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ float2 ld_keep2(const float2* p, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
                 : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}

// Programmatic dependent launch: the step path is a chain eval -> commit -> eval ... of grids
// that each fill the machine in one wave.  Every kernel lets its successor be scheduled at once
// (its CTAs land on SMs as this grid's tail drains) and waits for its predecessor's memory only
// right before its first global access, after the parameter-only prologue.
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct Quad {                      // operands of 4 consecutive pixels
    float4 ua, ub, iv, tv;
    float2 h0, h1, h2, h3;
};

// position of the thread's quad inside the image, advanced unit by unit (generic images)
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

template <bool WITH_T, bool RW>
__device__ __forceinline__ void load_quad(Quad& q, const float2* U, const float* I, const float* T,
                                          const float2* h, const Cursor& cu, int N, int P, int r, int c,
                                          uint64_t pf, uint64_t pl) {
    const size_t p = size_t(cu.y) * N + cu.x;
    const float4* Up = reinterpret_cast<const float4*>(U + p);
    if (RW) {
        q.ua = ld_rw4(Up, pf); q.ub = ld_rw4(Up + 1, pf);
        q.iv = ld_rw4(reinterpret_cast<const float4*>(I + p), pf);
    } else {
        q.ua = ld_stream4(Up, pf); q.ub = ld_stream4(Up + 1, pf);
        q.iv = ld_stream4(reinterpret_cast<const float4*>(I + p), pf);
    }
    if (WITH_T) q.tv = ld_stream4(reinterpret_cast<const float4*>(T + p), pf);
    int hy = cu.y - r; if (hy < 0) hy += P;
    int hx = cu.x - c; if (hx < 0) hx += P;
    const float2* hq = h + size_t(hy) * h_stride(P) + hx;
    q.h0 = ld_keep2(hq, pl); q.h1 = ld_keep2(hq + 1, pl);
    q.h2 = ld_keep2(hq + 2, pl); q.h3 = ld_keep2(hq + 3, pl);
}

// ---------------------------------------------------------------------------
// per-quad arithmetic, shared by k_eval_t and k_eval_bundle_t (identical results)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float fm(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fm(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ long long to_fix(float v) { return __float2ll_rn(v * FIX_SCALE); }
__device__ __forceinline__ long long to_fix(double v) { return __double2ll_rn(v * FIX_SCALE_D); }

template <typename R>
__device__ __forceinline__ void px_terms(float ur_, float ui_, float2 h_, float i_, float t_, R c2, R invFg,
                                         R& p1, R& p2, R& p3, R& p4) {
    const R ur = R(ur_), ui = R(ui_), hr = R(h_.x), hi = R(h_.y), iv = R(i_), tv = R(t_);
    const R a = fm(ur, hr, ui * hi) * c2;            // A
    const R m = fm(hr, hr, hi * hi) * invFg;         // M
    const R im = iv + m;
    p1 = fm(a, tv, p1);
    p2 = fm(m, tv, p2);
    p3 = fm(a, im + im, p3);
    p4 += fm(a, a, m * (iv + im));
}

// t[0] = fix(p3), t[1] = fix(p4)  (sum I^2: odd and even part in the sign)
// t[2] = fix(p1), t[3] = fix(p2)  (sum I T)
template <typename R>
__device__ __forceinline__ void quad_terms(const Quad& q, R c2, R invFg, long long (&t)[4]) {
    R p1 = R(0), p2 = R(0), p3 = R(0), p4 = R(0);
    px_terms<R>(q.ua.x, q.ua.y, q.h0, q.iv.x, q.tv.x, c2, invFg, p1, p2, p3, p4);
    px_terms<R>(q.ua.z, q.ua.w, q.h1, q.iv.y, q.tv.y, c2, invFg, p1, p2, p3, p4);
    px_terms<R>(q.ub.x, q.ub.y, q.h2, q.iv.z, q.tv.z, c2, invFg, p1, p2, p3, p4);
    px_terms<R>(q.ub.z, q.ub.w, q.h3, q.iv.w, q.tv.w, c2, invFg, p1, p2, p3, p4);
    t[0] = to_fix(p3); t[1] = to_fix(p4); t[2] = to_fix(p1); t[3] = to_fix(p2);
}

// NACC = 4: sign-agnostic accumulators (k_eval_t); NACC = 2: the sign (sg = +-1) is applied to the
// integer terms of every quad (k_eval_bundle_t keeps two accumulators per candidate).
template <typename R, int NACC>
__device__ __forceinline__ void accumulate_quad(const Quad& q, R c2, R invFg, int sg, long long* acc) {
    long long t[4];
    quad_terms<R>(q, c2, invFg, t);
    if (NACC == 4) {
        acc[0] += t[0]; acc[1] += t[1]; acc[2] += t[2]; acc[3] += t[3];
    } else {
        acc[0] += (sg < 0 ? -t[0] : t[0]) + t[1];
        acc[1] += (sg < 0 ? -t[2] : t[2]) + t[3];
    }
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

// PSNR of the running sums (tt.relativeLoss + tm.get_PSNR in closed form, float64) and the accept rules
__device__ __forceinline__ double psnr_from_sums(const DeltaArgs& a, double sii, double sit, double stt, size_t n2) {
    const double n = double(a.G) * double(n2);
    const double mse = a.relative ? (stt - sit * sit / sii) / n
                                  : (sii - 2.0 * sit + stt) / n;
    return -10.0 * log10(mse);
}
__device__ __forceinline__ int apply_rule(int rule, double psnr, double prev) {
    if (rule == RULE_ENV) return !(psnr - prev < 0.0);      // env.py:191
    if (rule == RULE_DBS) return psnr > prev;               // DBS.py:273
    return rule == RULE_ALWAYS;
}

__device__ __forceinline__ void write_result(const DeltaArgs& a, int k, const Decoded& d, long long act,
                                             int sg, long long sII, long long sIT, const double (&S)[4],
                                             size_t n2) {
    const double dII = double(sII) * FIX_INV, dIT = double(sIT) * FIX_INV;
    const double psnr = psnr_from_sums(a, S[0] + dII, S[1] + dIT, S[2], n2);
    const int acc = apply_rule(a.rule, psnr, S[3]);
    Result r; r.psnr_after = psnr; r.d_sii = dII; r.d_sit = dIT;
    r.action = act; r.accept = acc; r.sgn = sg;
    a.results[k] = r;
    if (a.results_host) a.results_host[k] = r;
    if (a.log_accept) a.log_accept[k] = uint8_t(acc);
    if (a.log_psnr) a.log_psnr[k] = psnr;
    if (a.dbs_s0) { a.dbs_s0[0] = S[0]; a.dbs_s0[1] = S[1]; a.dbs_s0[2] = S[2]; a.dbs_s0[3] = S[3]; }
    a.acc[2 * k] = 0ull; a.acc[2 * k + 1] = 0ull;
    (void)d;
}

// One CTA's exact partial sums (x: sum I^2, y: sum I T, sign applied) of task k go to the task's
// two accumulator words with ONE returning atomic each; the word carries the number of
// contributors, so the CTA whose atomics complete both words holds both totals in registers and
// finalises without another memory round trip.  If two different CTAs complete the two words
// (their atomics interleaved), the two meet at a ticket and the second one reads the words back.
__device__ __forceinline__ void contribute_and_finalise(const DeltaArgs& a, int k, const Decoded& d,
                                                        long long act, int sg, long long x, long long y,
                                                        unsigned n_ctas, size_t n2, const double (&S)[4]) {
    const unsigned long long cx = ((unsigned long long)x << CNT_BITS) + 1ull;
    const unsigned long long cy = ((unsigned long long)y << CNT_BITS) + 1ull;
    const unsigned long long ox = atomicAdd(a.acc + 2 * k, cx);
    const unsigned long long oy = atomicAdd(a.acc + 2 * k + 1, cy);
    const bool lx = unsigned(ox & CNT_MASK) == n_ctas - 1u, ly = unsigned(oy & CNT_MASK) == n_ctas - 1u;
    if (!lx && !ly) return;
    unsigned long long wx = ox + cx, wy = oy + cy;
    if (!(lx && ly)) {
        __threadfence();
        const unsigned t = atomicAdd(a.tickets + k, 1u);
        if (t == 0u) return;
        __threadfence();
        wx = __ldcg(a.acc + 2 * k); wy = __ldcg(a.acc + 2 * k + 1);
        a.tickets[k] = 0u;
    }
    write_result(a, k, d, act, sg, (long long)wx >> CNT_BITS, (long long)wy >> CNT_BITS, S, n2);
}

__device__ __forceinline__ void write_idle_result(const DeltaArgs& a, int k) {
    Result r; r.psnr_after = 0.0; r.d_sii = 0.0; r.d_sit = 0.0;
    r.action = -1; r.accept = 0; r.sgn = 0;
    a.results[k] = r;
    if (a.results_host) a.results_host[k] = r;
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
// response of a candidate is ONE table offset that advances by unit_dy rows (modulo P).  Scores L
// candidates of one frame per pass over units [w0, w1) of the task, UF units in flight.
// RW: U and I are read through the coherent path (a kernel that also writes them: k_rollout_t).
template <int L, int UF, typename R, int NACC, bool RW = false>
__device__ __forceinline__ void eval_run_rows(const DeltaArgs& a, const float2* U, const float* I,
                                              const float* T, const float2* h, const int* rr,
                                              const int* cc, const int* sg, long long (*acc)[NACC],
                                              int w0, int w1, int tid, R c2, R invFg,
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
            ua[k] = RW ? ld_rw4(Up, pf) : ld_stream4(Up, pf);
            ub[k] = RW ? ld_rw4(Up + 1, pf) : ld_stream4(Up + 1, pf);
            iv[k] = RW ? ld_rw4(reinterpret_cast<const float4*>(I + p), pf)
                       : ld_stream4(reinterpret_cast<const float4*>(I + p), pf);
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
                accumulate_quad<R, NACC>(q, c2, invFg, sg[i], acc[i]);
            }
        }
    }
#pragma unroll 1
    for (; w < w1; ++w) {
        Quad q;
        const float4* Up = reinterpret_cast<const float4*>(U + p);
        q.ua = RW ? ld_rw4(Up, pf) : ld_stream4(Up, pf);
        q.ub = RW ? ld_rw4(Up + 1, pf) : ld_stream4(Up + 1, pf);
        q.iv = RW ? ld_rw4(reinterpret_cast<const float4*>(I + p), pf)
                  : ld_stream4(reinterpret_cast<const float4*>(I + p), pf);
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
            accumulate_quad<R, NACC>(q, c2, invFg, sg[i], acc[i]);
        }
    }
}

// units in flight for a run of L candidates: about four quads of taps per thread
__host__ __device__ constexpr int run_uf(int L) { return L == 1 ? 3 : (L == 2 ? 2 : 1); }

template <int L, int B, typename R>
__device__ __forceinline__ void eval_run_dispatch(int len, const DeltaArgs& a, const float2* U,
                                                  const float* I, const float* T, const float2* h,
                                                  const int* rr, const int* cc, const int* sg,
                                                  long long (*acc)[2], int w0, int w1,
                                                  int tid, R c2, R invFg, uint64_t pf, uint64_t pl) {
    if (len == L || L == B)
        eval_run_rows<L, run_uf(L), R, 2>(a, U, I, T, h, rr, cc, sg, acc, w0, w1, tid, c2, invFg, pf, pl);
    else if constexpr (L < B)
        eval_run_dispatch<L + 1, B, R>(len, a, U, I, T, h, rr, cc, sg, acc, w0, w1, tid, c2, invFg, pf, pl);
}

// k_eval: streams U (8 B/px), I and T (4 B/px each) once per candidate and the
// shifted impulse response from L2: 16 N^2 algorithmic HBM bytes per candidate.
template <int UF, int MINB, typename R>
__global__ void __launch_bounds__(256, MINB)
k_eval_t(const DeltaArgs a) {
    __shared__ long long sh[4][8];
    pdl_release();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = a.N, P = a.P, upt = a.units_per_task;
    const long long total = (long long)a.n_tasks * upt;
    const long long beg = (long long)blockIdx.x * total / gridDim.x;
    const long long end = (long long)(blockIdx.x + 1) * total / gridDim.x;
    const R invFg = R(1) / R(a.Fg), c2 = R(2) / R(a.Fg);
    const uint64_t pf = policy_evict_first(), pl = policy_evict_last();
    const size_t n2 = size_t(N) * N;
    // the first task of this CTA: with inline tasks its decode is parameter-only and runs before
    // the predecessor's memory is waited for
    long long u = beg;
    bool waited = false;
    if (a.l2_prefetch_units > 0 && !a.offset_ptr && beg < end) {
        // The grid's CTAs land on the SMs while the predecessor's last CTAs drain, and block at the wait.
        // The task list (kernel parameters, or a device list nobody writes during the chain) is known, so
        // the first units of this CTA's stream are requested into L2 meanwhile: L2 is the coherence point,
        // a line the predecessor still writes is simply updated there.  After the wait the streaming loop
        // starts on L2 hits instead of a cold DRAM round trip.
        const int k0 = int(beg / upt);
        const long long act0 = task_action(a, k0);
        const Decoded d0 = decode_action(a, k0, act0);
        if (d0.active && a.unit_dx == 0) {
            const size_t px0 = size_t(beg - (long long)k0 * upt) * UNIT_PX;
            const char* Ub = reinterpret_cast<const char*>(a.U + (size_t(d0.env) * a.F + d0.f) * n2 + px0);
            const char* Ib = reinterpret_cast<const char*>(a.I + (size_t(d0.env) * a.G + d0.g) * n2 + px0);
            const char* Tb = reinterpret_cast<const char*>(a.T + (size_t(d0.env) * a.G + d0.g) * n2 + px0);
            long long avail = (long long)(k0 + 1) * upt - beg;            // units left in the first task
            if (avail > end - beg) avail = end - beg;
            const int pf = avail < a.l2_prefetch_units ? int(avail) : a.l2_prefetch_units;
            // per unit: 64 lines of U, 32 of I, 32 of T (128-byte lines); thread t takes line t of a unit pair
            for (int ln = tid; ln < pf * 128; ln += 256) {
                const int un = ln >> 7, l = ln & 127;
                const char* ptr = l < 64 ? Ub + size_t(un) * UNIT_PX * 8 + l * 128
                                : (l < 96 ? Ib + size_t(un) * UNIT_PX * 4 + (l - 64) * 128
                                          : Tb + size_t(un) * UNIT_PX * 4 + (l - 96) * 128);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
            }
        }
        pdl_wait(); waited = true;
    }
    while (u < end) {
        const int k = int(u / upt);
        const long long t_beg = (long long)k * upt;
        const long long seg_end = (t_beg + upt < end) ? t_beg + upt : end;
        if (!a.n_inline && !waited) { pdl_wait(); waited = true; }
        const long long act = task_action(a, k);
        const Decoded d = decode_action(a, k, act);
        if (!waited) { pdl_wait(); waited = true; }
        if (!d.active) {
            if (u == t_beg && tid == 0) write_idle_result(a, k);
            u = seg_end;
            continue;
        }
        const int sbyte = ld_state_issue(state_byte(a, d));          // used after the loop only
        double S[4] = {0.0, 0.0, 0.0, 0.0};
        if (tid == 0) {                                              // finaliser inputs, off the tail
            const double* Sp = a.sums + size_t(d.env) * 4;
            S[0] = Sp[0]; S[1] = Sp[1]; S[2] = Sp[2]; S[3] = Sp[3];
        }
        const float2* U = a.U + (size_t(d.env) * a.F + d.f) * n2;
        const float* I = a.I + (size_t(d.env) * a.G + d.g) * n2;
        const float* T = a.T + (size_t(d.env) * a.G + d.g) * n2;
        const float2* h = a.h + size_t(d.g) * P * a.HP;
        long long acc[1][4] = {{0, 0, 0, 0}};
        const int sg0 = 1;
        if (a.unit_dx == 0) {
            eval_run_rows<1, UF, R, 4>(a, U, I, T, h, &d.r, &d.c, &sg0, acc, int(u - t_beg),
                                       int(seg_end - t_beg), tid, c2, invFg, pf, pl);
        } else {
            Cursor cu; cu.init(int(u - t_beg), tid, N);
            long long v = u;
            for (; v + UF <= seg_end; v += UF) {         // UF units in flight per thread
                Quad q[UF];
#pragma unroll
                for (int i = 0; i < UF; ++i) {
                    load_quad<true, false>(q[i], U, I, T, h, cu, N, P, d.r, d.c, pf, pl); cu.next(a);
                }
#pragma unroll
                for (int i = 0; i < UF; ++i) accumulate_quad<R, 4>(q[i], c2, invFg, 1, acc[0]);
            }
            for (; v < seg_end; ++v) {
                Quad q0;
                load_quad<true, false>(q0, U, I, T, h, cu, N, P, d.r, d.c, pf, pl); cu.next(a);
                accumulate_quad<R, 4>(q0, c2, invFg, 1, acc[0]);
            }
        }
        const int sg = 1 - 2 * sbyte;
        long long aII = (sg < 0 ? -acc[0][0] : acc[0][0]) + acc[0][1];
        long long aIT = (sg < 0 ? -acc[0][2] : acc[0][2]) + acc[0][3];
        aII = warp_sum_ll(aII); aIT = warp_sum_ll(aIT);
        if (lane == 0) { sh[0][warp] = aII; sh[1][warp] = aIT; }
        __syncthreads();
        if (tid == 0) {
            long long x = 0, y = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { x += sh[0][i]; y += sh[1][i]; }
            const int first = cta_of_unit(t_beg, total, gridDim.x);
            const int last = cta_of_unit(t_beg + upt - 1, total, gridDim.x);
            contribute_and_finalise(a, k, d, act, sg, x, y, unsigned(last - first) + 1u, n2, S);
        }
        __syncthreads();
        u = seg_end;
    }
    if (!waited) pdl_wait();
}

// k_eval_bundle: candidate lists that revisit an environment (speculation windows of the greedy
// DBS, DBS.py:247-294; the candidate tables of env_group.py:96-119 and the sweeps) are scored B
// slots at a time.  Inside a bundle every run of candidates of one frame is ONE pass over the
// image: the thread that owns a quad loads U, I and T once per run and, per candidate, only the
// shifted impulse response (L2 resident).  HBM traffic per candidate falls from 16 N^2 B towards
// 16 N^2 / B; the per-quad terms and the 2^-40 fixed-point sums are those of k_eval_t, so
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

template <int B, int MINB, typename R>
__global__ void __launch_bounds__(256, MINB)
k_eval_bundle_t(const DeltaArgs a) {
    __shared__ long long sh[2 * B][8];
    __shared__ int s_key[SORT_WINDOW_MAX];
    __shared__ int s_order[SORT_WINDOW_MAX];
    pdl_release();
    pdl_wait();
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
    const R invFg = R(1) / R(a.Fg), c2 = R(2) / R(a.Fg);
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
            int rr[B], cc[B], sg[B]; long long acc[B][2];
            int len = 1;
            rr[0] = head.d.r; cc[0] = head.d.c; sg[0] = 1 - 2 * int(*state_byte(a, head.d));
            acc[0][0] = 0; acc[0][1] = 0;
#pragma unroll
            for (int i = 1; i < B; ++i) {
                rr[i] = 0; cc[i] = 0; sg[i] = 1; acc[i][0] = 0; acc[i][1] = 0;
                if (rb + i < slots && len == i) {
                    const BundleTask t = bundle_task(a, j * B + rb + i, sorted, s_order);
                    if (t.d.active && t.d.env == head.d.env && t.d.f == head.d.f) {
                        rr[i] = t.d.r; cc[i] = t.d.c; sg[i] = 1 - 2 * int(*state_byte(a, t.d));
                        len = i + 1;
                    }
                }
            }
            const float2* U = a.U + (size_t(head.d.env) * a.F + head.d.f) * n2;
            const float* I = a.I + (size_t(head.d.env) * a.G + head.d.g) * n2;
            const float* T = a.T + (size_t(head.d.env) * a.G + head.d.g) * n2;
            const float2* h = a.h + size_t(head.d.g) * P * a.HP;
            if (a.unit_dx == 0) {
                eval_run_dispatch<1, B, R>(len, a, U, I, T, h, rr, cc, sg, acc, int(u - t_beg),
                                           int(seg_end - t_beg), tid, c2, invFg, pf, pl);
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
                            accumulate_quad<R, 2>(q, c2, invFg, sg[i], acc[i]);
                        }
                    }
                    cu.next(a);
                }
            }
#pragma unroll
            for (int i = 0; i < B; ++i) {
                if (i < len) {
                    const long long x = warp_sum_ll(acc[i][0]), y = warp_sum_ll(acc[i][1]);
                    if (lane == 0) { sh[2 * i][warp] = x; sh[2 * i + 1][warp] = y; }
                }
            }
            __syncthreads();
            if (tid < len) {                                // one thread per candidate of the run
                const BundleTask t = bundle_task(a, j * B + rb + tid, sorted, s_order);
                long long x = 0, y = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { x += sh[2 * tid][w]; y += sh[2 * tid + 1][w]; }
                const double* Sp = a.sums + size_t(t.d.env) * 4;
                const double S[4] = {Sp[0], Sp[1], Sp[2], Sp[3]};
                int sgt = 1;
#pragma unroll
                for (int i = 0; i < B; ++i) if (i == tid) sgt = sg[i];
                contribute_and_finalise(a, t.task, t.d, t.act, sgt, x, y, n_ctas, n2, S);
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
// (n_tasks <= COMMIT_MAX_TASKS).  Every CTA first compacts the accepted tasks
// (one accept flag per thread, warp ballots), then the balanced unit partition runs over the
// accepted ones only, so a launch with one accepted flip out of K still uses the whole chip.
constexpr int COMMIT_MAX_TASKS = 256;

__device__ __forceinline__ float delta_px(float ur, float ui, float hr, float hi, float s2, float invFg) {
    const float a = fmaf(ur, hr, ui * hi);
    const float m = fmaf(hr, hr, hi * hi);
    return fmaf(s2, a, m * invFg);
}

__device__ __forceinline__ void commit_quad(const Quad& q, float2* U, float* I, size_t p, float s2,
                                            float sg, float invFg, uint64_t pf) {
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
    st_stream4(Up, ua, pf); st_stream4(Up + 1, ub, pf);
    st_stream4(reinterpret_cast<float4*>(I + p), iv, pf);
}

// Row-regular images: the commit loop over units [w0, w1) of one kept flip (fixed column per thread, one table
// offset advanced by rows), UF units in flight.  Shared by k_commit_t and k_rollout_t.
template <int UF>
__device__ __forceinline__ void commit_run_rows(const DeltaArgs& a, float2* U, float* I, const float2* h,
                                                int r, int c, float sg, int w0, int w1, int tid, float invFg,
                                                uint64_t pf, uint64_t pl) {
    const int N = a.N, P = a.P, HP = a.HP, dy = a.unit_dy;
    const float s2 = 2.f * sg * invFg;
    const int y0 = (tid * 4) / N, x = tid * 4 - y0 * N;
    const int y = w0 * dy + y0;
    size_t p = size_t(y) * N + x;
    const int hstep = dy * HP, hwrap = P * HP;
    int hx = x - c; if (hx < 0) hx += P;
    int hy = y - r; if (hy < 0) hy += P;
    int ho = hy * HP + hx;
    int w = w0;
#pragma unroll 1
    for (; w + UF <= w1; w += UF) {
        Quad q[UF];
#pragma unroll
        for (int i = 0; i < UF; ++i) {
            const float4* Up = reinterpret_cast<const float4*>(U + p + size_t(i) * UNIT_PX);
            q[i].ua = ld_rw4(Up, pf); q[i].ub = ld_rw4(Up + 1, pf);
            q[i].iv = ld_rw4(reinterpret_cast<const float4*>(I + p + size_t(i) * UNIT_PX), pf);
            const float2* hp = h + ho;
            q[i].h0 = ld_keep2(hp, pl); q[i].h1 = ld_keep2(hp + 1, pl);
            q[i].h2 = ld_keep2(hp + 2, pl); q[i].h3 = ld_keep2(hp + 3, pl);
            ho += hstep; if (ho >= hwrap) ho -= hwrap;
        }
#pragma unroll
        for (int i = 0; i < UF; ++i) commit_quad(q[i], U, I, p + size_t(i) * UNIT_PX, s2, sg, invFg, pf);
        p += size_t(UF) * UNIT_PX;
    }
#pragma unroll 1
    for (; w < w1; ++w) {
        Quad q;
        const float4* Up = reinterpret_cast<const float4*>(U + p);
        q.ua = ld_rw4(Up, pf); q.ub = ld_rw4(Up + 1, pf);
        q.iv = ld_rw4(reinterpret_cast<const float4*>(I + p), pf);
        const float2* hp = h + ho;
        q.h0 = ld_keep2(hp, pl); q.h1 = ld_keep2(hp + 1, pl);
        q.h2 = ld_keep2(hp + 2, pl); q.h3 = ld_keep2(hp + 3, pl);
        ho += hstep; if (ho >= hwrap) ho -= hwrap;
        commit_quad(q, U, I, p, s2, sg, invFg, pf);
        p += UNIT_PX;
    }
}

template <int UF, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_commit_t(const DeltaArgs a) {
    __shared__ int s_list[COMMIT_MAX_TASKS];
    __shared__ Result s_res[COMMIT_MAX_TASKS];
    __shared__ int s_wcnt[8];
    __shared__ int s_first;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_release();
    pdl_wait();
    // ONE round trip to global memory for everything the selection needs: the result record of task `tid`
    // (decision, action, sign, delta sums) goes to shared memory, the window's reference sums and the cursor
    // of the greedy DBS to registers -- the walk below and the per-segment set-up then read no global memory.
    int flag = 0;
    if (tid < a.n_tasks) {
        const Result r = a.results[tid];
        s_res[tid] = r;
        flag = r.accept != 0;
    }
    double s0[4] = {0.0, 0.0, 0.0, 0.0};
    long long cur0 = 0;
    if (a.dbs_cursor && tid == 0) {
        s0[0] = a.dbs_s0[0]; s0[1] = a.dbs_s0[1]; s0[2] = a.dbs_s0[2]; s0[3] = a.dbs_s0[3];
    }
    if (a.dbs_cursor && blockIdx.x == 0) cur0 = *a.dbs_cursor;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    if (tid == 0) s_first = 0x7fffffff;
    __syncthreads();
    int n_acc;
    if (a.dbs_cursor) {
        // speculative greedy DBS (DBS.py:247-294 order): all K candidates were scored against the same state S0.
        // Candidates before the first kept one were rejected against the right state.  After it, a candidate
        // of ANOTHER colour group still has exact delta sums (its frame's field and its group's I and T are
        // untouched; DBS_1024_24.py:324-352 re-simulates only the flipped group), only the running sums
        // moved: thread 0 walks on, re-decides such candidates in order with the updated sums (same float64
        // closed form, same order of additions as the sequential loop) and stops at the first candidate whose
        // group has been touched -- that one is scored again by the next batch.  Every CTA does the same walk
        // from the same inputs (results + S0 as published by the finalisers); block 0 logs and advances.
        if (bal && lane == 0) atomicMin(&s_first, warp * 32 + (__ffs(bal) - 1));
        __syncthreads();
        const int first = (s_first == 0x7fffffff) ? -1 : s_first;
        const bool log = (blockIdx.x == 0);
        long long off = 0, cnt = 0;
        if (log) {
            off = cur0;
            cnt = a.n_total - off;
            if (cnt > a.n_tasks) cnt = a.n_tasks;
        }
        if (tid == 0) {
            int kept = 0, used = -1;
            if (first >= 0) {
                const size_t n2w = size_t(a.N) * a.N;
                const Result r0 = s_res[first];
                double sii = s0[0] + r0.d_sii, sit = s0[1] + r0.d_sit, prev = r0.psnr_after;
                const double stt = s0[2];
                const unsigned all = (1u << a.G) - 1u;
                unsigned dirty = 1u << (int(r0.action / (long long)n2w) / a.Fg);
                s_list[0] = first; kept = 1;
                int k = first + 1;
                for (; k < a.n_tasks && dirty != all; ++k) {
                    const Result rk = s_res[k];
                    if (rk.action < 0) break;                        // idle slot: end of the list
                    const int g = int(rk.action / (long long)n2w) / a.Fg;
                    if ((dirty >> g) & 1u) break;
                    const double psnr = psnr_from_sums(a, sii + rk.d_sii, sit + rk.d_sit, stt, n2w);
                    const int acc = apply_rule(a.rule, psnr, prev);
                    if (log) {
                        a.dbs_accepted[off + k] = uint8_t(acc);
                        if (a.dbs_trace) a.dbs_trace[off + k] = psnr;
                    }
                    if (acc) { sii += rk.d_sii; sit += rk.d_sit; prev = psnr; dirty |= 1u << g; s_list[kept++] = k; }
                }
                used = k;
                if (log) {                                           // the sums after this batch (nobody reads them here)
                    double* S = a.sums + size_t(a.env_fixed) * 4;
                    S[0] = sii; S[1] = sit; S[3] = prev;
                }
            }
            s_wcnt[0] = kept; s_wcnt[1] = used;
        }
        __syncthreads();
        n_acc = s_wcnt[0];
        if (log && cnt > 0) {
            const int head = first >= 0 ? first + 1 : int(cnt);  // decided against S0 by the finalisers
            if (tid < head) {
                a.dbs_accepted[off + tid] = (tid == first) ? 1 : 0;
                if (a.dbs_trace) a.dbs_trace[off + tid] = s_res[tid].psnr_after;
            }
            if (tid == 0) {
                *a.dbs_count += n_acc;
                *a.dbs_cursor = off + (first >= 0 ? s_wcnt[1] : int(cnt));
            }
        }
    } else {
        int base = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) base += (w < warp) ? s_wcnt[w] : 0;
        if (flag) s_list[base + __popc(bal & ((1u << lane) - 1u))] = tid;
        n_acc = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) n_acc += s_wcnt[w];
    }
    __syncthreads();
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
        const Result res = s_res[k];
        const Decoded d = decode_action(a, k, res.action);
        float2* U = a.U + (size_t(d.env) * a.F + d.f) * n2;
        float* I = a.I + (size_t(d.env) * a.G + d.g) * n2;
        const float2* h = a.h + size_t(d.g) * P * a.HP;
        const float sg = float(res.sgn), s2 = 2.f * sg * invFg;
        if (u == t_beg && tid == 0) {                 // one thread per accepted flip: state byte + sums
            int8_t* st = a.state + (size_t(d.env) * a.F + d.f) * n2 + size_t(d.r) * N + d.c;
            *st = int8_t(res.sgn > 0 ? 1 : 0);
            if (!a.dbs_cursor) {                      // (the greedy DBS batch wrote its final sums above)
                double* S = a.sums + size_t(d.env) * 4;
                S[0] += res.d_sii;
                S[1] += res.d_sit;
                S[3] = res.psnr_after;
            }
            if (a.recon_stale) {
                uint8_t* sm = a.recon_stale + size_t(d.env) * RECON_MAX_BUFFERS;
#pragma unroll
                for (int b = 0; b < RECON_MAX_BUFFERS; ++b) sm[b] |= uint8_t(1 << d.g);
            }
        }
        if (a.unit_dx == 0) {
            commit_run_rows<UF>(a, U, I, h, d.r, d.c, sg, int(u - t_beg), int(seg_end - t_beg), tid, invFg, pf, pl);
        } else {
            Cursor cu; cu.init(int(u - t_beg), tid, N);
            long long v = u;
            for (; v + UF <= seg_end; v += UF) {
                Quad q[UF];
                size_t p[UF];
#pragma unroll
                for (int i = 0; i < UF; ++i) {
                    load_quad<false, true>(q[i], U, I, nullptr, h, cu, N, P, d.r, d.c, pf, pl);
                    p[i] = size_t(cu.y) * N + cu.x;
                    cu.next(a);
                }
#pragma unroll
                for (int i = 0; i < UF; ++i) commit_quad(q[i], U, I, p[i], s2, sg, invFg, pf);
            }
            for (; v < seg_end; ++v) {
                Quad q;
                load_quad<false, true>(q, U, I, nullptr, h, cu, N, P, d.r, d.c, pf, pl);
                commit_quad(q, U, I, size_t(cu.y) * N + cu.x, s2, sg, invFg, pf);
                cu.next(a);
            }
        }
        u = seg_end;
    }
}

// ---------------------------------------------------------------------------
// k_rollout: a whole open-loop rollout -- `steps` sequential flips per environment, actions known in advance
// (device-resident action lists: random-policy rollouts, replays, the candidate orders of a batched greedy
// DBS) -- in ONE persistent cooperative launch.  It removes the per-step fixed cost of the eval -> commit chain
// (two launches: CTA ramp, first DRAM round trip, tail, grid completion -- 10 of 48 us per vectorised step).
//
// Every environment is owned by `cpe` co-resident CTAs, each by a fixed slice of the image's units.  A step is
//     evaluate my slice (same per-quad arithmetic and 2^-40 fixed-point sums as k_eval_t)
//     -> one returning atomic per sum into the environment's ring slot (total << 10 | contributors)
//     -> spin until all cpe CTAs of THIS environment have contributed (no grid-wide barrier: environments drift
//        apart, the other CTA of the SM belongs to another environment and keeps the memory system busy)
//     -> every CTA turns the exact totals into PSNR and the decision itself (same float64 closed form)
//     -> commit my slice if kept (the same thread re-reads the quads it will evaluate next: program order)
// Decisions, results, U, I, state and sums are bit-identical to `steps` calls of k_eval_t + k_commit_t.
// Row-regular images only (N divides UNIT_PX); the host falls back to the two-kernel chain otherwise.
// The launch is cooperative (co-residency guaranteed or refused); a spin that exceeds ROLLOUT_SPIN_LIMIT polls
// raises *error and ends the kernel instead of hanging the device.
// ---------------------------------------------------------------------------
constexpr int ROLLOUT_RING = 4;
constexpr unsigned ROLLOUT_SPIN_LIMIT = 1u << 24;

struct RolloutArgs {
    DeltaArgs a;                       // fields, tables, sums, shape, rule (task lists unused)
    const int32_t* envs;               // [n_env] device env ids or nullptr (identity)
    const long long* actions;          // action of (step t, env slot e): actions[t * act_step + e * act_env]; < 0 = idle
    long long act_step, act_env;
    Result* results;                   // results[t * res_step + e * res_env] or nullptr
    long long res_step, res_env;
    uint8_t* log_accept; double* log_psnr;   // optional logs, indexed like results
    unsigned long long* ring;          // [n_env][ROLLOUT_RING][2], zero at launch
    int* error;                        // device flag, 0 at launch
    int n_env, steps, cpe;
    unsigned backoff_ns;               // sleep between two polls of a barrier word
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// LOOK: one step of look-ahead (few environments: nobody else streams while an environment sits at its barrier).
// After contributing to the barrier of step t a CTA evaluates the candidate of step t + 1 on its slice while the
// barrier resolves.  The sums of a candidate do not depend on the sign of its flip nor on the running sums (both
// enter afterwards), so they stay exact unless step t is kept AND touches the same colour group (its U_f / I_g
// changed): then the candidate is evaluated again after the commit.  Decisions are those of the sequential loop.
template <int UF, typename R, bool LOOK>
__global__ void __launch_bounds__(256, 2)
k_rollout_t(const RolloutArgs ra) {
    const DeltaArgs& a = ra.a;
    __shared__ long long sh[2][8];
    __shared__ int s_dec;                                  // 1 keep, 0 reject, -1 abort
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot_e = blockIdx.x / ra.cpe, part = blockIdx.x - slot_e * ra.cpe;
    const int env = ra.envs ? ra.envs[slot_e] : slot_e;
    const int N = a.N, P = a.P, upt = a.units_per_task;
    const int w0 = int((long long)part * upt / ra.cpe), w1 = int((long long)(part + 1) * upt / ra.cpe);
    const bool leader = (part == 0);
    const size_t n2 = size_t(N) * N;
    const R invFg = R(1) / R(a.Fg), c2 = R(2) / R(a.Fg);
    const float invFg_f = 1.f / float(a.Fg);
    const uint64_t pf = policy_evict_first(), pl = policy_evict_last();
    unsigned long long* ring = ra.ring + size_t(slot_e) * ROLLOUT_RING * 2;
    double S[4] = {0.0, 0.0, 0.0, 0.0};                    // running sums of the environment (thread 0 of every CTA)
    if (tid == 0) {
        const double* Sp = a.sums + size_t(env) * 4;
        S[0] = Sp[0]; S[1] = Sp[1]; S[2] = Sp[2]; S[3] = Sp[3];
    }
    long long prev_act = -1; int prev_sgn = 0, prev_keep = 0;
    int ring_step = 0;                                     // counts the active (non-idle) steps
    long long spec[4] = {0, 0, 0, 0};                      // look-ahead: this thread's sums of the next candidate
    bool have_spec = false;
    auto act_of = [&](int t) { return ra.actions[(long long)t * ra.act_step + (long long)slot_e * ra.act_env]; };
    auto evaluate = [&](long long act, long long (&out)[4]) {
        const int f = int(act / (long long)n2);
        const int pix = int(act - (long long)f * (long long)n2);
        const int r = pix / N, c = pix - r * N, g = f / a.Fg;
        const float2* U = a.U + (size_t(env) * a.F + f) * n2;
        const float* I = a.I + (size_t(env) * a.G + g) * n2;
        const float* T = a.T + (size_t(env) * a.G + g) * n2;
        const float2* h = a.h + size_t(g) * P * a.HP;
        long long acc[1][4] = {{0, 0, 0, 0}};
        const int sg0 = 1;
        eval_run_rows<1, UF, R, 4, true>(a, U, I, T, h, &r, &c, &sg0, acc, w0, w1, tid, c2, invFg, pf, pl);
        out[0] = acc[0][0]; out[1] = acc[0][1]; out[2] = acc[0][2]; out[3] = acc[0][3];
    };
    for (int t = 0; t < ra.steps; ++t) {
        const long long act = act_of(t);
        const long long ri = (long long)t * ra.res_step + (long long)slot_e * ra.res_env;
        if (act < 0) {                                     // idle slot: every CTA of the environment sees the same
            if (leader && tid == 0 && ra.results) {
                Result r; r.psnr_after = 0.0; r.d_sii = 0.0; r.d_sit = 0.0; r.action = -1; r.accept = 0; r.sgn = 0;
                ra.results[ri] = r;
            }
            continue;
        }
        const int f = int(act / (long long)n2);
        const int pix = int(act - (long long)f * (long long)n2);
        const int r = pix / N, c = pix - r * N, g = f / a.Fg;
        int8_t* stp = a.state + (size_t(env) * a.F + f) * n2 + size_t(r) * N + c;
        // sign of the flip: the resident state byte -- unless the previous step touched the same pixel (its byte
        // is written by the leader CTA after ITS barrier, not ordered with this read): then it follows from
        // the previous step's sign and decision.  Bytes written two or more steps ago are ordered by the
        // barrier in between.  (Read here, after the previous barrier: the sums below do not need it.)
        int sbyte;
        if (act == prev_act) sbyte = prev_keep ? (prev_sgn > 0 ? 1 : 0) : (prev_sgn > 0 ? 0 : 1);
        else sbyte = ld_state_relaxed(stp);
        long long acc[4];
        if (LOOK && have_spec) { acc[0] = spec[0]; acc[1] = spec[1]; acc[2] = spec[2]; acc[3] = spec[3]; }
        else evaluate(act, acc);
        have_spec = false;
        const int sg = 1 - 2 * sbyte;
        long long aII = (sg < 0 ? -acc[0] : acc[0]) + acc[1];
        long long aIT = (sg < 0 ? -acc[2] : acc[2]) + acc[3];
        aII = warp_sum_ll(aII); aIT = warp_sum_ll(aIT);
        if (lane == 0) { sh[0][warp] = aII; sh[1][warp] = aIT; }
        __syncthreads();
        unsigned long long* wd = ring + size_t(ring_step % ROLLOUT_RING) * 2;
        if (tid == 0) {
            long long x = 0, y = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { x += sh[0][i]; y += sh[1][i]; }
            __threadfence();                               // my earlier stores (ring zeroing, state byte) first
            atomicAdd(wd, ((unsigned long long)x << CNT_BITS) + 1ull);
            atomicAdd(wd + 1, ((unsigned long long)y << CNT_BITS) + 1ull);
        }
        // look-ahead: the next candidate's sums while the barrier resolves
        int g_next = -1;
        if (LOOK && t + 1 < ra.steps) {
            const long long an = act_of(t + 1);
            if (an >= 0) {
                g_next = int(an / (long long)n2) / a.Fg;
                evaluate(an, spec);
                have_spec = true;
            }
        }
        if (tid == 0) {
            unsigned long long wx = 0, wy = 0;
            unsigned spins = 0;
            bool ok = false;
            // poll the first word until it is complete, then the second (one load per poll)
            while (true) {
                wx = ld_acquire_u64(wd);
                if (unsigned(wx & CNT_MASK) == unsigned(ra.cpe)) {
                    wy = ld_acquire_u64(wd + 1);
                    if (unsigned(wy & CNT_MASK) == unsigned(ra.cpe)) { ok = true; break; }
                }
                if (++spins > ROLLOUT_SPIN_LIMIT) break;
                if ((spins & 255u) == 0u && *reinterpret_cast<volatile int*>(ra.error)) break;
                if (ra.backoff_ns) __nanosleep(ra.backoff_ns);
            }
            int keep = -1;
            if (ok) {
                const double dII = double((long long)wx >> CNT_BITS) * FIX_INV;
                const double dIT = double((long long)wy >> CNT_BITS) * FIX_INV;
                const double psnr = psnr_from_sums(a, S[0] + dII, S[1] + dIT, S[2], n2);
                keep = apply_rule(a.rule, psnr, S[3]);
                if (leader) {
                    if (ra.results) {
                        Result res; res.psnr_after = psnr; res.d_sii = dII; res.d_sit = dIT;
                        res.action = act; res.accept = keep; res.sgn = sg;
                        ra.results[ri] = res;
                    }
                    if (ra.log_accept) ra.log_accept[ri] = uint8_t(keep);
                    if (ra.log_psnr) ra.log_psnr[ri] = psnr;
                    // the slot used two active steps from now: everyone has read it (they contributed to this step)
                    unsigned long long* wz = ring + size_t((ring_step + 2) % ROLLOUT_RING) * 2;
                    wz[0] = 0ull; wz[1] = 0ull;
                    if (keep) {
                        *stp = int8_t(sg > 0 ? 1 : 0);
                        if (a.recon_stale) {
                            uint8_t* sm = a.recon_stale + size_t(env) * RECON_MAX_BUFFERS;
#pragma unroll
                            for (int b = 0; b < RECON_MAX_BUFFERS; ++b) sm[b] |= uint8_t(1 << g);
                        }
                    }
                }
                if (keep) { S[0] += dII; S[1] += dIT; S[3] = psnr; }
            } else {
                atomicExch(ra.error, 1);
            }
            s_dec = keep;
        }
        __syncthreads();
        const int keep = s_dec;
        if (keep < 0) break;                               // abort: CTA-uniform
        if (keep) {
            float2* U = a.U + (size_t(env) * a.F + f) * n2;
            float* I = a.I + (size_t(env) * a.G + g) * n2;
            const float2* h = a.h + size_t(g) * P * a.HP;
            commit_run_rows<UF>(a, U, I, h, r, c, float(sg), w0, w1, tid, invFg_f, pf, pl);
            if (LOOK && g_next == g) have_spec = false;    // the next candidate reads what this flip changed: again
        }
        prev_act = act; prev_sgn = sg; prev_keep = keep;
        ++ring_step;
    }
    if (leader && tid == 0) {
        double* Sp = a.sums + size_t(env) * 4;
        Sp[0] = S[0]; Sp[1] = S[1]; Sp[3] = S[3];
    }
}

// ---------------------------------------------------------------------------
// observation path: obs["recon_image"] of every environment of a step (env.py:176-181,
// appendix B-2: a rejected step still shows the reconstruction WITH the rejected flip).
//
// The host (or device) observation block out[E][G][N][N] is kept up to date plane by plane: a
// step changes one colour plane per environment (the flipped group), so only that plane is
// rewritten -- plus planes that went stale earlier: a rejected candidate shown by this buffer in
// an earlier step (has to return to the committed I), or flips kept while the other buffer of a
// double-buffered pair was current.  stale[buffer][env] holds one bit per plane.
//
// k_recon_plan (one CTA) turns the step's results into a per-environment plan and advances the
// stale masks; k_recon_batch (grid: chunks x planes x tasks) writes the planes: I_g, or
// I_g + dI(candidate) for the plane of a rejected flip, straight to `out` -- mapped pinned host
// memory (the stores travel over PCIe, no staging copy and no host round trip to learn the sign)
// or device memory (zero-copy observations for a policy on the same GPU).
// ---------------------------------------------------------------------------
struct ReconPlan {            // one per task
    int env, mask, cand_g, f, r, c, sgn, pad_;
};

struct ReconArgs {
    const float2* U; const float* I; const float2* h; const int8_t* state;
    const Result* results;        // device; nullptr: committed reconstruction only
    const int32_t* envs;          // device env ids or nullptr (identity / inline)
    int n_inline; int inl_envs[32];
    uint8_t* stale;               // [E][RECON_MAX_BUFFERS]
    unsigned long long* planes_written;   // running count of planes written (bytes accounting of the bench)
    ReconPlan* plan;              // [n_tasks]
    float* out;                   // [E][G][N][N]
    int n_tasks, E, N, P, HP, F, G, Fg, buffer, full;
};

__global__ void __launch_bounds__(256)
k_recon_plan(const ReconArgs a) {
    for (int k = threadIdx.x; k < a.n_tasks; k += blockDim.x) {
        const int env = a.n_inline ? a.inl_envs[k] : (a.envs ? a.envs[k] : k);
        ReconPlan pl; pl.env = env; pl.cand_g = -1; pl.f = pl.r = pl.c = pl.sgn = 0; pl.pad_ = 0;
        const int all = (1 << a.G) - 1;
        uint8_t* sm = a.stale + size_t(env) * RECON_MAX_BUFFERS + a.buffer;
        int mask = a.full ? all : (int(*sm) & all);       // kept flips were marked by k_commit
        int after = 0;
        if (a.results) {
            const Result res = a.results[k];
            if (res.action >= 0 && !res.accept) {
                const int n2 = a.N * a.N;
                const int f = int(res.action / n2);
                const int pix = int(res.action - (long long)f * n2);
                const int g = f / a.Fg;
                mask |= 1 << g;
                pl.cand_g = g; pl.f = f; pl.r = pix / a.N; pl.c = pix - pl.r * a.N; pl.sgn = res.sgn;
                after = 1 << g;              // this buffer shows a rejected flip: refresh next time
            }
        }
        pl.mask = mask;
        *sm = uint8_t(after);
        if (a.planes_written) atomicAdd(a.planes_written, (unsigned long long)__popc(mask));
        a.plan[k] = pl;
    }
}

// grid (chunks, G, n_tasks); 256 threads, 4 px per thread and iteration
__global__ void __launch_bounds__(256)
k_recon_batch(const ReconArgs a) {
    const ReconPlan pl = a.plan[blockIdx.z];
    const int g = blockIdx.y;
    if (!((pl.mask >> g) & 1)) return;
    const int N = a.N, P = a.P, HP = a.HP;
    const size_t n2 = size_t(N) * N;
    const float* I = a.I + (size_t(pl.env) * a.G + g) * n2;
    float* out = a.out + (size_t(pl.env) * a.G + g) * n2;
    const bool cand = (pl.cand_g == g);
    const float2* U = a.U + (size_t(pl.env) * a.F + pl.f) * n2;
    const float2* h = a.h + size_t(g) * P * HP;
    const float invFg = 1.f / float(a.Fg), s2 = 2.f * float(pl.sgn) * invFg;
    for (size_t p = (size_t(blockIdx.x) * 256 + threadIdx.x) * 4; p < n2; p += size_t(gridDim.x) * 1024) {
        float4 iv = *reinterpret_cast<const float4*>(I + p);
        if (cand) {
            const int y = int(p / N), x = int(p - size_t(y) * N);
            int hy = y - pl.r; if (hy < 0) hy += P;
            int hx = x - pl.c; if (hx < 0) hx += P;
            const float2* hq = h + size_t(hy) * HP + hx;
            const float4 ua = *reinterpret_cast<const float4*>(U + p);
            const float4 ub = *reinterpret_cast<const float4*>(U + p + 2);
            const float2 h0 = __ldg(hq), h1 = __ldg(hq + 1), h2 = __ldg(hq + 2), h3 = __ldg(hq + 3);
            iv.x += delta_px(ua.x, ua.y, h0.x, h0.y, s2, invFg);
            iv.y += delta_px(ua.z, ua.w, h1.x, h1.y, s2, invFg);
            iv.z += delta_px(ub.x, ub.y, h2.x, h2.y, s2, invFg);
            iv.w += delta_px(ub.z, ub.w, h3.x, h3.y, s2, invFg);
        }
        *reinterpret_cast<float4*>(out + p) = iv;
    }
}

// k_recon_candidate: out_g += dI of one (uncommitted) candidate flip (single-plane helper of
// bh_get_recon; the sign is read from the resident state on the device).
__global__ void __launch_bounds__(256)
k_recon_candidate(const float2* __restrict__ U, const float2* __restrict__ h, const int8_t* __restrict__ sbyte,
                  float* __restrict__ out_g, int N, int P, int r, int c, int Fg) {
    const float sgn = 1.f - 2.f * float(*sbyte);
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

}  // namespace bh
