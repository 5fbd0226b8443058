// Shared-memory tile FFT primitives for the hologram propagation kernels (sm_100a).
//
// A "tile" is W = 8 independent length-N complex sequences in shared memory, in
// one of two skewed layouts (see "shared-memory layouts" below): sequence-major
// for the row passes (global rows copy in and out coalesced), interleaved for the
// column pass (a 64-byte row segment of 8 columns copies in and out directly).
//
// One Stockham autosort pass of radix R (Ns = product of the radices already
// applied) does, for butterfly j in [0, N/R):
//     k  = j mod Ns
//     v[r] = in[j + r*N/R] * exp(-+2 pi i r k / (Ns R))        r = 0..R-1
//     V = DFT_R(v)
//     out[(j / Ns) * Ns * R + k + r * Ns] = V[r]
// All butterflies of a thread are held in registers between the read and the
// write, so the pass is in place with one barrier either side.
//
// Everything here is __host__ __device__ so tests/host_check.cu can run the
// exact index math and butterflies on the CPU.
#pragma once
#include <cuda_runtime.h>

#ifndef BH_HD
#define BH_HD __host__ __device__ __forceinline__
#endif

namespace bh {

// Complex arithmetic on float2.  On sm_100 the device path uses the packed FP32x2 instructions (FADD2 / FMUL2 /
// FFMA2: one issue slot for the real and the imaginary part; the swap and per-half negation of a complex
// product are operand modifiers of FFMA2), which cuts the floating-point instructions of a transform by ~55 %.
// The host path (tests/native/host_check.cu) and BH_SCALAR_FFT builds keep the scalar forms; the two differ only
// in the rounding order of a complex product.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(BH_SCALAR_FFT)
#define BH_PACKED_FFT 1
#else
#define BH_PACKED_FFT 0
#endif

BH_HD float2 cmul(float2 a, float2 b) {
#if BH_PACKED_FFT
    const float2 t = __fmul2_rn(a, make_float2(b.x, b.x));
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), t);
#else
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
#endif
}
BH_HD float2 cadd(float2 a, float2 b) {
#if BH_PACKED_FFT
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
BH_HD float2 csub(float2 a, float2 b) {
#if BH_PACKED_FFT
    return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// multiply by -i (forward) or +i (inverse)
template <bool INV> BH_HD float2 mul_mi(float2 a) {
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// a + (-+i) b and a - (-+i) b: the rotation by a quarter turn folds into the operand modifiers
template <bool INV> BH_HD float2 cadd_mi(float2 a, float2 b) {
#if BH_PACKED_FFT
    return __ffma2_rn(make_float2(b.y, b.x), INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f), a);
#else
    return cadd(a, mul_mi<INV>(b));
#endif
}
template <bool INV> BH_HD float2 csub_mi(float2 a, float2 b) {
#if BH_PACKED_FFT
    return __ffma2_rn(make_float2(b.y, b.x), INV ? make_float2(1.f, -1.f) : make_float2(-1.f, 1.f), a);
#else
    return csub(a, mul_mi<INV>(b));
#endif
}
template <bool INV> BH_HD float2 cmulw(float2 a, float wr, float wi_fwd) {
    // multiply by (wr + i*wi) where wi = wi_fwd (forward) or -wi_fwd (inverse)
    const float wi = INV ? -wi_fwd : wi_fwd;
#if BH_PACKED_FFT
    const float2 t = __fmul2_rn(a, make_float2(wr, wr));
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-wi, wi), t);
#else
    return make_float2(fmaf(a.x, wr, -a.y * wi), fmaf(a.x, wi, a.y * wr));
#endif
}

// ---------------------------------------------------------------------------
// small DFTs in registers; forward kernel exp(-2 pi i r q / R), INV conjugates.
// Elements are addressed as v[off + idx*stride] so composite sizes can recurse.
// ---------------------------------------------------------------------------
template <bool INV> BH_HD void dft2(float2& a, float2& b) {
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <bool INV> BH_HD void dft4(float2& a, float2& b, float2& c, float2& d) {
    const float2 s0 = cadd(a, c), d0 = csub(a, c);
    const float2 s1 = cadd(b, d), t1 = csub(b, d);
    a = cadd(s0, s1);
    b = cadd_mi<INV>(d0, t1);
    c = csub(s0, s1);
    d = csub_mi<INV>(d0, t1);
}

// natural-order in, natural-order out
template <bool INV> BH_HD void dft8(float2* v) {
    // R1 = 2 (a), R2 = 4 (b): r = a*4 + b ; X[c + 2 d]
    const float h = 0.70710678118654752440f;
#pragma unroll
    for (int b = 0; b < 4; ++b) dft2<INV>(v[b], v[4 + b]);
    // twiddle Y[c=1][b] *= W8^b
    v[5] = cmulw<INV>(v[5], h, -h);
    v[6] = mul_mi<INV>(v[6]);
    v[7] = cmulw<INV>(v[7], -h, -h);
    dft4<INV>(v[0], v[1], v[2], v[3]);   // c = 0 -> X[0], X[2], X[4], X[6]
    dft4<INV>(v[4], v[5], v[6], v[7]);   // c = 1 -> X[1], X[3], X[5], X[7]
    const float2 x0 = v[0], x2 = v[1], x4 = v[2], x6 = v[3];
    const float2 x1 = v[4], x3 = v[5], x5 = v[6], x7 = v[7];
    v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3;
    v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
}

template <bool INV> BH_HD void dft16(float2* v) {
    // R1 = 4 (a), R2 = 4 (b): r = a*4 + b ; X[c + 4 d]
    const float c1 = 0.92387953251128675613f;  // cos(pi/8)
    const float s1 = 0.38268343236508977173f;  // sin(pi/8)
    const float h = 0.70710678118654752440f;
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // Y[c][b] = v[c*4 + b]; twiddle by W16^(c*b) = exp(-2 pi i c b / 16)
    v[5] = cmulw<INV>(v[5], c1, -s1);     // 1
    v[6] = cmulw<INV>(v[6], h, -h);       // 2
    v[7] = cmulw<INV>(v[7], s1, -c1);     // 3
    v[9] = cmulw<INV>(v[9], h, -h);       // 2
    v[10] = mul_mi<INV>(v[10]);           // 4
    v[11] = cmulw<INV>(v[11], -h, -h);    // 6
    v[13] = cmulw<INV>(v[13], s1, -c1);   // 3
    v[14] = cmulw<INV>(v[14], -h, -h);    // 6
    v[15] = cmulw<INV>(v[15], -c1, s1);   // 9
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    // v[c*4 + d] holds X[c + 4 d]  -> transpose the 4x4
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) {
            const float2 t = v[c * 4 + d];
            v[c * 4 + d] = v[d * 4 + c];
            v[d * 4 + c] = t;
        }
}

template <bool INV> BH_HD void dft7(float2* v) {
    // direct 7-point DFT, exploiting the conjugate symmetry of the roots
    const float c1 = 0.62348980185873353053f, s1 = 0.78183148246802980871f;
    const float c2 = -0.22252093395631440429f, s2 = 0.97492791218182360702f;
    const float c3 = -0.90096886790241912624f, s3 = 0.43388373911755812048f;
    const float2 p1 = cadd(v[1], v[6]), m1 = csub(v[1], v[6]);
    const float2 p2 = cadd(v[2], v[5]), m2 = csub(v[2], v[5]);
    const float2 p3 = cadd(v[3], v[4]), m3 = csub(v[3], v[4]);
    const float2 x0 = v[0];
    // real-coefficient parts
    const float2 a1 = make_float2(x0.x + c1 * p1.x + c2 * p2.x + c3 * p3.x,
                                  x0.y + c1 * p1.y + c2 * p2.y + c3 * p3.y);
    const float2 a2 = make_float2(x0.x + c2 * p1.x + c3 * p2.x + c1 * p3.x,
                                  x0.y + c2 * p1.y + c3 * p2.y + c1 * p3.y);
    const float2 a3 = make_float2(x0.x + c3 * p1.x + c1 * p2.x + c2 * p3.x,
                                  x0.y + c3 * p1.y + c1 * p2.y + c2 * p3.y);
    // sine parts: b_q = sum_r sin(2 pi r q / 7) m_r
    const float2 b1 = make_float2(s1 * m1.x + s2 * m2.x + s3 * m3.x,
                                  s1 * m1.y + s2 * m2.y + s3 * m3.y);
    const float2 b2 = make_float2(s2 * m1.x - s3 * m2.x - s1 * m3.x,
                                  s2 * m1.y - s3 * m2.y - s1 * m3.y);
    const float2 b3 = make_float2(s3 * m1.x - s1 * m2.x + s2 * m3.x,
                                  s3 * m1.y - s1 * m2.y + s2 * m3.y);
    // forward: X[q] = a_q - i b_q ; X[7-q] = a_q + i b_q   (inverse swaps)
    v[0] = make_float2(x0.x + p1.x + p2.x + p3.x, x0.y + p1.y + p2.y + p3.y);
    const float2 ib1 = mul_mi<INV>(b1), ib2 = mul_mi<INV>(b2), ib3 = mul_mi<INV>(b3);
    v[1] = cadd(a1, ib1); v[6] = csub(a1, ib1);
    v[2] = cadd(a2, ib2); v[5] = csub(a2, ib2);
    v[3] = cadd(a3, ib3); v[4] = csub(a3, ib3);
}

template <bool INV> BH_HD void dft32(float2* v) {
    // R1 = 2 (a), R2 = 16 (b): r = a*16 + b ; X[c + 2 d]
    const float cs[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                          0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                          0.19509032201612826785f, 0.f, -0.19509032201612826785f, -0.38268343236508977173f,
                          -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                          -0.92387953251128675613f, -0.98078528040323044913f};
    const float sn[16] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                          0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                          0.98078528040323044913f, 1.f, 0.98078528040323044913f, 0.92387953251128675613f,
                          0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                          0.38268343236508977173f, 0.19509032201612826785f};
#pragma unroll
    for (int b = 0; b < 16; ++b) dft2<INV>(v[b], v[16 + b]);
#pragma unroll
    for (int b = 1; b < 16; ++b) v[16 + b] = cmulw<INV>(v[16 + b], cs[b], -sn[b]);   // W32^b
    dft16<INV>(v);          // c = 0 -> X[0], X[2], ..., X[30] in v[0..15]
    dft16<INV>(v + 16);     // c = 1 -> X[1], X[3], ..., X[31] in v[16..31]
    float2 t[32];
#pragma unroll
    for (int d = 0; d < 16; ++d) { t[2 * d] = v[d]; t[2 * d + 1] = v[16 + d]; }
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = t[k];
}

template <bool INV> BH_HD void dft28(float2* v) {
    // R1 = 4 (a), R2 = 7 (b): r = a*7 + b ; X[c + 4 d]
    const float cs[19] = {1.00000000000000000000f, 0.97492791218182361934f, 0.90096886790241914600f, 0.78183148246802980363f, 0.62348980185873359439f, 0.43388373911755817591f, 0.22252093395631444839f, 0.00000000000000006123f, -0.22252093395631433737f, -0.43388373911755806489f, -0.62348980185873348336f, -0.78183148246802947057f, -0.90096886790241903498f, -0.97492791218182373036f, -1.00000000000000000000f, -0.97492791218182373036f, -0.90096886790241914600f, -0.78183148246802958159f, -0.62348980185873370541f};
    const float sn[19] = {0.00000000000000000000f, 0.22252093395631439288f, 0.43388373911755812040f, 0.62348980185873348336f, 0.78183148246802980363f, 0.90096886790241914600f, 0.97492791218182361934f, 1.00000000000000000000f, 0.97492791218182361934f, 0.90096886790241914600f, 0.78183148246802991466f, 0.62348980185873392745f, 0.43388373911755823142f, 0.22252093395631408757f, 0.00000000000000012246f, -0.22252093395631383776f, -0.43388373911755800938f, -0.62348980185873381643f, -0.78183148246802969261f};
#pragma unroll
    for (int b = 0; b < 7; ++b) dft4<INV>(v[b], v[7 + b], v[14 + b], v[21 + b]);
#pragma unroll
    for (int c = 1; c < 4; ++c)
#pragma unroll
        for (int b = 1; b < 7; ++b) v[c * 7 + b] = cmulw<INV>(v[c * 7 + b], cs[c * b], -sn[c * b]);   // W28^(c b)
#pragma unroll
    for (int c = 0; c < 4; ++c) dft7<INV>(v + 7 * c);     // v[c*7 + d] = X[c + 4 d]
    float2 t[28];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = 0; d < 7; ++d) t[c + 4 * d] = v[c * 7 + d];
#pragma unroll
    for (int k = 0; k < 28; ++k) v[k] = t[k];
}

template <int R, bool INV> BH_HD void dft(float2* v) {
    if (R == 32) dft32<INV>(v);
    else if (R == 28) dft28<INV>(v);
    else if (R == 2) dft2<INV>(v[0], v[1]);
    else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) dft8<INV>(v);
    else if (R == 16) dft16<INV>(v);
    else if (R == 7) dft7<INV>(v);
}

// ---------------------------------------------------------------------------
// shared-memory layouts.  Element e of a sequence lives at
//     s[ (e + (e >> SK)) * ES + base ]
// i.e. one padding slot after every 2^SK elements (a skew that breaks the
// power-of-two strides of the Stockham scatter):
//   row layout     ES = 1, base = w * SEQ      threads along the sequence
//                  (16 consecutive butterflies per half-warp -> 16 distinct banks)
//   column layout  ES = W (8 or 16), base = w   threads across the W sequences first
//                  (W = 8: 8 sequences x 2 butterflies per half-warp, the skew makes the two
//                   butterflies land 8 float2 apart modulo 16; W = 16: one butterfly per
//                   half-warp, 16 consecutive float2)
// ---------------------------------------------------------------------------
template <int SK> BH_HD int padded(int e) { return e + (e >> SK); }
template <int N, int SK> struct SeqLen { static constexpr int value = N + (N >> SK) + 1; };

// ---------------------------------------------------------------------------
// radix plans.  The product of the radices is N; larger radices first keeps the
// twiddle-free first pass the most expensive one.
// ---------------------------------------------------------------------------
template <int N> struct Plan;
template <> struct Plan<32>   { static constexpr int n = 2; static constexpr int r[3] = {8, 4, 1}; };
template <> struct Plan<64>   { static constexpr int n = 2; static constexpr int r[3] = {8, 8, 1}; };
template <> struct Plan<128>  { static constexpr int n = 2; static constexpr int r[3] = {16, 8, 1}; };
template <> struct Plan<256>  { static constexpr int n = 2; static constexpr int r[3] = {16, 16, 1}; };
template <> struct Plan<512>  { static constexpr int n = 3; static constexpr int r[3] = {8, 8, 8}; };
template <> struct Plan<896>  { static constexpr int n = 2; static constexpr int r[3] = {32, 28, 1}; };   // two passes
template <> struct Plan<1024> { static constexpr int n = 2; static constexpr int r[3] = {32, 32, 1}; };   // two passes
template <> struct Plan<1792> { static constexpr int n = 3; static constexpr int r[3] = {16, 16, 7}; };
template <> struct Plan<2048> { static constexpr int n = 3; static constexpr int r[3] = {16, 16, 8}; };

// Twiddle table of a plan: for pass p >= 1 with radix R and Ns = product of the
// earlier radices, entries tw[off_p + (r - 1) * Ns + k] = exp(-2 pi i r k / (Ns R)),
// r = 1..R-1, k = 0..Ns-1.  Consecutive k are contiguous, so the threads of a warp
// (consecutive butterflies) read one or two 128-byte lines per twiddle.
template <int N> struct TwLayout {
    using P = Plan<N>;
    static constexpr int off1 = 0;
    static constexpr int len1 = (P::r[1] - 1) * P::r[0];
    static constexpr int off2 = len1;
    static constexpr int len2 = (P::n == 3) ? (P::r[2] - 1) * P::r[0] * P::r[1] : 0;
    static constexpr int total = len1 + len2;
    // two-pass plans with unequal radices append the block of the reversed plan (bh_fft2.cuh)
    static constexpr int len_rev = (P::n == 2 && P::r[0] != P::r[1]) ? (P::r[0] - 1) * P::r[1] : 0;
    static constexpr int total_all = total + len_rev;
};

// ---------------------------------------------------------------------------
// one Stockham pass split in two phases around a barrier
// ---------------------------------------------------------------------------
template <int N, int R, int Ns, int Q>
struct PassShape {
    static constexpr int NBF = N / R;                 // butterflies per sequence
    static constexpr int NB = (NBF + Q - 1) / Q;      // butterflies per thread
};

// phase 1: gather + twiddle + butterfly into registers.
// s points at the sequence base; q is the thread's slot in [0,Q).
// twp: this pass's twiddle block (see TwLayout), forward sign.
template <int N, int R, int Ns, int Q, int ES, int SK, bool INV>
BH_HD void pass_read(const float2* s, int q, const float2* __restrict__ twp,
                     float2 (&v)[PassShape<N, R, Ns, Q>::NB][R]) {
    constexpr int NBF = N / R;
    constexpr int NB = PassShape<N, R, Ns, Q>::NB;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = q + b * Q;
        if ((NBF % Q == 0) || j < NBF) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[b][r] = s[padded<SK>(j + r * NBF) * ES];
            if (Ns > 1) {
                const int k = j % Ns;
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    float2 w = twp[(r - 1) * Ns + k];
                    if (INV) w.y = -w.y;
                    v[b][r] = cmul(v[b][r], w);
                }
            }
            dft<R, INV>(v[b]);
        }
    }
}

// phase 2: scatter to the autosorted positions.
template <int N, int R, int Ns, int Q, int ES, int SK>
BH_HD void pass_write(float2* s, int q, const float2 (&v)[PassShape<N, R, Ns, Q>::NB][R]) {
    constexpr int NBF = N / R;
    constexpr int NB = PassShape<N, R, Ns, Q>::NB;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int j = q + b * Q;
        if ((NBF % Q == 0) || j < NBF) {
            const int k = j % Ns;
            const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) s[padded<SK>(j0 + r * Ns) * ES] = v[b][r];
        }
    }
}

#ifdef __CUDACC__
// Full in-place tile FFT executed cooperatively by the CTA (device only).
template <int N, int Q, int ES, int SK, bool INV, int R, int Ns>
__device__ __forceinline__ void tile_pass(float2* s, int q, const float2* __restrict__ twp) {
    float2 v[PassShape<N, R, Ns, Q>::NB][R];
    pass_read<N, R, Ns, Q, ES, SK, INV>(s, q, twp, v);
    __syncthreads();
    pass_write<N, R, Ns, Q, ES, SK>(s, q, v);
    __syncthreads();
}

template <int N, int Q, int ES, int SK, bool INV>
__device__ __forceinline__ void tile_fft(float2* s, int q, const float2* __restrict__ tw) {
    using P = Plan<N>;
    using L = TwLayout<N>;
    constexpr int R0 = P::r[0], R1 = P::r[1], R2 = P::r[2];
    tile_pass<N, Q, ES, SK, INV, R0, 1>(s, q, tw);
    tile_pass<N, Q, ES, SK, INV, R1, R0>(s, q, tw + L::off1);
    if constexpr (P::n == 3) tile_pass<N, Q, ES, SK, INV, R2, R0 * R1>(s, q, tw + L::off2);
}
#endif

}  // namespace bh
