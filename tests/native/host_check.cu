// CPU execution of the __host__ __device__ FFT pass functions of bh_fft.cuh:
// the same index math, twiddle lookups and register butterflies the CUDA
// kernels run, emulating the CTA as a loop over thread slots with the barrier
// between pass_read and pass_write.  Compared against a double-precision DFT.
// Also checks the host-side table builder (bh_tables.hpp) against its own DFT.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../binary_hologram_reinforcement_learning_b200/csrc/bh_fft.cuh"
#include "../../binary_hologram_reinforcement_learning_b200/csrc/bh_tables.hpp"

using namespace bh;

template <int N, int Q, int ES, int SK, bool INV, int R, int Ns>
static void host_pass(std::vector<float2>& s, int base, const float2* twp) {
    constexpr int NB = PassShape<N, R, Ns, Q>::NB;
    std::vector<float2> regs(size_t(Q) * NB * R);
    for (int q = 0; q < Q; ++q) {
        float2 v[NB][R];
        for (int b = 0; b < NB; ++b) for (int r = 0; r < R; ++r) v[b][r] = make_float2(0, 0);
        pass_read<N, R, Ns, Q, ES, SK, INV>(s.data() + base, q, twp, v);
        for (int b = 0; b < NB; ++b) for (int r = 0; r < R; ++r) regs[(size_t(q) * NB + b) * R + r] = v[b][r];
    }
    for (int q = 0; q < Q; ++q) {
        float2 v[NB][R];
        for (int b = 0; b < NB; ++b) for (int r = 0; r < R; ++r) v[b][r] = regs[(size_t(q) * NB + b) * R + r];
        pass_write<N, R, Ns, Q, ES, SK>(s.data() + base, q, v);
    }
}

template <int N, int Q, int ES, int SK, bool INV>
static void host_tile_fft(std::vector<float2>& s, int base, const float2* tw) {
    using P = Plan<N>;
    using L = TwLayout<N>;
    constexpr int R0 = P::r[0], R1 = P::r[1], R2 = P::r[2];
    host_pass<N, Q, ES, SK, INV, R0, 1>(s, base, tw);
    host_pass<N, Q, ES, SK, INV, R1, R0>(s, base, tw + L::off1);
    if constexpr (P::n == 3) host_pass<N, Q, ES, SK, INV, R2, R0 * R1>(s, base, tw + L::off2);
}

constexpr int ilog2_h(int v) { return v <= 1 ? 0 : 1 + ilog2_h(v >> 1); }

// COL = false: row layout (ES = 1, base = w * SEQ, skew 4); COL = true: column layout
// (ES = 8, base = w, skew log2(R0)) -- the two layouts of the CUDA passes
template <int N, int Q, bool COL>
static double check_layout() {
    constexpr int W = 8;
    constexpr int SK = COL ? ilog2_h(Plan<N>::r[0]) : 4;
    constexpr int ES = COL ? W : 1;
    constexpr int SEQ = SeqLen<N, SK>::value;
    std::vector<float> twf = build_twiddles(N);
    if (int(twf.size()) != 2 * TwLayout<N>::total_all) { printf("twiddle table size mismatch N=%d\n", N); return 1.0; }
    std::vector<int> rad = plan_radices(N);
    if (int(rad.size()) != Plan<N>::n || rad[0] != Plan<N>::r[0] || rad[1] != Plan<N>::r[1] ||
        (Plan<N>::n == 3 && rad[2] != Plan<N>::r[2])) { printf("plan mismatch N=%d\n", N); return 1.0; }
    twf.push_back(0.f); twf.push_back(0.f);
    const float2* tw = reinterpret_cast<const float2*>(twf.data());
    auto at = [&](int e, int w) { return COL ? padded<SK>(e) * W + w : w * SEQ + padded<SK>(e); };
    std::vector<float2> s(size_t(SEQ) * W + 8, make_float2(0, 0));
    srand(N);
    for (int i = 0; i < N; ++i)
        for (int w = 0; w < W; ++w)
            s[at(i, w)] = make_float2(float(rand() / double(RAND_MAX) - 0.5), float(rand() / double(RAND_MAX) - 0.5));
    double worst = 0;
    for (int inv = 0; inv < 2; ++inv) {
        std::vector<float2> t = s;
        for (int w = 0; w < W; ++w) {
            const int base = COL ? w : w * SEQ;
            if (inv) host_tile_fft<N, Q, ES, SK, true>(t, base, tw);
            else host_tile_fft<N, Q, ES, SK, false>(t, base, tw);
        }
        const double sg = inv ? 1.0 : -1.0;
        for (int w = 0; w < W; w += 3) {
            double scale = 0, err = 0;
            for (int k = 0; k < N; ++k) {
                std::complex<double> acc = 0;
                for (int i = 0; i < N; ++i) {
                    const double ph = sg * 6.283185307179586 * double((long long)i * k % N) / N;
                    acc += std::complex<double>(s[at(i, w)].x, s[at(i, w)].y) *
                           std::complex<double>(std::cos(ph), std::sin(ph));
                }
                const std::complex<double> got(t[at(k, w)].x, t[at(k, w)].y);
                err = std::max(err, std::abs(got - acc));
                scale = std::max(scale, std::abs(acc));
            }
            worst = std::max(worst, err / scale);
        }
    }
    printf("fft N=%d Q=%d layout=%s rel_err=%.3e\n", N, Q, COL ? "col" : "row", worst);
    return worst;
}

template <int N, int Q>
static double check_size() { return std::max(check_layout<N, Q, false>(), check_layout<N, Q, true>()); }

static double check_tables(int P) {
    // h must be the inverse DFT of H*P^2 (H is stored pre-scaled by 1/P^2)
    auto t = build_tables(P, 515e-9, 7.56e-6, 2e-3, 0);
    double worst = 0;
    const int ys[3] = {0, 1, P - 3}, xs[3] = {0, 2, P - 1};
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            std::complex<double> acc = 0;
            for (int ky = 0; ky < P; ++ky)
                for (int kx = 0; kx < P; ++kx) {
                    const double ph = 6.283185307179586 * (double((long long)ky * ys[a] % P) + double((long long)kx * xs[b] % P)) / P;
                    acc += std::complex<double>(t->H[2 * (size_t(ky) * P + kx)], t->H[2 * (size_t(ky) * P + kx) + 1]) *
                           std::complex<double>(std::cos(ph), std::sin(ph));
                }
            const size_t i = size_t(ys[a]) * P + xs[b];
            worst = std::max(worst, std::abs(acc - std::complex<double>(t->h[2 * i], t->h[2 * i + 1])));
        }
    // |H| * P^2 == 1 (pure phase at the reference's geometry)
    double dev = 0;
    for (size_t i = 0; i < size_t(P) * P; ++i)
        dev = std::max(dev, std::fabs(std::hypot(t->H[2 * i], t->H[2 * i + 1]) * P * P - 1.0));
    printf("tables P=%d h_abs_err=%.3e |H|dev=%.3e\n", P, worst, dev);
    return std::max(worst, dev * 1e-3);
}

int main() {
    double worst = 0;
    worst = std::max(worst, check_size<32, 32>());
    worst = std::max(worst, check_size<64, 32>());
    worst = std::max(worst, check_size<128, 32>());
    worst = std::max(worst, check_size<256, 32>());
    worst = std::max(worst, check_size<512, 32>());
    worst = std::max(worst, check_size<896, 32>());
    worst = std::max(worst, check_size<1024, 32>());
    worst = std::max(worst, check_size<1792, 64>());
    worst = std::max(worst, check_size<2048, 64>());
    double tw = std::max(check_tables(64), check_tables(112));
    const bool ok = worst < 5e-6 && tw < 1e-7;
    printf("%s worst_fft=%.3e worst_tables=%.3e\n", ok ? "HOST_CHECK_OK" : "HOST_CHECK_FAIL", worst, tw);
    return ok ? 0 : 1;
}
