// Host-side (double precision) construction of the per-wavelength tables:
//   H  angular-spectrum transfer function on the P x P FFT grid
//   h  = ifft2(H), the field of a unit pixel (the impulse the delta kernel shifts)
//   tw forward twiddles exp(-2 pi i m / P)
// Restates what tt.simulate multiplies by (reference call sites env.py:127,172;
// SURVEY.md 8c).  The phase 2 pi z sqrt(1/wl^2 - f^2) is ~2.4e4 rad, so it is
// formed and reduced in double and only then rounded to float.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

namespace bh {

using cd = std::complex<double>;

template <typename Fn>
inline void parallel_for(int n, Fn fn) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = int(std::min<unsigned>(hw ? hw : 1u, 32u));
    nt = std::min(nt, n);
    if (nt <= 1) { for (int i = 0; i < n; ++i) fn(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=]() { for (int i = t; i < n; i += nt) fn(i); });
    for (auto& x : th) x.join();
}

// recursive decimation-in-time FFT for any n (O(n * sum of prime factors)).
// roots[i] = exp(sign * 2 pi i * i / N0); rs = N0 / n.
inline void fft_rec(const cd* in, cd* out, int n, int stride, const cd* roots, int N0, int rs) {
    if (n == 1) { out[0] = in[0]; return; }
    int p = 2;
    while (n % p) ++p;
    const int m = n / p;
    for (int q = 0; q < p; ++q) fft_rec(in + size_t(q) * stride, out + size_t(q) * m, m, stride * p, roots, N0, rs * p);
    cd tbuf[32], xbuf[32];
    std::vector<cd> tvec, xvec;
    cd *t = tbuf, *x = xbuf;
    if (p > 32) { tvec.resize(p); xvec.resize(p); t = tvec.data(); x = xvec.data(); }
    for (int k = 0; k < m; ++k) {
        for (int q = 0; q < p; ++q)
            t[q] = out[size_t(q) * m + k] * roots[(int64_t(q) * k * rs) % N0];
        for (int j = 0; j < p; ++j) {
            cd acc = t[0];
            for (int q = 1; q < p; ++q) acc += t[q] * roots[(int64_t(q) * j * m * rs) % N0];
            x[j] = acc;
        }
        for (int j = 0; j < p; ++j) out[size_t(k) + size_t(m) * j] = x[j];
    }
}

// in-place unnormalised 2-D transform of a P x P array, sign = -1 forward, +1 inverse
inline void fft2_host(std::vector<cd>& a, int P, int sign) {
    std::vector<cd> roots(P);
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < P; ++i) roots[i] = cd(std::cos(two_pi * i / P), sign * std::sin(two_pi * i / P));
    parallel_for(P, [&](int y) {
        std::vector<cd> tmp(P);
        fft_rec(&a[size_t(y) * P], tmp.data(), P, 1, roots.data(), P, 1);
        std::copy(tmp.begin(), tmp.end(), a.begin() + size_t(y) * P);
    });
    parallel_for(P, [&](int x) {
        std::vector<cd> col(P), tmp(P);
        for (int y = 0; y < P; ++y) col[y] = a[size_t(y) * P + x];
        fft_rec(col.data(), tmp.data(), P, 1, roots.data(), P, 1);
        for (int y = 0; y < P; ++y) a[size_t(y) * P + x] = tmp[y];
    });
}

struct HostTables {
    int P = 0;
    std::vector<float> H;   // interleaved complex, pre-scaled by 1/P^2
    std::vector<float> h;   // interleaved complex
};

inline double fft_freq(int i, int P, double dx) {
    const int k = (i < (P + 1) / 2) ? i : i - P;       // numpy.fft.fftfreq ordering
    return double(k) / (double(P) * dx);
}

inline std::shared_ptr<HostTables> build_tables(int P, double wl, double dx, double z, int method) {
    auto t = std::make_shared<HostTables>();
    t->P = P;
    std::vector<cd> Hd(size_t(P) * P);
    const double two_pi = 6.283185307179586476925286766559;
    const double pi = 3.1415926535897932384626433832795;
    parallel_for(P, [&](int y) {
        const double fy = fft_freq(y, P, dx);
        for (int x = 0; x < P; ++x) {
            const double fx = fft_freq(x, P, dx);
            cd v(0.0, 0.0);
            if (method == 0) {
                const double rad = 1.0 / (wl * wl) - fx * fx - fy * fy;
                if (rad > 0.0) {
                    const double ph = two_pi * z * std::sqrt(rad);
                    v = cd(std::cos(ph), std::sin(ph));
                }
            } else {
                const double ph = two_pi * z / wl - pi * wl * z * (fx * fx + fy * fy);
                v = cd(std::cos(ph), std::sin(ph));
            }
            Hd[size_t(y) * P + x] = v;
        }
    });
    const double inv = 1.0 / (double(P) * double(P));
    t->H.resize(size_t(P) * P * 2);
    for (size_t i = 0; i < size_t(P) * P; ++i) {
        t->H[2 * i] = float(Hd[i].real() * inv);
        t->H[2 * i + 1] = float(Hd[i].imag() * inv);
    }
    fft2_host(Hd, P, +1);
    t->h.resize(size_t(P) * P * 2);
    for (size_t i = 0; i < size_t(P) * P; ++i) {
        t->h[2 * i] = float(Hd[i].real() * inv);
        t->h[2 * i + 1] = float(Hd[i].imag() * inv);
    }
    return t;
}

// process-wide cache: contexts of the same geometry share the host tables
inline std::shared_ptr<HostTables> get_tables(int P, double wl, double dx, double z, int method) {
    using Key = std::tuple<int, double, double, double, int>;
    static std::mutex mu;
    static std::map<Key, std::shared_ptr<HostTables>> cache;
    std::lock_guard<std::mutex> lk(mu);
    const Key key(P, wl, dx, z, method);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto t = build_tables(P, wl, dx, z, method);
    cache[key] = t;
    return t;
}

// Spectra of the even kernels of the correlation sweep, derived from the float impulse
// response the delta kernel uses: K3 = fft2(h|h|^2), K4 = fft2(|h|^2), K5 = fft2(h^2),
// each pre-scaled by 1/P^2 like H; m4 = sum |h|^4.
struct SweepTables {
    std::vector<float> K3, K4, K5, K6;   // interleaved complex [P][P]; K6 = fft2(|h|^4) (pad = 2 only)
    double m4 = 0.0;
};

inline std::shared_ptr<SweepTables> build_sweep_tables(const HostTables& t) {
    const int P = t.P;
    const size_t n = size_t(P) * P;
    auto out = std::make_shared<SweepTables>();
    std::vector<cd> k3(n), k4(n), k5(n), k6(n);
    double m4 = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const cd h(double(t.h[2 * i]), double(t.h[2 * i + 1]));
        const double a2 = std::norm(h);
        k3[i] = h * a2; k4[i] = cd(a2, 0.0); k5[i] = h * h; k6[i] = cd(a2 * a2, 0.0);
        m4 += a2 * a2;
    }
    out->m4 = m4;
    const double inv = 1.0 / (double(P) * double(P));
    auto run = [&](std::vector<cd>& k, std::vector<float>& K) {
        fft2_host(k, P, -1);
        K.resize(2 * n);
        for (size_t i = 0; i < n; ++i) { K[2 * i] = float(k[i].real() * inv); K[2 * i + 1] = float(k[i].imag() * inv); }
    };
    run(k3, out->K3); run(k4, out->K4); run(k5, out->K5); run(k6, out->K6);
    return out;
}

// radix plan of an FFT side; must mirror Plan<N> in bh_fft.cuh (tests/native/host_check.cu
// runs both against each other)
inline std::vector<int> plan_radices(int P) {
    switch (P) {
        case 32: return {8, 4};
        case 64: return {8, 8};
        case 128: return {16, 8};
        case 256: return {16, 16};
        case 512: return {8, 8, 8};
        case 896: return {32, 28};
        case 1024: return {32, 32};
        case 1792: return {16, 16, 7};
        case 2048: return {16, 16, 8};
    }
    return {};
}

// Per-pass twiddle blocks (TwLayout in bh_fft.cuh): for pass p >= 1 with radix R and
// Ns = product of the earlier radices, tw[off_p + (r-1)*Ns + k] = exp(-2 pi i r k / (Ns R)).
inline std::vector<float> build_twiddles(int P) {
    const std::vector<int> rad = plan_radices(P);
    std::vector<float> tw;
    const double two_pi = 6.283185307179586476925286766559;
    int Ns = rad.empty() ? 1 : rad[0];
    for (size_t p = 1; p < rad.size(); ++p) {
        const int R = rad[p];
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < Ns; ++k) {
                const double ph = two_pi * double(r) * double(k) / (double(Ns) * double(R));
                tw.push_back(float(std::cos(ph)));
                tw.push_back(float(-std::sin(ph)));
            }
        Ns *= R;
    }
    // two-pass plans with unequal radices (896 = 32 x 28): block of the REVERSED plan (R1, R0) used by the
    // register-resident column pass (bh_fft2.cuh, Plan2<P, true>): tw[(r-1)*R1 + k] = exp(-2 pi i r k / P)
    if (rad.size() == 2 && rad[0] != rad[1]) {
        const int RA = rad[1], RB = rad[0];
        for (int r = 1; r < RB; ++r)
            for (int k = 0; k < RA; ++k) {
                const double ph = two_pi * double(r) * double(k) / double(P);
                tw.push_back(float(std::cos(ph)));
                tw.push_back(float(-std::sin(ph)));
            }
    }
    return tw;
}

}  // namespace bh
