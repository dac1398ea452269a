// Host-side planning for the mixed-radix / Bluestein FFT engine (fft_engine.cuh).
// Tables are computed in long double and rounded once to binary64.
#pragma once
#include <complex>
#include <vector>
#include <cmath>
#include "fft_engine.cuh"

namespace cbplan {

typedef std::complex<long double> lcplx;

static inline lcplx cis_ratio(long long num, long long den) {
    // exp(-2 pi i num / den), argument reduced in integers
    num %= den;
    if (num < 0) num += den;
    const long double two_pi = 6.283185307179586476925286766559005768L;
    long double a = two_pi * (long double)num / (long double)den;
    return lcplx(cosl(a), -sinl(a));
}

// Radix list for a 2/3/5-smooth N (empty if N has another prime factor).
static inline std::vector<int> smooth_radices(int N) {
    int c2 = 0, c3 = 0, c5 = 0, m = N;
    while (m % 2 == 0) { m /= 2; ++c2; }
    while (m % 3 == 0) { m /= 3; ++c3; }
    while (m % 5 == 0) { m /= 5; ++c5; }
    std::vector<int> r;
    if (m != 1 || N < 2) return r;
    while (c2 >= 3) { r.push_back(8); c2 -= 3; }
    if (c2 == 2) r.push_back(4);
    if (c2 == 1) r.push_back(2);
    for (int i = 0; i < c5; ++i) r.push_back(5);
    for (int i = 0; i < c3; ++i) r.push_back(3);
    // 640 = 8 * 10 * 8: the factors 2 and 5 become one radix-10 pass (Dft<10>, Good-Thomas inside), three passes
    // instead of four -- the Bluestein length of the 301-node windows of BASELINE config 5.  The radix-10 pass sits
    // in the MIDDLE: the pruned first / last passes and the fused filter pass then all run the radix-8 butterfly with
    // its linear padded addressing (8.8.10, radix 10 in the fused pass, measured 55.1 ms against 45.2 ms for the
    // 36100 windows of a 4096^2 lattice at 20 km stride, gpurun r02v).  Only this length, so that every other size
    // keeps the plan its static engine was built for.
    if (N == 640) { r.clear(); r.push_back(8); r.push_back(10); r.push_back(8); }
    return r;
}

// Cost model used to pick the padded Bluestein length: every pass streams the line once;
// forward + backward share one fused pass.
static inline double plan_cost(int N, const std::vector<int> &radices) {
    return (double)N * (2.0 * (double)radices.size() - 1.0);
}

static inline bool static_engine_length(int M) { return M == 256 || M == 320 || M == 512 || M == 640; }

static inline int bluestein_length(int n) {
    // The chirp filter is even (b[m] = b[-m]), so the circular length only has to reach 2n - 2: the two taps that
    // alias there, m = +-(n-1), are equal.  Taken when it lands on a compile-time engine (n = 129, 161, 257, 321:
    // PlusOneFft in fft_engine.cuh carries the last node around the pruned passes); the next usable length would be
    // 1.25x longer and fill the lanes of every pass worse.
    if (n >= 3 && static_engine_length(2 * n - 2)) return 2 * n - 2;
    int lo = 2 * n - 1, best = 0;
    double best_cost = 0;
    for (int M = lo; M <= 2 * lo + 16; ++M) {
        std::vector<int> r = smooth_radices(M);
        if (r.empty() || (int)r.size() > CB_MAX_PASS) continue;
        double c = plan_cost(M, r);
        // lengths served by a compile-time engine (fft_engine.cuh: 8.8.4, 8.8.5, 8.8.8, 8.10.8 with fused, pruned
        // first / last passes) run ~1.5x faster per streamed point than the runtime-radix engine
        if (!static_engine_length(M)) c *= 1.5;
        if (!best || c < best_cost) { best = M; best_cost = c; }
    }
    return best;
}

static inline bool build_fft_plan(int N, FftPlan &pl, std::vector<double2> &tw) {
    std::vector<int> radices = smooth_radices(N);
    if (radices.empty() || (int)radices.size() > CB_MAX_PASS) return false;
    pl.n = N;
    pl.npass = (int)radices.size();
    tw.clear();
    int L = N;
    for (int p = 0; p < pl.npass; ++p) {
        FftPass &ps = pl.pass[p];
        ps.radix = radices[p];
        ps.len = L;
        ps.sub = L / ps.radix;
        ps.magic = ps.sub > 1 ? (uint32_t)(((1ull << 32) + (uint64_t)ps.sub - 1) / (uint64_t)ps.sub) : 0u;
        if (ps.sub > 1) {
            ps.tw_off = (int32_t)tw.size();
            for (int s = 1; s < ps.radix; ++s)
                for (int j = 0; j < ps.sub; ++j) {
                    lcplx w = cis_ratio((long long)j * s, L);
                    tw.push_back(cplx((double)w.real(), (double)w.imag()));
                }
        } else {
            ps.tw_off = -1;
        }
        L = ps.sub;
    }
    for (int p = pl.npass; p < CB_MAX_PASS; ++p) pl.pass[p] = FftPass{0, 0, 0, 0u, -1};
    pl.tw_count = (int32_t)tw.size();
    return true;
}

// Reference DIF in long double, same pass structure as the device engine: natural in,
// digit-reversed out.
static inline void host_dif(std::vector<lcplx> &a, const FftPlan &pl) {
    const int N = pl.n;
    for (int p = 0; p < pl.npass; ++p) {
        const int R = pl.pass[p].radix, L = pl.pass[p].len, sub = pl.pass[p].sub;
        for (int base = 0; base < N; base += L)
            for (int j = 0; j < sub; ++j) {
                lcplx v[16], y[16];
                for (int t = 0; t < R; ++t) v[t] = a[base + j + t * sub];
                for (int s = 0; s < R; ++s) {
                    lcplx acc(0, 0);
                    for (int t = 0; t < R; ++t) acc += v[t] * cis_ratio((long long)s * t, R);
                    y[s] = acc * cis_ratio((long long)j * s, L);
                }
                for (int s = 0; s < R; ++s) a[base + j + s * sub] = y[s];
            }
    }
}

// pos[k] = slot that holds X[k] after the DIF passes.
static inline std::vector<int32_t> digit_reversed_slots(const FftPlan &pl) {
    std::vector<int32_t> pos(pl.n);
    for (int p = 0; p < pl.n; ++p) {
        int k = 0, mult = 1, rem = p;
        for (int q = 0; q < pl.npass; ++q) {
            int s = rem / pl.pass[q].sub;
            rem -= s * pl.pass[q].sub;
            k += s * mult;
            mult *= pl.pass[q].radix;
        }
        pos[k] = p;
    }
    return pos;
}

// chirp[j] = exp(-i pi j^2 / n) (j < n);  filt = DIF(b)/M in digit-reversed order and padded
// addressing (CB_PAD), where b[m] = conj(chirp[|m|]) wrapped circularly into length M.
static inline void build_bluestein_tables(int n, const FftPlan &pl, std::vector<double2> &chirp,
                                          std::vector<double2> &filt) {
    const int M = pl.n;
    std::vector<lcplx> w(n);
    chirp.resize(n);
    for (int j = 0; j < n; ++j) {
        long long q = ((long long)j * j) % (2LL * n);          // j^2 mod 2n keeps the phase exact
        w[j] = cis_ratio(q, 2LL * n);                          // exp(-2 pi i q / 2n) = exp(-i pi j^2 / n)
        chirp[j] = cplx((double)w[j].real(), (double)w[j].imag());
    }
    std::vector<lcplx> b(M, lcplx(0, 0));
    b[0] = std::conj(w[0]);
    for (int m = 1; m < n; ++m) {
        b[m] = std::conj(w[m]);
        b[M - m] = std::conj(w[m]);
    }
    host_dif(b, pl);
    filt.assign(fft_line_slots(M), cplx(0.0, 0.0));
    for (int p = 0; p < M; ++p) {
        lcplx h = b[p] / (long double)M;
        filt[CB_PAD(p)] = cplx((double)h.real(), (double)h.imag());
    }
}

}  // namespace cbplan
