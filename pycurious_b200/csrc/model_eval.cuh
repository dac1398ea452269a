// Warp-cooperative evaluation of the Bouligand model over all wavenumbers of one window.
// This is the hot loop of the fit and MCMC kernels (pycurious/optimise.py:223 executed ~50 x per fit,
// 101,000 x per chain).  One warp owns one window; its shared-memory slab holds the window's
// wavenumber tables, the model outputs and the per-evaluation order tables.
//
//   * the (ascending) wavenumbers are split into three contiguous classes by x = k dz, so that a
//     warp pass never mixes code paths:
//       x <= 2          Temme's series (few points)
//       2 < x < x_skip  K_mu, K_mu+1 by Miller's backward recurrence on Temme's three-term
//                       relation z_{n-1} = b_n z_n + a_{n+1} z_{n+1},  b_n = 2(n+x),
//                       a_n = -((n-1/2)^2 - mu^2), normalised by S = sum C_n z_n  -- the same
//                       continued fraction CF2 that Steed's algorithm evaluates forward, but with
//                       NO division inside the loop and a trip count that is uniform across the
//                       warp (N = 12 + 140/x_min, validated against mpmath: 6e-16 relative)
//       x >= x_skip     the K term is below 2^-60 of the cosh term, i.e. it cannot change the
//                       rounded bracket: Phi = C - 2 k zt - (beta-1) ln k + ln(lead Gamma(nu)/4),
//                       dPhi/ddz = 0 (cosh overflow at x > 710.47 still gives +inf like numpy)
//   * the bracket is evaluated in the scaled form  e^-x [..] = Gamma(nu)/4 (1 + e^-2x) - e^-x K hx,
//     so the "-x + log(bracket)" pair of grid.py:598 collapses to one log and nothing overflows
//   * everything that depends on the order only lives in per-evaluation tables in shared memory,
//     built lane-parallel: ModelCoef (lane o computes the coefficients of parameter vector o),
//     C_n / a_n (prefix product by warp shuffles), 1/(j-mu), 1/(j+mu)
//   * everything that depends on the wavenumber only (ln k, 1/k, sqrt(pi/2k)) is tabulated once per
//     window, so the per-point work has no division by x and no sqrt
//
// Falls back to the per-point generic path (bouligand.cuh) for unsorted k, dz <= 0, beta < 0,
// very large orders or non-finite parameters, where the reference's nan/inf pattern matters more
// than speed.  The heavy functions are deliberately NOT inlined on the device: the first version
// of the fit kernel inlined them at every call site, grew to 770 KB of SASS and spent 75 % of its
// issue slots waiting on the instruction cache (profiles/r01_fit_icache.md).
#pragma once
#include <string.h>
#include "bouligand.cuh"

#define CB_MILLER_NMAX 97                     /* tables hold n = 0..97 */
#define CB_TEMME_TAB 28
#define CB_SKIP_LOG 43.0                      /* > 60 ln 2 + ln(4 sqrt(pi/4)) */
#define CB_COSH_OVERFLOW 710.4758600739439    /* np.cosh overflows above this */
#ifndef CB_MH_BLOCK
#define CB_MH_BLOCK 8
#endif                         /* proposals prepared together by the MCMC kernel */
#define CB_MAX_ORDERS 3                       /* parameter vectors evaluated per call group: x, x+h e_beta, x-h e_beta */

#if defined(__CUDACC__) && !defined(CB_EMU)
#define CB_NOINLINE_BODY static __device__ __noinline__
#else
#define CB_NOINLINE_BODY inline
#endif

// ModelCoef plus what the fast path needs, one per parameter vector, in shared memory.
struct EvalCoef {
    ModelCoef m;
    double a1;          // 1/4 - mu^2
    double rdz;         // 1/dz
    double rsdz;        // 1/sqrt(dz)
    double lgnu;        // ln Gamma(nu)
    double lgc;         // ln(lead Gamma(nu)/4)
    double hp, hq;      // 1/(2 Gamma(1+mu)), 1/(2 Gamma(1-mu))
    double rmu;         // 1/mu (unused when mu ln(2/x) is tiny)
    int32_t fast;       // 1: the class-split path applies
    int32_t pad_;
};

struct WarpSlab {          // views into one warp's shared-memory slab
    // window-independent tables, built once per launch in global memory (ktable_fill), read through L1
    const double *k, *logk, *rk, *sq;   // [nk] k, ln k, 1/k, sqrt(pi/(2k))
    const double *rn;                   // [CB_MILLER_NMAX + 1] 1/n
    // per-window arrays in the warp's shared-memory slab
    double *obs;                        // [nk] observed Phi
    double *p0, *pp, *pm, *dz;          // [nk] model values (p0) and, for the fit, Phi(beta+-h), dPhi/ddz
    double *kb;                         // [nk] K_mu+1 between the phases of one evaluation (K_mu sits in out_phi)
    double *res;                        // [nk] fit only: residuals during the reduction
    double2 *CA;                        // [CB_MILLER_NMAX + 1] (C_n, a_{n+1}) of the order being evaluated
    double2 *rpq;                       // [CB_TEMME_TAB] (1/(j-mu), 1/(j+mu))
    double *rnj;                        // [CB_TEMME_TAB] 1/j (shared copy of rn for the Temme loop)
    EvalCoef *coef;                     // [ncoef]: CB_MAX_ORDERS (fit) or CB_MH_BLOCK (MCMC) parameter vectors
    double *scratch;                    // solver state (fit) behind the tables
    int nk;
    bool sorted;
    double kmin, kmax;
    double cu;                          // fit: solver coordinate u = exp(-cu dz) (0 = none), see fit_body.cuh
};

#define CB_SLAB_COEF_DOUBLES ((int)((sizeof(EvalCoef) + 7) / 8))

CB_HD int warp_slab_doubles(int nk, bool fit, int scratch_doubles = 0, int ncoef = CB_MAX_ORDERS) {
    int n = (fit ? 6 : 3) * nk + 2 * (CB_MILLER_NMAX + 1) + 3 * CB_TEMME_TAB +
            ncoef * CB_SLAB_COEF_DOUBLES + scratch_doubles;
    return (n + 1) & ~1;      // keep every slab 16-byte aligned
}

// Window-independent tables of one launch: [0,nk) ln k, [nk,2nk) 1/k, [2nk,3nk) sqrt(pi/(2k)),
// then CB_MILLER_NMAX+1 values 1/n, then (sorted flag, k_min, k_max).
CB_HD int ktable_doubles(int nk) { return 3 * nk + (CB_MILLER_NMAX + 1) + 3; }

// Entry i of the table (i < ktable_doubles(nk)); every entry is independent so that the fill is a
// plain parallel loop on the device and a serial one in the CPU harness.
CB_HD double ktable_entry(const double *kh, int nk, int i) {
    if (i < nk) return cb_log(kh[i]);
    if (i < 2 * nk) return 1.0 / kh[i - nk];
    if (i < 3 * nk) return sqrt(0.5 * CB_PI / kh[i - 2 * nk]);
    i -= 3 * nk;
    if (i <= CB_MILLER_NMAX) return i ? 1.0 / (double)i : 0.0;
    i -= CB_MILLER_NMAX + 1;
    double kmin = INFINITY, kmax = 0.0;
    bool asc = true;
    for (int j = 0; j < nk; ++j) {
        const double k = kh[j];
        if (j + 1 < nk && !(kh[j + 1] > k)) asc = false;
        if (!(k > 0.0)) asc = false;
        kmin = fmin(kmin, k);
        kmax = fmax(kmax, k);
    }
    return i == 0 ? (asc ? 1.0 : 0.0) : (i == 1 ? kmin : kmax);
}

// The table part of the layout -- everything model_eval_warp touches besides its array arguments -- follows from
// (base, kh, ktab, nk) alone: the evaluation core re-derives it from those four registers instead of reading a
// WarpSlab that lives in its caller's local memory (one LDL per field use).  Returns the first slot behind it.
CB_HD double *warp_slab_tables(WarpSlab &W, double *base, const double *kh, const double *ktab, int nk) {
    W.nk = nk;
    W.k = kh; W.logk = ktab; W.rk = ktab + nk; W.sq = ktab + 2 * nk; W.rn = ktab + 3 * nk;
    double *t = base;
    W.CA = reinterpret_cast<double2 *>(t); t += 2 * (CB_MILLER_NMAX + 1);
    W.rpq = reinterpret_cast<double2 *>(t); t += 2 * CB_TEMME_TAB;
    W.rnj = t; t += CB_TEMME_TAB;
    W.coef = reinterpret_cast<EvalCoef *>(t);
    return t;
}

CB_HD WarpSlab warp_slab_carve(double *base, int nk, bool fit, const double *kh, const double *ktab,
                               int ncoef = CB_MAX_ORDERS) {
    WarpSlab W;
    const double *meta = ktab + 3 * nk + (CB_MILLER_NMAX + 1);
    W.sorted = meta[0] != 0.0;
    W.kmin = meta[1];
    W.kmax = meta[2];
    double *t = warp_slab_tables(W, base, kh, ktab, nk);
    t += ncoef * CB_SLAB_COEF_DOUBLES;
    W.obs = t; W.p0 = t + nk; W.kb = t + 2 * nk;
    t += 3 * nk;
    if (fit) { W.pp = t; W.pm = t + nk; W.dz = t + 2 * nk; W.res = W.pm; t += 3 * nk; }   // res reuses the spent pm slot
    else { W.pp = W.pm = W.dz = W.res = nullptr; }
    W.scratch = t;
    W.cu = 0.0;
    return W;
}

// Stage the window's spectrum.
template <class Ctx>
CB_BODY void warp_slab_load(const Ctx &cx, WarpSlab &W, const double *phi_row, int lane) {
    const int L = cx.team_lanes();
#pragma unroll 1
    for (int i = lane; i < W.nk; i += L) W.obs[i] = phi_row[i];
    cx.sync_warp();
}

CB_HD EvalCoef eval_coef_make(double beta, double zt, double dz, double C) {
    EvalCoef e;
    e.m = model_coef(beta, zt, dz, C);
    const ModelCoef &m = e.m;
    e.a1 = 0.25 - m.ord.mu * m.ord.mu;
    e.rdz = 1.0 / m.dz;
    e.rsdz = 1.0 / sqrt(m.dz);
    e.lgnu = cb_log(m.gnu);
    e.lgc = cb_log(m.lead * m.gnu * 0.25);
    e.hp = 0.5 / m.ord.gampl;
    e.hq = 0.5 / m.ord.gammi;
    e.rmu = 1.0 / m.ord.mu;
    e.fast = (m.dz > 0.0 && m.dz < 1e300 && m.nu >= 0.5 && m.nu <= 8.5 && isfinite(m.zt) && isfinite(m.C) &&
              m.gnu > 0.0 && m.lead > 0.0 && isfinite(m.gnu) && isfinite(m.lead)) ? 1 : 0;
    e.pad_ = 0;
    return e;
}

// Coefficients of `count` parameter vectors x + off[idx] e_beta, idx = first .. first+count-1, with
// off = (0, +hb, -hb): lane l computes vector first+l (the lanes run the same code on different data).
template <class Ctx>
CB_NOINLINE_BODY void model_coefs_core(const Ctx &cx, EvalCoef *coef, const double *x, double hb, int first, int count,
                                       int lane) {
    const int L = cx.team_lanes();
    for (int l = lane; l < count; l += L) {
        const int idx = first + l;
        const double beta = x[0] + (idx == 1 ? hb : (idx == 2 ? -hb : 0.0));
        coef[idx] = eval_coef_make(beta, x[1], x[2], x[3]);
    }
    cx.sync_warp();
}
template <class Ctx>
CB_BODY void model_coefs_warp(const Ctx &cx, const WarpSlab &W, const double *x, double hb, int first, int count,
                              int lane) {
    model_coefs_core(cx, W.coef, x, hb, first, count, lane);      // (pointer argument: see model_eval_core)
}

// Coefficients of `count` unrelated parameter vectors props[l][0..3] -> slot l, one lane each (the
// Metropolis-Hastings kernel prepares a block of proposals at once).
template <class Ctx>
CB_NOINLINE_BODY void model_coefs_lanes_core(const Ctx &cx, EvalCoef *coef, const double *props, int count, int lane) {
    const int L = cx.team_lanes();
    for (int l = lane; l < count; l += L) coef[l] = eval_coef_make(props[4 * l], props[4 * l + 1], props[4 * l + 2], props[4 * l + 3]);
    cx.sync_warp();
}
template <class Ctx>
CB_BODY void model_coefs_lanes(const Ctx &cx, const WarpSlab &W, const double *props, int count, int lane) {
    model_coefs_lanes_core(cx, W.coef, props, count, lane);
}

// 2^-e for |v| = m 2^e, m in [1, 2) (1.0 for zero, subnormal, inf or nan): an exact scale factor.
CB_HD double cb_pow2_inverse_of(double v) {
    unsigned long long b;
    memcpy(&b, &v, sizeof(b));
    const unsigned long long ex = (b >> 52) & 0x7ffull;
    if (ex == 0ull || ex >= 0x7feull) return 1.0;
    const unsigned long long r = (0x7feull - ex) << 52;          // biased exponent of 2^-(ex-1023) is 2046 - ex
    double o;
    memcpy(&o, &r, sizeof(o));
    return o;
}

// Reference-order arithmetic for one point (the slow, always-valid path).
CB_NOINLINE_BODY double generic_point(const ModelCoef *m, double k, double logk, double *ddz) {
    if (ddz) return bouligand_point_ddz(*m, k, logk, *ddz);
    return bouligand_point(*m, k, logk);
}

// Phi(k_i) of parameter vector `idx` (coefficients prepared by model_coefs_warp) for every
// wavenumber of the slab -> out_phi[i] (and dPhi/ddz -> out_ddz[i] if not NULL).
// Phases, each a single copy of its code (exp / log / division sites are not duplicated):
//   S  x >= x_skip      closed form (first: it needs no Bessel function)
//   then per chunk of <= 32 points of one class
//   T  x <= 2           Temme series          -> K_mu,      K_mu+1        (unscaled)
//   M  2 < x < x_skip   Miller recurrence     -> K_mu e^x,  K_mu+1 e^x
//   F  scale by e^-x resp. e^-2x, upward recurrence in the order, bracket, log
// Returns 0.5 sum_i (Phi_i - obs_i)^2 when `obs` is given (0 otherwise); with a finite `bound` the
// evaluation may stop early and return a partial sum that already exceeds the bound.
// OBS = false (the fit: values and derivatives only) compiles the residual sums and bound checks out -- each
// warp_sum is ten shuffles wrapped in warp-sync sequences inside a non-inlined function, a fifth of its size.
template <bool OBS, class Ctx>
CB_NOINLINE_BODY double model_eval_core(const Ctx &cx, double *slab_base, const double *kh, const double *ktab, int nk_,
                                        int sorted, double cu, double *kb, int idx, int lane, double *out_phi,
                                        double *out_ddz, const double *obs_, double bound) {
    WarpSlab W;                      // register-resident view: only the fields set here are used below
    warp_slab_tables(W, slab_base, kh, ktab, nk_);
    W.sorted = sorted != 0;
    W.cu = cu;
    W.kb = kb;
    const double *obs = OBS ? obs_ : nullptr;
    const EvalCoef &e = W.coef[idx];
    const ModelCoef &m = e.m;
    const int nk = W.nk, L = cx.team_lanes();
    if (!(W.sorted && e.fast)) {                       // generic per-point path, reference nan/inf semantics
        double accg = 0.0;
#pragma unroll 1
        for (int i = lane; i < nk; i += L) {
            const double ph = generic_point(&m, Ctx::ldg(W.k + (i)), Ctx::ldg(W.logk + (i)), out_ddz ? out_ddz + i : nullptr);
            out_phi[i] = ph;
            if (OBS && obs) { const double r = ph - obs[i]; accg = fma(r, r, accg); }
        }
        cx.sync_warp();
        if (OBS && obs) return 0.5 * cx.warp_sum(accg);
        return 0.0;
    }
    const double mu = m.ord.mu, a1 = e.a1, dz = m.dz;

    // ---- classes: [0,nT) Temme, [nT,nS) Miller, [nS,nk) K negligible ---------------------------
    // With residuals wanted (OBS: Metropolis-Hastings, min_func) the closed-form class is evaluated in the same
    // pass, from the k and ln k this lane already holds: ~90 % of the proposals are rejected on that partial
    // misfit, so for them this loop is the whole evaluation.
    int lastT = -1, lastC = -1, firstSkip = nk;
    double accS = 0.0;
    const double cudz = out_ddz ? W.cu * dz : 0.0;
    const auto OBSV = Ctx::shared_view(obs);
#pragma unroll 1
    for (int i = lane; i < nk; i += L) {
        const double k = Ctx::ldg(W.k + (i)), lk = Ctx::ldg(W.logk + (i));
        const double xi = k * dz;
        if (xi <= 2.0) lastT = i;
        // the K term relative to the cosh term is ~ exp(-(2x - nu ln(x/2) + ln Gamma(nu))).  When the caller
        // iterates in u = exp(-cu dz) the derivative d/du = -(1/(cu u)) d/ddz magnifies it by exp(cu dz):
        // W.cu dz is taken off the exponent so that dPhi/du keeps its relative accuracy on the flat tail
        const bool skip = xi > 2.0 && (2.0 * xi - cudz - m.nu * (lk + m.lhalfdz) + e.lgnu > CB_SKIP_LOG);
        if (!skip) lastC = i;
        else if (OBS && obs) {
            if (i < firstSkip) firstSkip = i;
            const double ph = xi > CB_COSH_OVERFLOW ? INFINITY : fma(-2.0 * k, m.zt, m.C) - (m.beta - 1.0) * lk + e.lgc;
            const double r = ph - OBSV[i];
            accS = fma(r, r, accS);
        }
    }
    const int nT = cx.warp_max_i(lastT) + 1;
    int nS = cx.warp_max_i(lastC) + 1;
    if (nS < nT) nS = nT;

    // Residuals against `obs` are accumulated as the values appear (acc: this lane's share of sum r^2) and,
    // when the caller gave a finite `bound`, the partial misfit is checked after the cheap class and after
    // every chunk: 0.5 sum r^2 only grows, so a proposal whose partial misfit is already above the bound is
    // rejected without the Bessel work of the remaining points (Metropolis-Hastings rejects ~90 %).
    double acc = 0.0;
    const bool bounded = OBS && obs != nullptr && bound < 1e300;
    // the one-pass sums stand unless a skipped point lies below nS (skip flags not monotone in k: it is then
    // evaluated with the Bessel classes and the closed-form class is redone on its own)
    const bool fused = OBS && obs != nullptr && !cx.any(firstSkip < nS);
    if (fused) acc = accS;

    // ---- S: the Bessel term cannot change the rounded bracket (no K_nu needed: done first) --------
    if (!fused) {
#pragma unroll 1
        for (int i = nS + lane; i < nk; i += L) {
            const double k = Ctx::ldg(W.k + (i));
            const bool over = k * dz > CB_COSH_OVERFLOW;
            const double ph = over ? INFINITY : fma(-2.0 * k, m.zt, m.C) - (m.beta - 1.0) * Ctx::ldg(W.logk + (i)) + e.lgc;
            out_phi[i] = ph;
            if (out_ddz) out_ddz[i] = over ? NAN : 0.0;
            if (OBS && obs) { const double r = ph - obs[i]; acc = fma(r, r, acc); }
        }
    }
    if (bounded && nS < nk) {
        const double part = 0.5 * cx.warp_sum(acc);
        if (!(part <= bound)) { cx.sync_warp(); return part; }       // also leaves on nan / inf
    }

    // ---- order-only tables, built lane-parallel ------------------------------------------------
    if (nS > nT) {
        // C_0 = 1, C_1 = a1, C_n = C_{n-1} (a1 + n(n-1)) / n;  CA[n].y = a_{n+1} = -(a1 + n(n+1))
        // lane l owns n = 3l+2 .. 3l+4; local prefix, product scan over lanes, expand
        if (L == 1) {                                   // serial emulation: one lane builds the whole table
            double c = a1;
            W.CA[0] = make_double2(1.0, -a1);
            W.CA[1] = make_double2(a1, -(a1 + 2.0));
            for (int n = 2; n <= CB_MILLER_NMAX; ++n) {
                c *= fma(a1, Ctx::ldg(W.rn + (n)), (double)(n - 1));
                W.CA[n] = make_double2(c, -(a1 + (double)n * (double)(n + 1)));
            }
        } else {
            double q[3];
            double run = 1.0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int n = 3 * lane + 2 + j;
                const double g = n <= CB_MILLER_NMAX ? fma(a1, Ctx::ldg(W.rn + (n)), (double)(n - 1)) : 1.0;
                run *= g;
                q[j] = run;
            }
            double incl = run;
#pragma unroll 1
            for (int d = 1; d < L; d <<= 1) {
                const double t = cx.shfl_up(incl, d);
                if (lane >= d) incl *= t;
            }
            double excl = cx.shfl_up(incl, 1);
            if (lane == 0) excl = 1.0;
            const double base = a1 * excl;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int n = 3 * lane + 2 + j;
                if (n <= CB_MILLER_NMAX) W.CA[n] = make_double2(base * q[j], -(a1 + (double)n * (double)(n + 1)));
            }
            if (lane == 0) {
                W.CA[0] = make_double2(1.0, -a1);
                W.CA[1] = make_double2(a1, -(a1 + 2.0));
            }
        }
    }
    if (nT > 0) {
#pragma unroll 1
        for (int j = 1 + lane; j < CB_TEMME_TAB; j += L) {
            W.rpq[j] = make_double2(1.0 / ((double)j - mu), 1.0 / ((double)j + mu));
            W.rnj[j] = Ctx::ldg(W.rn + (j));
        }
    }
    cx.sync_warp();

    // ---- chunks of up to L consecutive points of [0,nS): a lane's point is in the Temme class (i < nT) or the
    //      Miller class; the chunk that straddles nT runs both branches (lanes diverge) and shares ONE phase F.
    //      K_mu, K_mu+1 (phase T or M), then scale / raise the order / bracket (phase F) ----------------
    const double qg = 0.25 * m.gnu;
    const int nl = m.ord.nl;
#pragma unroll 1
    for (int c0 = 0; c0 < nS;) {
        const int c1 = c0 + L < nS ? c0 + L : nS;
        const int i = c0 + lane;
        const bool temme = i < nT;
        if (temme) {
            if (i < c1) {
                const double d = -(Ctx::ldg(W.logk + (i)) + m.lhalfdz);      // ln(2/x)
                const double sg = mu * d;
                const double ee = cb_exp(sg), ei = 1.0 / ee;
                // sinh(sg)/sg * d = sinh(sg)/mu
                const double shd = fabs(sg) < 1e-4 ? d * (1.0 + sg * sg * (1.0 / 6.0)) : 0.5 * (ee - ei) * e.rmu;
                double ff = m.ord.sinc * (m.ord.gam1 * 0.5 * (ee + ei) + m.ord.gam2 * shd);
                double p = e.hp * ee, q = e.hq * ei, c = 1.0;
                const double x2 = 0.5 * Ctx::ldg(W.k + (i)) * dz, dd = x2 * x2;
                double sum = ff, sum1 = p;
                const auto RPQ = Ctx::shared_view(W.rpq);
                const auto RNJ = Ctx::shared_view(W.rnj);
#pragma unroll 1
                for (int j = 1; j < CB_TEMME_TAB; ++j) {
                    const double fj = (double)j;
                    const double2 r = RPQ[j];
                    ff = (fj * ff + p + q) * (r.x * r.y);
                    c *= dd * RNJ[j];
                    p *= r.x;
                    q *= r.y;
                    const double del = c * ff;
                    sum += del;
                    sum1 = fma(c, p - fj * ff, sum1);
                    if (fabs(del) < fabs(sum) * CB_KNU_EPS) break;
                }
                out_phi[i] = sum;                       // K_mu; phase F turns it into Phi in place
                W.kb[i] = sum1 * (2.0 * Ctx::ldg(W.rk + (i)) * e.rdz);
            }
        } else if (c1 > nT) {
            // Miller backward recurrence, trip count uniform over the chunk's Miller points (from the smallest
            // x among them: k ascending)
            const int ii = i < c1 ? i : c1 - 1;
            const double xi = Ctx::ldg(W.k + (ii)) * dz;
            int N = (int)(12.0 + 140.0 * (Ctx::ldg(W.rk + (c0 > nT ? c0 : nT)) * e.rdz)) + 1;
            if (N > CB_MILLER_NMAX - 1) N = CB_MILLER_NMAX - 1;
            // (x: weight of y_n in the normalising sum, y: recurrence coefficient); one shared load per step
            const auto CA = Ctx::shared_view(W.CA);
            double2 ca = CA[N];
            double yp = 0.0, y = 1.0, S = ca.x;
            double bn = 2.0 * ((double)N + xi);
#pragma unroll 2
            for (int n = N; n >= 1; --n) {
                const double2 cm = CA[n - 1];
                const double ym = fma(bn, y, ca.y * yp);
                yp = y;
                y = ym;
                S = fma(cm.x, y, S);
                bn -= 2.0;
                ca = cm;
            }
            if (i < c1) {
                const double invx = Ctx::ldg(W.rk + (ii)) * e.rdz;
                // y, yp and S have grown together to ~1e100+: take S's exponent off all three (exact) so that
                // the division stays on the hardware's fast path instead of the out-of-range slow path
                const double sc = cb_pow2_inverse_of(S);
                const double cS = Ctx::ldg(W.sq + (ii)) * e.rsdz / (S * sc);        // sqrt(pi/(2x)) / S, scaled
                const double ys = y * sc, yps = yp * sc;
                out_phi[i] = cS * ys;
                W.kb[i] = cS * fma(mu + xi + 0.5, ys, -a1 * yps) * invx;
            }
        }
        // ---- F (same lane, same point: no barrier needed): bracket e^-x = Gamma(nu)/4 (1 + e^-2x) - hx (K_nu e^-x)
        if (i < c1) {
            const double k = Ctx::ldg(W.k + (i)), lk = Ctx::ldg(W.logk + (i)), xi = k * dz;
            const double ex = cb_exp(temme ? -xi : -2.0 * xi);     // scales K_mu, K_mu+1 to K e^-x in both classes
            const double e2 = temme ? ex * ex : ex;
            double a = out_phi[i] * ex, b = W.kb[i] * ex;
            const double tox = 2.0 * Ctx::ldg(W.rk + (i)) * e.rdz;
#pragma unroll 1
            for (int j = 1; j < nl; ++j) {
                const double c = fma((mu + (double)j) * tox, b, a);
                a = b;
                b = c;
            }
            const double hx = cb_exp(m.nu * (lk + m.lhalfdz));
            const double br = fma(qg, 1.0 + e2, -b * hx);
            const double ph = fma(-2.0 * k, m.zt, m.C) - (m.beta - 1.0) * lk + cb_log(m.lead * br);
            out_phi[i] = ph;
            // k [ (Gamma/4 (1 - e2) + hx K_{nu-1} e^-x) / br - 1 ] with the subtraction done analytically
            if (out_ddz) out_ddz[i] = k * fma(hx, a + b, -2.0 * qg * e2) / br;
            if (OBS && obs) { const double r = ph - obs[i]; acc = fma(r, r, acc); }
        }
        c0 = c1;
        if (bounded && c0 < nS) {
            const double part = 0.5 * cx.warp_sum(acc);
            if (!(part <= bound)) { cx.sync_warp(); return part; }
        }
    }
    cx.sync_warp();
    if (OBS && obs) return 0.5 * cx.warp_sum(acc);
    return 0.0;
}

template <bool OBS, class Ctx>
CB_BODY double model_eval_warp(const Ctx &cx, const WarpSlab &W, int idx, int lane, double *out_phi, double *out_ddz,
                               const double *obs = nullptr, double bound = INFINITY) {
    return model_eval_core<OBS>(cx, reinterpret_cast<double *>(W.CA), W.k, W.logk, W.nk, W.sorted ? 1 : 0, W.cu, W.kb,
                                idx, lane, out_phi, out_ddz, obs, bound);
}

// ---- fused evaluation of the fit: (trial, beta+h [, beta-h]) -> the 15 sums of one Gauss-Newton step ---------------
// The MAP fit needs Phi at a trial point and at its beta +- h neighbours (dPhi/dbeta has no closed form).  The
// vectors share zt, dz and C, hence x = k dz and the class split; evaluating them one after the other runs every
// class pass once per vector, each with few live lanes (the Bessel classes of a 99-bin window hold ~40 points: one
// full chunk with the longest Miller recurrence plus a quarter-full one).  Here the work items are (point, vector)
// PAIRS: lane = (point within the chunk) * nv + vector, so the chunks are full, the pairs of one chunk have
// neighbouring x (the Miller trip count follows the chunk's smallest x) and the number of passes drops from 2 nv to
// ~nv + 1.  The model values stay in shared memory (4 nk doubles per window); from them every lane accumulates its
// points' contributions to
//     sums[0] = r.r;  sums[1..4] = r.J_a;  sums[5..14] = J_a.J_b (packed upper triangle)
// with r = Phi - obs and J = (dPhi/dbeta, dPhi/dzt = -2k, dPhi/ddz, dPhi/dC = 1), and the lanes are summed by a
// transposing butterfly (16 double shuffles).  Per-warp tables in shared memory (doubles): CT [3][CB_MILLER_NMAX+1]
// weights C_n of the normalising sum per vector (a_{n+1} is updated inside the recurrence), RPQ [3][CB_TEMME_TAB]
// double2, RNJ [CB_TEMME_TAB] 1/j, COEF [3][FC_COUNT].
enum {
    FC_MU, FC_A1, FC_RMU, FC_HP, FC_HQ, FC_SINC, FC_GAM1, FC_GAM2, FC_QG, FC_NU, FC_BM1, FC_LGC, FC_LEAD, FC_LGNU,
    FC_NL, FC_FAST, FC_BETA, FC_ZT, FC_C, FC_DZ, FC_RDZ, FC_RSDZ, FC_LHDZ,
    FC_AUX,            // vector 0: the beta step hb; vector 1: the solver coordinate's cu; vector 2: number of vectors
    FC_COUNT
};
#define CB_FIT_CT_DOUBLES (3 * (CB_MILLER_NMAX + 1))
#define CB_FIT_RPQ_DOUBLES (2 * 3 * CB_TEMME_TAB)
#define CB_FIT_TABLE_DOUBLES (CB_FIT_CT_DOUBLES + CB_FIT_RPQ_DOUBLES + CB_TEMME_TAB + 3 * FC_COUNT)
#define CB_FIT_COEF_OFFSET (CB_FIT_CT_DOUBLES + CB_FIT_RPQ_DOUBLES + CB_TEMME_TAB)

// Coefficients of vector v of the trial point x (beta offset by 0, +hb, -hb) -> c[FC_COUNT].
CB_HD void fit_coef_fill(double *c, const double *x, double hb, int v) {
    const double beta = x[0] + (v == 1 ? hb : (v == 2 ? -hb : 0.0));
    const EvalCoef e = eval_coef_make(beta, x[1], x[2], x[3]);
    c[FC_MU] = e.m.ord.mu; c[FC_A1] = e.a1; c[FC_RMU] = e.rmu; c[FC_HP] = e.hp; c[FC_HQ] = e.hq;
    c[FC_SINC] = e.m.ord.sinc; c[FC_GAM1] = e.m.ord.gam1; c[FC_GAM2] = e.m.ord.gam2; c[FC_QG] = 0.25 * e.m.gnu;
    c[FC_NU] = e.m.nu; c[FC_BM1] = e.m.beta - 1.0; c[FC_LGC] = e.lgc; c[FC_LEAD] = e.m.lead; c[FC_LGNU] = e.lgnu;
    c[FC_NL] = (double)e.m.ord.nl; c[FC_FAST] = e.fast ? 1.0 : 0.0; c[FC_BETA] = e.m.beta; c[FC_ZT] = e.m.zt;
    c[FC_C] = e.m.C; c[FC_DZ] = e.m.dz; c[FC_RDZ] = e.rdz; c[FC_RSDZ] = e.rsdz; c[FC_LHDZ] = e.m.lhalfdz;
    c[FC_AUX] = 0.0;
}

// One point's contribution to the 15 sums (p0, pp, pm = Phi at beta, beta + hb, beta - hb; jd = dPhi/ddz).
CB_HD void fit_sums_add(double *v, double p0, double pp, double pm, bool use_pm, double rh, double jd, double k,
                        double obs) {
    const bool fp = isfinite(pp), fm = use_pm && isfinite(pm);
    const double jb = (fp && fm) ? (pp - pm) * (0.5 * rh) : (fp ? (pp - p0) * rh : (fm ? (p0 - pm) * rh : 0.0));
    const double r = p0 - obs, jk = -2.0 * k;
    v[0] = fma(r, r, v[0]); v[1] = fma(r, jb, v[1]); v[2] = fma(r, jk, v[2]); v[3] = fma(r, jd, v[3]); v[4] += r;
    v[5] = fma(jb, jb, v[5]); v[6] = fma(jb, jk, v[6]); v[7] = fma(jb, jd, v[7]); v[8] += jb;
    v[9] = fma(jk, jk, v[9]); v[10] = fma(jk, jd, v[10]); v[11] += jk;
    v[12] = fma(jd, jd, v[12]); v[13] += jd;
    v[14] += 1.0;
}

// exp / log inside the fused evaluation are inlined (three exp and one log site): a call in the chunk loop makes the
// 16 running sums and the loop state live across it, and at 80 registers they spill around every call.
CB_HD double cb_log_inline(double x) { return 2.0 * cb_half_log(x); }

// Cold path of the fused evaluation: one point, all vectors, through the reference-order arithmetic (unsorted k,
// dz <= 0, beta < 0, non-finite parameters ...).  out = (Phi, Phi(beta + hb), Phi(beta - hb), dPhi/ddz).
CB_NOINLINE_BODY void fit_generic_point(const double *coef, int nv, double k, double lk, double *out) {
    out[2] = 0.0; out[3] = 0.0;
    for (int u = 0; u < nv; ++u) {
        const double *c = coef + u * FC_COUNT;
        const ModelCoef m = model_coef(c[FC_BETA], c[FC_ZT], c[FC_DZ], c[FC_C]);
        out[u] = generic_point(&m, k, lk, u == 0 ? out + 3 : nullptr);
    }
}

// 16 per-lane sums over the lanes by a transposing butterfly: every level halves the sums a lane carries (16 double
// shuffles instead of 15 x 5); lane l ends with sum number (bits 4..1 of l, most significant first) and writes it.
template <class Ctx>
CB_BODY void fit_sums_reduce(const Ctx &cx, double *v, int lane, double *sums_out) {
    if (cx.team_lanes() == 32) {
        int mine = 0;
#pragma unroll
        for (int lv = 0; lv < 4; ++lv) {
            const int m = 16 >> lv, n = 8 >> lv;
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int j = 0; j < n; ++j) {
                const double send = up ? v[j] : v[j + n], keep = up ? v[j + n] : v[j];
                v[j] = keep + cx.shfl_xor(send, m);
            }
            mine = 2 * mine + (up ? 1 : 0);
        }
        v[0] += cx.shfl_xor(v[0], 1);
        if ((lane & 1) == 0 && mine < 15) sums_out[mine] = v[0];
    } else {
        for (int q = 0; q < 15; ++q) sums_out[q] = v[q];
    }
}

// The 15 sums of one window from stored model values (fit_eval_values) and its spectrum.
template <class Ctx>
CB_BODY void fit_sums_from_values(const Ctx &cx, const double *store, const double *kh, int nk, bool use_pm, double hb,
                                  int lane, const double *obs, double *sums_out) {
    const int L = cx.team_lanes();
    const double rh = 1.0 / hb;
    double v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = 0.0;
#pragma unroll 1
    for (int i = lane; i < nk; i += L)
        fit_sums_add(v, store[i], store[nk + i], store[2 * nk + i], use_pm, rh, store[3 * nk + i], Ctx::ldg(kh + i), Ctx::ldg(obs + i));
    fit_sums_reduce(cx, v, lane, sums_out);
    cx.sync_warp();
}

// tab: the warp's tables with COEF filled for nv vectors (FC_AUX: hb, cu, nv); obs: the window's spectrum (global);
// sums_out[16] receives the sums (any address space; written by 15 lanes after the butterfly).
// The model values go to store[0..4 nk) in shared memory (Phi, Phi(beta + hb), Phi(beta - hb), dPhi/ddz, nk apart),
// every lane writing the value of its own (point, vector) pair; the 15 sums are then formed from them by
// fit_sums_from_values -- ONE summation order for every way a window can be evaluated (a warp of the batched
// evaluation kernel, the four warps of a straggler's CTA, the shared first evaluation), so that a window's result
// does not depend on how many other windows are unfinished, i.e. on how a survey is split over GPUs.
// part / nparts: the warps of a CTA that work on ONE window together (the stragglers' finish kernel) take the
// chunks and the closed-form points round robin, warp `part` of `nparts`; the tables are built by part 0 behind a
// block barrier.  The caller synchronises (warp or block) before the values are read.
template <class Ctx>
CB_BODY void fit_eval_values(const Ctx &cx, double *tab, const double *kh, const double *ktab, int nk, int sorted,
                             int lane, double *store, int part = 0, int nparts = 1) {
    const int L = cx.team_lanes();
    double *CTg = tab;
    double2 *RPQg = reinterpret_cast<double2 *>(tab + CB_FIT_CT_DOUBLES);
    const double *RNJg = tab + CB_FIT_CT_DOUBLES + CB_FIT_RPQ_DOUBLES;
    double *COEF = tab + CB_FIT_COEF_OFFSET;
    const double *logk = ktab, *rk = ktab + nk, *sq = ktab + 2 * nk, *rn = ktab + 3 * nk;
    const auto C0 = Ctx::shared_view((const double *)COEF);
    const double hb = C0[FC_AUX], cu = C0[FC_COUNT + FC_AUX];
    const int nv = (int)C0[2 * FC_COUNT + FC_AUX];
    const bool use_pm = nv == 3;
    (void)hb;

    bool fast = sorted != 0;
    for (int u = 0; u < nv; ++u) fast = fast && C0[u * FC_COUNT + FC_FAST] != 0.0;
    if (!fast) {                                          // generic per-point path, reference nan/inf semantics
#pragma unroll 1
        for (int i = lane + L * part; i < nk; i += L * nparts) {
            double vals[4];
            fit_generic_point(COEF, nv, Ctx::ldg(kh + i), Ctx::ldg(logk + i), vals);
            store[i] = vals[0]; store[nk + i] = vals[1]; store[2 * nk + i] = vals[2]; store[3 * nk + i] = vals[3];
        }
    } else {
        const double dz = C0[FC_DZ], zt = C0[FC_ZT], Cc = C0[FC_C], lhdz = C0[FC_LHDZ], rdz = C0[FC_RDZ];
        // ---- classes from vector 0 (one split for all vectors: the difference quotient in beta must not straddle
        //      two evaluation methods): [0,nT) Temme, [nT,nS) Miller, [nS,nk) K negligible ------------------------
        int lastT = -1, lastC = -1;
        {
            const double nu0 = C0[FC_NU], lgnu0 = C0[FC_LGNU], cudz = cu * dz;
#pragma unroll 1
            for (int i = lane; i < nk; i += L) {
                const double k = Ctx::ldg(kh + i), lk = Ctx::ldg(logk + i);
                const double xi = k * dz;
                if (xi <= 2.0) lastT = i;
                const bool skip = xi > 2.0 && (2.0 * xi - cudz - nu0 * (lk + lhdz) + lgnu0 > CB_SKIP_LOG);
                if (!skip) lastC = i;
            }
        }
        const int nT = cx.warp_max_i(lastT) + 1;
        int nS = cx.warp_max_i(lastC) + 1;
        if (nS < nT) nS = nT;

        // ---- closed-form class: every vector from the k and ln k the lane holds; dPhi/ddz = 0 -------------------
        {
            const double b0 = C0[FC_BM1], g0 = C0[FC_LGC], b1 = C0[FC_COUNT + FC_BM1], g1 = C0[FC_COUNT + FC_LGC];
            const double b2 = use_pm ? C0[2 * FC_COUNT + FC_BM1] : 0.0, g2 = use_pm ? C0[2 * FC_COUNT + FC_LGC] : 0.0;
#pragma unroll 1
            for (int i = nS + lane + L * part; i < nk; i += L * nparts) {
                const double k = Ctx::ldg(kh + i), lk = Ctx::ldg(logk + i);
                const bool over = k * dz > CB_COSH_OVERFLOW;
                const double lin = fma(-2.0 * k, zt, Cc);
                const double p0 = over ? INFINITY : lin - b0 * lk + g0, pp = over ? INFINITY : lin - b1 * lk + g1;
                const double pm = over ? INFINITY : lin - b2 * lk + g2;
                store[i] = p0; store[nk + i] = pp; store[2 * nk + i] = pm; store[3 * nk + i] = over ? NAN : 0.0;
            }
        }

        // ---- order-only tables of the nv vectors, built side by side (the product scans overlap) -----------------
        if (nS > nT && part == 0) {
            if (L == 1) {
                for (int u = 0; u < nv; ++u) {
                    const double a1 = COEF[u * FC_COUNT + FC_A1];
                    double *CT = CTg + u * (CB_MILLER_NMAX + 1);
                    double c = a1;
                    CT[0] = 1.0; CT[1] = a1;
                    for (int n = 2; n <= CB_MILLER_NMAX; ++n) { c *= fma(a1, Ctx::ldg(rn + n), (double)(n - 1)); CT[n] = c; }
                }
            } else {
                double a1v[3], q[3][3], incl[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    a1v[u] = C0[(u < nv ? u : 0) * FC_COUNT + FC_A1];
                    double run = 1.0;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int n = 3 * lane + 2 + j;
                        const double g = n <= CB_MILLER_NMAX ? fma(a1v[u], Ctx::ldg(rn + n), (double)(n - 1)) : 1.0;
                        run *= g;
                        q[u][j] = run;
                    }
                    incl[u] = run;
                }
#pragma unroll 1
                for (int d = 1; d < L; d <<= 1) {
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const double t = cx.shfl_up(incl[u], d);
                        if (lane >= d) incl[u] *= t;
                    }
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    double excl = cx.shfl_up(incl[u], 1);
                    if (lane == 0) excl = 1.0;
                    const double base = a1v[u] * excl;
                    if (u < nv) {
                        double *CT = CTg + u * (CB_MILLER_NMAX + 1);
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const int n = 3 * lane + 2 + j;
                            if (n <= CB_MILLER_NMAX) CT[n] = base * q[u][j];
                        }
                        if (lane == 0) { CT[0] = 1.0; CT[1] = a1v[u]; }
                    }
                }
            }
        }
        if (nT > 0 && part == 0) {
#pragma unroll 1
            for (int j = 1 + lane; j < CB_TEMME_TAB; j += L) {
#pragma unroll 1
                for (int u = 0; u < nv; ++u) {
                    const double mu = C0[u * FC_COUNT + FC_MU];
                    RPQg[u * CB_TEMME_TAB + j] = make_double2(1.0 / ((double)j - mu), 1.0 / ((double)j + mu));
                }
            }
        }
        if (nparts > 1) cx.sync_block();
        else cx.sync_warp();

        // ---- chunks of P = L / nv consecutive points x nv vectors: lane = point * nv + vector ---------------------
        const int VL = L >= nv ? nv : 1;              // vectors side by side in the lanes (1 under serial emulation)
        const int P = L / VL;
        const int il = lane / VL, vl = lane - il * VL;
        const auto RNJ = Ctx::shared_view(RNJg);
#pragma unroll 1
        for (int c0 = part * P; c0 < nS; c0 += nparts * P) {
            const int c1 = c0 + P < nS ? c0 + P : nS;
            const int i = c0 + il;
            const bool act = il < P && i < c1;
            const bool temme = i < nT;
#pragma unroll 1
            for (int u = vl; u < nv; u += VL) {
                const auto CV = Ctx::shared_view((const double *)(COEF + u * FC_COUNT));
                const double mu = CV[FC_MU], a1 = CV[FC_A1];
                double a = 0.0, b = 0.0;              // K_mu, K_mu+1 (Temme: unscaled; Miller: times e^x)
                if (temme) {
                    if (act) {
                        const double d = -(Ctx::ldg(logk + i) + lhdz);      // ln(2/x)
                        const double sg = mu * d;
                        const double ee = exp(sg), ei = 1.0 / ee;
                        const double shd = fabs(sg) < 1e-4 ? d * (1.0 + sg * sg * (1.0 / 6.0)) : 0.5 * (ee - ei) * CV[FC_RMU];
                        double ff = CV[FC_SINC] * (CV[FC_GAM1] * 0.5 * (ee + ei) + CV[FC_GAM2] * shd);
                        double p = CV[FC_HP] * ee, q = CV[FC_HQ] * ei, c = 1.0;
                        const double x2 = 0.5 * Ctx::ldg(kh + i) * dz, dd = x2 * x2;
                        double sum = ff, sum1 = p;
                        const auto RPQ = Ctx::shared_view((const double2 *)(RPQg + u * CB_TEMME_TAB));
#pragma unroll 1
                        for (int j = 1; j < CB_TEMME_TAB; ++j) {
                            const double fj = (double)j;
                            const double2 r = RPQ[j];
                            ff = (fj * ff + p + q) * (r.x * r.y);
                            c *= dd * RNJ[j];
                            p *= r.x;
                            q *= r.y;
                            const double del = c * ff;
                            sum += del;
                            sum1 = fma(c, p - fj * ff, sum1);
                            if (fabs(del) < fabs(sum) * CB_KNU_EPS) break;
                        }
                        a = sum;
                        b = sum1 * (2.0 * Ctx::ldg(rk + i) * rdz);
                    }
                } else if (c1 > nT) {
                    // Miller backward recurrence, trip count uniform over the chunk (from its smallest Miller x)
                    const int ii = act ? i : c1 - 1;
                    const double xi = Ctx::ldg(kh + ii) * dz;
                    int N = (int)(12.0 + 140.0 * (Ctx::ldg(rk + (c0 > nT ? c0 : nT)) * rdz)) + 1;
                    if (N > CB_MILLER_NMAX - 1) N = CB_MILLER_NMAX - 1;
                    // one 8-byte shared load per step (the weight C_n); a_{n+1} = -(a1 + n(n+1)) and b_n = 2(n + x) are
                    // updated in registers: tabulating a_{n+1} beside C_n (one 16-byte load, two additions less per step)
                    // was measured 11-13 % slower with and without spills -- the loop is bound by shared-memory loads
                    const auto CT = Ctx::shared_view((const double *)(CTg + u * (CB_MILLER_NMAX + 1)));
                    double yp = 0.0, y = 1.0, S = CT[N];
                    double tn = 2.0 * (double)N;                              // 2n
                    double an1 = -(a1 + (double)N * (double)(N + 1));         // a_{n+1}
                    const double twox = 2.0 * xi;
                    double bn = tn + twox;                                    // b_n = 2(n + x)
#pragma unroll 2
                    for (int n = N; n >= 1; --n) {
                        const double cm = CT[n - 1];
                        const double ym = fma(bn, y, an1 * yp);
                        yp = y;
                        y = ym;
                        S = fma(cm, y, S);
                        an1 += tn;                                            // a_n = a_{n+1} + 2n
                        tn -= 2.0;
                        bn = tn + twox;
                    }
                    if (act) {
                        const double invx = Ctx::ldg(rk + ii) * rdz;
                        const double sc = cb_pow2_inverse_of(S);
                        const double cS = Ctx::ldg(sq + ii) * CV[FC_RSDZ] / (S * sc);
                        const double ys = y * sc, yps = yp * sc;
                        a = cS * ys;
                        b = cS * fma(mu + xi + 0.5, ys, -a1 * yps) * invx;
                    }
                }
                // ---- F: scale to K e^-x, raise the order, bracket e^-x = Gamma(nu)/4 (1 + e^-2x) - hx (K_nu e^-x), log
                if (act) {
                    const double k = Ctx::ldg(kh + i), lk = Ctx::ldg(logk + i), xi = k * dz;
                    const double ex = exp(temme ? -xi : -2.0 * xi);
                    const double e2 = temme ? ex * ex : ex;
                    a *= ex;
                    b *= ex;
                    const double tox = 2.0 * Ctx::ldg(rk + i) * rdz;
                    const int nl = (int)CV[FC_NL];
#pragma unroll 1
                    for (int j = 1; j < nl; ++j) {
                        const double c = fma((mu + (double)j) * tox, b, a);
                        a = b;
                        b = c;
                    }
                    const double hx = exp(CV[FC_NU] * (lk + lhdz));
                    const double qg = CV[FC_QG];
                    const double br = fma(qg, 1.0 + e2, -b * hx);
                    store[u * nk + i] = fma(-2.0 * k, zt, Cc) - CV[FC_BM1] * lk + cb_log_inline(CV[FC_LEAD] * br);
                    if (u == 0) store[3 * nk + i] = k * fma(hx, a + b, -2.0 * qg * e2) / br;
                }
            }
        }
    }

    (void)use_pm;
}

// 0.5 sum (p0 - obs)^2 over the slab (every lane returns the total).
template <class Ctx>
CB_BODY double slab_half_ssq(const Ctx &cx, const WarpSlab &W, int lane) {
    double acc = 0.0;
#pragma unroll 1
    for (int i = lane; i < W.nk; i += cx.team_lanes()) {
        const double r = W.p0[i] - W.obs[i];
        acc = fma(r, r, acc);
    }
    return 0.5 * cx.warp_sum(acc);
}

// min_func for the staged spectrum (optimise.py:220-230).
template <class Ctx>
CB_BODY double misfit_slab(const Ctx &cx, const WarpSlab &W, const double *x, const Priors &pr, int lane) {
    model_coefs_warp(cx, W, x, 0.0, 0, 1, lane);
    const double f = model_eval_warp<true>(cx, W, 0, lane, W.p0, nullptr, W.obs);
    if (!isfinite(f)) return CB_BIG_MISFIT;
    return f + prior_penalty(pr, x);
}

// min_func when only "is it <= bound?" matters: the return value is exact when it is <= bound, and some
// value > bound (a partial sum, or 1e99 for a non-finite model) otherwise.
template <class Ctx>
CB_BODY double misfit_slab_bounded(const Ctx &cx, const WarpSlab &W, const double *x, const Priors &pr, int lane,
                                   double bound) {
    const double pen = prior_penalty(pr, x);
    model_coefs_warp(cx, W, x, 0.0, 0, 1, lane);
    const double f = model_eval_warp<true>(cx, W, 0, lane, W.p0, nullptr, W.obs, bound - pen);
    if (!isfinite(f)) return CB_BIG_MISFIT;
    return f + pen;
}
