// Warp-cooperative evaluation of the Bouligand model over all wavenumbers of one window.
// This is the hot loop of the fit and MCMC kernels (pycurious/optimise.py:223 executed ~70 x per fit,
// 101,000 x per chain).  Compared with calling bouligand_point() per lane it
//
//   * splits the (ascending) wavenumbers into three contiguous classes by x = k dz, so that a
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
//   * hoists everything that depends on the order only into per-evaluation tables in shared
//     memory, built lane-parallel: C_n (prefix product by warp shuffles), 1/(i-mu), 1/(i+mu), 1/i
//
// Falls back to the per-point generic path (bouligand.cuh) for unsorted k, dz <= 0, beta < 0,
// very large orders or non-finite parameters, where the reference's nan/inf pattern matters more
// than speed.
#pragma once
#include "bouligand.cuh"

#define CB_MILLER_NMAX 97
#define CB_TEMME_TAB 28
#define CB_SKIP_LOG 43.0                      /* > 60 ln 2 + ln(4 sqrt(pi/4)) */
#define CB_COSH_OVERFLOW 710.4758600739439    /* np.cosh overflows above this */

struct WarpSlab {          // views into one warp's shared-memory slab
    double *k, *logk, *obs;        // [nk] wavenumber, ln k, observed Phi
    double *p0, *pp, *pm, *dz;     // [nk] model values (p0) and, for the fit, Phi(beta+-h), dPhi/ddz
    double *C;                     // [CB_MILLER_NMAX + 1]
    double *rp, *rq, *ri;          // [CB_TEMME_TAB]
    int nk;
    bool sorted;
};

CB_HD int warp_slab_doubles(int nk, bool fit) {
    return (fit ? 7 : 4) * nk + (CB_MILLER_NMAX + 1) + 3 * CB_TEMME_TAB;
}

CB_HD WarpSlab warp_slab_carve(double *base, int nk, bool fit) {
    WarpSlab W;
    W.nk = nk;
    W.k = base; W.logk = base + nk; W.obs = base + 2 * nk; W.p0 = base + 3 * nk;
    double *t = base + 4 * nk;
    if (fit) { W.pp = t; W.pm = t + nk; W.dz = t + 2 * nk; t += 3 * nk; }
    else { W.pp = W.pm = W.dz = nullptr; }
    W.C = t; W.rp = t + CB_MILLER_NMAX + 1; W.rq = W.rp + CB_TEMME_TAB; W.ri = W.rq + CB_TEMME_TAB;
    W.sorted = false;
    return W;
}

// Stage k, ln k and the observed spectrum; detect ascending k.
template <class Ctx>
CB_BODY void warp_slab_load(const Ctx &cx, WarpSlab &W, const double *kh, const double *phi_row, int lane) {
    const int L = cx.team_lanes();
    bool asc = true;
    for (int i = lane; i < W.nk; i += L) {
        const double k = kh[i];
        W.k[i] = k;
        W.logk[i] = log(k);
        W.obs[i] = phi_row[i];
        if (i + 1 < W.nk && !(kh[i + 1] > k)) asc = false;
        if (!(k > 0.0)) asc = false;
    }
    W.sorted = cx.all(asc);
    cx.sync_warp();
}

// K_nu, K_nu-1 from K_mu, K_mu+1 (nl >= 1), then the model value and d/ddz.
CB_HD double finish_point(const ModelCoef &m, double k, double logk, double x, double invx, double emx,
                          double kmu, double kmu1, double *ddz) {
    double a = kmu, b = kmu1;
    const double tox = 2.0 * invx;
    for (int j = 1; j < m.ord.nl; ++j) {
        const double c = fma((m.ord.mu + (double)j) * tox, b, a);
        a = b;
        b = c;
    }
    const double hx = exp(m.nu * (logk + m.lhalfdz));
    const double ex = 1.0 / emx;
    const double ch = 0.5 * (ex + emx), sh = 0.5 * (ex - emx);
    const double bracket = 0.5 * ch * m.gnu - b * hx;
    if (ddz) *ddz = -k + k * (0.5 * sh * m.gnu + hx * a) / bracket;
    return m.C - 2.0 * k * m.zt - (m.beta - 1.0) * logk - x + log(m.lead * bracket);
}

// Phi(k_i) for every wavenumber of the slab -> out_phi[i] (and dPhi/ddz -> out_ddz[i] if not NULL).
template <class Ctx>
CB_BODY void model_eval_warp(const Ctx &cx, const WarpSlab &W, const double *x, int lane, double *out_phi,
                             double *out_ddz) {
    const ModelCoef m = model_coef(x[0], x[1], x[2], x[3]);
    const int nk = W.nk, L = cx.team_lanes();
    const bool fast = W.sorted && m.dz > 0.0 && m.dz < 1e300 && m.nu >= 0.5 && m.nu <= 8.5 &&
                      isfinite(m.zt) && isfinite(m.C) && m.gnu > 0.0 && m.lead > 0.0 && isfinite(m.gnu) &&
                      isfinite(m.lead);
    if (!fast) {                       // generic per-point path, reference nan/inf semantics
        for (int i = lane; i < nk; i += L) {
            if (out_ddz) out_phi[i] = bouligand_point_ddz(m, W.k[i], W.logk[i], out_ddz[i]);
            else out_phi[i] = bouligand_point(m, W.k[i], W.logk[i]);
        }
        cx.sync_warp();
        return;
    }
    const double mu = m.ord.mu, a1 = 0.25 - mu * mu;

    // ---- classes: [0,nT) Temme, [nT,nS) Miller, [nS,nk) K negligible ---------------------------
    const double lgnu = log(m.gnu);
    int lastT = -1, lastC = -1;
    for (int i = lane; i < nk; i += L) {
        const double xi = W.k[i] * m.dz;
        if (xi <= 2.0) lastT = i;
        const bool skip = xi > 2.0 && (2.0 * xi - m.nu * (W.logk[i] + m.lhalfdz) + lgnu > CB_SKIP_LOG);
        if (!skip) lastC = i;
    }
    const int nT = cx.warp_max_i(lastT) + 1;
    int nS = cx.warp_max_i(lastC) + 1;
    if (nS < nT) nS = nT;

    // ---- order-only tables, built lane-parallel ----------------------------------------------------
    if (nS > nT) {                     // C_0 = 1, C_1 = a1, C_n = C_{n-1} (a1 + n(n-1)) / n
        if (lane == 0) { W.C[0] = 1.0; W.C[1] = a1; }
        double carry = a1;
        for (int n0 = 2; n0 <= CB_MILLER_NMAX; n0 += L) {
            const int n = n0 + lane;
            double g = (n <= CB_MILLER_NMAX) ? (a1 + (double)n * (double)(n - 1)) / (double)n : 1.0;
            for (int d = 1; d < L; d <<= 1) {
                const double t = cx.shfl_up(g, d);
                if (lane >= d) g *= t;
            }
            g *= carry;
            if (n <= CB_MILLER_NMAX) W.C[n] = g;
            carry = cx.shfl(g, L - 1);
        }
    }
    if (nT > 0) {
        for (int i = 1 + lane; i < CB_TEMME_TAB; i += L) {
            W.rp[i] = 1.0 / ((double)i - mu);
            W.rq[i] = 1.0 / ((double)i + mu);
            W.ri[i] = 1.0 / (double)i;
        }
    }
    cx.sync_warp();

    // ---- x <= 2: Temme's series ---------------------------------------------------------------
    const double hp = 0.5 / m.ord.gampl, hq = 0.5 / m.ord.gammi;
    for (int i = lane; i < nT; i += L) {
        const double k = W.k[i], lk = W.logk[i], xi = k * m.dz;
        const double d = -(lk + m.lhalfdz);             // ln(2/x)
        const double sg = mu * d;
        const double e = exp(sg), ei = 1.0 / e;
        const double shc = fabs(sg) < 1e-4 ? 1.0 + sg * sg * (1.0 / 6.0) : 0.5 * (e - ei) / sg;
        double ff = m.ord.sinc * (m.ord.gam1 * 0.5 * (e + ei) + m.ord.gam2 * shc * d);
        double p = hp * e, q = hq * ei, c = 1.0;
        const double x2 = 0.5 * xi, dd = x2 * x2;
        double sum = ff, sum1 = p;
        for (int j = 1; j < CB_TEMME_TAB; ++j) {
            const double fj = (double)j, rpj = W.rp[j], rqj = W.rq[j];
            ff = (fj * ff + p + q) * (rpj * rqj);
            c *= dd * W.ri[j];
            p *= rpj;
            q *= rqj;
            const double del = c * ff;
            sum += del;
            sum1 = fma(c, p - fj * ff, sum1);
            if (fabs(del) < fabs(sum) * CB_KNU_EPS) break;
        }
        const double invx = 1.0 / xi;
        out_phi[i] = finish_point(m, k, lk, xi, invx, exp(-xi), sum, sum1 * (2.0 * invx), out_ddz ? out_ddz + i : nullptr);
    }

    // ---- 2 < x < x_skip: Miller backward recurrence, uniform trip count per pass -------------------
    for (int base = nT; base < nS; base += L) {
        const int i = base + lane;
        const bool act = i < nS;
        const int ii = act ? i : nS - 1;
        const double k = W.k[ii], lk = W.logk[ii], xi = k * m.dz;
        const double xmin = W.k[base] * m.dz;          // smallest x of this pass (k ascending)
        int N = (int)(12.0 + 140.0 / xmin) + 1;
        if (N > CB_MILLER_NMAX - 1) N = CB_MILLER_NMAX - 1;
        double yp = 0.0, y = 1.0, S = W.C[N];
        double bn = 2.0 * ((double)N + xi);
        double an1 = -(a1 + (double)(N + 1) * (double)N);
        for (int n = N; n >= 1; --n) {
            const double ym = fma(bn, y, an1 * yp);
            yp = y;
            y = ym;
            S = fma(W.C[n - 1], y, S);
            bn -= 2.0;
            an1 += 2.0 * (double)n;
        }
        const double invx = 1.0 / xi, emx = exp(-xi);
        const double kmu = sqrt(CB_PI * 0.5 * invx) * emx * y / S;
        const double kmu1 = kmu * (mu + xi + 0.5 - a1 * yp / y) * invx;
        double dv;
        const double ph = finish_point(m, k, lk, xi, invx, emx, kmu, kmu1, out_ddz ? &dv : nullptr);
        if (act) {
            out_phi[i] = ph;
            if (out_ddz) out_ddz[i] = dv;
        }
    }

    // ---- x >= x_skip: the Bessel term cannot change the rounded bracket ----------------------------
    const double lgc = log(m.lead * m.gnu * 0.25);
    for (int i = nS + lane; i < nk; i += L) {
        const double k = W.k[i];
        const bool over = k * m.dz > CB_COSH_OVERFLOW;
        out_phi[i] = over ? INFINITY : fma(-2.0 * k, m.zt, m.C) - (m.beta - 1.0) * W.logk[i] + lgc;
        if (out_ddz) out_ddz[i] = over ? NAN : 0.0;
    }
    cx.sync_warp();
}

// min_func for the staged spectrum (optimise.py:220-230).
template <class Ctx>
CB_BODY double misfit_slab(const Ctx &cx, const WarpSlab &W, const double *x, const Priors &pr, int lane) {
    model_eval_warp(cx, W, x, lane, W.p0, nullptr);
    double acc = 0.0;
    for (int i = lane; i < W.nk; i += cx.team_lanes()) {
        const double r = W.p0[i] - W.obs[i];
        acc = fma(r, r, acc);
    }
    cx.sync_warp();
    const double f = 0.5 * cx.warp_sum(acc);
    if (!isfinite(f)) return CB_BIG_MISFIT;
    return f + prior_penalty(pr, x);
}
