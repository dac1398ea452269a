// Batched MAP fit of (beta, zt, dz, C), one window per warp.
// Replaces CurieOptimise.optimise / optimise_routine (pycurious/optimise.py:232-301, :446-509),
// i.e. scipy.optimize.minimize(min_func, x0, bounds=self.bounds) = L-BFGS-B with forward
// differences, by a bounded quasi-Newton method on the same objective:
//
//   * lanes own wavenumbers; one evaluation gives the misfit, its gradient and the Gauss-Newton
//     matrix J^T J in one pass (warp-shuffle reductions of 15 scalars)
//   * dPhi/dC = 1, dPhi/dzt = -2k and dPhi/ddz (via K_{nu-1}) are analytic; dPhi/dbeta is a
//     central difference of the model in beta (no closed form for dK_nu/dnu)
//   * projected BFGS: the Hessian model starts from J^T J + prior precision and receives BFGS
//     secant updates; variables sitting on a bound with the gradient pointing outward are frozen
//     (two-metric projection), the line search is a projected Armijo backtracking
//   * non-finite model values give 1e99 exactly like min_func (optimise.py:226-227)
//
// The reference's stopping point is set by its own noisy forward-difference gradient (factr 1e7,
// pgtol 1e-5); this solver converges to the same bounded minimum more tightly, so on windows where
// the reference is converged the parameters agree to ~1e-5 relative (tests/test_fit_parity.py).
#pragma once
#include "model_eval.cuh"

struct FitParams {
    const double *kh;       // [nk]
    int32_t nk;
    const double *phi;      // [nwin, nk]
    int64_t nwin;
    cb_fit_options opt;
    const double *x0_dev;   // [nwin,4] or NULL
    double *x_out;          // [nwin,4]
    double *f_out;          // [nwin] or NULL
    int32_t *niter;         // [nwin] or NULL
    int32_t *status;        // [nwin] or NULL
    int32_t *nfev;          // [nwin] or NULL: model evaluations (one = nk forward-model points)
    int32_t warps_per_block;
};

enum { FIT_CONVERGED_PG = 0, FIT_CONVERGED_F = 1, FIT_LINESEARCH = 2, FIT_MAXITER = 3, FIT_BAD_START = 4 };

struct FitEval {
    double f;        // objective incl. priors (1e99 when non-finite)
    double g[4];
    double H[10];    // packed upper triangle of J^T J + prior precision: (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)(1,3)(2,2)(2,3)(3,3)
    bool ok;
};

CB_HD int tri(int i, int j) {   // i <= j
    return i * 4 - (i * (i - 1)) / 2 + (j - i);
}

template <class Ctx>
CB_BODY void fit_eval_full(const Ctx &cx, const WarpSlab &W, const double *x, const Priors &pr, int lane,
                           FitEval &ev) {
    const double hb = 1e-6 * fmax(1.0, fabs(x[0]));
    double xs[4] = {x[0], x[1], x[2], x[3]};
    model_eval_warp(cx, W, xs, lane, W.p0, W.dz);
    xs[0] = x[0] + hb;
    model_eval_warp(cx, W, xs, lane, W.pp, nullptr);
    xs[0] = x[0] - hb;
    model_eval_warp(cx, W, xs, lane, W.pm, nullptr);
    double acc[15];
#pragma unroll
    for (int q = 0; q < 15; ++q) acc[q] = 0.0;
    for (int i = lane; i < W.nk; i += cx.team_lanes()) {
        const double p0 = W.p0[i], pp = W.pp[i], pm = W.pm[i];
        double J[4];
        if (isfinite(pp) && isfinite(pm)) J[0] = (pp - pm) / (2.0 * hb);
        else if (isfinite(pp)) J[0] = (pp - p0) / hb;
        else if (isfinite(pm)) J[0] = (p0 - pm) / hb;
        else J[0] = 0.0;
        J[1] = -2.0 * W.k[i];
        J[2] = W.dz[i];
        J[3] = 1.0;
        const double r = p0 - W.obs[i];
        acc[0] = fma(r, r, acc[0]);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            acc[1 + a] = fma(r, J[a], acc[1 + a]);
#pragma unroll
            for (int b = a; b < 4; ++b) acc[5 + tri(a, b)] = fma(J[a], J[b], acc[5 + tri(a, b)]);
        }
    }
    cx.sync_warp();
#pragma unroll
    for (int q = 0; q < 15; ++q) acc[q] = cx.warp_sum(acc[q]);
    double f = 0.5 * acc[0];
    bool ok = isfinite(f);
#pragma unroll
    for (int q = 1; q < 15; ++q) ok = ok && isfinite(acc[q]);
    ev.ok = ok;
    if (!ok) {
        ev.f = CB_BIG_MISFIT;
#pragma unroll
        for (int a = 0; a < 4; ++a) ev.g[a] = 0.0;
#pragma unroll
        for (int q = 0; q < 10; ++q) ev.H[q] = 0.0;
        return;
    }
    ev.f = f + prior_penalty(pr, x);
#pragma unroll
    for (int a = 0; a < 4; ++a) ev.g[a] = acc[1 + a];
#pragma unroll
    for (int q = 0; q < 10; ++q) ev.H[q] = acc[5 + q];
#pragma unroll
    for (int a = 0; a < 4; ++a)
        if (!isnan(pr.mu[a])) {
            const double w = 1.0 / (pr.sigma[a] * pr.sigma[a]);
            ev.g[a] += (x[a] - pr.mu[a]) * w;
            ev.H[tri(a, a)] += w;
        }
}

// Solve B d = -g on the free variables (frozen ones get d = 0) by Cholesky; returns false when the
// reduced matrix is not positive definite.
CB_HD bool solve_free(const double *B /*packed*/, const double *g, const bool *frozen, double *d) {
    double A[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double v = B[i <= j ? tri(i, j) : tri(j, i)];
            A[i][j] = (frozen[i] || frozen[j]) ? (i == j ? 1.0 : 0.0) : v;
        }
    double L[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) L[i][j] = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double s = A[j][j];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < j) s -= L[j][k] * L[j][k];
        if (!(s > 0.0) || !isfinite(s)) return false;
        const double ljj = sqrt(s);
        L[j][j] = ljj;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i > j) {
                double t = A[i][j];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < j) t -= L[i][k] * L[j][k];
                L[i][j] = t / ljj;
            }
    }
    double y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double t = frozen[i] ? 0.0 : -g[i];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < i) t -= L[i][k] * y[k];
        y[i] = t / L[i][i];
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        double t = y[i];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k > i) t -= L[k][i] * d[k];
        d[i] = t / L[i][i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (frozen[i]) d[i] = 0.0;
    return true;
}

CB_HD double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

template <class Ctx>
CB_BODY void fit_body(const Ctx &cx, const FitParams &P) {
    const int tid = cx.tid();
    const int warp = tid >> 5, lane = tid & 31;
    const int nk = P.nk;
    const int64_t w = (int64_t)cx.bid() * P.warps_per_block + warp;
    if (w >= P.nwin) return;     // whole warp leaves together

    // per-warp slab in shared memory: k, ln k, Phi, model values and the order tables
    WarpSlab W = warp_slab_carve(reinterpret_cast<double *>(cx.smem()) + (int64_t)warp * warp_slab_doubles(nk, true),
                                 nk, true);
    warp_slab_load(cx, W, P.kh, P.phi + w * nk, lane);

    const cb_fit_options &o = P.opt;
    Priors pr;
    double lb[4], ub[4], x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        pr.mu[j] = o.prior_mu[j];
        pr.sigma[j] = o.prior_sigma[j];
        lb[j] = o.lb[j];
        ub[j] = o.ub[j];
        const double s = P.x0_dev ? P.x0_dev[w * 4 + j] : o.x0[j];
        x[j] = clampd(s, lb[j], ub[j]);     // scipy clips x0 into the bounds (_lbfgsb_py.py:359)
    }
    const int max_iter = o.max_iter > 0 ? o.max_iter : 200;
    const double gtol = o.gtol > 0.0 ? o.gtol : 1e-9;
    const double ftol = o.ftol > 0.0 ? o.ftol : 1e-15;

    FitEval ev;
    fit_eval_full(cx, W, x, pr, lane, ev);
    int status = FIT_MAXITER, it = 0, nfev = 3;
    if (!ev.ok) {
        status = FIT_BAD_START;
    } else {
        double B[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) B[q] = ev.H[q];
        bool fresh = true;      // B is the Gauss-Newton matrix of the current point
        int stall = 0;
        for (it = 0; it < max_iter; ++it) {
            bool frozen[4];
            double pgn = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                frozen[j] = (x[j] <= lb[j] && ev.g[j] > 0.0) || (x[j] >= ub[j] && ev.g[j] < 0.0);
                if (!frozen[j]) pgn = fmax(pgn, fabs(ev.g[j]));
            }
            if (pgn <= gtol) { status = FIT_CONVERGED_PG; break; }
#ifdef CB_FIT_TRACE
            if (lane == 0) printf("it %3d f %.15g pgn %.3e x %.6f %.6f %.6f %.6f fresh %d nfev %d\n", it, ev.f, pgn, x[0], x[1], x[2], x[3], (int)fresh, nfev);
#endif

            double d[4];
            bool have = solve_free(B, ev.g, frozen, d);
            double gd = 0.0;
            if (have) {
#pragma unroll
                for (int j = 0; j < 4; ++j) gd += ev.g[j] * d[j];
                have = (gd < 0.0) && isfinite(gd);
            }
            if (!have && !fresh) {       // fall back to the Gauss-Newton metric of this point
#pragma unroll
                for (int q = 0; q < 10; ++q) B[q] = ev.H[q];
                fresh = true;
                have = solve_free(B, ev.g, frozen, d);
                gd = 0.0;
                if (have) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) gd += ev.g[j] * d[j];
                    have = (gd < 0.0) && isfinite(gd);
                }
            }
            if (!have) {                  // steepest descent, scaled by the diagonal
                gd = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double hjj = fmax(ev.H[tri(j, j)], 1e-12);
                    d[j] = frozen[j] ? 0.0 : -ev.g[j] / hjj;
                    gd += ev.g[j] * d[j];
                }
            }

            // Newton decrement: the model predicts a decrease of -gd/2; below the rounding noise of
            // the objective no line search can confirm progress, so this is convergence.
            if (-gd <= 1e-13 * fmax(1.0, fabs(ev.f))) { status = FIT_CONVERGED_F; break; }

            // projected Armijo backtracking
            double alpha = 1.0, xt[4], ft = CB_BIG_MISFIT;
            bool accepted = false;
            for (int ls = 0; ls < 40; ++ls) {
                double dec = 0.0, move = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    xt[j] = clampd(x[j] + alpha * d[j], lb[j], ub[j]);
                    dec += ev.g[j] * (xt[j] - x[j]);
                    move = fmax(move, fabs(xt[j] - x[j]));
                }
                if (move == 0.0) break;
                ft = misfit_slab(cx, W, xt, pr, lane);
                ++nfev;
                if (ft <= ev.f + 1e-4 * dec && ft < 1e98) { accepted = true; break; }
                if (ft >= 1e98) alpha *= 0.25;
                else {
                    // minimiser of the quadratic through f(0), f'(0), f(alpha)
                    const double den = 2.0 * (ft - ev.f - dec);
                    double an = den > 0.0 ? -dec * alpha / den : 0.5 * alpha;
                    alpha = clampd(an, 0.1 * alpha, 0.5 * alpha);
                }
            }
            if (!accepted) {
                if (!fresh) {             // retry once from the Gauss-Newton metric
#pragma unroll
                    for (int q = 0; q < 10; ++q) B[q] = ev.H[q];
                    fresh = true;
                    continue;
                }
                status = FIT_LINESEARCH;
                break;
            }

            FitEval en;
            fit_eval_full(cx, W, xt, pr, lane, en);
            nfev += 3;
            if (!en.ok) { status = FIT_LINESEARCH; break; }
            double s[4], y[4], sy = 0.0, ss = 0.0, yy = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[j] = xt[j] - x[j];
                y[j] = en.g[j] - ev.g[j];
                sy += s[j] * y[j];
                ss += s[j] * s[j];
                yy += y[j] * y[j];
            }
            // BFGS secant update, skipped (reset to Gauss-Newton) when the curvature condition fails
            bool updated = false;
            if (sy > 1e-10 * sqrt(ss * yy) && sy > 0.0) {
                double Bs[4], sBs = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double t = 0.0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) t += B[i <= j ? tri(i, j) : tri(j, i)] * s[j];
                    Bs[i] = t;
                    sBs += t * s[i];
                }
                if (sBs > 0.0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = i; j < 4; ++j)
                            B[tri(i, j)] += y[i] * y[j] / sy - Bs[i] * Bs[j] / sBs;
                    updated = true;
                    fresh = false;
                }
            }
            if (!updated) {
#pragma unroll
                for (int q = 0; q < 10; ++q) B[q] = en.H[q];
                fresh = true;
            }
            const double fold = ev.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = xt[j];
            ev = en;
            const double rel = (fold - ev.f) / fmax(fmax(fabs(fold), fabs(ev.f)), 1.0);
            stall = (rel <= ftol) ? stall + 1 : 0;
            if (stall >= 3) { status = FIT_CONVERGED_F; ++it; break; }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) P.x_out[w * 4 + j] = x[j];
        if (P.f_out) P.f_out[w] = ev.f;
        if (P.niter) P.niter[w] = it;
        if (P.status) P.status[w] = status;
        if (P.nfev) P.nfev[w] = nfev;
    }
}
