// Batched MAP fit of (beta, zt, dz, C), one window per warp.
// Replaces CurieOptimise.optimise / optimise_routine (pycurious/optimise.py:232-301, :446-509),
// i.e. scipy.optimize.minimize(min_func, x0, bounds=self.bounds) = L-BFGS-B with forward
// differences, by a bounded quasi-Newton method on the same objective:
//
//   * lanes own wavenumbers; one evaluation gives the misfit, its gradient and the Gauss-Newton
//     matrix J^T J in one pass (warp-shuffle reductions of 15 scalars); the 4x4 algebra between
//     evaluations is done by lane 0 on a small state block in shared memory
//   * dPhi/dC = 1, dPhi/dzt = -2k and dPhi/ddz (via K_{nu-1}) are analytic; dPhi/dbeta is a
//     central difference of the model in beta (no closed form for dK_nu/dnu)
//   * projected BFGS: the Hessian model starts from J^T J + prior precision and receives BFGS
//     secant updates; variables sitting on a bound with the gradient pointing outward are frozen
//     (two-metric projection), the line search is a projected Armijo backtracking whose accepted
//     trial evaluation is reused as the next iterate's model values
//   * dz is additionally bounded by the cosh-overflow wall 710.4758/k_max, beyond which min_func is
//     1e99: windows whose misfit valley runs flat towards large dz end on that wall instead of
//     creeping up to it through rejected steps
//   * non-finite model values give 1e99 exactly like min_func (optimise.py:226-227)
//
// The reference's stopping point is set by its own noisy forward-difference gradient (factr 1e7,
// pgtol 1e-5); this solver converges to the same bounded minimum more tightly, so on windows where
// the reference is converged the parameters agree to ~1e-5 relative (tests/test_gpu_parity.py::_check_fit,
// tests/test_gpu_config3.py; reference self-sensitivity on the rest: tests/golden/make_golden_cfg3.py).
#pragma once
#include "model_eval.cuh"

struct FitParams {
    const double *kh;       // [nk]
    const double *ktab;     // [ktable_doubles(nk)] window-independent tables (ktable_entry)
    int32_t nk;
    const double *phi;      // [nwin, nk]
    int64_t nwin;
    cb_fit_options opt;
    const double *x0_dev;   // [nwin,4] or NULL
    const double *prior_mu_dev;   // [nwin,4] or NULL: per-row prior means (sensitivity), NaN = opt.prior_mu[j]
    double *x_out;          // [nwin,4]
    double *f_out;          // [nwin] or NULL
    int32_t *niter;         // [nwin] or NULL
    int32_t *status;        // [nwin] or NULL
    int32_t *nfev;          // [nwin] or NULL: model evaluations (one = nk forward-model points)
    int32_t warps_per_block;
};

enum { FIT_CONVERGED_PG = 0, FIT_CONVERGED_F = 1, FIT_LINESEARCH = 2, FIT_MAXITER = 3, FIT_BAD_START = 4 };

CB_HD int tri(int i, int j) {   // i <= j
    return i * 4 - (i * (i - 1)) / 2 + (j - i);
}

// Solver state of one window, in the warp's shared-memory slab.  The 4x4 algebra is done by lane 0
// (the other lanes wait at the next warp barrier); the model evaluations are done by all lanes.
struct FitState {
    double x[4], lb[4], ub[4];
    double f, g[4], H[10];        // current iterate: objective, gradient, Gauss-Newton matrix (+ prior precision)
    double B[10];                 // quasi-Newton metric (packed upper triangle)
    double d[4], gd;              // search direction and g.d
    double xt[4], dec, move, pen_t, hb;   // trial point, g.(xt-x), max |xt-x|, prior penalty at xt, beta step
    double fn, gn[4], Hn[10];     // full evaluation at the accepted trial point
    double pmu[4], pw[4];         // prior mean and inverse variance (0 = no prior)
    double one;                   // 1.0 (the dPhi/dC column of the Jacobian, read with stride 0)
    double pt[4];                 // physical parameters (beta, zt, dz, C) of the trial point
    double cu, sdz;               // solver coordinate u = exp(-cu dz) (cu = 0: u = dz); sdz = d dz / d u at the trial point
    int32_t okn, fresh, stall, status, action, it, nfev, mask;
    int32_t central;              // derivative evaluations use the central difference in beta from now on
    int32_t gcentral;             // the gradient of the current iterate was taken with the central difference
    int32_t use_pm;               // the evaluation being reduced staged Phi(beta - hb) (else forward difference)
    int32_t buf_at_x;             // p0 / dPhi/ddz / Phi(beta + hb) in the slab belong to the current iterate
};
#define CB_FIT_STATE_DOUBLES ((int)((sizeof(FitState) + 7) / 8))
enum { FIT_GO = 0, FIT_DONE = 1, FIT_REFINE = 2 };
// dPhi/dbeta is a difference quotient in the order of K_nu (no closed form).  Far from the minimum the
// forward difference (one extra evaluation) is accurate enough -- its O(h) error is ~1e-6 of the
// gradient; once the projected gradient falls below this fraction of max(1, f) the solver re-takes the
// gradient of the current iterate with the central difference and keeps using it, so every convergence
// decision is made on central-difference gradients.
#define CB_FIT_CENTRAL_BELOW 1.0e-3
CB_HD int fit_slab_doubles(int nk) { return warp_slab_doubles(nk, true, CB_FIT_STATE_DOUBLES); }

// Solve B d = -g on the free variables (frozen ones get d = 0) by Cholesky; returns false when the
// reduced matrix is not positive definite.
CB_HD_NOINLINE bool solve_free(const double *B /*packed*/, const double *g, const bool *frozen, double *d) {
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
        const double ljj = sqrt(s), rjj = 1.0 / ljj;
        L[j][j] = rjj;                 // the diagonal holds the reciprocal
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i > j) {
                double t = A[i][j];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < j) t -= L[i][k] * L[j][k];
                L[i][j] = t * rjj;
            }
    }
    double y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double t = frozen[i] ? 0.0 : -g[i];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < i) t -= L[i][k] * y[k];
        y[i] = t * L[i][i];
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        double t = y[i];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k > i) t -= L[k][i] * d[k];
        d[i] = t * L[i][i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (frozen[i]) d[i] = 0.0;
    return true;
}

CB_HD double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

CB_HD double fit_prior_penalty(const FitState *S, const double *x) {
    double c = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double dl = x[j] - S->pmu[j];
        c = fma(0.5 * dl * dl, S->pw[j], c);
    }
    return c;
}

// Solver coordinates.  The misfit valley of this model runs flat towards large dz like
// f_inf + A exp(-2 k_0 dz) (only the lowest wavenumbers still feel the base of the layer): in dz a
// Newton-type step on that tail has constant length 1/(2 k_0) and 30 % of the windows of a fractal
// grid crawl for dozens of iterations.  In u = exp(-2 k_0 dz) the tail is linear, the Gauss-Newton
// model is exact there and the step lands on the bound (the cosh wall) at once.  Bounds, priors and
// results stay in dz; only the iteration runs in (beta, zt, u, C).
CB_HD void fit_to_physical(FitState *S, const double *y, double *p) {
    p[0] = y[0]; p[1] = y[1]; p[3] = y[3];
    p[2] = S->cu > 0.0 ? -cb_log(y[2]) / S->cu : y[2];
}

CB_HD void fit_set_trial(FitState *S) {
    fit_to_physical(S, S->xt, S->pt);
    S->sdz = S->cu > 0.0 ? -1.0 / (S->cu * S->xt[2]) : 1.0;
    S->hb = 1e-6 * fmax(1.0, fabs(S->pt[0]));
    S->pen_t = fit_prior_penalty(S, S->pt);
}

// ---- lane-0 phases ------------------------------------------------------------------------------
// Start: clip x0 into the bounds (scipy does, _lbfgsb_py.py:359) and into the cosh wall.
CB_NOINLINE_BODY void fit_init(FitState *S, const cb_fit_options *o, const double *x0, const double *row_mu,
                               double kmin, double kmax) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        S->lb[j] = o->lb[j];
        S->ub[j] = o->ub[j];
        const bool has = !isnan(o->prior_mu[j]);           // a prior exists iff the options carry one
        const double mu = (row_mu && !isnan(row_mu[j])) ? row_mu[j] : o->prior_mu[j];
        S->pmu[j] = has ? mu : 0.0;
        S->pw[j] = has ? 1.0 / (o->prior_sigma[j] * o->prior_sigma[j]) : 0.0;
    }
    if (kmax > 0.0) {
        double wall = CB_COSH_OVERFLOW / kmax;
        while (kmax * wall > CB_COSH_OVERFLOW) wall = nextafter(wall, 0.0);
        S->ub[2] = fmin(S->ub[2], wall);
    }
    double p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = clampd(x0[j], S->lb[j], S->ub[j]);
    // u = exp(-cu dz) needs a finite upper bound on dz (the wall) and ascending positive k
    S->cu = (kmin > 0.0 && kmax > 0.0 && S->ub[2] < 1e300 && !(o->flags & CB_FIT_FLAG_LINEAR_DZ)) ? 2.0 * kmin : 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) S->xt[j] = p[j];
    if (S->cu > 0.0) {
        const double dlo = S->lb[2], dhi = S->ub[2];
        double ulo = exp(-S->cu * dhi);                       // dz = dhi  <->  smallest u
        while (-log(ulo) / S->cu > dhi) ulo = nextafter(ulo, 1.0);
        double uhi = dlo > -1e300 ? exp(-S->cu * dlo) : (double)INFINITY;
        if (uhi < (double)INFINITY) while (-log(uhi) / S->cu < dlo) uhi = nextafter(uhi, 0.0);
        S->lb[2] = ulo;
        S->ub[2] = uhi;
        S->xt[2] = clampd(exp(-S->cu * p[2]), ulo, uhi);
    }
    fit_set_trial(S);
    S->fresh = 1; S->stall = 0; S->status = FIT_MAXITER; S->action = FIT_GO; S->it = 0; S->nfev = 0; S->mask = 0;
    S->central = 0; S->gcentral = 0; S->use_pm = 0; S->buf_at_x = 0;
}

// Make the evaluated trial point the current iterate (first evaluation: B = Gauss-Newton metric).
CB_NOINLINE_BODY void fit_accept_first(FitState *S) {
    if (!S->okn) { S->status = FIT_BAD_START; S->action = FIT_DONE; }
    S->gcentral = S->use_pm;
    S->buf_at_x = 1;
    S->f = S->fn;
#pragma unroll
    for (int j = 0; j < 4; ++j) { S->x[j] = S->xt[j]; S->g[j] = S->gn[j]; }
#pragma unroll
    for (int q = 0; q < 10; ++q) { S->H[q] = S->Hn[q]; S->B[q] = S->Hn[q]; }
}

// Convergence test on the projected gradient, then the quasi-Newton direction on the free variables.
CB_NOINLINE_BODY void fit_direction(FitState *S, double gtol) {
    bool frozen[4];
    double pgn = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        frozen[j] = (S->x[j] <= S->lb[j] && S->g[j] > 0.0) || (S->x[j] >= S->ub[j] && S->g[j] < 0.0);
        // no slope and no curvature: the objective does not depend on this variable here (dz far out on
        // the flat tail, where the Bessel term is below the rounding of the bracket) -- leave it alone
        if (S->g[j] == 0.0 && S->H[tri(j, j)] == 0.0) frozen[j] = true;
        if (!frozen[j]) pgn = fmax(pgn, fabs(S->g[j]));
    }
    if (!S->gcentral && pgn <= CB_FIT_CENTRAL_BELOW * fmax(1.0, fabs(S->f))) {
        S->central = 1;
        S->action = FIT_REFINE;          // re-take this iterate's gradient with the central difference first
        return;
    }
    if (pgn <= gtol) { S->status = FIT_CONVERGED_PG; S->action = FIT_DONE; return; }
    // the secant pairs were collected on another face of the box: when the active set changes, restart
    // the metric from the Gauss-Newton matrix of this point
    int mask = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) mask |= frozen[j] ? (1 << j) : 0;
    if (mask != S->mask) {
        S->mask = mask;
        if (!S->fresh) {
#pragma unroll
            for (int q = 0; q < 10; ++q) S->B[q] = S->H[q];
            S->fresh = 1;
        }
    }
#ifdef CB_FIT_TRACE
    {
        double pp_[4];
        fit_to_physical(S, S->x, pp_);
        printf("it %3d f %.15g pgn %.3e x %.6f %.6f %.6f %.6f fresh %d nfev %d\n", S->it, S->f, pgn, pp_[0], pp_[1],
               pp_[2], pp_[3], S->fresh, S->nfev);
    }
#endif
    // Direction from the quasi-Newton metric; a metric that fails (not positive definite, not a
    // descent direction) or that predicts no measurable decrease is replaced ONCE by the
    // Gauss-Newton matrix of this point before either conclusion is drawn.
    double d[4], gd = 0.0;
    bool have = false;
    for (int attempt = 0; attempt < 2; ++attempt) {
        have = solve_free(S->B, S->g, frozen, d);
        gd = 0.0;
        if (have) {
#pragma unroll
            for (int j = 0; j < 4; ++j) gd += S->g[j] * d[j];
            have = (gd < 0.0) && isfinite(gd);
        }
        // Newton decrement: the model predicts a decrease of -gd/2; below the rounding noise of the
        // objective no line search can confirm progress
        const bool tiny = have && -gd <= 1e-13 * fmax(1.0, fabs(S->f));
        if ((have && !tiny) || S->fresh) break;
#pragma unroll
        for (int q = 0; q < 10; ++q) S->B[q] = S->H[q];
        S->fresh = 1;
    }
    if (!have) {                      // steepest descent, scaled by the diagonal
        gd = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double hjj = fmax(S->H[tri(j, j)], 1e-12);
            d[j] = frozen[j] ? 0.0 : -S->g[j] / hjj;
            gd += S->g[j] * d[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) S->d[j] = d[j];
    S->gd = gd;
    if (!S->gcentral && -gd <= 1e-13 * fmax(1.0, fabs(S->f))) { S->central = 1; S->action = FIT_REFINE; return; }
#ifdef CB_FIT_TRACE
    printf("      g %.3e %.3e %.3e %.3e  d %.3e %.3e %.3e %.3e  gd %.3e mask %d Hdiag %.3e %.3e %.3e %.3e Bdiag %.3e %.3e %.3e %.3e\n", S->g[0], S->g[1], S->g[2], S->g[3], d[0], d[1], d[2], d[3], gd, S->mask,
           S->H[0], S->H[4], S->H[7], S->H[9], S->B[0], S->B[4], S->B[7], S->B[9]);
#endif
    if (-gd <= 1e-13 * fmax(1.0, fabs(S->f))) { S->status = FIT_CONVERGED_F; S->action = FIT_DONE; }
}

CB_NOINLINE_BODY void fit_trial(FitState *S, double alpha) {
    double dec = 0.0, move = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double t = clampd(S->x[j] + alpha * S->d[j], S->lb[j], S->ub[j]);
        S->xt[j] = t;
        dec += S->g[j] * (t - S->x[j]);
        move = fmax(move, fabs(t - S->x[j]));
    }
    S->dec = dec;
    S->move = move;
    fit_set_trial(S);
}

// The line search failed: retry once from the Gauss-Newton metric, else give up.
CB_NOINLINE_BODY void fit_linesearch_failed(FitState *S) {
    if (!S->fresh) {
#pragma unroll
        for (int q = 0; q < 10; ++q) S->B[q] = S->H[q];
        S->fresh = 1;
    } else {
        S->status = FIT_LINESEARCH;
        S->action = FIT_DONE;
    }
}

// BFGS secant update (skipped, i.e. reset to Gauss-Newton, when the curvature condition fails), move
// to the accepted point, stall test on the relative decrease.
CB_NOINLINE_BODY void fit_update(FitState *S, double ftol, int max_iter) {
    if (!S->okn) { S->status = FIT_LINESEARCH; S->action = FIT_DONE; return; }
    double s[4], y[4], sy = 0.0, ss = 0.0, yy = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s[j] = S->xt[j] - S->x[j];
        y[j] = S->gn[j] - S->g[j];
        sy += s[j] * y[j];
        ss += s[j] * s[j];
        yy += y[j] * y[j];
    }
    bool updated = false;
    if (sy > 1e-10 * sqrt(ss * yy) && sy > 0.0) {
        double Bs[4], sBs = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) t += S->B[i <= j ? tri(i, j) : tri(j, i)] * s[j];
            Bs[i] = t;
            sBs += t * s[i];
        }
        if (sBs > 0.0) {
            const double rsy = 1.0 / sy, rsBs = 1.0 / sBs;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = i; j < 4; ++j) S->B[tri(i, j)] += y[i] * y[j] * rsy - Bs[i] * Bs[j] * rsBs;
            updated = true;
            S->fresh = 0;
        }
    }
    if (!updated) {
#pragma unroll
        for (int q = 0; q < 10; ++q) S->B[q] = S->Hn[q];
        S->fresh = 1;
    }
    S->gcentral = S->use_pm;
    S->buf_at_x = 1;
    const double fold = S->f;
    S->f = S->fn;
#pragma unroll
    for (int j = 0; j < 4; ++j) { S->x[j] = S->xt[j]; S->g[j] = S->gn[j]; }
#pragma unroll
    for (int q = 0; q < 10; ++q) S->H[q] = S->Hn[q];
    S->it += 1;
    const double rel = (fold - S->f) / fmax(fmax(fabs(fold), fabs(S->f)), 1.0);
    S->stall = (rel <= ftol) ? S->stall + 1 : 0;
    if (S->stall >= 3) { S->status = FIT_CONVERGED_F; S->action = FIT_DONE; }
    else if (S->it >= max_iter) { S->status = FIT_MAXITER; S->action = FIT_DONE; }
}

// ---- all-lane phases ----------------------------------------------------------------------------
// Objective, gradient and Gauss-Newton matrix at S->xt from the staged model values
// (p0, dPhi/ddz at xt; pp, pm at beta +- hb).  The 15 sums are sum_i u_a[i] u_b[i] over the five
// per-point vectors u = (r, dPhi/dbeta, k, dPhi/ddz, 1) (dPhi/dzt = -2k is applied as a scale).
// Pass 1 stages r and the central difference in beta, lane-parallel over the points; in pass 2
// lane q < 15 (and lane 16+q for the second half of the points) owns ONE of the 15 products and
// walks the points with broadcast-free strided reads -- no 15-wide butterfly, ~100 instructions.
// The arrays arrive as one pointer (obs: the arrays obs, p0, kb, pp, pm, dz, res follow each other nk apart) and
// the k table, not as a WarpSlab reference: see model_eval_core.
struct FitSlabView { const double *k; double *obs, *p0, *kb, *pp, *pm, *dz, *res; int nk; };

template <class Ctx>
CB_NOINLINE_BODY void fit_reduce_core(const Ctx &cx, const double *kh, double *obs0, int nk_, FitState *S, int lane) {
    FitSlabView W;
    W.k = kh; W.nk = nk_;
    W.obs = obs0; W.p0 = obs0 + nk_; W.kb = obs0 + 2 * nk_; W.pp = obs0 + 3 * nk_; W.pm = obs0 + 4 * nk_;
    W.dz = obs0 + 5 * nk_; W.res = W.pm;
    const int L = cx.team_lanes(), nk = W.nk;
    const double hb = S->hb;
    const double rh = 1.0 / hb, rh2 = 0.5 * rh;
    const bool use_pm = S->use_pm != 0;
#pragma unroll 1
    for (int i = lane; i < nk; i += L) {
        const double p0 = W.p0[i], pp = W.pp[i], pm = use_pm ? W.pm[i] : 0.0;
        const bool fp = isfinite(pp), fm = use_pm && isfinite(pm);
        W.kb[i] = (fp && fm) ? (pp - pm) * rh2 : (fp ? (pp - p0) * rh : (fm ? (p0 - pm) * rh : 0.0));
        W.res[i] = p0 - W.obs[i];             // Phi(beta - hb) is consumed: its slot holds the residuals for pass 2
    }
    if (lane == 0) S->one = 1.0;
    cx.sync_warp();
    const bool split = (L == 32);
    const int half = split ? (lane >> 4) : 0;
    const int lo = half ? (nk + 1) / 2 : 0, hi = split ? (half ? nk : (nk + 1) / 2) : nk;
    const int qstep = split ? 16 : L;
#pragma unroll 1
    for (int q = split ? (lane & 15) : lane; q < 16; q += qstep) {
        // q = 0: r.r;  1..4: r.J_a;  5..14: J_a.J_b packed upper triangle;  15: idle (keeps the shuffle convergent)
        int va = 0, vb = 0;
        if (q >= 1 && q < 5) vb = q;
        else if (q >= 5) {
            const int t = q - 5;
            const int a = t < 4 ? 0 : (t < 7 ? 1 : (t < 9 ? 2 : 3));
            va = a + 1;
            vb = t - (a * 4 - (a * (a - 1)) / 2) + a + 1;
        }
        const double *A = va == 0 ? W.res : (va == 1 ? W.kb : (va == 2 ? W.k : (va == 3 ? W.dz : &S->one)));
        const double *B = vb == 0 ? W.res : (vb == 1 ? W.kb : (vb == 2 ? W.k : (vb == 3 ? W.dz : &S->one)));
        const int sa = va == 4 ? 0 : 1, sb = vb == 4 ? 0 : 1;
        double acc = 0.0;
        if (q < 15) {
#pragma unroll 4
            for (int i = lo; i < hi; ++i) acc = fma(A[i * sa], B[i * sb], acc);
        }
        if (split) acc += cx.shfl_xor(acc, 16);
        acc *= (va == 2 ? -2.0 : 1.0) * (vb == 2 ? -2.0 : 1.0);
        if (half == 0 && q < 15) {
            if (q == 0) S->fn = 0.5 * acc;
            else if (q < 5) S->gn[q - 1] = acc;
            else S->Hn[q - 5] = acc;
        }
    }
    cx.sync_warp();
    if (lane == 0) {
        bool ok = isfinite(S->fn);
#pragma unroll 1
        for (int a = 0; a < 4; ++a) ok = ok && isfinite(S->gn[a]);
#pragma unroll 1
        for (int q = 0; q < 10; ++q) ok = ok && isfinite(S->Hn[q]);
        S->okn = ok ? 1 : 0;
        if (!ok) {
            S->fn = CB_BIG_MISFIT;
#pragma unroll 1
            for (int a = 0; a < 4; ++a) S->gn[a] = 0.0;
#pragma unroll 1
            for (int q = 0; q < 10; ++q) S->Hn[q] = 0.0;
        } else {
            S->fn += S->pen_t;
            // chain rule for the solver coordinate of dz, then the priors (Gaussian in the physical parameters)
            const double sd = S->sdz;
            S->gn[2] *= sd;
            S->Hn[tri(0, 2)] *= sd; S->Hn[tri(1, 2)] *= sd; S->Hn[tri(2, 3)] *= sd;
            S->Hn[tri(2, 2)] *= sd * sd;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double ja = a == 2 ? sd : 1.0;
                S->gn[a] = fma((S->pt[a] - S->pmu[a]) * ja, S->pw[a], S->gn[a]);
                S->Hn[tri(a, a)] = fma(ja * ja, S->pw[a], S->Hn[tri(a, a)]);
            }
        }
    }
    cx.sync_warp();
}

template <class Ctx>
CB_BODY void fit_reduce_full(const Ctx &cx, const WarpSlab &W, FitState *S, int lane) {
    fit_reduce_core(cx, W.k, W.obs, W.nk, S, lane);
}

template <class Ctx>
CB_BODY int fit_poll(const Ctx &cx, const FitState *S) {
    cx.sync_warp();
    const int a = S->action;
    cx.sync_warp();
    return a;
}

// Lane 0: the re-taken gradient replaces the iterate's (same point, same f up to rounding).
CB_NOINLINE_BODY void fit_refined(FitState *S) {
    if (!S->okn) { S->status = FIT_LINESEARCH; S->action = FIT_DONE; return; }
    S->f = S->fn;
#pragma unroll
    for (int j = 0; j < 4; ++j) S->g[j] = S->gn[j];
#pragma unroll
    for (int q = 0; q < 10; ++q) S->H[q] = S->Hn[q];
    S->gcentral = 1;
    S->buf_at_x = 1;
    S->action = FIT_GO;
}

// Lane 0: make the current iterate the trial point again (refinement after a failed line search).
CB_NOINLINE_BODY void fit_trial_at_x(FitState *S) {
#pragma unroll
    for (int j = 0; j < 4; ++j) S->xt[j] = S->x[j];
    fit_set_trial(S);
}

template <class Ctx>
CB_BODY void fit_body(const Ctx &cx, const FitParams &P) {
    const int tid = cx.tid();
    const int warp = tid >> 5, lane = tid & 31;
    const int nk = P.nk;
    const int64_t w = (int64_t)cx.bid() * P.warps_per_block + warp;
    if (w >= P.nwin) return;     // whole warp leaves together

    // per-warp slab in shared memory: wavenumber tables, model values, order tables, solver state
    WarpSlab W = warp_slab_carve(reinterpret_cast<double *>(cx.smem()) +
                                     (int64_t)warp * fit_slab_doubles(nk), nk, true, P.kh, P.ktab);
    warp_slab_load(cx, W, P.phi + w * nk, lane);
    FitState *S = reinterpret_cast<FitState *>(W.scratch);
    const cb_fit_options *o = &P.opt;
    const int max_iter = o->max_iter > 0 ? o->max_iter : 200;
    const double gtol = o->gtol > 0.0 ? o->gtol : 1e-9;
    const double ftol = o->ftol > 0.0 ? o->ftol : 1e-15;

    if (lane == 0)
        fit_init(S, o, P.x0_dev ? P.x0_dev + w * 4 : o->x0, P.prior_mu_dev ? P.prior_mu_dev + w * 4 : nullptr,
                 W.sorted ? W.kmin : 0.0, W.sorted ? W.kmax : 0.0);
    cx.sync_warp();
    W.cu = S->cu;

    // first evaluation: forward difference in beta (see CB_FIT_CENTRAL_BELOW)
    model_coefs_warp(cx, W, S->pt, S->hb, 0, 2, lane);
    model_eval_warp<false>(cx, W, 0, lane, W.p0, W.dz);
    model_eval_warp<false>(cx, W, 1, lane, W.pp, nullptr);
    fit_reduce_full(cx, W, S, lane);
    if (lane == 0) { fit_accept_first(S); S->nfev = 2; }

    // S->action is written by lane 0 only; every read is fenced by a barrier on both sides so that
    // lane 0 cannot change it again before all lanes have seen the same value
    while (fit_poll(cx, S) == FIT_GO) {
        if (lane == 0) fit_direction(S, gtol);
        const int act = fit_poll(cx, S);
        if (act == FIT_DONE) break;
        if (act == FIT_REFINE) {
            int extra = 1;
            if (S->buf_at_x) {            // Phi, dPhi/ddz and Phi(beta + hb) of this iterate are still staged
                model_coefs_warp(cx, W, S->pt, S->hb, 2, 1, lane);
            } else {
                cx.sync_warp();
                if (lane == 0) fit_trial_at_x(S);
                cx.sync_warp();
                model_coefs_warp(cx, W, S->pt, S->hb, 0, 3, lane);
                model_eval_warp<false>(cx, W, 0, lane, W.p0, W.dz);
                model_eval_warp<false>(cx, W, 1, lane, W.pp, nullptr);
                extra = 3;
            }
            model_eval_warp<false>(cx, W, 2, lane, W.pm, nullptr);
            if (lane == 0) S->use_pm = 1;
            cx.sync_warp();
            fit_reduce_full(cx, W, S, lane);
            if (lane == 0) { S->nfev += extra; fit_refined(S); }
            continue;
        }

        // projected Armijo backtracking
        const double f0 = S->f;
        double alpha = 1.0;
        bool accepted = false;
        int nls = 0;
        for (int ls = 0; ls < 40; ++ls) {
            if (lane == 0) { fit_trial(S, alpha); S->buf_at_x = 0; }
            cx.sync_warp();
            const double dec = S->dec;
            if (S->move == 0.0) break;
            // coefficients of the trial point AND of its beta +- hb neighbours in one call (one lane each, same
            // instruction count as the trial alone): the derivative evaluations of an accepted trial need no second pass
            model_coefs_warp(cx, W, S->pt, S->hb, 0, S->central ? 3 : 2, lane);
            model_eval_warp<false>(cx, W, 0, lane, W.p0, W.dz);
            ++nls;
            double ft = slab_half_ssq(cx, W, lane);
            ft = isfinite(ft) ? ft + S->pen_t : CB_BIG_MISFIT;
            if (ft <= f0 + 1e-4 * dec && ft < 1e98) { accepted = true; break; }
            if (ft >= 1e98) alpha *= 0.25;
            else {
                // minimiser of the quadratic through f(0), f'(0), f(alpha)
                const double den = 2.0 * (ft - f0 - dec);
                const double an = den > 0.0 ? -dec * alpha / den : 0.5 * alpha;
                alpha = clampd(an, 0.1 * alpha, 0.5 * alpha);
            }
            cx.sync_warp();
        }
        if (!accepted) {
            cx.sync_warp();
            if (lane == 0) { S->nfev += nls; fit_linesearch_failed(S); }
            continue;
        }
        // derivatives at the accepted point: beta + hb (and beta - hb once the central difference is on);
        // p0 and dPhi/ddz are already staged
        const int central = S->central;
        model_eval_warp<false>(cx, W, 1, lane, W.pp, nullptr);
        if (central) model_eval_warp<false>(cx, W, 2, lane, W.pm, nullptr);
        if (lane == 0) S->use_pm = central;
        cx.sync_warp();
        fit_reduce_full(cx, W, S, lane);
        if (lane == 0) { S->nfev += nls + (central ? 2 : 1); fit_update(S, ftol, max_iter); }
    }
    if (lane == 0) {
        double pf[4];
        fit_to_physical(S, S->x, pf);
#pragma unroll
        for (int j = 0; j < 4; ++j) P.x_out[w * 4 + j] = pf[j];
        if (P.f_out) P.f_out[w] = S->f;
        if (P.niter) P.niter[w] = S->it;
        if (P.status) P.status[w] = S->status;
        if (P.nfev) P.nfev[w] = S->nfev;
    }
}
