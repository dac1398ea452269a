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
// Execution structure (round 2).  A window advances in ROUNDS; every round is
//     fit_eval_values : ONE fused evaluation of (trial, beta+h [, beta-h]) over (point, vector) pairs by a whole
//                       warp, ending in the 15 sums of a Gauss-Newton step (model_eval.cuh) -- the derivative
//                       evaluations of a trial are done speculatively WITH the trial (the first trial of a line
//                       search is accepted ~90 % of the time);
//     fit_step_window : the scalar part -- Armijo test, then [BFGS update, next direction and trial] or
//                       [backtracked trial], and the order-only coefficients of the next evaluation -- by ONE
//                       THREAD per window.
// The two halves are separate kernels over the list of unfinished windows (fit_eval_kernel: one warp per window;
// fit_step_kernel: one thread per window, 32 windows per warp), the solver state of a window living in global
// memory between them (~1.5 KB per window and round: 1 % of the run time).  Round 1 ran both halves in one
// persistent warp per window: the scalar half then executes on one lane of 32 and streams ~30 KB of straight-line
// code through a 32 KB instruction cache that 24 independent warps share (ncu: 40 % of the stall samples were
// instruction fetch, 18 of 32 lanes active).  Windows that need more rounds than the batched phase runs are
// finished by fit_finish_kernel, which applies the same two functions in a per-warp loop.
//
// The reference's stopping point is set by its own noisy forward-difference gradient (factr 1e7,
// pgtol 1e-5); this solver converges to the same bounded minimum more tightly, so on windows where
// the reference is converged the parameters agree to ~1e-5 relative (tests/test_gpu_parity.py::_check_fit,
// tests/test_gpu_config3.py; reference self-sensitivity on the rest: tests/golden/make_golden_cfg3.py).
#pragma once
#include <stddef.h>
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
    // scratch of the batched rounds (library-owned, see cb_fit_batch_rows): per window the solver state, the
    // coefficients of its next evaluation and the sums of its last one; two lists of unfinished windows
    double *state;          // [CB_FIT_PERSIST_DOUBLES, stride]: field-major, so that the threads of the step kernel
                            //      (one window each) read and write whole lines
    int64_t stride;         // windows the scratch arrays are laid out for (>= nwin)
    double *tmpl;           // [CB_FIT_STATE_DOUBLES]: the initial state for opt.x0 and opt.prior_mu, built once per
                            //      launch (fit_template_body); carries the launch's constants (bounds in the solver
                            //      coordinate, cosh wall, prior weights)
    double *coef;           // [nwin, 3 FC_COUNT]
    double *sums;           // [nwin, 16]
    int32_t *list_in, *list_out;       // unfinished windows of this round / of the next one
    unsigned int *count;               // [6]: list lengths of the rounds r, r+1, r+2 (rotating), then the
                                       //      work counters of the evaluation kernels of those rounds
    int32_t round;
    int32_t handover;       // the batched rounds stop once fewer windows than this are unfinished: fit_finish_body
                            // takes them, a CTA per window (few windows make a round latency bound)
};

enum { FIT_CONVERGED_PG = 0, FIT_CONVERGED_F = 1, FIT_LINESEARCH = 2, FIT_MAXITER = 3, FIT_BAD_START = 4 };

CB_HD int tri(int i, int j) {   // i <= j
    return i * 4 - (i * (i - 1)) / 2 + (j - i);
}

// Solver state of one window, in the warp's shared-memory slab.  The 4x4 algebra is done by lane 0
// (the other lanes wait at the next warp barrier); the model evaluations are done by all lanes.
struct FitState {
    // ---- what a window carries from round to round (saved in the scratch arrays, CB_FIT_PERSIST_DOUBLES doubles) ----
    double x[4];
    double f, g[4], H[10];        // current iterate: objective, gradient, Gauss-Newton matrix (+ prior precision)
    double B[10];                 // quasi-Newton metric (packed upper triangle)
    double d[4], gd;              // search direction and g.d
    double xt[4], dec, move, pen_t, hb;   // trial point, g.(xt-x), max |xt-x|, prior penalty at xt, beta step
    double pt[4];                 // physical parameters (beta, zt, dz, C) of the trial point
    double sdz;                   // d dz / d u at the trial point (solver coordinate u = exp(-cu dz); cu = 0: u = dz)
    double alpha;                 // step length of the staged trial
    int32_t fresh, stall, status, action, it, nfev, mask;
    int32_t central;              // derivative evaluations use the central difference in beta from now on
    int32_t gcentral;             // the gradient of the current iterate was taken with the central difference
    int32_t use_pm;               // the evaluation being reduced staged Phi(beta - hb) (else forward difference)
    int32_t kind;                 // what the staged trial point is: first point, line-search trial, refinement
    int32_t nv;                   // vectors of the round's fused evaluation: 2 (x, beta+hb) or 3 (+ beta-hb)
    int32_t nls;                  // trials of the current line search so far
    int32_t pad_;
    // ---- constants of the launch (pmu: of the window when per-row prior means are given) ----
    double lb[4], ub[4];
    double pmu[4], pw[4];         // prior mean and inverse variance (0 = no prior)
    double cu;
    // ---- transients of one solver step ----
    double fn, gn[4], Hn[10];     // full evaluation at the accepted trial point
    int32_t okn, pad2_;
};
#define CB_FIT_PERSIST_DOUBLES ((int)(offsetof(FitState, lb) / 8))
enum { FIT_KIND_FIRST = 0, FIT_KIND_TRIAL = 1, FIT_KIND_REFINE = 2 };
#define CB_FIT_STATE_DOUBLES ((int)((sizeof(FitState) + 7) / 8))
enum { FIT_GO = 0, FIT_DONE = 1, FIT_REFINE = 2 };
// dPhi/dbeta is a difference quotient in the order of K_nu (no closed form).  Far from the minimum the
// forward difference (one extra evaluation) is accurate enough -- its O(h) error is ~1e-6 of the
// gradient; once the projected gradient falls below this fraction of max(1, f) the solver re-takes the
// gradient of the current iterate with the central difference and keeps using it, so every convergence
// decision is made on central-difference gradients.
#define CB_FIT_CENTRAL_BELOW 1.0e-3
// per warp: the order tables and the model values of one evaluation (4 nk doubles)
CB_HD int fit_eval_warp_doubles(int nk) { return (CB_FIT_TABLE_DOUBLES + 4 * nk + 1) & ~1; }
CB_HD size_t fit_eval_smem_bytes(int warps, int nk) { return (size_t)warps * fit_eval_warp_doubles(nk) * sizeof(double); }
static_assert(CB_FIT_TABLE_DOUBLES % 2 == 0 && CB_FIT_CT_DOUBLES % 2 == 0, "double2 tables must stay 16-byte aligned");

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
    S->central = (o->flags & CB_FIT_FLAG_FORWARD_FIRST) ? 0 : 1; S->gcentral = 0; S->use_pm = 0;
    S->kind = FIT_KIND_FIRST; S->nv = S->central ? 3 : 2; S->nls = 0; S->alpha = 1.0;
}

// Make the evaluated trial point the current iterate (first evaluation: B = Gauss-Newton metric).
CB_NOINLINE_BODY void fit_accept_first(FitState *S) {
    if (!S->okn) { S->status = FIT_BAD_START; S->action = FIT_DONE; }
    S->gcentral = S->use_pm;
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

// Objective, gradient and Gauss-Newton matrix at S->xt from the 15 sums of fit_sums_from_values (model_eval.cuh).
CB_NOINLINE_BODY void fit_take_sums(FitState *S, const double *sums) {
    S->fn = 0.5 * sums[0];
    bool ok = isfinite(S->fn);
#pragma unroll
    for (int a = 0; a < 4; ++a) { S->gn[a] = sums[1 + a]; ok = ok && isfinite(S->gn[a]); }
#pragma unroll
    for (int q = 0; q < 10; ++q) { S->Hn[q] = sums[5 + q]; ok = ok && isfinite(S->Hn[q]); }
    S->okn = ok ? 1 : 0;
    if (!ok) {
        S->fn = CB_BIG_MISFIT;
#pragma unroll
        for (int a = 0; a < 4; ++a) S->gn[a] = 0.0;
#pragma unroll
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

// Lane 0: the re-taken gradient replaces the iterate's (same point, same f up to rounding).
CB_NOINLINE_BODY void fit_refined(FitState *S) {
    if (!S->okn) { S->status = FIT_LINESEARCH; S->action = FIT_DONE; return; }
    S->f = S->fn;
#pragma unroll
    for (int j = 0; j < 4; ++j) S->g[j] = S->gn[j];
#pragma unroll
    for (int q = 0; q < 10; ++q) S->H[q] = S->Hn[q];
    S->gcentral = 1;
    S->action = FIT_GO;
}

// Lane 0: make the current iterate the trial point again (refinement: its gradient is re-taken with the
// central difference).
CB_NOINLINE_BODY void fit_trial_at_x(FitState *S) {
#pragma unroll
    for (int j = 0; j < 4; ++j) S->xt[j] = S->x[j];
    fit_set_trial(S);
}

// Lane 0, after an accepted point: next direction and its first trial (or the refinement point, or the end).
CB_NOINLINE_BODY void fit_next_step(FitState *S, double gtol) {
    for (int guard = 0; guard < 4 && S->action == FIT_GO; ++guard) {
        fit_direction(S, gtol);
        if (S->action == FIT_DONE) return;
        if (S->action == FIT_REFINE) {
            fit_trial_at_x(S);
            S->kind = FIT_KIND_REFINE; S->nv = 3; S->action = FIT_GO;
            return;
        }
        S->alpha = 1.0; S->nls = 0;
        fit_trial(S, 1.0);
        S->kind = FIT_KIND_TRIAL; S->nv = S->central ? 3 : 2;
        if (S->move != 0.0) return;
        fit_linesearch_failed(S);        // no move possible: once more from the Gauss-Newton metric, else give up
    }
    if (S->action == FIT_GO) { S->status = FIT_LINESEARCH; S->action = FIT_DONE; }
}

// Lane 0, after a rejected trial (ft = its objective value): shorter step along the same direction.
CB_NOINLINE_BODY void fit_backtrack(FitState *S, double ft, double gtol) {
    const double f0 = S->f, dec = S->dec;
    double alpha = S->alpha;
    if (ft >= 1e98) alpha *= 0.25;
    else {
        // minimiser of the quadratic through f(0), f'(0), f(alpha)
        const double den = 2.0 * (ft - f0 - dec);
        const double an = den > 0.0 ? -dec * alpha / den : 0.5 * alpha;
        alpha = clampd(an, 0.1 * alpha, 0.5 * alpha);
    }
    S->alpha = alpha;
    S->nls += 1;
    if (S->nls < 40) {
        fit_trial(S, alpha);
        if (S->move != 0.0) return;
    }
    fit_linesearch_failed(S);
    if (S->action == FIT_GO) fit_next_step(S, gtol);
}

// Coefficients of the staged trial point for the next evaluation (nv vectors) -> c[3 FC_COUNT].
CB_NOINLINE_BODY void fit_stage_coefs(const FitState *S, double *c) {
    const int nv = S->nv;
    for (int v = 0; v < nv; ++v) fit_coef_fill(c + v * FC_COUNT, S->pt, S->hb, v);
    c[FC_AUX] = S->hb;
    c[FC_COUNT + FC_AUX] = S->cu;
    c[2 * FC_COUNT + FC_AUX] = (double)nv;
}

// One thread: the scalar half of a round.  `sums` = the 15 sums of the evaluation of the staged trial point.
// Returns true when the window is finished; otherwise a new trial point and its coefficients are staged.
CB_BODY bool fit_step_window(FitState *S, const double *sums, double gtol, double ftol, int max_iter, double *coef) {
    const int kind = S->kind, nv = S->nv;
    S->nfev += nv;
    bool take = true;
    double ft = 0.0;
    if (kind == FIT_KIND_TRIAL) {                        // projected Armijo test
        ft = 0.5 * sums[0];
        ft = isfinite(ft) ? ft + S->pen_t : CB_BIG_MISFIT;
        take = ft <= S->f + 1e-4 * S->dec && ft < 1e98;
    }
    if (take) {
        S->use_pm = nv == 3;
        fit_take_sums(S, sums);
        if (kind == FIT_KIND_FIRST) fit_accept_first(S);
        else if (kind == FIT_KIND_TRIAL) fit_update(S, ftol, max_iter);
        else fit_refined(S);
        if (S->action == FIT_GO) fit_next_step(S, gtol);
    } else {
        fit_backtrack(S, ft, gtol);
    }
    if (S->action != FIT_GO) return true;
    if (coef) fit_stage_coefs(S, coef);
    return false;
}

CB_BODY void fit_write_result(const FitParams &P, FitState *S, int64_t w) {
    double pf[4];
    fit_to_physical(S, S->x, pf);
#pragma unroll
    for (int j = 0; j < 4; ++j) P.x_out[w * 4 + j] = pf[j];
    if (P.f_out) P.f_out[w] = S->f;
    if (P.niter) P.niter[w] = S->it;
    if (P.status) P.status[w] = S->status;
    if (P.nfev) P.nfev[w] = S->nfev;
}

CB_BODY void fit_solver_constants(const FitParams &P, int &max_iter, double &gtol, double &ftol) {
    const cb_fit_options *o = &P.opt;
    max_iter = o->max_iter > 0 ? o->max_iter : 200;
    gtol = o->gtol > 0.0 ? o->gtol : 1e-9;
    ftol = o->ftol > 0.0 ? o->ftol : 1e-15;
}

// ---- kernels' bodies ----------------------------------------------------------------------------------------
// The per-window state in the scratch arrays is field-major; a thread (or the lanes of a warp) moves one window's
// persistent part between it and a FitState.
CB_BODY void fit_state_load(const FitParams &P, int64_t w, FitState *S, int first, int step) {
    double *d = reinterpret_cast<double *>(S);
    for (int i = first; i < CB_FIT_PERSIST_DOUBLES; i += step) d[i] = P.state[(int64_t)i * P.stride + w];
}
CB_BODY void fit_state_store(const FitParams &P, int64_t w, const FitState *S) {
    const double *d = reinterpret_cast<const double *>(S);
    for (int i = 0; i < CB_FIT_PERSIST_DOUBLES; ++i) P.state[(int64_t)i * P.stride + w] = d[i];
}
// constants of the launch from the template; the window's own prior means where the caller gave per-row ones
CB_BODY void fit_state_constants(const FitParams &P, int64_t w, FitState *S) {
    const FitState *T = reinterpret_cast<const FitState *>(P.tmpl);
#pragma unroll
    for (int j = 0; j < 4; ++j) { S->lb[j] = T->lb[j]; S->ub[j] = T->ub[j]; S->pmu[j] = T->pmu[j]; S->pw[j] = T->pw[j]; }
    S->cu = T->cu;
    if (P.prior_mu_dev) {
        const double *row = P.prior_mu_dev + w * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (!isnan(P.opt.prior_mu[j]) && !isnan(row[j])) S->pmu[j] = row[j];
    }
}

// template: one thread -- the initial state for the launch's common start point and priors.
CB_BODY void fit_template_body(const FitParams &P) {
    const double *meta = P.ktab + 3 * P.nk + (CB_MILLER_NMAX + 1);
    const bool sorted = meta[0] != 0.0;
    FitState S;
    memset(&S, 0, sizeof(S));
    fit_init(&S, &P.opt, P.opt.x0, nullptr, sorted ? meta[1] : 0.0, sorted ? meta[2] : 0.0);
    const double *src = reinterpret_cast<const double *>(&S);
    for (int i = 0; i < CB_FIT_STATE_DOUBLES; ++i) P.tmpl[i] = src[i];
    P.count[0] = (unsigned)P.nwin;
    for (int q = 1; q < 6; ++q) P.count[q] = 0u;
}

// begin: one thread per window -- initial state (the template, or its own start point / prior means), coefficients
// of the first evaluation, the identity list.
CB_BODY void fit_begin_window(const FitParams &P, int64_t w) {
    FitState S;
    if (P.x0_dev || P.prior_mu_dev) {
        const double *meta = P.ktab + 3 * P.nk + (CB_MILLER_NMAX + 1);
        const bool sorted = meta[0] != 0.0;
        memset(&S, 0, sizeof(S));
        fit_init(&S, &P.opt, P.x0_dev ? P.x0_dev + w * 4 : P.opt.x0, P.prior_mu_dev ? P.prior_mu_dev + w * 4 : nullptr,
                 sorted ? meta[1] : 0.0, sorted ? meta[2] : 0.0);
    } else {
        double *d = reinterpret_cast<double *>(&S);
        for (int i = 0; i < CB_FIT_STATE_DOUBLES; ++i) d[i] = P.tmpl[i];
    }
    fit_stage_coefs(&S, P.coef + w * (3 * FC_COUNT));
    fit_state_store(P, w, &S);
    P.list_in[w] = (int32_t)w;
}
template <class Ctx>
CB_BODY void fit_begin_body(const Ctx &cx, const FitParams &P) {
    const int64_t w = (int64_t)cx.bid() * cx.nthreads() + cx.tid();
    if (w < P.nwin) fit_begin_window(P, w);
}

// eval: one warp per unfinished window; the warps of a persistent grid claim list entries from a counter.
template <class Ctx>
CB_BODY void fit_eval_body(const Ctx &cx, const FitParams &P) {
    const int tid = cx.tid(), warp = tid >> 5, lane = tid & 31, L = cx.team_lanes();
    const unsigned n = P.count[P.round % 3];
    unsigned int *next = P.count + 3 + P.round % 3;
    if (cx.bid() == 0 && tid == 0) {       // what the round after next uses: nobody reads or writes these now
        P.count[(P.round + 2) % 3] = 0u;
        P.count[3 + (P.round + 2) % 3] = 0u;
    }
    if ((int64_t)n < (int64_t)P.handover) return;       // left to the finish kernel (the step kernel passes the list on)
    double *tab = reinterpret_cast<double *>(cx.smem()) + (int64_t)warp * fit_eval_warp_doubles(P.nk);
    double *vals = tab + CB_FIT_TABLE_DOUBLES;
    const double *meta = P.ktab + 3 * P.nk + (CB_MILLER_NMAX + 1);
    const int sorted = meta[0] != 0.0 ? 1 : 0;
    for (int j = lane; j < CB_TEMME_TAB; j += L) tab[CB_FIT_CT_DOUBLES + CB_FIT_RPQ_DOUBLES + j] = P.ktab[3 * P.nk + j];
#pragma unroll 1
    for (;;) {
        unsigned e = 0;
        if (lane == 0) e = cx.atomic_inc_u32(next);
        e = (unsigned)cx.shfl_i((int)e, 0);
        if (e >= n) break;
        const int64_t w = P.list_in[e];
        cx.sync_warp();
        const double *cg = P.coef + w * (3 * FC_COUNT);
        for (int j = lane; j < 3 * FC_COUNT; j += L) tab[CB_FIT_COEF_OFFSET + j] = cg[j];
        cx.sync_warp();
        fit_eval_values(cx, tab, P.kh, P.ktab, P.nk, sorted, lane, vals);
        cx.sync_warp();
        fit_sums_from_values(cx, vals, P.kh, P.nk, tab[CB_FIT_COEF_OFFSET + 2 * FC_COUNT + FC_AUX] == 3.0,
                             tab[CB_FIT_COEF_OFFSET + FC_AUX], lane, P.phi + w * P.nk, P.sums + w * 16);
    }
}

// first round of a launch whose windows share the start point (no x0_dev): the model values are the same for every
// window, so each CTA evaluates them once (warp 0, into shared memory) and every window only forms its sums.
CB_HD size_t fit_first_smem_bytes(int nk) { return (size_t)(CB_FIT_TABLE_DOUBLES + 4 * nk) * sizeof(double); }
template <class Ctx>
CB_BODY void fit_first_body(const Ctx &cx, const FitParams &P) {
    const int tid = cx.tid(), warp = tid >> 5, lane = tid & 31, L = cx.team_lanes();
    const int wpb = cx.nthreads() >> 5, nk = P.nk;
    double *tab = reinterpret_cast<double *>(cx.smem()), *vals = tab + CB_FIT_TABLE_DOUBLES;
    const double *meta = P.ktab + 3 * nk + (CB_MILLER_NMAX + 1);
    const int sorted = meta[0] != 0.0 ? 1 : 0;
    if (warp == 0) {
        for (int j = lane; j < CB_TEMME_TAB; j += L) tab[CB_FIT_CT_DOUBLES + CB_FIT_RPQ_DOUBLES + j] = P.ktab[3 * nk + j];
        for (int j = lane; j < 3 * FC_COUNT; j += L) tab[CB_FIT_COEF_OFFSET + j] = P.coef[j];      // window 0's = everybody's
        cx.sync_warp();
        fit_eval_values(cx, tab, P.kh, P.ktab, nk, sorted, lane, vals);
    }
    cx.sync_block();
    const double hb = tab[CB_FIT_COEF_OFFSET + FC_AUX];
    const bool use_pm = tab[CB_FIT_COEF_OFFSET + 2 * FC_COUNT + FC_AUX] == 3.0;
#pragma unroll 1
    for (int64_t w = (int64_t)cx.bid() * wpb + warp; w < P.nwin; w += (int64_t)cx.nblocks() * wpb)
        fit_sums_from_values(cx, vals, P.kh, nk, use_pm, hb, lane, P.phi + w * nk, P.sums + w * 16);
}

// step: one thread per unfinished window, on a private copy of its state.  (Measured slower: four lanes per window
// with the three coefficient sets side by side, 3.4 vs 2.3 ms over the rounds of config 3; the private copy in
// shared memory at an odd stride instead of local memory, 2.8 vs 2.3 ms.)
template <class Ctx>
CB_BODY void fit_step_entry(const Ctx &cx, const FitParams &P, int64_t e) {
    int max_iter;
    double gtol, ftol;
    fit_solver_constants(P, max_iter, gtol, ftol);
    const int64_t w = P.list_in[e];
    FitState S;
    fit_state_load(P, w, &S, 0, 1);
    fit_state_constants(P, w, &S);
    const bool done = fit_step_window(&S, P.sums + w * 16, gtol, ftol, max_iter, P.coef + w * (3 * FC_COUNT));
    if (done) fit_write_result(P, &S, w);
    else {
        fit_state_store(P, w, &S);
        P.list_out[cx.atomic_inc_u32(P.count + (P.round + 1) % 3)] = (int32_t)w;
    }
}
template <class Ctx>
CB_BODY void fit_step_body(const Ctx &cx, const FitParams &P) {
    const int tid = cx.tid(), nt = cx.nthreads();
    const unsigned n = P.count[P.round % 3];
    if ((int64_t)n < (int64_t)P.handover) {              // handed over: the list travels on unchanged
        for (int64_t e = (int64_t)cx.bid() * nt + tid; e < (int64_t)n; e += (int64_t)cx.nblocks() * nt) P.list_out[e] = P.list_in[e];
        if (cx.bid() == 0 && tid == 0) P.count[(P.round + 1) % 3] = n;
        return;
    }
#pragma unroll 1
    for (int64_t e = (int64_t)cx.bid() * nt + tid; e < (int64_t)n; e += (int64_t)cx.nblocks() * nt) fit_step_entry(cx, P, e);
}

// finish: the windows still unfinished after the batched rounds, run to the end -- one CTA per window: its warps
// share the chunks of every evaluation (fit_eval_values part / nparts), warp 0 forms the sums, takes the
// solver step on lane 0 and fills the coefficients of the next evaluation on lanes 0..nv-1.  With few windows left
// a round is latency bound (one warp needs ~50 us for an evaluation and ~20 us for the scalar step).
// Shared memory (doubles): tables, state, sums[16], model values [4 nk].
CB_HD size_t fit_finish_smem_bytes(int nk) {
    return (size_t)(((CB_FIT_TABLE_DOUBLES + CB_FIT_STATE_DOUBLES + 1) & ~1) + 16 + 4 * nk) * sizeof(double);
}
template <class Ctx>
CB_BODY void fit_finish_body(const Ctx &cx, const FitParams &P) {
    const int tid = cx.tid(), nthr = cx.nthreads(), warp = tid >> 5, lane = tid & 31, L = cx.team_lanes();
    const int nwarps = nthr >> 5;
    const unsigned n = P.count[P.round % 3];
    int max_iter;
    double gtol, ftol;
    fit_solver_constants(P, max_iter, gtol, ftol);
    double *tab = reinterpret_cast<double *>(cx.smem());
    double *sd = tab + CB_FIT_TABLE_DOUBLES;
    double *sums = tab + ((CB_FIT_TABLE_DOUBLES + CB_FIT_STATE_DOUBLES + 1) & ~1), *vals = sums + 16;
    FitState *S = reinterpret_cast<FitState *>(sd);
    const double *meta = P.ktab + 3 * P.nk + (CB_MILLER_NMAX + 1);
    const int sorted = meta[0] != 0.0 ? 1 : 0;
    if (warp == 0)
        for (int j = lane; j < CB_TEMME_TAB; j += L) tab[CB_FIT_CT_DOUBLES + CB_FIT_RPQ_DOUBLES + j] = P.ktab[3 * P.nk + j];
#pragma unroll 1
    for (int64_t e = cx.bid(); e < (int64_t)n; e += cx.nblocks()) {
        const int64_t w = P.list_in[e];
        cx.sync_block();
        if (warp == 0) {
            const double *cg = P.coef + w * (3 * FC_COUNT);
            fit_state_load(P, w, S, lane, L);
            if (lane == 0) fit_state_constants(P, w, S);
            for (int j = lane; j < 3 * FC_COUNT; j += L) tab[CB_FIT_COEF_OFFSET + j] = cg[j];
        }
        cx.sync_block();
        for (;;) {
            fit_eval_values(cx, tab, P.kh, P.ktab, P.nk, sorted, lane, vals, warp, nwarps);
            cx.sync_block();
            if (warp == 0) {
                fit_sums_from_values(cx, vals, P.kh, P.nk, tab[CB_FIT_COEF_OFFSET + 2 * FC_COUNT + FC_AUX] == 3.0,
                                     tab[CB_FIT_COEF_OFFSET + FC_AUX], lane, P.phi + w * P.nk, sums);
                if (lane == 0 && fit_step_window(S, sums, gtol, ftol, max_iter, nullptr)) fit_write_result(P, S, w);
                cx.sync_warp();
                if (S->action == FIT_GO) {
                    const int nv = S->nv;
                    double *c = tab + CB_FIT_COEF_OFFSET;
                    for (int v = lane; v < nv; v += L) fit_coef_fill(c + v * FC_COUNT, S->pt, S->hb, v);
                    cx.sync_warp();
                    if (lane == 0) { c[FC_AUX] = S->hb; c[FC_COUNT + FC_AUX] = S->cu; c[2 * FC_COUNT + FC_AUX] = (double)nv; }
                }
            }
            cx.sync_block();
            if (S->action != FIT_GO) break;
        }
    }
}
