// Bouligand et al. (2009) eq. 4 forward model and the CurieOptimise objective.
// Replaces pycurious/grid.py:552-600 (bouligand2009) and pycurious/optimise.py:159-230
// (objective_routine / objective_function / min_func).
//
//   Phi(k) = C - 2 k zt - (beta-1) ln k - k dz + ln A
//   A      = sqrt(pi)/Gamma(1+beta/2) * [ cosh(k dz)/2 * Gamma(nu) - K_nu(k dz) (k dz/2)^nu ],  nu=(1+beta)/2
//
// The arithmetic follows the reference literally (cosh overflow -> inf, dz = 0 -> nan, the
// small-x cancellation inside the bracket), because min_func maps every non-finite misfit to
// 1e99 (optimise.py:226-227) and the samplers rely on that.
#pragma once
#include "besselk.cuh"
#include "exec_ctx.cuh"

#define CB_SQRT_PI 1.77245385090551602729816748334114518
#define CB_BIG_MISFIT 1e99

// Everything that depends on the parameter vector but not on the wavenumber.
struct ModelCoef {
    double beta, zt, dz, C;
    double nu;         // (1+beta)/2, signed as in the reference
    KnuOrder ord;      // |nu| prepared for K
    double gnu;        // Gamma(nu)            (grid.py:594)
    double lead;       // sqrt(pi)/Gamma(1+beta/2)  (grid.py:591-592)
    double lhalfdz;    // ln(dz/2)
};

CB_HD ModelCoef model_coef(double beta, double zt, double dz, double C) {
    ModelCoef m;
    m.beta = beta; m.zt = zt; m.dz = dz; m.C = C;
    m.nu = 0.5 * (1.0 + beta);
    m.ord = knu_order(m.nu);
    // Gamma(nu) from the 1/Gamma(1+mu) the order already holds: Gamma(mu+nl) = Gamma(1+mu) prod_{i<nl} (mu+i)
    if (m.nu >= 0.5 && m.ord.nl >= 1 && m.ord.nl <= 12) {
        double prod = 1.0;
        for (int i = 1; i < m.ord.nl; ++i) prod *= m.ord.mu + (double)i;
        m.gnu = prod / m.ord.gampl;
    } else {
        m.gnu = cb_gamma(m.nu);
    }
    m.lead = CB_SQRT_PI / cb_gamma(1.0 + 0.5 * beta);
    m.lhalfdz = cb_log(0.5 * dz);
    return m;
}

// (x/2)^nu with x = k dz, from the cached ln k:  exp(nu (ln k + ln(dz/2))).  dz <= 0 falls back
// to pow() so that the reference's nan / inf pattern is kept.
CB_HD double half_x_pow(const ModelCoef &m, double x, double logk) {
    if (m.dz > 0.0) return exp(m.nu * (logk + m.lhalfdz));
    return pow(0.5 * x, m.nu);
}

// Phi(k).  logk = ln k is passed in because it is constant per wavenumber for a whole fit / chain.
CB_HD double bouligand_point(const ModelCoef &m, double k, double logk) {
    const double x = k * m.dz;
    double kn, knm1;
    knu_pair(m.ord, x, kn, knm1);
    const double bracket = 0.5 * cosh(x) * m.gnu - kn * half_x_pow(m, x, logk);
    return m.C - 2.0 * k * m.zt - (m.beta - 1.0) * logk - x + log(m.lead * bracket);
}

// Phi(k) and its analytic partial derivative with respect to dz:
//   d/dx [K_nu(x) (x/2)^nu] = -(x/2)^nu K_{nu-1}(x)
//   dPhi/ddz = -k + k [ sinh(x)/2 Gamma(nu) + (x/2)^nu K_{nu-1}(x) ] / bracket
CB_HD double bouligand_point_ddz(const ModelCoef &m, double k, double logk, double &dphi_ddz) {
    const double x = k * m.dz;
    double kn, knm1;
    knu_pair(m.ord, x, kn, knm1);
    const double hx = half_x_pow(m, x, logk);
    const double ex = exp(x), emx = 1.0 / ex;
    const double ch = 0.5 * (ex + emx), sh = 0.5 * (ex - emx);
    const double bracket = 0.5 * ch * m.gnu - kn * hx;
    dphi_ddz = -k + k * (0.5 * sh * m.gnu + hx * knm1) / bracket;
    return m.C - 2.0 * k * m.zt - (m.beta - 1.0) * logk - x + log(m.lead * bracket);
}

struct Priors {
    double mu[4], sigma[4];    // mu NaN = no prior
};

CB_HD double prior_penalty(const Priors &pr, const double *x) {
    double c = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (!isnan(pr.mu[j])) {
            const double d = x[j] - pr.mu[j];
            c += 0.5 * d * d / (pr.sigma[j] * pr.sigma[j]);
        }
    return c;
}

// min_func for one spectrum, evaluated cooperatively by the lanes of a warp (optimise.py:220-230).
template <class Ctx>
CB_BODY double misfit_warp(const Ctx &cx, const double *kh, const double *logk, const double *phi, int nk,
                           const double *x, const Priors &pr, int lane) {
    const ModelCoef m = model_coef(x[0], x[1], x[2], x[3]);
    double acc = 0.0;
    for (int i = lane; i < nk; i += cx.team_lanes()) {
        const double r = bouligand_point(m, kh[i], logk[i]) - phi[i];
        acc = fma(r, r, acc);
    }
    double f = 0.5 * cx.warp_sum(acc);
    if (!isfinite(f)) return CB_BIG_MISFIT;
    return f + prior_penalty(pr, x);
}
