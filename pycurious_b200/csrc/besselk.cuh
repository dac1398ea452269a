// Modified Bessel function of the second kind K_nu(x), real order, x > 0, in binary64.
// Replaces scipy.special.kv (AMOS zbesk, not vendored by the reference) behind
// pycurious/grid.py:595, and scipy.special.gamma behind grid.py:592-594.
//
// Method (Temme 1975, "On the numerical evaluation of the modified Bessel function of the
// third kind", J. Comput. Phys. 19; Thompson & Barnett 1987 for the continued fraction):
//   nu = nl + mu with |mu| <= 1/2
//   x <= 2 : Temme's series for K_mu, K_mu+1
//   x >  2 : Steed's algorithm on the continued fraction CF2 for K_mu, K_mu+1
//   then the (stable) upward recurrence K_{m+1} = K_{m-1} + (2m/x) K_m.
// Everything that depends on the order only (mu, 1/Gamma(1+-mu), Gamma_1, Gamma_2) is split
// into KnuOrder: in the fit / MCMC kernels the order is the same for every wavenumber of a
// window, so it is computed once per model evaluation, not once per point.
#pragma once
#include "common.cuh"
#include "fastlog.cuh"

#define CB_PI 3.14159265358979323846264338327950288
#define CB_KNU_EPS 1.0e-16
#define CB_KNU_MAXIT 400
#define CB_KNU_MAX_ORDER 2000

// Code-size discipline: the fit / MCMC kernels run ~16 independent warps per SM through this code,
// each at a different place, and the SM's instruction cache holds 32 KB.  Helpers that are used from
// several places are therefore real functions on the device (one copy), not inlined.
#if defined(__CUDACC__) && !defined(CB_EMU)
#define CB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define CB_HD_NOINLINE CB_HD
#endif
CB_HD_NOINLINE double cb_exp(double x) { return exp(x); }
// ln x through the spectrum kernels' short logarithm (fastlog.cuh: < 1 ulp, ~40 instructions instead of the
// library's ~75; 0, subnormal, negative, inf and nan arguments take the library call): one call per forward-model
// point plus the order-only coefficients -- it was 9 % of fit_kernel's instructions.
CB_HD_NOINLINE double cb_log(double x) { return 2.0 * cb_half_log(x); }

// 1/Gamma(1+m) = sum_k c_k m^k  (mpmath, 40 digits, rounded), split into even / odd parts:
// E(m) = sum_{k even} c_k m^k, O(m) = sum_{k odd} c_k m^(k-1), so that
// 1/Gamma(1+m) = E + m O and 1/Gamma(1-m) = E - m O.
CB_HD_NOINLINE void cb_rgamma_parts(double m, double &even, double &odd) {
    const double m2 = m * m;
    double e = -2.058326053566506783222e-14, o = -5.34812253942301798237e-15;
#define CB_RG(ce, co) e = fma(e, m2, ce); o = fma(o, m2, co);
    CB_RG(-3.696805618642205708188e-12, 5.100370287454475979015e-13)
    CB_RG(1.043426711691100510492e-10, 7.78226343990507125405e-12)
    CB_RG(5.002007644469222930056e-9, -1.181274570487020144588e-9)
    CB_RG(-2.05633841697760710345e-7, 6.116095104481415817862e-9)
    CB_RG(-0.000001250493482142670657345, 0.000001133027231981695882374)
    CB_RG(0.0001280502823881161861532, -0.00002013485478078823865569)
    CB_RG(-0.001165167591859065112114, -0.0002152416741149509728157)
    CB_RG(-0.009621971527876973562115, 0.007218943246663099542395)
    CB_RG(0.1665386113822914895017, -0.04219773455554433674821)
    CB_RG(-0.655878071520253881077, -0.042002635034095235529)
    CB_RG(1.0, 0.5772156649015328606065)
#undef CB_RG
    even = e;
    odd = o;
}

// Poles, negative and large arguments: the C library tgamma, which has the same pole conventions
// as scipy.special.gamma (gamma(0) = inf, gamma(-n) = nan).  Cold path.
CB_HD_NOINLINE double cb_gamma_slow(double z) {
    if (z == 0.0) return copysign(INFINITY, z);
    if (z < 0.0 && z == floor(z)) return NAN;
    return tgamma(z);
}

// Gamma(z).  0 < z <= 24: shift into [0.5, 1.5] and use the reciprocal-gamma series (a few ulp).
CB_HD_NOINLINE double cb_gamma(double z) {
    if (!(z > 0.0) || z > 24.0) return cb_gamma_slow(z);
    double w = z, num = 1.0, den = 1.0;
    while (w > 1.5) { w -= 1.0; num *= w; }
    while (w < 0.5) { den *= w; w += 1.0; }
    double e, o;
    const double m = w - 1.0;
    cb_rgamma_parts(m, e, o);
    return num / (den * fma(m, o, e));
}

struct KnuOrder {
    double nu;       // |order|
    double mu;       // fractional part in [-1/2, 1/2]
    int nl;          // recurrence steps: K_nu = K_{mu+nl}
    double gampl;    // 1/Gamma(1+mu)
    double gammi;    // 1/Gamma(1-mu)
    double gam1;     // (gammi - gampl) / (2 mu)
    double gam2;     // (gammi + gampl) / 2
    double sinc;     // pi mu / sin(pi mu)
};

CB_HD double cb_sinpi(double t) {          // |t| <= 1/2
#ifdef __CUDA_ARCH__
    return sinpi(t);
#else
    return sin(CB_PI * t);
#endif
}

CB_HD KnuOrder knu_order(double nu_signed) {
    KnuOrder o;
    o.nu = fabs(nu_signed);
    if (o.nu < 0.5) {
        // use the order -nu in (-1/2, 0]: K_mu = K_nu and K_{mu+1} = K_{nu-1}
        o.nl = 0;
        o.mu = -o.nu;
    } else if (o.nu < 1.0e6) {
        o.nl = (int)floor(o.nu + 0.5);
        o.mu = o.nu - (double)o.nl;
    } else {                      // absurd orders (a runaway trial step): K overflows, see knu_pair
        o.nl = 1000000;
        o.mu = 0.0;
    }
    double e, od;
    cb_rgamma_parts(o.mu, e, od);
    o.gampl = fma(o.mu, od, e);
    o.gammi = fma(-o.mu, od, e);
    o.gam1 = -od;
    o.gam2 = e;
    const double pm = CB_PI * o.mu;
    o.sinc = (fabs(pm) < 1e-8) ? 1.0 : pm / cb_sinpi(o.mu);
    return o;
}

// K_mu(x) and K_{mu+1}(x), |mu| <= 1/2, by Temme's series (x <= 2).
CB_HD void knu_temme_series(const KnuOrder &o, double x, double &kmu, double &kmu1) {
    const double mu = o.mu, mu2 = mu * mu;
    const double x2 = 0.5 * x;
    const double d = -log(x2);                 // ln(2/x)
    const double sg = mu * d;                  // sigma
    const double shc = (fabs(sg) < 1e-8) ? 1.0 : sinh(sg) / sg;
    double ff = o.sinc * (o.gam1 * cosh(sg) + o.gam2 * shc * d);
    const double ee = exp(sg);
    double p = 0.5 * ee / o.gampl;
    double q = 0.5 / (ee * o.gammi);
    double c = 1.0;
    const double dd = x2 * x2;
    double sum = ff, sum1 = p;
    for (int i = 1; i <= CB_KNU_MAXIT; ++i) {
        const double fi = (double)i;
        ff = (fi * ff + p + q) / (fi * fi - mu2);
        c *= dd / fi;
        p /= (fi - mu);
        q /= (fi + mu);
        const double del = c * ff;
        sum += del;
        sum1 += c * (p - fi * ff);
        if (fabs(del) < fabs(sum) * CB_KNU_EPS) break;
    }
    kmu = sum;
    kmu1 = sum1 / x2;
}

// K_mu(x) and K_{mu+1}(x), |mu| <= 1/2, by Steed's algorithm on CF2 (x > 2).
CB_HD void knu_steed_cf2(const KnuOrder &o, double x, double &kmu, double &kmu1) {
    const double mu = o.mu;
    double b = 2.0 * (1.0 + x);
    double d = 1.0 / b;
    double h = d, delh = d;
    double q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25 - mu * mu;
    double q = a1, c = a1, a = -a1;
    double s = 1.0 + q * delh;
    for (int i = 2; i <= CB_KNU_MAXIT; ++i) {
        a -= (double)(2 * (i - 1));
        c = -a * c / (double)i;
        const double qnew = (q1 - b * q2) / a;
        q1 = q2;
        q2 = qnew;
        q += c * qnew;
        b += 2.0;
        d = 1.0 / (b + a * d);
        delh = (b * d - 1.0) * delh;
        h += delh;
        const double dels = q * delh;
        s += dels;
        if (fabs(dels) < fabs(s) * CB_KNU_EPS) break;
    }
    h = a1 * h;
    kmu = sqrt(CB_PI / (2.0 * x)) * exp(-x) / s;
    kmu1 = kmu * (mu + x + 0.5 - h) / x;
}

// K_nu(x) and K_{nu-1}(x) for the prepared order.  x <= 0: scipy conventions
// (kv(nu, 0) = inf, kv(nu, x<0) = nan for real x).
CB_HD void knu_pair(const KnuOrder &o, double x, double &k_nu, double &k_num1) {
    if (!(x > 0.0)) {
        k_nu = k_num1 = (x == 0.0) ? INFINITY : NAN;
        return;
    }
    if (o.nl > CB_KNU_MAX_ORDER || isnan(o.nu)) {
        // Gamma(nu) overflows at nu = 171.6, so the reference's bracket is already non-finite
        // (min_func -> 1e99) far below this order; do not walk a recurrence of absurd length
        k_nu = k_num1 = isnan(o.nu) ? NAN : INFINITY;
        return;
    }
    double a, b;
    if (x <= 2.0) knu_temme_series(o, x, a, b);
    else knu_steed_cf2(o, x, a, b);
    if (o.nl == 0) {          // order mu = -nu: a = K_nu, b = K_{1-nu} = K_{nu-1}
        k_nu = a;
        k_num1 = b;
        return;
    }
    const double tox = 2.0 / x;
    for (int i = 1; i < o.nl; ++i) {
        const double cnew = fma((o.mu + (double)i) * tox, b, a);
        a = b;
        b = cnew;
    }
    k_nu = b;
    k_num1 = a;
}

CB_HD double cb_kv(double nu, double x) {
    if (isnan(nu) || isnan(x)) return NAN;
    KnuOrder o = knu_order(nu);
    double k, km1;
    knu_pair(o, x, k, km1);
    return k;
}
