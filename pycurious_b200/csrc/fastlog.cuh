// 0.5 * ln(x) for the log-modulus sink of the spectrum kernels: ln|F| = 0.5 ln(re^2 + im^2) is taken once per
// retained Fourier coefficient (2.4e9 times per config-3 step) and the math library's log() costs ~75 issue slots
// there, most of them constant moves and special-case plumbing -- a quarter of colspec_kernel's instructions.
// This one is ~40: the classic reduction x = 2^k m, m in [sqrt(1/2), sqrt(2)), s = f / (2 + f) with f = m - 1, the
// degree-14 even minimax polynomial in s (published fdlibm / msun e_log.c coefficients, error < 1 ulp), the one
// division as a hardware reciprocal seed plus two Newton steps (2 + f lies in [1.7, 2.42): no special cases), and
// the coefficients as constant-bank operands.  Zero, subnormal, negative, inf and nan arguments take the library
// call, so log(0) = -inf propagates exactly like numpy's (grid.py:283).
#pragma once
#include "common.cuh"
#include <string.h>

#define CB_FASTLOG_COEFS                                                                                      \
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,   \
    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,                             \
    6.93147180369123816490e-01 /* ln2 hi */, 1.90821492927058770002e-10 /* ln2 lo */

#if defined(__CUDACC__) && !defined(CB_EMU)
static __constant__ double cb_fastlog_k[9] = {CB_FASTLOG_COEFS};
#if defined(__CUDA_ARCH__)
#define CB_FASTLOG_DEVICE 1            /* device pass of the product build */
#endif
#endif

#if defined(__CUDACC__) && !defined(CB_EMU)
static __device__ __noinline__ double cb_half_log_special(double x) { return 0.5 * log(x); }
#endif

CB_HD double cb_half_log(double x) {
#ifdef CB_FASTLOG_DEVICE
    const double *c = cb_fastlog_k;
    const long long bits = __double_as_longlong(x);
#else
    const double c[9] = {CB_FASTLOG_COEFS};
    long long bits;
    memcpy(&bits, &x, 8);
#endif
    int hi = (int)(bits >> 32);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) {          // 0, subnormal, < 0, inf, nan
#ifdef CB_FASTLOG_DEVICE
        return cb_half_log_special(x);
#else
        return 0.5 * log(x);
#endif
    }
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int up = (hi + 0x95f64) & 0x100000;                   // mantissa >= sqrt(2): halve it, bump the exponent
    k += up >> 20;
    const long long mb = ((long long)(hi | (up ^ 0x3ff00000)) << 32) | (bits & 0xffffffffll);
    double m, r;
#ifdef CB_FASTLOG_DEVICE
    m = __longlong_as_double(mb);
#else
    memcpy(&m, &mb, 8);
#endif
    const double f = m - 1.0, y = 2.0 + f;
#ifdef CB_FASTLOG_DEVICE
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
#else
    r = (double)(float)(1.0 / y);                               // emulation: a seed of comparable (~2^-23) accuracy
#endif
    r = fma(fma(-y, r, 1.0), r, r);
    r = fma(fma(-y, r, 1.0), r, r);
    double s = f * r;
    s = fma(fma(-y, s, f), r, s);                               // f / (2 + f) to within an ulp
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, c[5], c[3]), c[1]);
    const double t2 = z * fma(w, fma(w, fma(w, c[6], c[4]), c[2]), c[0]);
    const double R = t1 + t2, hfsq = 0.5 * f * f, dk = (double)k;
    return 0.5 * (dk * c[7] - ((hfsq - fma(s, hfsq + R, dk * c[8])) - f));
}
