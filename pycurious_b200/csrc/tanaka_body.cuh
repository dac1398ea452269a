// Batched Tanaka et al. (1999) line fits, one window per warp.
// Replaces, for every row of a batch of radial spectra (power = 0.5),
//   tanaka1999      pycurious/grid.py:603-674   two weighted straight-line fits over k-bands
//   ComputeTanaka   pycurious/grid.py:677-699   Zb = |2 Zo - Zt|, eZb = 2 dZo + dZt
// The arithmetic follows the reference literally, including its weights 1/E^2 with E already a
// variance (the sigma^4 quirk, grid.py:630-635) and the centroid band's
// Y = ln(exp(S)/(2 pi sf)), E = ln(exp(sigma^2)/(2 pi sf)) (grid.py:667-668).
#pragma once
#include "exec_ctx.cuh"

struct TanakaParams {
    const double *k;        // [nk] rad/km
    int32_t nk;
    const double *lnphi;    // [nwin, nk]
    const double *sigma;    // [nwin, nk]
    int64_t nwin;
    double top_lo, top_hi;  // kmin_range (spatial frequency k / 2 pi)
    double cen_lo, cen_hi;  // kmax_range
    double *out;            // [nwin, 8]: Ztr, btr, dZtr, Zor, bor, dZor, Zb, eZb
    int32_t warps_per_block;
};

// compute_coefficients of grid.py:632-653 from the five weighted sums
CB_HD void tanaka_line(double sx, double sy, double s1, double sxx, double sxy, double *Z, double *b, double *dZ) {
    const double TL = sx * sy - sxy * s1;
    const double BL = sx * sx - sxx * s1;
    *Z = TL / BL;
    *b = (sxy - *Z * sxx) / sx;
    *dZ = sqrt(s1 / (sxx * s1 - sx * sx));
}

template <class Ctx>
CB_BODY void tanaka_body(const Ctx &cx, const TanakaParams &P) {
    const int tid = cx.tid();
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t w = (int64_t)cx.bid() * P.warps_per_block + warp;
    if (w >= P.nwin) return;
    const double two_pi = 2.0 * 3.14159265358979323846264338327950288;
    double a[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) a[q] = 0.0;
    for (int i = lane; i < P.nk; i += cx.team_lanes()) {
        const double sf = P.k[i] / two_pi;
        const double S = P.lnphi[w * P.nk + i];
        const double sg = P.sigma[w * P.nk + i];
        const double var = sg * sg;
        if (sf >= P.top_lo && sf <= P.top_hi) {
            const double wt = 1.0 / (var * var);
            a[0] += sf * wt; a[1] += S * wt; a[2] += wt; a[3] += sf * sf * wt; a[4] += sf * S * wt;
        }
        if (sf >= P.cen_lo && sf <= P.cen_hi) {
            const double Y = log(exp(S) / (sf * two_pi));
            const double E = log(exp(var) / (sf * two_pi));
            const double wt = 1.0 / (E * E);
            a[5] += sf * wt; a[6] += Y * wt; a[7] += wt; a[8] += sf * sf * wt; a[9] += sf * Y * wt;
        }
    }
#pragma unroll
    for (int q = 0; q < 10; ++q) a[q] = cx.warp_sum(a[q]);
    if (lane == 0) {
        double *o = P.out + w * 8;
        tanaka_line(a[0], a[1], a[2], a[3], a[4], o + 0, o + 1, o + 2);
        tanaka_line(a[5], a[6], a[7], a[8], a[9], o + 3, o + 4, o + 5);
        o[6] = fabs(2.0 * o[3] - o[0]);
        o[7] = 2.0 * o[5] + o[2];
    }
}
