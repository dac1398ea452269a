// Radon transform: ray/cell intersection lengths (CSR), sinogram, and the 1-D spectral tail of
// CurieGrid.azimuthal_spectrum.
// Replaces src/cradon.c:4-246 (the C ray tracer), the sparse product of src/radon.pyx:110-115 and
// pycurious/grid.py:393-401.
//
// One thread traces one ray.  The walk is inherently sequential per ray and every floating-point
// expression is kept in the operand order of cradon.c; this translation unit is compiled with
// -fmad=false so that no multiply-add is contracted and the emitted (cell, length) pairs are
// bit-identical to the C code.  Cells are unit squares with edges at n - 1/2 (cradon.c:21-31); the
// flat cell id is ix*ny + iy (cradon.c:66).  The CSR depends only on (nw, theta), so it is traced
// once per plan and reused for every window of a batch.
#pragma once
#include "exec_ctx.cuh"

#define CB_RADON_TINY 1.e-10          /* cradon.c:11 */
#define CB_PI_RADON 3.14159265358979323846264338327950288

CB_HD double radon_edge(int64_t n) { return -0.5 + (double)n; }

CB_HD int64_t radon_locate(double v, int64_t n) {      // cradon.c:63-64
    int64_t i;
    for (i = 0; i < n; ++i)
        if (v < radon_edge(i + 1)) break;
    return i;
}

// Walks one ray.  With len == NULL it only counts.  Returns the number of (cell, length) pairs.
CB_HD int64_t radon_walk(double xs, double ys, double xr, double yr, int64_t nx, int64_t ny,
                         double *len, int64_t *cell) {
    int64_t n = 0;
#define CB_EMIT(c, l) do { if (len) { len[n] = (l); cell[n] = (c); } ++n; } while (0)
    if (xs > xr) {                                   // cradon.c:47-55
        double t = xs; xs = xr; xr = t;
        t = ys; ys = yr; yr = t;
    }
    double x = xs, y = ys;
    int64_t ix, iy;
    if (fabs(ys - yr) < CB_RADON_TINY) {             // constant y, cradon.c:61-90
        ix = radon_locate(x, nx);
        iy = radon_locate(y, ny);
        while (x < xr) {
            const double e = radon_edge(ix + 1);
            const double stop = e < xr ? e : xr;
            CB_EMIT(ix * ny + iy, stop - x);
            ++ix;
            x = radon_edge(ix);
        }
    } else if (fabs(xs - xr) < CB_RADON_TINY) {      // constant x, cradon.c:91-126
        if (ys > yr) { double t = ys; ys = yr; yr = t; }
        y = ys;
        ix = radon_locate(x, nx);
        iy = radon_locate(y, ny);
        while (y < yr) {
            const double e = radon_edge(iy + 1);
            const double stop = e < yr ? e : yr;
            CB_EMIT(ix * ny + iy, stop - y);
            ++iy;
            y = radon_edge(iy);
        }
    } else {                                         // oblique, cradon.c:127-219
        const double m = (yr - ys) / (xr - xs);
        const double b = yr - m * xr;
        const bool rising = m > 0;
        ix = radon_locate(x, nx);
        iy = radon_locate(y, ny);
        while (x < xr) {
            const double ycross = m * radon_edge(ix + 1) + b;
            if (rising) {
                while (y < ycross && y < yr) {
                    const double e = radon_edge(iy + 1);
                    double ye = e < ycross ? e : ycross;
                    ye = ye < yr ? ye : yr;
                    const double xe = (ye - b) / m;
                    const double dlx = xe - x, dly = ye - y;
                    CB_EMIT(ix * ny + iy, sqrt(dlx * dlx + dly * dly));
                    x = xe;
                    y = ye;
                    if (fabs(y - radon_edge(iy + 1)) < CB_RADON_TINY) ++iy;
                }
            } else {
                while (y > ycross && y > yr) {
                    const double e = radon_edge(iy);
                    double ye = e > ycross ? e : ycross;
                    ye = ye > yr ? ye : yr;
                    const double xe = (ye - b) / m;
                    const double dlx = xe - x, dly = ye - y;
                    CB_EMIT(ix * ny + iy, sqrt(dlx * dlx + dly * dly));
                    x = xe;
                    y = ye;
                    if (fabs(y - radon_edge(iy)) < CB_RADON_TINY) --iy;
                }
            }
            ++ix;
            x = radon_edge(ix);
        }
    }
#undef CB_EMIT
    return n;
}

struct RadonTraceParams {
    const double *tx, *rx;     // [nray, 2]
    int64_t nray;
    int32_t nw;
    int64_t *count;            // pass 1: [nray] entries per ray
    const int64_t *offset;     // pass 2: [nray] global start of each ray
    double *data;              // pass 2
    int64_t *indices;          // pass 2
};

template <class Ctx>
CB_BODY void radon_trace_body(const Ctx &cx, const RadonTraceParams &P) {
    const int64_t r = (int64_t)cx.bid() * cx.nthreads() + cx.tid();
    if (r >= P.nray) return;
    const double xs = P.tx[2 * r], ys = P.tx[2 * r + 1], xr = P.rx[2 * r], yr = P.rx[2 * r + 1];
    if (P.data == nullptr) {
        P.count[r] = radon_walk(xs, ys, xr, yr, P.nw, P.nw, nullptr, nullptr);
    } else {
        const int64_t o = P.offset[r];
        radon_walk(xs, ys, xr, yr, P.nw, P.nw, P.data + o, P.indices + o);
    }
}

struct RadonApplyParams {
    const double *grid;
    int64_t ld;
    const int32_t *iy0, *ix0;
    int64_t nwin;
    int32_t nw, ntheta;
    const int64_t *offset;     // [ntheta*nw + 1] global CSR row starts
    const double *data;
    const int64_t *indices;
    double *sino;              // [nwin, nw, ntheta]
    int32_t window_fastest;    // thread order: 1 = consecutive threads take consecutive windows
};

// sinogram[w, r, t] = sum_q data[q] * window_w.flat[indices[q]], accumulated sequentially in storage
// order like scipy's csr_matvec (radon.pyx:114-115).  Cell id ix*ny+iy addresses the C-order
// flattened window, i.e. row ix, column iy.
template <class Ctx>
CB_BODY void radon_apply_body(const Ctx &cx, const RadonApplyParams &P) {
    const int64_t nrow = (int64_t)P.ntheta * P.nw;
    const int64_t total = nrow * P.nwin;
    const int64_t g = (int64_t)cx.bid() * cx.nthreads() + cx.tid();
    if (g >= total) return;
    int64_t w, row;
    if (P.window_fastest) { w = g % P.nwin; row = g / P.nwin; }
    else { row = g % nrow; w = g / nrow; }
    const int64_t t = row / P.nw, r = row - t * P.nw;
    const double *win = P.grid + (int64_t)P.iy0[w] * P.ld + P.ix0[w];
    double acc = 0.0;
    for (int64_t q = P.offset[row]; q < P.offset[row + 1]; ++q) {
        const int64_t c = P.indices[q];
        const int64_t a = c / P.nw, b = c - a * P.nw;
        acc += P.data[q] * win[a * P.ld + b];
    }
    P.sino[(w * P.nw + r) * P.ntheta + t] = acc;
}

struct AzTailParams {
    const double *sino;        // [nwin, nw, ntheta]
    int32_t nw, ntheta, nkb;
    int64_t nwin;
    const double *taper;       // [nw] or NULL
    double power;
    double *S;                 // [nwin, ntheta, nkb]
};

// One CTA per (window, angle): S[j] = power * log | sum_r taper[r] x[r] exp(-2 pi i r (j+1) / nk) |,
// nk = 1 + 2 nkb, r < min(nw, nk)  (np.fft.fft(..., n=nk) crops or zero-pads, grid.py:398-401).
// Shared memory: nk complex roots of unity, then the tapered samples.
template <class Ctx>
CB_BODY void az_tail_body(const Ctx &cx, const AzTailParams &P) {
    const int nk = 1 + 2 * P.nkb;
    const int nr = P.nw < nk ? P.nw : nk;
    const int64_t wt = cx.bid();
    const int64_t w = wt / P.ntheta, t = wt - w * P.ntheta;
    double2 *root = reinterpret_cast<double2 *>(cx.smem());
    double *x = reinterpret_cast<double *>(root + nk);
    for (int m = cx.tid(); m < nk; m += cx.nthreads()) {
        const double a = 2.0 * CB_PI_RADON * (double)m / (double)nk;
        root[m] = cplx(cos(a), -sin(a));
    }
    for (int r = cx.tid(); r < nr; r += cx.nthreads()) {
        const double v = P.sino[(w * P.nw + r) * P.ntheta + t];
        x[r] = P.taper ? P.taper[r] * v : v;
    }
    cx.sync_block();
    for (int j = cx.tid(); j < P.nkb; j += cx.nthreads()) {
        double re = 0.0, im = 0.0;
        int m = 0;
        for (int r = 0; r < nr; ++r) {
            re += x[r] * root[m].x;
            im += x[r] * root[m].y;
            m += j + 1;
            if (m >= nk) m -= nk;
        }
        P.S[(w * P.ntheta + t) * P.nkb + j] = P.power * log(sqrt(re * re + im * im));
    }
}
