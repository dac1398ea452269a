/*
 * CPU oracle for the Radon ray tracer  (TEST INFRASTRUCTURE ONLY).
 *
 * Plain-C restatement of the cell walker of the reference, src/cradon.c:4-246
 * (paths relative to /root/reference).  It is the checker for
 * pycurious_b200/csrc/radon.cu and is never linked into the product.
 * Parity pin: tests/golden/radon_*.npz were produced by the reference's own
 * cradon.c (built into oracle/_ref by oracle/Makefile) and
 * tests/test_oracle_golden.py compares this file against them bit for bit.
 *
 * Geometry (cradon.c:21-31): cell (ix,iy) spans [ix-0.5, ix+0.5) x [iy-0.5, iy+0.5);
 * the flattened cell id is ix*ny + iy (cradon.c:66).  Rays are ordered so that
 * x increases (cradon.c:47-55).  Three cases: |dy| < 1e-10 "horizontal"
 * (cradon.c:61-90), |dx| < 1e-10 "vertical" (cradon.c:91-126), else oblique
 * (cradon.c:127-219).  Every floating-point expression below is kept in the
 * operand order of the reference so that the emitted lengths are bit-identical.
 *
 * Two-pass use: call with data == NULL to count (fills indptr), then with
 * buffers of that size.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define EDGE(n) (-0.5 + (double)(n))        /* grx[n] / gry[n], cradon.c:21-31 */
#define TINY 1.e-10                          /* cradon.c:11 */

typedef struct {
    double  *len;
    int64_t *cell;
    int64_t  cap;
    int64_t  n;
} sink_t;

static inline void emit(sink_t *s, int64_t cell, double len)
{
    if (s->len && s->n < s->cap) {
        s->len[s->n] = len;
        s->cell[s->n] = cell;
    }
    s->n++;
}

/* first cell index whose upper edge is beyond v  (cradon.c:63-64) */
static inline int64_t locate(double v, int64_t n)
{
    int64_t i;
    for (i = 0; i < n; ++i)
        if (v < EDGE(i + 1)) break;
    return i;
}

static void walk_ray(double xs, double ys, double xr, double yr,
                     int64_t nx, int64_t ny, sink_t *out)
{
    if (xs > xr) {                           /* cradon.c:47-55 */
        double t = xs; xs = xr; xr = t;
        t = ys; ys = yr; yr = t;
    }
    double x = xs, y = ys;
    int64_t ix, iy;

    if (fabs(ys - yr) < TINY) {              /* constant y: cradon.c:61-90 */
        ix = locate(x, nx);
        iy = locate(y, ny);
        while (x < xr) {
            double stop = EDGE(ix + 1) < xr ? EDGE(ix + 1) : xr;
            emit(out, ix * ny + iy, stop - x);
            ++ix;
            x = EDGE(ix);
        }
    } else if (fabs(xs - xr) < TINY) {       /* constant x: cradon.c:91-126 */
        if (ys > yr) { double t = ys; ys = yr; yr = t; }
        y = ys;
        ix = locate(x, nx);
        iy = locate(y, ny);
        while (y < yr) {
            double stop = EDGE(iy + 1) < yr ? EDGE(iy + 1) : yr;
            emit(out, ix * ny + iy, stop - y);
            ++iy;
            y = EDGE(iy);
        }
    } else {                                 /* oblique: cradon.c:127-219 */
        double m = (yr - ys) / (xr - xs);
        double b = yr - m * xr;
        int rising = m > 0;
        ix = locate(x, nx);
        iy = locate(y, ny);
        while (x < xr) {
            double ycross = m * EDGE(ix + 1) + b;    /* y where the ray leaves column ix */
            if (rising) {
                while (y < ycross && y < yr) {
                    double ye = EDGE(iy + 1) < ycross ? EDGE(iy + 1) : ycross;
                    ye = ye < yr ? ye : yr;
                    double xe = (ye - b) / m;
                    double dlx = xe - x, dly = ye - y;
                    emit(out, ix * ny + iy, sqrt(dlx * dlx + dly * dly));
                    x = xe;
                    y = ye;
                    if (fabs(y - EDGE(iy + 1)) < TINY) ++iy;
                }
            } else {
                while (y > ycross && y > yr) {
                    double ye = EDGE(iy) > ycross ? EDGE(iy) : ycross;
                    ye = ye > yr ? ye : yr;
                    double xe = (ye - b) / m;
                    double dlx = xe - x, dly = ye - y;
                    emit(out, ix * ny + iy, sqrt(dlx * dlx + dly * dly));
                    x = xe;
                    y = ye;
                    if (fabs(y - EDGE(iy)) < TINY) --iy;
                }
            }
            ++ix;
            x = EDGE(ix);
        }
    }
}

/* Tx, Rx: [nray][2] (x,y) endpoints.  Returns nnz; indptr has nray+1 entries. */
int64_t oracle_radon_trace(const double *Tx, const double *Rx, int64_t nray,
                           int64_t nx, int64_t ny,
                           double *data, int64_t *indices, int64_t *indptr, int64_t cap)
{
    sink_t s = { data, indices, cap, 0 };
    for (int64_t r = 0; r < nray; ++r) {
        indptr[r] = s.n;
        walk_ray(Tx[2 * r], Tx[2 * r + 1], Rx[2 * r], Rx[2 * r + 1], nx, ny, &s);
    }
    indptr[nray] = s.n;
    return s.n;
}
