// Spectrum stage kernel body: window gather -> [plane detrend] -> taper -> 2-D DFT of the real
// window (row pairs packed into one complex transform, Hermitian half plane only) ->
// log-modulus -> radial bin mean / std.
//
// Replaces, for a whole batch of windows, the reference's
//   CurieGrid.subgrid            pycurious/grid.py:112-157   (the gather; index rule on the host)
//   CurieGrid.remove_trend_linear   grid.py:199-216          (closed-form plane, optional)
//   CurieGrid._taper_spectrum    grid.py:218-246             (outer-product taper)
//   CurieGrid._FFT_spectrum      grid.py:264-288             (fft2, abs, fftshift, bin loop)
//
// Work decomposition: one persistent CTA owns one window at a time; each warp of the CTA is an
// FFT "team" with a private shared-memory line.
//   phase A  row pairs (y, y+1): gather + detrend + taper + chirp, one complex transform, untangle
//            the two real-row spectra, store G^T[v][y] (v = 0..n/2) to the CTA's L2-resident slab
//   phase B  columns v: transform G^T[v][:], store log|F|[v][u] to the slab
//   phase C  one warp per radial bin: warp-shuffle segmented reduction over the host-built CSR
#pragma once
#include "exec_ctx.cuh"
#include "fft_engine.cuh"
#include "fastlog.cuh"

struct SpectrumParams {
    // input
    const double *grid;       // row-major, leading dimension ld
    int64_t ld;
    const int32_t *iy0, *ix0; // window origins
    int64_t nwin;
    // geometry / plan
    int32_t n;                // window edge
    int32_t nvh;              // n/2 + 1 half-plane columns
    int32_t npad;             // n rounded up to even (row stride of G^T)
    int32_t bluestein;        // 1: chirp-z on length fft.n; 0: direct, fft.n == n
    FftPlan fft;
    const double2 *tw;        // fft.tw_count twiddles
    const double2 *chirp;     // n           (bluestein)
    const double2 *filt;      // padded line (bluestein)
    const int32_t *slot;      // n: slot of X[k] after DIF (direct mode)
    const double *taper_r, *taper_c;   // n each, or NULL
    const int32_t *bin_ptr, *bin_idx;  // folded CSR, ids v*n+u
    const int32_t *col_span;           // [nvh] (lo_end | hi_start << 16): u <= lo_end or u >= hi_start belong to a bin (0xffff: none)
    int32_t nbins;
    int32_t detrend;
    double power;
    // output
    double *phi, *sigma;      // [nwin, nbins]
    // scratch: per-CTA slabs
    double2 *gt;              // [nvh][npad]
    double *lh;               // [nvh][n]
    int64_t gt_stride, lh_stride;   // elements per CTA
    int64_t lh_half;                // colspec_body: offset of the CTA's second log-modulus slab
    // shared-memory carve-up (in double2 slots)
    int32_t sm_tw, sm_filt, sm_chirp, sm_cw, sm_buf, line_slots;
    // shared-row-spectra path (rowspec_body / colspec_body): windows with the same ix0 share the
    // x-transforms of the grid rows they overlap
    const int32_t *group_of_win;   // [nwin] group (= distinct ix0) of each window
    const int32_t *win_order;      // [nwin] or NULL: processing order (windows of one group together, by iy0)
    const int32_t *group_ix0;      // [ngroups]
    int32_t ngroups;
    int32_t row_lo;                // first grid row held in rt
    int32_t nrows;                 // rows per (group, v) line of rt, even
    int32_t nrows_valid;           // rows [row_lo, row_lo + nrows_valid) exist in the grid
    double2 *rt;                   // [ngroups][nvh][nrows] half spectra of (grid row x column taper)
    const double2 *plane_t0, *plane_t1;   // [nvh] x-transforms of taper_c and (x - ctr) taper_c (detrend)
    double2 *rowsum;               // [ngroups][nrows] (sum_x d, sum_x (x - ctr) d) of each row segment (detrend) or NULL
};

CB_HD double2 bluestein_out(const double2 *buf, const double2 *chirp, int k) {
    // X[k] = chirp[k] * (a (*) b)[k]; the line holds the conjugate of the convolution
    return cmul(chirp[k], cconj(buf[CB_PAD(k)]));
}

// Least-squares plane of one window in closed form (grid.py:199-216): c0 + cx1 (x - ctr) + cy1 (y - ctr).
// Block-wide; s_red = 96 doubles of shared scratch.
template <class Ctx>
CB_BODY void window_plane(const Ctx &cx, const SpectrumParams &P, const double *src, double *s_red, int warp, int lane,
                          int nwarps, double &c0, double &cx1, double &cy1) {
    const int n = P.n;
    const double ctr = 0.5 * (double)(n - 1);
    double s0 = 0.0, sx = 0.0, sy = 0.0;
    for (int y = warp; y < n; y += nwarps) {
        const double *row = src + (int64_t)y * P.ld;
        double r0 = 0.0, rx = 0.0;
        for (int x = lane; x < n; x += CB_WARP) {
            double d = Ctx::ldg(row + x);
            r0 += d;
            rx = fma((double)x - ctr, d, rx);
        }
        s0 += r0;
        sx += rx;
        sy = fma((double)y - ctr, r0, sy);
    }
    s0 = cx.warp_sum(s0);
    sx = cx.warp_sum(sx);
    sy = cx.warp_sum(sy);
    if (lane == 0) { s_red[warp] = s0; s_red[32 + warp] = sx; s_red[64 + warp] = sy; }
    cx.sync_block();
    s0 = sx = sy = 0.0;
    for (int i = 0; i < nwarps; ++i) { s0 += s_red[i]; sx += s_red[32 + i]; sy += s_red[64 + i]; }
    const double nn = (double)n;
    const double sxx = nn * nn * (nn * nn - 1.0) / 12.0;   // sum over the window of (x-ctr)^2
    c0 = s0 / (nn * nn);
    cx1 = sx / sxx;
    cy1 = sy / sxx;
    cx.sync_block();
}

// The same plane from the per-row sums rowspec_body leaves for every (group, grid row): the windows of a group
// share their row segments exactly like they share the row spectra, so a window adds n pairs instead of reading
// its n^2 nodes again.  `rs` points at the window's first row.
template <class Ctx>
CB_BODY void window_plane_rows(const Ctx &cx, const SpectrumParams &P, const double2 *rs, double *s_red, int warp,
                               int lane, int nwarps, double &c0, double &cx1, double &cy1) {
    const int n = P.n;
    const double ctr = 0.5 * (double)(n - 1);
    double s0 = 0.0, sx = 0.0, sy = 0.0;
    for (int y = warp * CB_WARP + lane; y < n; y += nwarps * CB_WARP) {
        const double2 r = Ctx::ldcg(rs + y);
        s0 += r.x;
        sx += r.y;
        sy = fma((double)y - ctr, r.x, sy);
    }
    s0 = cx.warp_sum(s0);
    sx = cx.warp_sum(sx);
    sy = cx.warp_sum(sy);
    if (lane == 0) { s_red[warp] = s0; s_red[32 + warp] = sx; s_red[64 + warp] = sy; }
    cx.sync_block();
    s0 = sx = sy = 0.0;
    for (int i = 0; i < nwarps; ++i) { s0 += s_red[i]; sx += s_red[32 + i]; sy += s_red[64 + i]; }
    const double nn = (double)n;
    const double sxx = nn * nn * (nn * nn - 1.0) / 12.0;   // sum over the window of (x-ctr)^2
    c0 = s0 / (nn * nn);
    cx1 = sx / sxx;
    cy1 = sy / sxx;
    cx.sync_block();
}

// Phase C: radial bins (grid.py:281-286), mean / population std of the members of each bin, one
// warp per bin over the folded CSR.  The members are gathered through two dependent L2 loads (CSR
// index, then the log-power), so the walk over 128-member chunks is software pipelined two deep: while
// chunk c is reduced, the values of chunk c+1 and the indices of chunk c+2 -- of this bin or of the
// warp's next bins -- are in flight.  Sum and sum of squares are taken about a pivot member (the first
// one): the population variance is shift invariant and the pivot removes the mean^2 cancellation.
struct BinChunk { int b, lo, hi, eb; };

template <class Ctx>
CB_BODY void bin_phase(const Ctx &cx, const SpectrumParams &P, const double *Lh, int64_t w, int warp, int lane,
                       int nwarps) {
    constexpr int CHUNK = 4 * CB_WARP;
    if (warp >= P.nbins) return;
    auto next_of = [&](BinChunk c) -> BinChunk {
        c.eb += CHUNK;
        if (c.eb >= c.hi) {
            c.b += nwarps;
            c.lo = c.hi = 0;
            if (c.b < P.nbins) { c.lo = Ctx::ldg(P.bin_ptr + c.b); c.hi = Ctx::ldg(P.bin_ptr + c.b + 1); }
            c.eb = c.lo;
        }
        return c;
    };
    auto load_idx = [&](const BinChunk &c, int (&idx)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = c.eb + j * CB_WARP + lane;
            idx[j] = e < c.hi ? Ctx::ldg(P.bin_idx + e) : -1;
        }
    };
    auto load_val = [&](const int (&idx)[4], double (&v)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = idx[j] >= 0 ? Ctx::ldcg(Lh + idx[j]) : 0.0;
    };
    BinChunk c0{warp, Ctx::ldg(P.bin_ptr + warp), Ctx::ldg(P.bin_ptr + warp + 1), 0};
    c0.eb = c0.lo;
    int idx0[4], idx1[4];
    double v0[4];
    load_idx(c0, idx0);
    BinChunk c1 = next_of(c0);
    load_idx(c1, idx1);
    load_val(idx0, v0);
    double s = 0.0, q = 0.0, pivot = 0.0;
    while (c0.b < P.nbins) {
        const BinChunk c2 = next_of(c1);
        int idx2[4];
        double v1[4];
        load_idx(c2, idx2);
        load_val(idx1, v1);
        if (c0.eb == c0.lo) {
            pivot = cx.shfl(v0[0], 0);               // member `lo` (an empty bin has no pivot: 0)
            if (!isfinite(pivot) || c0.lo >= c0.hi) pivot = 0.0;   // log(0) = -inf members: plain sums, numpy's -inf / nan result
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double d = idx0[j] >= 0 ? v0[j] - pivot : 0.0;
            s += d;
            q = fma(d, d, q);
        }
        if (c0.eb + CHUNK >= c0.hi) {                // last chunk of bin c0.b
            s = cx.warp_sum(s);
            q = cx.warp_sum(q);
            const double cnt = (double)(c0.hi - c0.lo);
            const double md = s / cnt;                  // mean - pivot
            const double mean = pivot + md;
            q -= s * md;                                // sum (x - mean)^2
            if (q < 0.0) q = 0.0;                       // (a nan stays a nan)
            if (lane == 0) {
                P.phi[w * P.nbins + c0.b] = P.power * mean;
                P.sigma[w * P.nbins + c0.b] = fabs(P.power) * sqrt(q / cnt);
            }
            s = q = 0.0;
        }
        c0 = c1; c1 = c2;
#pragma unroll
        for (int j = 0; j < 4; ++j) { idx0[j] = idx1[j]; idx1[j] = idx2[j]; v0[j] = v1[j]; }
    }
}

template <class Engine, class Ctx>
CB_BODY void spectrum_body(const Ctx &cx, const SpectrumParams &P) {
    const int tid = cx.tid(), nthr = cx.nthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int n = P.n, nvh = P.nvh, npad = P.npad, M = P.fft.n;
    double2 *sm = reinterpret_cast<double2 *>(cx.smem());
    double2 *s_tw = sm + P.sm_tw;
    double2 *s_filt = sm + P.sm_filt;
    double2 *s_chirp = sm + P.sm_chirp;
    double2 *s_cw = sm + P.sm_cw;
    double2 *buf = sm + P.sm_buf + (int64_t)warp * P.line_slots;
    double *s_red = reinterpret_cast<double *>(sm);   // 3*32 doubles reserved in front of sm_tw

    for (int i = tid; i < P.fft.tw_count; i += nthr) s_tw[i] = P.tw[i];
    if (P.bluestein) {
        for (int i = tid; i < P.line_slots; i += nthr) s_filt[i] = P.filt[i];
        for (int i = tid; i < n; i += nthr) s_chirp[i] = P.chirp[i];
    }
    for (int i = tid; i < n; i += nthr) {
        const double tc = P.taper_c ? P.taper_c[i] : 1.0;
        s_cw[i] = P.bluestein ? cscale(P.chirp[i], tc) : cplx(tc, 0.0);
    }
    cx.sync_block();

    double2 *Gt = P.gt + (int64_t)cx.bid() * P.gt_stride;
    double *Lh = P.lh + (int64_t)cx.bid() * P.lh_stride;
    const double ctr = 0.5 * (double)(n - 1);

    for (int64_t w = cx.bid(); w < P.nwin; w += cx.nblocks()) {
        const double *src = P.grid + (int64_t)P.iy0[w] * P.ld + P.ix0[w];

        // ---- optional least-squares plane (centred regressors are orthogonal) -------------
        double c0 = 0.0, cx1 = 0.0, cy1 = 0.0;
        if (P.detrend) window_plane(cx, P, src, s_red, warp, lane, nwarps, c0, cx1, cy1);

        // ---- phase A: row pairs ---------------------------------------------------------------
        const int npairs = (n + 1) >> 1;
        for (int p = warp; p < npairs; p += nwarps) {
            const int y0 = 2 * p, y1 = y0 + 1;
            const bool has1 = y1 < n;
            const double *r0 = src + (int64_t)y0 * P.ld;
            const double *r1 = src + (int64_t)(has1 ? y1 : y0) * P.ld;
            const double t0 = P.taper_r ? Ctx::ldg(P.taper_r + y0) : 1.0;
            const double t1 = (P.taper_r && has1) ? Ctx::ldg(P.taper_r + y1) : 1.0;
            const double pl0 = c0 + cy1 * ((double)y0 - ctr), pl1 = c0 + cy1 * ((double)y1 - ctr);
            bool fused = false;
            if constexpr (Engine::FUSED_SPAN > 0) if (P.bluestein) {
                fused = true;
                // the first FFT pass pulls its inputs straight from the grid (coalesced: lane j reads
                // x = j + t*SUB), the last one leaves the n live outputs in the line
                auto ld = [&](int x) -> double2 {
                    if (x >= n) return cplx(0.0, 0.0);
                    double av = Ctx::ldg(r0 + x);
                    double bv = has1 ? Ctx::ldg(r1 + x) : 0.0;
                    if (P.detrend) {
                        const double px = cx1 * ((double)x - ctr);
                        av -= pl0 + px;
                        bv = has1 ? bv - (pl1 + px) : 0.0;
                    }
                    return cmul(cplx(av * t0, bv * t1), s_cw[x]);
                };
                auto sk = [&](int k, double2 z) {
                    if (k < n) buf[CB_PAD(k)] = z;
                };
                Engine::convolve_io(cx, buf, s_tw, s_filt, lane, s_chirp, n, ld, sk);
                cx.sync_warp();
            }
            if (!fused) {
                // all global loads of a chunk are issued before the first use (4 x 2 loads in flight per lane)
                for (int xb = 0; xb < n; xb += 4 * CB_WARP) {
                    double a[4], b[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int x = xb + j * CB_WARP + lane;
                        a[j] = x < n ? Ctx::ldg(r0 + x) : 0.0;
                        b[j] = (x < n && has1) ? Ctx::ldg(r1 + x) : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int x = xb + j * CB_WARP + lane;
                        if (x < n) {
                            double av = a[j], bv = b[j];
                            if (P.detrend) {
                                const double px = cx1 * ((double)x - ctr);
                                av -= pl0 + px;
                                bv = has1 ? bv - (pl1 + px) : 0.0;
                            }
                            // s_cw = column taper (x chirp for Bluestein), staged once per CTA
                            const double2 cw = s_cw[x];
                            buf[CB_PAD(x)] = P.bluestein ? cmul(cplx(av * t0, bv * t1), cw) : cplx(av * t0 * cw.x, bv * t1 * cw.x);
                        }
                    }
                }
                for (int x = n + lane; x < M; x += CB_WARP) buf[CB_PAD(x)] = cplx(0.0, 0.0);
                if (P.bluestein) Engine::convolve(cx, buf, s_tw, s_filt, P.fft, lane);
                else Engine::forward(cx, buf, s_tw, P.fft, lane);
            }
            // untangle: row y0 spectrum A[v] = (Z[v] + conj Z[-v])/2, row y1: B[v] = (Z[v] - conj Z[-v])/(2i)
            for (int v = lane; v < nvh; v += CB_WARP) {
                const int vm = v ? n - v : 0;
                double2 zv, zm;
                if (P.bluestein) {
                    zv = bluestein_out(buf, s_chirp, v);
                    zm = bluestein_out(buf, s_chirp, vm);
                } else {
                    zv = buf[CB_PAD(Ctx::ldg(P.slot + v))];
                    zm = buf[CB_PAD(Ctx::ldg(P.slot + vm))];
                }
                double2 A = cplx(0.5 * (zv.x + zm.x), 0.5 * (zv.y - zm.y));
                double2 B = cplx(0.5 * (zv.y + zm.y), -0.5 * (zv.x - zm.x));
                double2 *dst = Gt + (int64_t)v * npad + y0;
                Ctx::stcg(dst, A);
                if (has1) Ctx::stcg(dst + 1, B);
            }
            cx.sync_warp();
        }
        cx.sync_block();

        // ---- phase B: columns -------------------------------------------------------------------
        for (int v = warp; v < nvh; v += nwarps) {
            const double2 *col = Gt + (int64_t)v * npad;
            double *out = Lh + (int64_t)v * n;
            bool fused = false;
            if constexpr (Engine::FUSED_SPAN > 0) if (P.bluestein) {
                fused = true;
                auto ld = [&](int y) -> double2 {
                    if (y >= n) return cplx(0.0, 0.0);
                    return cmul(Ctx::ldcg(col + y), s_chirp[y]);
                };
                // |chirp| = 1, so the modulus needs no post-multiply; outputs go straight to the slab.
                // The log is only taken where a radial bin will read it (the corners of the plane lie outside).
                const int span = Ctx::ldg(P.col_span + v);
                const int lo_end = (span & 0xffff) == 0xffff ? -1 : (span & 0xffff), hi_start = (int)((unsigned)span >> 16);
                auto sk = [&](int u, double2 z) {
                    if ((u <= lo_end || u >= hi_start) && u < n) Ctx::stcg(out + u, cb_half_log(fma(z.x, z.x, z.y * z.y)));
                };
                Engine::convolve_io(cx, buf, s_tw, s_filt, lane, s_chirp, n, ld, sk);
            }
            if (!fused) {
                for (int yb = 0; yb < n; yb += 4 * CB_WARP) {
                    double2 z[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int y = yb + j * CB_WARP + lane;
                        z[j] = y < n ? Ctx::ldcg(col + y) : cplx(0.0, 0.0);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int y = yb + j * CB_WARP + lane;
                        if (y < n) buf[CB_PAD(y)] = P.bluestein ? cmul(z[j], s_chirp[y]) : z[j];
                    }
                }
                for (int y = n + lane; y < M; y += CB_WARP) buf[CB_PAD(y)] = cplx(0.0, 0.0);
                if (P.bluestein) Engine::convolve(cx, buf, s_tw, s_filt, P.fft, lane);
                else Engine::forward(cx, buf, s_tw, P.fft, lane);
                for (int u = lane; u < n; u += CB_WARP) {
                    // |chirp| = 1, so the modulus needs no post-multiply
                    const double2 z = P.bluestein ? buf[CB_PAD(u)] : buf[CB_PAD(Ctx::ldg(P.slot + u))];
                    Ctx::stcg(out + u, cb_half_log(fma(z.x, z.x, z.y * z.y)));
                }
            }
            cx.sync_warp();
        }
        cx.sync_block();

        // ---- phase C: radial bins ----------------------------------------------------------------
        bin_phase(cx, P, Lh, w, warp, lane, nwarps);
        // phase C of this window may overlap phase A of the next: A writes only Gt, C reads only Lh;
        // the barrier after the next phase A orders C(w) before B(w+1).
    }
}


// =====================================================================================
// Shared row spectra.  On a centroid lattice every window that starts at the same column ix0
// needs the x-transform of (grid row x column taper) for the rows it overlaps -- the row taper is a
// per-row scalar and the detrending plane a rank-2 correction, so they commute with that transform.
// A 201-node window on a 10-node stride shares each row with 20 other windows:
//   rowspec_body  one warp per (group = distinct ix0, row pair): the phase-A transform, ONCE, into
//                 rt[group][v][row]   (HBM resident: config 3 = 390 x 101 x 4096 x 16 B = 2.6 GB)
//   colspec_body  one CTA per window: phase B reads its 201-row slice of rt[group][v] (contiguous),
//                 applies row taper, chirp and the plane correction, transforms, bins (phase C)
// Per window this halves the FFT work; the row pass costs 1/20 of the old phase A.
// =====================================================================================
template <class Ctx>
CB_BODY void spectrum_stage_tables(const Ctx &cx, const SpectrumParams &P, double2 *s_tw, double2 *s_filt,
                                   double2 *s_chirp, double2 *s_cw, const double *taper) {
    const int tid = cx.tid(), nthr = cx.nthreads(), n = P.n;
    for (int i = tid; i < P.fft.tw_count; i += nthr) s_tw[i] = P.tw[i];
    if (P.bluestein) {
        for (int i = tid; i < P.line_slots; i += nthr) s_filt[i] = P.filt[i];
        for (int i = tid; i < n; i += nthr) s_chirp[i] = P.chirp[i];
    }
    for (int i = tid; i < n; i += nthr) {
        const double t = taper ? taper[i] : 1.0;
        s_cw[i] = P.bluestein ? cscale(P.chirp[i], t) : cplx(t, 0.0);
    }
    cx.sync_block();
}

template <class Engine, class Ctx>
CB_BODY void rowspec_body(const Ctx &cx, const SpectrumParams &P) {
    const int tid = cx.tid(), nthr = cx.nthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int n = P.n, nvh = P.nvh, M = P.fft.n;
    double2 *sm = reinterpret_cast<double2 *>(cx.smem());
    double2 *s_tw = sm + P.sm_tw, *s_filt = sm + P.sm_filt, *s_chirp = sm + P.sm_chirp, *s_cw = sm + P.sm_cw;
    double2 *buf = sm + P.sm_buf + (int64_t)warp * P.line_slots;
    spectrum_stage_tables(cx, P, s_tw, s_filt, s_chirp, s_cw, P.taper_c);

    const int npairs = P.nrows >> 1;
    const int64_t total = (int64_t)P.ngroups * npairs;
    for (int64_t item = (int64_t)cx.bid() * nwarps + warp; item < total; item += (int64_t)cx.nblocks() * nwarps) {
        const int g = (int)(item / npairs), p = (int)(item - (int64_t)g * npairs);
        const bool has0 = 2 * p < P.nrows_valid, has1 = 2 * p + 1 < P.nrows_valid;
        const double *r0 = P.grid + (int64_t)(P.row_lo + (has0 ? 2 * p : 0)) * P.ld + Ctx::ldg(P.group_ix0 + g);
        const double *r1 = has1 ? r0 + P.ld : r0;
        if (P.rowsum) {                          // detrend: zeroth and first moment of both row segments
            const double ctr = 0.5 * (double)(n - 1);
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            for (int x = lane; x < n; x += CB_WARP) {
                const double av = has0 ? Ctx::ldg(r0 + x) : 0.0, bv = has1 ? Ctx::ldg(r1 + x) : 0.0;
                const double xc = (double)x - ctr;
                a0 += av; a1 = fma(xc, av, a1);
                b0 += bv; b1 = fma(xc, bv, b1);
            }
            a0 = cx.warp_sum(a0); a1 = cx.warp_sum(a1); b0 = cx.warp_sum(b0); b1 = cx.warp_sum(b1);
            if (lane == 0) {
                double2 *rs = P.rowsum + (int64_t)g * P.nrows + 2 * p;
                rs[0] = cplx(a0, a1);
                rs[1] = cplx(b0, b1);
            }
        }
        bool fused = false;
        if constexpr (Engine::FUSED_SPAN > 0) if (P.bluestein) {
            fused = true;
            auto ld = [&](int x) -> double2 {
                if (x >= n) return cplx(0.0, 0.0);
                const double av = has0 ? Ctx::ldg(r0 + x) : 0.0, bv = has1 ? Ctx::ldg(r1 + x) : 0.0;
                return cmul(cplx(av, bv), s_cw[x]);
            };
            auto sk = [&](int k, double2 z) {
                if (k < n) buf[CB_PAD(k)] = z;
            };
            Engine::convolve_io(cx, buf, s_tw, s_filt, lane, s_chirp, n, ld, sk);
            cx.sync_warp();
        }
        if (!fused) {
            cx.sync_warp();
            for (int x = lane; x < M; x += CB_WARP) {
                double2 z = cplx(0.0, 0.0);
                if (x < n) {
                    const double av = has0 ? Ctx::ldg(r0 + x) : 0.0, bv = has1 ? Ctx::ldg(r1 + x) : 0.0;
                    const double2 cw = s_cw[x];
                    z = P.bluestein ? cmul(cplx(av, bv), cw) : cplx(av * cw.x, bv * cw.x);
                }
                buf[CB_PAD(x)] = z;
            }
            if (P.bluestein) Engine::convolve(cx, buf, s_tw, s_filt, P.fft, lane);
            else Engine::forward(cx, buf, s_tw, P.fft, lane);
        }
        // untangle the two real rows; rt[g][v][2p], rt[g][v][2p+1] are adjacent (one 32-byte sector)
        double2 *dst0 = P.rt + ((int64_t)g * nvh) * P.nrows + 2 * p;
        for (int v = lane; v < nvh; v += CB_WARP) {
            const int vm = v ? n - v : 0;
            double2 zv, zm;
            if (P.bluestein) {
                zv = bluestein_out(buf, s_chirp, v);
                zm = bluestein_out(buf, s_chirp, vm);
            } else {
                zv = buf[CB_PAD(Ctx::ldg(P.slot + v))];
                zm = buf[CB_PAD(Ctx::ldg(P.slot + vm))];
            }
            double2 *dst = dst0 + (int64_t)v * P.nrows;
            Ctx::stcg(dst, cplx(0.5 * (zv.x + zm.x), 0.5 * (zv.y - zm.y)));
            Ctx::stcg(dst + 1, cplx(0.5 * (zv.y + zm.y), -0.5 * (zv.x - zm.x)));
        }
        cx.sync_warp();
    }
}

template <class Engine, class Ctx>
CB_BODY void colspec_body(const Ctx &cx, const SpectrumParams &P) {
    const int tid = cx.tid(), nthr = cx.nthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int n = P.n, nvh = P.nvh, M = P.fft.n;
    double2 *sm = reinterpret_cast<double2 *>(cx.smem());
    double2 *s_tw = sm + P.sm_tw, *s_filt = sm + P.sm_filt, *s_chirp = sm + P.sm_chirp, *s_rw = sm + P.sm_cw;
    double2 *buf = sm + P.sm_buf + (int64_t)warp * P.line_slots;
    double *s_red = reinterpret_cast<double *>(sm);
    spectrum_stage_tables(cx, P, s_tw, s_filt, s_chirp, s_rw, P.taper_r);     // s_rw = row taper (x chirp)

    // two log-modulus slabs per CTA: the bins of window w (phase C) are read from one while the columns
    // of window w+1 (phase B) already fill the other -- one block barrier per window, and the gather
    // latency of phase C hides behind the transforms of the warps that are already ahead
    double *Lh2[2] = {P.lh + (int64_t)cx.bid() * P.lh_stride, P.lh + (int64_t)cx.bid() * P.lh_stride + P.lh_half};
    const double ctr = 0.5 * (double)(n - 1);
    int parity = 0;
    for (int64_t wi = cx.bid(); wi < P.nwin; wi += cx.nblocks(), parity ^= 1) {
        // CTAs that run side by side take neighbouring windows of the SAME column group: their row slices of
        // rt overlap by (nw - stride)/nw, so the slab is read from L2 instead of HBM
        const int64_t w = P.win_order ? (int64_t)P.win_order[wi] : wi;
        double *Lh = Lh2[parity];
        const int iy0 = P.iy0[w], g = P.group_of_win[w];
        double c0 = 0.0, cx1 = 0.0, cy1 = 0.0;
        if (P.detrend)
        {
            if (P.rowsum) window_plane_rows(cx, P, P.rowsum + (int64_t)g * P.nrows + (iy0 - P.row_lo), s_red, warp, lane, nwarps, c0, cx1, cy1);
            else window_plane(cx, P, P.grid + (int64_t)iy0 * P.ld + P.ix0[w], s_red, warp, lane, nwarps, c0, cx1, cy1);
        }
        const double2 *rt_w = P.rt + ((int64_t)g * nvh) * P.nrows + (iy0 - P.row_lo);
        if constexpr (Engine::PREFETCH) if (P.bluestein) {
            // Static engines: the raw row-spectrum slice of the NEXT column is requested before the current
            // column is transformed, so its L2 / HBM latency hides behind a whole column's passes (ncu r01s:
            // 18 % of this kernel's stall samples sat on the first use of these loads).
            double2 nxt[Engine::NIT][Engine::HALF];
            auto raw_of = [&](const double2 *c) {
                return [=](int y) -> double2 { return y < n ? Ctx::ldcg(c + y) : cplx(0.0, 0.0); };
            };
            if (warp < nvh) Engine::gather(nxt, lane, raw_of(rt_w + (int64_t)warp * P.nrows));
            for (int v = warp; v < nvh; v += nwarps) {
                double2 cur[Engine::NIT][Engine::HALF];
#pragma unroll
                for (int it = 0; it < Engine::NIT; ++it)
#pragma unroll
                    for (int t = 0; t < Engine::HALF; ++t) cur[it][t] = nxt[it][t];
                if (v + nwarps < nvh) Engine::gather(nxt, lane, raw_of(rt_w + (int64_t)(v + nwarps) * P.nrows));
                double2 t0 = cplx(0.0, 0.0), t1 = cplx(0.0, 0.0);
                if (P.detrend) {
                    t0 = Ctx::ldg(P.plane_t0 + v);
                    t1 = cscale(Ctx::ldg(P.plane_t1 + v), cx1);
                }
#pragma unroll
                for (int it = 0; it < Engine::NIT; ++it)
#pragma unroll
                    for (int t = 0; t < Engine::HALF; ++t) {
                        const int y = Engine::input_index(it, t, lane);
                        double2 z = cur[it][t];
                        if (y < n) {
                            if (P.detrend) {
                                const double a = c0 + cy1 * ((double)y - ctr);
                                z = cplx(z.x - fma(a, t0.x, t1.x), z.y - fma(a, t0.y, t1.y));
                            }
                            z = cmul(z, s_rw[y]);
                        }
                        cur[it][t] = z;
                    }
                double *out = Lh + (int64_t)v * n;
                const int span = Ctx::ldg(P.col_span + v);
                const int lo_end = (span & 0xffff) == 0xffff ? -1 : (span & 0xffff), hi_start = (int)((unsigned)span >> 16);
                auto sk = [&](int u, double2 z) {
                    if ((u <= lo_end || u >= hi_start) && u < n) Ctx::stcg(out + u, cb_half_log(fma(z.x, z.x, z.y * z.y)));
                };
                Engine::convolve_pre(cx, buf, s_tw, s_filt, lane, cur, sk);
                cx.sync_warp();
            }
        }
        if (!(Engine::PREFETCH && P.bluestein))
        for (int v = warp; v < nvh; v += nwarps) {
            const double2 *col = rt_w + (int64_t)v * P.nrows;
            double *out = Lh + (int64_t)v * n;
            double2 t0 = cplx(0.0, 0.0), t1 = cplx(0.0, 0.0);
            if (P.detrend) {
                t0 = Ctx::ldg(P.plane_t0 + v);
                t1 = cscale(Ctx::ldg(P.plane_t1 + v), cx1);
            }
            // row y of the window, x-transformed, minus the transformed plane, times the row taper
            auto element = [&](int y) -> double2 {
                double2 z = Ctx::ldcg(col + y);
                if (P.detrend) {
                    const double a = c0 + cy1 * ((double)y - ctr);
                    z = cplx(z.x - fma(a, t0.x, t1.x), z.y - fma(a, t0.y, t1.y));
                }
                const double2 rw = s_rw[y];
                return P.bluestein ? cmul(z, rw) : cscale(z, rw.x);
            };
            bool fused = false;
            if constexpr (Engine::FUSED_SPAN > 0) if (P.bluestein) {
                fused = true;
                auto ld = [&](int y) -> double2 { return y < n ? element(y) : cplx(0.0, 0.0); };
                const int span = Ctx::ldg(P.col_span + v);
                const int lo_end = (span & 0xffff) == 0xffff ? -1 : (span & 0xffff), hi_start = (int)((unsigned)span >> 16);
                auto sk = [&](int u, double2 z) {
                    if ((u <= lo_end || u >= hi_start) && u < n) Ctx::stcg(out + u, cb_half_log(fma(z.x, z.x, z.y * z.y)));
                };
                Engine::convolve_io(cx, buf, s_tw, s_filt, lane, s_chirp, n, ld, sk);
            }
            if (!fused) {
                cx.sync_warp();
                for (int y = lane; y < M; y += CB_WARP) buf[CB_PAD(y)] = y < n ? element(y) : cplx(0.0, 0.0);
                if (P.bluestein) Engine::convolve(cx, buf, s_tw, s_filt, P.fft, lane);
                else Engine::forward(cx, buf, s_tw, P.fft, lane);
                for (int u = lane; u < n; u += CB_WARP) {
                    const double2 z = P.bluestein ? buf[CB_PAD(u)] : buf[CB_PAD(Ctx::ldg(P.slot + u))];
                    Ctx::stcg(out + u, cb_half_log(fma(z.x, z.x, z.y * z.y)));
                }
            }
            cx.sync_warp();
        }
        cx.sync_block();       // all columns of w are in Lh; every warp has left phase C of w-1 (other slab)
        bin_phase(cx, P, Lh, w, warp, lane, nwarps);
    }
}

// Detrended windows themselves (CurieGrid.remove_trend_linear as a batched op).
struct DetrendParams {
    const double *grid;
    int64_t ld;
    const int32_t *iy0, *ix0;
    int64_t nwin;
    int32_t n;
    double *out;   // [nwin, n, n]
};

template <class Ctx>
CB_BODY void detrend_body(const Ctx &cx, const DetrendParams &P) {
    const int tid = cx.tid(), nthr = cx.nthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int n = P.n;
    double *s_red = reinterpret_cast<double *>(cx.smem());
    const double ctr = 0.5 * (double)(n - 1);
    for (int64_t w = cx.bid(); w < P.nwin; w += cx.nblocks()) {
        const double *src = P.grid + (int64_t)P.iy0[w] * P.ld + P.ix0[w];
        double s0 = 0.0, sx = 0.0, sy = 0.0;
        for (int y = warp; y < n; y += nwarps) {
            const double *row = src + (int64_t)y * P.ld;
            double r0 = 0.0, rx = 0.0;
            for (int x = lane; x < n; x += CB_WARP) {
                double d = Ctx::ldg(row + x);
                r0 += d;
                rx = fma((double)x - ctr, d, rx);
            }
            s0 += r0;
            sx += rx;
            sy = fma((double)y - ctr, r0, sy);
        }
        s0 = cx.warp_sum(s0);
        sx = cx.warp_sum(sx);
        sy = cx.warp_sum(sy);
        if (lane == 0) { s_red[warp] = s0; s_red[32 + warp] = sx; s_red[64 + warp] = sy; }
        cx.sync_block();
        s0 = sx = sy = 0.0;
        for (int i = 0; i < nwarps; ++i) { s0 += s_red[i]; sx += s_red[32 + i]; sy += s_red[64 + i]; }
        const double nn = (double)n;
        const double sxx = nn * nn * (nn * nn - 1.0) / 12.0;
        const double c0 = s0 / (nn * nn), cx1 = sx / sxx, cy1 = sy / sxx;
        double *dst = P.out + w * (int64_t)n * n;
        for (int y = warp; y < n; y += nwarps) {
            const double *row = src + (int64_t)y * P.ld;
            const double py = c0 + cy1 * ((double)y - ctr);
            for (int x = lane; x < n; x += CB_WARP)
                dst[(int64_t)y * n + x] = Ctx::ldg(row + x) - (py + cx1 * ((double)x - ctr));
        }
        cx.sync_block();
    }
}
