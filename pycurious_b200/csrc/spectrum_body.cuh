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
    int32_t nbins;
    int32_t detrend;
    double power;
    // output
    double *phi, *sigma;      // [nwin, nbins]
    // scratch: per-CTA slabs
    double2 *gt;              // [nvh][npad]
    double *lh;               // [nvh][n]
    int64_t gt_stride, lh_stride;   // elements per CTA
    // shared-memory carve-up (in double2 slots)
    int32_t sm_tw, sm_filt, sm_chirp, sm_cw, sm_buf, line_slots;
};

CB_HD double2 bluestein_out(const double2 *buf, const double2 *chirp, int k) {
    // X[k] = chirp[k] * (a (*) b)[k]; the line holds the conjugate of the convolution
    return cmul(chirp[k], cconj(buf[CB_PAD(k)]));
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
        if (P.detrend) {
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
                Engine::convolve_io(cx, buf, s_tw, s_filt, lane, ld, sk);
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
                // |chirp| = 1, so the modulus needs no post-multiply; outputs go straight to the slab
                auto sk = [&](int u, double2 z) {
                    if (u < n) Ctx::stcg(out + u, 0.5 * log(fma(z.x, z.x, z.y * z.y)));
                };
                Engine::convolve_io(cx, buf, s_tw, s_filt, lane, ld, sk);
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
                    Ctx::stcg(out + u, 0.5 * log(fma(z.x, z.x, z.y * z.y)));
                }
            }
            cx.sync_warp();
        }
        cx.sync_block();

        // ---- phase C: radial bins (grid.py:281-286), mean / population std --------------
        for (int b = warp; b < P.nbins; b += nwarps) {
            const int lo = Ctx::ldg(P.bin_ptr + b), hi = Ctx::ldg(P.bin_ptr + b + 1);
            // one pass over the members: sum and sum of squares about a pivot member (the population
            // variance is shift invariant; the pivot removes the mean^2 cancellation), 4 gathers in flight
            double pivot = lo < hi ? Ctx::ldcg(Lh + Ctx::ldg(P.bin_idx + lo)) : 0.0;
            if (!isfinite(pivot)) pivot = 0.0;          // log(0) = -inf members: plain sums, numpy's -inf / nan result
            double s = 0.0, q = 0.0;
            for (int eb = lo; eb < hi; eb += 4 * CB_WARP) {
                double v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = eb + j * CB_WARP + lane;
                    v[j] = e < hi ? Ctx::ldcg(Lh + Ctx::ldg(P.bin_idx + e)) : pivot;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double d = v[j] - pivot;      // padding lanes contribute exactly 0
                    s += d;
                    q = fma(d, d, q);
                }
            }
            s = cx.warp_sum(s);
            q = cx.warp_sum(q);
            const double cnt = (double)(hi - lo);
            const double md = s / cnt;                  // mean - pivot
            const double mean = pivot + md;
            q -= s * md;                                // sum (x - mean)^2
            if (q < 0.0) q = 0.0;                       // (a nan stays a nan)
            if (lane == 0) {
                P.phi[w * P.nbins + b] = P.power * mean;
                P.sigma[w * P.nbins + b] = fabs(P.power) * sqrt(q / cnt);
            }
        }
        // phase C of this window may overlap phase A of the next: A writes only Gt, C reads only Lh;
        // the barrier after the next phase A orders C(w) before B(w+1).
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
