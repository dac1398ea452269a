// Whole-array FFT-domain filtering: out = real(ifft2(filter * fft2(data))).
// Replaces the numpy.fft.fft2 / ifft2 pairs of CurieGrid.reduce_to_pole (pycurious/grid.py:489-490) and
// CurieGrid.upward_continuation (grid.py:545-546) -- SURVEY.md section 8f item 4.  The filter itself is
// built on the host with the reference's numpy expressions (it depends on (ny, nx, dx, dy) and a few
// scalars only) and multiplied in on the device.
//
// One generic kernel does a batch of 1-D DFTs, one warp per line, through the same shared-memory FFT engine
// as the spectrum stage (mixed radix for 2/3/5-smooth lengths, Bluestein otherwise).  A 2-D transform is
// four launches: rows, columns (x filter), columns inverse, rows inverse (-> real part).  The inverse
// uses ifft(z) = conj(fft(conj(z))) / n.  Lines are addressed by (element stride, line stride), so columns
// need no transpose pass.
//
// The same kernel serves a BATCH of equally sized images (the windows of a survey, filtered before their
// spectra are taken -- the `process_subgrid` stage of CurieOptimise.optimise, optimise.py:285-287): line
// index = image * lines_per_img + line-in-image, images `in_bs` / `out_bs` elements apart, the first pass
// gathering each window straight from the survey grid at its (iy0, ix0) origin.
#pragma once
#include "exec_ctx.cuh"
#include "fft_engine.cuh"
#include <string.h>

struct LineFftParams {
    const double *in_real;     // real input (first pass) or NULL
    const double2 *in_cplx;    // complex input or NULL
    int64_t in_es, in_ls;      // element / line strides of the input (in elements)
    double2 *out_cplx;         // complex output or NULL
    double *out_real;          // real-part output (last pass) or NULL
    int64_t out_es, out_ls;
    const double2 *mul;        // optional pointwise multiplier applied to the transform (same strides as out)
    int64_t nlines;            // over all images
    int64_t lines_per_img;     // 0: one image
    int64_t in_bs, out_bs;     // image strides (elements)
    const int32_t *iy0, *ix0;  // first pass of a window batch: image i starts at in_real + iy0[i]*in_ld + ix0[i]
    int64_t in_ld;
    int32_t n;                 // line length
    // Real data have Hermitian spectra, and a filter applied as real(ifft2(op * fft2(x))) acts through its
    // Hermitian part only; the window batch uses both (cb_fft2_filter_windows):
    //   pair = 1  forward pass over PAIRS of real lines (lines 2p, 2p+1 of an image in one complex transform,
    //             untangled); only elements k < half of each spectrum are written
    //   pair = 2  inverse pass from two Hermitian half lines (elements k < half given, the rest by conjugate
    //             symmetry) to two real lines
    //   pair = 0  plain complex lines (the column passes then run over the half columns only)
    int32_t pair, half;
    // twice = 1 (Bluestein lines, pair = 0): the two column passes of a filter in one visit -- forward transform,
    // times `mul`, inverse transform (scale applies to it), all in the warp's line; in and out may be the same array.
    // Bit for bit what the two separate passes give (the value that would travel through global memory is the same).
    int32_t twice;
    int32_t inverse;           // 1: conj in, conj out, scale
    double scale;
    int32_t bluestein;
    FftPlan fft;
    const double2 *tw, *chirp, *filt;   // global tables (filt padded like a line)
    const int32_t *slot;                // direct mode: slot of X[k]
    int32_t line_slots;
    int32_t stage_slots;                // shared-memory slots holding tw + filt in front of the lines (0: read from global)
};

// Engine = GenericFft (any length, tables read from global memory) or a StaticFft whose pass list matches the plan
// (the Bluestein lengths of the BASELINE window sizes): twiddles and Bluestein filter are then staged in shared
// memory in front of the lines (P.stage_slots double2 slots) and the passes run with compile-time strides.
template <class Engine, class Ctx>
CB_BODY void line_fft_body(const Ctx &cx, const LineFftParams &P) {
    const int tid = cx.tid(), nthr = cx.nthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int n = P.n, M = P.fft.n;
    double2 *sm = reinterpret_cast<double2 *>(cx.smem());
    const double2 *tw = P.tw, *filt = P.filt;
    if (P.stage_slots > 0) {
        for (int i = tid; i < P.fft.tw_count; i += nthr) sm[i] = P.tw[i];
        for (int i = tid; i < P.line_slots; i += nthr) sm[P.fft.tw_count + i] = P.filt[i];
        cx.sync_block();
        tw = sm;
        filt = sm + P.fft.tw_count;
    }
    double2 *buf = sm + P.stage_slots + (int64_t)warp * P.line_slots;
    for (int64_t line = (int64_t)cx.bid() * nwarps + warp; line < P.nlines; line += (int64_t)cx.nblocks() * nwarps) {
        cx.sync_warp();
        int64_t img = 0, l = line;
        if (P.lines_per_img) { img = line / P.lines_per_img; l = line - img * P.lines_per_img; }
        const int64_t l0 = P.pair ? 2 * l : l;                       // first (or only) line of this transform
        const bool has1 = P.pair && l0 + 1 < P.n;                     // (pair modes: square images, n lines of n elements)
        int64_t in0 = img * P.in_bs + l0 * P.in_ls;
        if (P.iy0) in0 += (int64_t)Ctx::ldg(P.iy0 + img) * P.in_ld + Ctx::ldg(P.ix0 + img);
        // element x < n of the line as the transform takes it (pair packing, conjugation for the inverse, chirp)
        auto element = [&](int x) -> double2 {
            double2 z;
            if (P.pair == 1) {
                z = cplx(P.in_real[in0 + (int64_t)x * P.in_es], has1 ? P.in_real[in0 + P.in_ls + (int64_t)x * P.in_es] : 0.0);
            } else if (P.pair == 2) {
                // Z[x] = A0[x] + i A1[x], the half lines extended by A[n - x] = conj A[x]
                const bool up = x >= P.half;
                const int64_t xi = up ? n - x : x;
                double2 a = P.in_cplx[in0 + xi * P.in_es];
                double2 b = has1 ? P.in_cplx[in0 + P.in_ls + xi * P.in_es] : cplx(0.0, 0.0);
                if (up) { a = cconj(a); b = cconj(b); }
                z = cplx(a.x - b.y, a.y + b.x);
            } else if (P.in_real) z = cplx(P.in_real[in0 + (int64_t)x * P.in_es], 0.0);
            else z = P.in_cplx[in0 + (int64_t)x * P.in_es];
            if (P.inverse) z = cconj(z);
            if (P.bluestein) z = cmul(z, Ctx::ldg(P.chirp + x));
            return z;
        };
        // Static engines on lines that fit their pruned passes: the first pass pulls its inputs straight from
        // `element`, the last one leaves the n live outputs in the line (no fill pass, no zero padding in memory).
        bool pruned = false;
        if constexpr (Engine::FUSED_SPAN > 0) pruned = P.bluestein && n <= Engine::FUSED_SPAN;
        auto keep = [&](int k, double2 z) {
            if (k < n) buf[CB_PAD(k)] = z;
        };
        if (pruned) {
            if constexpr (Engine::FUSED_SPAN > 0) {
                auto ld = [&](int x) -> double2 { return x < n ? element(x) : cplx(0.0, 0.0); };
                Engine::convolve_io(cx, buf, tw, filt, lane, P.chirp, n, ld, keep);
                cx.sync_warp();
            }
        } else {
            for (int x = lane; x < M; x += CB_WARP) buf[CB_PAD(x)] = x < n ? element(x) : cplx(0.0, 0.0);
            if (P.bluestein) Engine::convolve(cx, buf, tw, filt, P.fft, lane);
            else Engine::forward(cx, buf, tw, P.fft, lane);
        }
        auto X = [&](int k) -> double2 {
            return P.bluestein ? cmul(Ctx::ldg(P.chirp + k), cconj(buf[CB_PAD(k)])) : buf[CB_PAD(Ctx::ldg(P.slot + k))];
        };
        if (P.pair == 1) {
            // untangle: line l0 has A[k] = (Z[k] + conj Z[-k]) / 2, line l0 + 1 has B[k] = (Z[k] - conj Z[-k]) / (2i)
            for (int k = lane; k < P.half; k += CB_WARP) {
                const double2 zv = X(k), zm = X(k ? n - k : 0);
                const int64_t o = img * P.out_bs + l0 * P.out_ls + (int64_t)k * P.out_es;
                P.out_cplx[o] = cplx(0.5 * (zv.x + zm.x), 0.5 * (zv.y - zm.y));
                if (has1) P.out_cplx[o + P.out_ls] = cplx(0.5 * (zv.y + zm.y), -0.5 * (zv.x - zm.x));
            }
            continue;
        }
        if (P.twice) {
            // forward output X[k] -> Y[k] = X[k] mul[k] -> input of the inverse transform, conj(Y[k]) chirp[k]
            auto again = [&](int k) -> double2 {
                const int64_t oi = l0 * P.out_ls + (int64_t)k * P.out_es;
                const double2 c = Ctx::ldg(P.chirp + k);
                return cmul(cconj(cmul(cmul(c, cconj(buf[CB_PAD(k)])), Ctx::ldg(P.mul + oi))), c);
            };
            if (pruned) {
                // a first-pass butterfly reads the slots it writes, and they were filled by this lane's own sink
                if constexpr (Engine::FUSED_SPAN > 0) {
                    auto ld = [&](int x) -> double2 { return x < n ? again(x) : cplx(0.0, 0.0); };
                    Engine::convolve_io(cx, buf, tw, filt, lane, P.chirp, n, ld, keep);
                    cx.sync_warp();
                }
            } else {
                // written back to slot k (each lane rewrites the slots it read); the zero padding is restored
                for (int k = lane; k < M; k += CB_WARP) buf[CB_PAD(k)] = k < n ? again(k) : cplx(0.0, 0.0);
                Engine::convolve(cx, buf, tw, filt, P.fft, lane);
            }
            for (int k = lane; k < n; k += CB_WARP)
                P.out_cplx[img * P.out_bs + l0 * P.out_ls + (int64_t)k * P.out_es] = cscale(cconj(X(k)), P.scale);
            continue;
        }
        for (int k = lane; k < n; k += CB_WARP) {
            double2 z = X(k);
            if (P.inverse) z = cscale(cconj(z), P.scale);
            const int64_t oi = l0 * P.out_ls + (int64_t)k * P.out_es;     // offset within the image
            const int64_t o = img * P.out_bs + oi;
            if (P.mul) z = cmul(z, Ctx::ldg(P.mul + oi));
            if (P.pair == 2) {
                P.out_real[o] = z.x;
                if (has1) P.out_real[o + P.out_ls] = z.y;
            } else if (P.out_real) P.out_real[o] = z.x;
            else P.out_cplx[o] = z;
        }
    }
}

// The passes of the window batch (cb_fft2_filter_windows and its emulation): windows gathered from the grid
// -> A[img][y][v] (v < half) -> B[img][u][v] = op_sym * column transform -> A (column inverse) -> real stack.
// op_half = the Hermitian part of the operator on the half plane, [n][half].  Bluestein lines (every odd BASELINE
// window) take the two column passes in one visit, A -> A (pass 1 with fused = true, pass 2 skipped, no B).
static inline void fft2_windows_pass(int pass, int n, int64_t nwin, const double *grid, int64_t ld, const int32_t *iy0,
                                     const int32_t *ix0, double2 *A, double2 *B, const double2 *op_half, double *out,
                                     LineFftParams &P, bool fused = false) {
    const int64_t half = n / 2 + 1, npairs = (n + 1) / 2, img = (int64_t)n * half;
    memset(&P, 0, sizeof(P));
    P.n = n; P.half = (int32_t)half; P.scale = 1.0;
    if (pass == 0) {            // row pairs forward, gathered from the grid
        P.pair = 1; P.in_real = grid; P.in_es = 1; P.in_ls = ld; P.in_ld = ld; P.iy0 = iy0; P.ix0 = ix0;
        P.out_cplx = A; P.out_es = 1; P.out_ls = half; P.out_bs = img; P.lines_per_img = npairs; P.nlines = nwin * npairs;
    } else if (pass == 1) {     // half columns forward, times the operator (fused: and back, in place)
        P.in_cplx = A; P.in_es = half; P.in_ls = 1; P.in_bs = img; P.out_cplx = fused ? A : B; P.out_es = half; P.out_ls = 1; P.out_bs = img;
        P.lines_per_img = half; P.nlines = nwin * half; P.mul = op_half;
        if (fused) { P.twice = 1; P.scale = 1.0 / (double)n; }
    } else if (pass == 2) {     // half columns inverse
        P.in_cplx = B; P.in_es = half; P.in_ls = 1; P.in_bs = img; P.out_cplx = A; P.out_es = half; P.out_ls = 1; P.out_bs = img;
        P.lines_per_img = half; P.nlines = nwin * half; P.inverse = 1; P.scale = 1.0 / (double)n;
    } else {                    // row pairs inverse -> two real rows each
        P.pair = 2; P.in_cplx = A; P.in_es = 1; P.in_ls = half; P.in_bs = img; P.out_real = out; P.out_es = 1; P.out_ls = n;
        P.out_bs = (int64_t)n * n; P.lines_per_img = npairs; P.nlines = nwin * npairs; P.inverse = 1; P.scale = 1.0 / (double)n;
    }
}

// Hermitian part of an [n][n] operator on the half plane: op_half[u][v] = (op[u][v] + conj op[-u][-v]) / 2, v < half.
CB_HD double2 op_half_entry(const double2 *op, int n, int64_t i) {
    const int half = n / 2 + 1;
    const int u = (int)(i / half), v = (int)(i - (int64_t)u * half);
    const double2 a = op[(int64_t)u * n + v], b = op[(int64_t)((n - u) % n) * n + (n - v) % n];
    return cplx(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
}
template <class Ctx>
CB_BODY void op_half_body(const Ctx &cx, const double2 *op, int n, double2 *op_half) {
    const int64_t total = (int64_t)n * (n / 2 + 1);
    for (int64_t i = (int64_t)cx.bid() * cx.nthreads() + cx.tid(); i < total; i += (int64_t)cx.nblocks() * cx.nthreads())
        op_half[i] = op_half_entry(op, n, i);
}
