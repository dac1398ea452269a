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
    int32_t inverse;           // 1: conj in, conj out, scale
    double scale;
    int32_t bluestein;
    FftPlan fft;
    const double2 *tw, *chirp, *filt;   // global tables (filt padded like a line)
    const int32_t *slot;                // direct mode: slot of X[k]
    int32_t line_slots;
};

template <class Ctx>
CB_BODY void line_fft_body(const Ctx &cx, const LineFftParams &P) {
    const int tid = cx.tid(), nthr = cx.nthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int n = P.n, M = P.fft.n;
    double2 *buf = reinterpret_cast<double2 *>(cx.smem()) + (int64_t)warp * P.line_slots;
    for (int64_t line = (int64_t)cx.bid() * nwarps + warp; line < P.nlines; line += (int64_t)cx.nblocks() * nwarps) {
        cx.sync_warp();
        int64_t img = 0, l = line;
        if (P.lines_per_img) { img = line / P.lines_per_img; l = line - img * P.lines_per_img; }
        int64_t in0 = img * P.in_bs + l * P.in_ls;
        if (P.iy0) in0 += (int64_t)Ctx::ldg(P.iy0 + img) * P.in_ld + Ctx::ldg(P.ix0 + img);
        for (int x = lane; x < M; x += CB_WARP) {
            double2 z = cplx(0.0, 0.0);
            if (x < n) {
                if (P.in_real) z = cplx(P.in_real[in0 + (int64_t)x * P.in_es], 0.0);
                else z = P.in_cplx[in0 + (int64_t)x * P.in_es];
                if (P.inverse) z = cconj(z);
                if (P.bluestein) z = cmul(z, Ctx::ldg(P.chirp + x));
            }
            buf[CB_PAD(x)] = z;
        }
        if (P.bluestein) GenericFft::convolve(cx, buf, P.tw, P.filt, P.fft, lane);
        else GenericFft::forward(cx, buf, P.tw, P.fft, lane);
        for (int k = lane; k < n; k += CB_WARP) {
            double2 z = P.bluestein ? cmul(Ctx::ldg(P.chirp + k), cconj(buf[CB_PAD(k)]))
                                    : buf[CB_PAD(Ctx::ldg(P.slot + k))];
            if (P.inverse) z = cscale(cconj(z), P.scale);
            const int64_t oi = l * P.out_ls + (int64_t)k * P.out_es;     // offset within the image
            const int64_t o = img * P.out_bs + oi;
            if (P.mul) z = cmul(z, Ctx::ldg(P.mul + oi));
            if (P.out_real) P.out_real[o] = z.x;
            else P.out_cplx[o] = z;
        }
    }
}
