// CPU emulation harness for the kernel bodies (TEST INFRASTRUCTURE ONLY).
// Compiles pycurious_b200/csrc/*_body.cuh with CB_EMU so that the same templates that the
// CUDA kernels instantiate with DevCtx run here with HostCtx: one pthread per CUDA thread,
// pthread barriers for __syncthreads/__syncwarp, a scratch array for shuffles.  It exists to
// check kernel LOGIC in a container without a GPU; nothing in pycurious_b200/ loads it.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include "spectrum_plan.h"
#include "fit_body.cuh"
#include "mh_body.cuh"
#include "tanaka_body.cuh"
#include "fft2d_body.cuh"
#include "radon_body.cuh"

void cb_set_error(const char *fmt, ...) { (void)fmt; }
int cb_check_device() { return CB_ERR_NO_DEVICE; }

template <class F>
static void emu_launch(int nblocks, int nthreads, size_t smem_bytes, F body) {
    for (int b = 0; b < nblocks; ++b) {
        EmuBlock blk;
        blk.nthreads = nthreads; blk.nblocks = nblocks; blk.bid = b;
        pthread_barrier_init(&blk.block_bar, nullptr, nthreads);
        int nw = (nthreads + 31) / 32;
        blk.warp_bar.resize(nw);
        for (int w = 0; w < nw; ++w) {
            int cnt = std::min(32, nthreads - 32 * w);
            pthread_barrier_init(&blk.warp_bar[w], nullptr, cnt);
        }
        blk.shfl_buf.assign(nthreads, 0.0);
        blk.vote_buf.assign(nthreads, 0);
        blk.smem.assign(smem_bytes + 64, 0);
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&blk, t, &body]() { HostCtx cx{&blk, t}; body(cx); });
        for (auto &x : th) x.join();
        pthread_barrier_destroy(&blk.block_bar);
        for (auto &wb : blk.warp_bar) pthread_barrier_destroy(&wb);
    }
}

// Shared-row-spectra path (rowspec_body + colspec_body); groups are built here like the host side of
// pycurious_b200/_plans.py does.  max_teams < 0 forces the generic engine.
template <class Engine>
static void emu_run_shared(int nblocks, const SpectrumHostPlan &hp, const SpectrumParams &P) {
    emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { rowspec_body<Engine>(cx, P); });
    emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { colspec_body<Engine>(cx, P); });
}

static std::vector<double> emu_ktable(const double *kh, int nk) {
    std::vector<double> t(ktable_doubles(nk));
    for (int i = 0; i < (int)t.size(); ++i) t[i] = ktable_entry(kh, nk, i);
    return t;
}

extern "C" {

// plan introspection without a device: FFT length, radices, folded CSR size
int emu_spectrum_plan_info(int nw, int nbins, const int32_t *bin_ptr, const int32_t *bin_idx,
                           int32_t *out /* bluestein, fft_len, npass, radix[12], nnz_full, nnz_folded, teams */) {
    SpectrumHostPlan hp;
    if (!spectrum_plan_build_host(hp, nw, nbins, nullptr, nullptr, bin_ptr, bin_idx, 232448)) return -1;
    out[0] = hp.bluestein; out[1] = hp.fft.n; out[2] = hp.fft.npass;
    for (int p = 0; p < CB_MAX_PASS; ++p) out[3 + p] = hp.fft.pass[p].radix;
    out[15] = hp.nnz_full; out[16] = (int32_t)hp.bin_idx.size(); out[17] = hp.teams;
    return 0;
}

// 1-D DFT of length n through the device FFT engine (one emulated warp), Bluestein or direct.
int emu_dft(int n, const double *in_re_im, double *out_re_im) {
    FftPlan fft;
    std::vector<double2> tw, chirp, filt;
    std::vector<int32_t> slot;
    bool direct = cbplan::build_fft_plan(n, fft, tw);
    if (direct) slot = cbplan::digit_reversed_slots(fft);
    else {
        int M = cbplan::bluestein_length(n);
        if (!M || !cbplan::build_fft_plan(M, fft, tw)) return -1;
        cbplan::build_bluestein_tables(n, fft, chirp, filt);
    }
    std::vector<double2> line(fft_line_slots(fft.n) + 2, cplx(0, 0));
    emu_launch(1, 32, 0, [&](const HostCtx &cx) {
        const int lane = cx.tid();
        for (int x = lane; x < fft.n; x += 32) {
            double2 z = cplx(0, 0);
            if (x < n) {
                z = cplx(in_re_im[2 * x], in_re_im[2 * x + 1]);
                if (!direct) z = cmul(z, chirp[x]);
            }
            line[CB_PAD(x)] = z;
        }
        if (direct) fft_forward_dif(cx, line.data(), tw.data(), fft, lane);
        else if (StaticFft<8, 8, 8>::matches(fft)) StaticFft<8, 8, 8>::convolve(cx, line.data(), tw.data(), filt.data(), fft, lane);
        else if (StaticFft<8, 8, 4>::matches(fft)) StaticFft<8, 8, 4>::convolve(cx, line.data(), tw.data(), filt.data(), fft, lane);
        else if (StaticFft<8, 8, 5>::matches(fft)) StaticFft<8, 8, 5>::convolve(cx, line.data(), tw.data(), filt.data(), fft, lane);
        else if (StaticFft<8, 10, 8>::matches(fft)) StaticFft<8, 10, 8>::convolve(cx, line.data(), tw.data(), filt.data(), fft, lane);
        else bluestein_convolve(cx, line.data(), tw.data(), filt.data(), fft, lane);
        for (int k = lane; k < n; k += 32) {
            double2 z = direct ? line[CB_PAD(slot[k])] : bluestein_out(line.data(), chirp.data(), k);
            out_re_im[2 * k] = z.x;
            out_re_im[2 * k + 1] = z.y;
        }
    });
    return direct ? 0 : fft.n;
}

int emu_spectrum(int nw, int nbins, const double *taper_r, const double *taper_c, const int32_t *bin_ptr,
                 const int32_t *bin_idx, const double *grid, int64_t ld, const int32_t *iy0,
                 const int32_t *ix0, int64_t nwin, int detrend, double power, double *phi, double *sigma,
                 int nblocks, int max_teams) {
    SpectrumHostPlan hp;
    if (!spectrum_plan_build_host(hp, nw, nbins, taper_r, taper_c, bin_ptr, bin_idx, 232448)) {
        fprintf(stderr, "emu_spectrum: %s\n", hp.error.c_str());
        return -1;
    }
    const int mt = max_teams < 0 ? -max_teams : max_teams;
    if (mt > 0 && hp.teams > mt) {
        hp.teams = mt;
        hp.threads = 32 * mt;
        hp.smem_bytes = 16 * ((int64_t)hp.sm_buf + (int64_t)hp.teams * hp.line_slots);
    }
    SpectrumParams P;
    memset(&P, 0, sizeof(P));
    spectrum_fill_params(hp, P, hp.tw.data(), hp.chirp.data(), hp.filt.data(), hp.slot.data(),
                         hp.taper_r.data(), hp.taper_c.data(), hp.bin_ptr.data(), hp.bin_idx.data(), hp.col_span.data());
    P.grid = grid; P.ld = ld; P.iy0 = iy0; P.ix0 = ix0; P.nwin = nwin;
    P.detrend = detrend; P.power = power; P.phi = phi; P.sigma = sigma;
    std::vector<double2> gt((size_t)nblocks * hp.gt_stride);
    std::vector<double> lh((size_t)nblocks * hp.lh_stride);
    P.gt = gt.data(); P.lh = lh.data();
    // same engine selection as the product (spectrum.cu:pick_spectrum_kernel); max_teams < 0 forces the generic engine
    const bool generic = max_teams < 0 || !hp.bluestein;
    if (!generic && PlusOneFft<StaticFft<8, 8, 4>>::serves(hp.fft, hp.n))
        emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { spectrum_body<PlusOneFft<StaticFft<8, 8, 4>>>(cx, P); });
    else if (!generic && PlusOneFft<StaticFft<8, 8, 5>>::serves(hp.fft, hp.n))
        emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { spectrum_body<PlusOneFft<StaticFft<8, 8, 5>>>(cx, P); });
    else if (!generic && PlusOneFft<StaticFft<8, 8, 8>>::serves(hp.fft, hp.n))
        emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { spectrum_body<PlusOneFft<StaticFft<8, 8, 8>>>(cx, P); });
    else if (!generic && PlusOneFft<StaticFft<8, 10, 8>>::serves(hp.fft, hp.n))
        emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { spectrum_body<PlusOneFft<StaticFft<8, 10, 8>>>(cx, P); });
    else if (!generic && StaticFft<8, 8, 8>::matches(hp.fft) && hp.n <= StaticFft<8, 8, 8>::FUSED_SPAN)
        emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { spectrum_body<StaticFft<8, 8, 8>>(cx, P); });
    else if (!generic && StaticFft<8, 8, 4>::matches(hp.fft) && hp.n <= StaticFft<8, 8, 4>::FUSED_SPAN)
        emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { spectrum_body<StaticFft<8, 8, 4>>(cx, P); });
    else if (!generic && StaticFft<8, 8, 5>::matches(hp.fft) && hp.n <= StaticFft<8, 8, 5>::FUSED_SPAN)
        emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { spectrum_body<StaticFft<8, 8, 5>>(cx, P); });
    else if (!generic && StaticFft<8, 10, 8>::matches(hp.fft) && hp.n <= StaticFft<8, 10, 8>::FUSED_SPAN)
        emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { spectrum_body<StaticFft<8, 10, 8>>(cx, P); });
    else
        emu_launch(nblocks, hp.threads, (size_t)hp.smem_bytes, [&](const HostCtx &cx) { spectrum_body<GenericFft>(cx, P); });
    return 0;
}

int emu_spectrum_shared(int nw, int nbins, const double *taper_r, const double *taper_c, const int32_t *bin_ptr,
                        const int32_t *bin_idx, const double *grid, int64_t ny, int64_t ld, const int32_t *iy0,
                        const int32_t *ix0, int64_t nwin, int detrend, double power, double *phi, double *sigma,
                        int nblocks, int max_teams) {
    SpectrumHostPlan hp;
    if (!spectrum_plan_build_host(hp, nw, nbins, taper_r, taper_c, bin_ptr, bin_idx, 232448)) {
        fprintf(stderr, "emu_spectrum_shared: %s\n", hp.error.c_str());
        return -1;
    }
    const int mt = max_teams < 0 ? -max_teams : max_teams;
    if (mt > 0 && hp.teams > mt) {
        hp.teams = mt;
        hp.threads = 32 * mt;
        hp.smem_bytes = 16 * ((int64_t)hp.sm_buf + (int64_t)hp.teams * hp.line_slots);
    }
    std::vector<int32_t> gix, gof(nwin);
    int row_lo = 1 << 30, row_hi = 0;
    for (int64_t w = 0; w < nwin; ++w) {
        auto it = std::find(gix.begin(), gix.end(), ix0[w]);
        if (it == gix.end()) { gix.push_back(ix0[w]); it = gix.end() - 1; }
        gof[w] = (int32_t)(it - gix.begin());
        row_lo = std::min(row_lo, (int)iy0[w]);
        row_hi = std::max(row_hi, (int)iy0[w] + nw);
    }
    SpectrumParams P;
    memset(&P, 0, sizeof(P));
    spectrum_fill_params(hp, P, hp.tw.data(), hp.chirp.data(), hp.filt.data(), hp.slot.data(),
                         hp.taper_r.data(), hp.taper_c.data(), hp.bin_ptr.data(), hp.bin_idx.data(), hp.col_span.data());
    P.grid = grid; P.ld = ld; P.iy0 = iy0; P.ix0 = ix0; P.nwin = nwin;
    P.detrend = detrend; P.power = power; P.phi = phi; P.sigma = sigma;
    P.group_of_win = gof.data(); P.group_ix0 = gix.data(); P.ngroups = (int)gix.size();
    // processing order as the host builds it: by (group, iy0) -- here reversed as well, to exercise the indirection
    std::vector<int32_t> order(nwin);
    for (int64_t w = 0; w < nwin; ++w) order[w] = (int32_t)w;
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        return gof[a] != gof[b] ? gof[a] > gof[b] : iy0[a] > iy0[b];
    });
    P.win_order = order.data();
    P.row_lo = row_lo; P.nrows_valid = row_hi - row_lo; P.nrows = (P.nrows_valid + 1) & ~1;
    std::vector<double2> rt((size_t)P.ngroups * hp.nvh * P.nrows);
    std::vector<double> lh((size_t)nblocks * 2 * hp.lh_stride);
    P.rt = rt.data(); P.lh = lh.data(); P.lh_stride = 2 * hp.lh_stride; P.lh_half = hp.lh_stride;
    P.plane_t0 = hp.plane_t0.data(); P.plane_t1 = hp.plane_t1.data();
    std::vector<double2> rowsum((size_t)P.ngroups * P.nrows);
    if (detrend) P.rowsum = rowsum.data();
    (void)ny;
    const bool generic = max_teams < 0 || !hp.bluestein;
    if (!generic && PlusOneFft<StaticFft<8, 8, 4>>::serves(hp.fft, hp.n)) emu_run_shared<PlusOneFft<StaticFft<8, 8, 4>>>(nblocks, hp, P);
    else if (!generic && PlusOneFft<StaticFft<8, 8, 5>>::serves(hp.fft, hp.n)) emu_run_shared<PlusOneFft<StaticFft<8, 8, 5>>>(nblocks, hp, P);
    else if (!generic && PlusOneFft<StaticFft<8, 8, 8>>::serves(hp.fft, hp.n)) emu_run_shared<PlusOneFft<StaticFft<8, 8, 8>>>(nblocks, hp, P);
    else if (!generic && PlusOneFft<StaticFft<8, 10, 8>>::serves(hp.fft, hp.n)) emu_run_shared<PlusOneFft<StaticFft<8, 10, 8>>>(nblocks, hp, P);
    else if (!generic && StaticFft<8, 8, 8>::matches(hp.fft) && hp.n <= StaticFft<8, 8, 8>::FUSED_SPAN) emu_run_shared<StaticFft<8, 8, 8>>(nblocks, hp, P);
    else if (!generic && StaticFft<8, 8, 4>::matches(hp.fft) && hp.n <= StaticFft<8, 8, 4>::FUSED_SPAN) emu_run_shared<StaticFft<8, 8, 4>>(nblocks, hp, P);
    else if (!generic && StaticFft<8, 8, 5>::matches(hp.fft) && hp.n <= StaticFft<8, 8, 5>::FUSED_SPAN) emu_run_shared<StaticFft<8, 8, 5>>(nblocks, hp, P);
    else if (!generic && StaticFft<8, 10, 8>::matches(hp.fft) && hp.n <= StaticFft<8, 10, 8>::FUSED_SPAN) emu_run_shared<StaticFft<8, 10, 8>>(nblocks, hp, P);
    else emu_run_shared<GenericFft>(nblocks, hp, P);
    return 0;
}

int emu_detrend(const double *grid, int64_t ld, int nw, const int32_t *iy0, const int32_t *ix0,
                int64_t nwin, double *out) {
    DetrendParams P{grid, ld, iy0, ix0, nwin, nw, out};
    emu_launch(1, 128, 96 * sizeof(double), [&](const HostCtx &cx) { detrend_body(cx, P); });
    return 0;
}

// ---- forward model hooks ---------------------------------------------------------------------
void emu_besselk(const double *nu, const double *x, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = cb_kv(nu[i], x[i]);
}

void emu_gamma(const double *z, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = cb_gamma(z[i]);
}

void emu_bouligand(const double *kh, int nk, const double *params, int64_t m, double *out) {
    for (int64_t r = 0; r < m; ++r) {
        const double *p = params + 4 * r;
        ModelCoef mc = model_coef(p[0], p[1], p[2], p[3]);
        for (int j = 0; j < nk; ++j) out[r * nk + j] = bouligand_point(mc, kh[j], log(kh[j]));
    }
}

// analytic dPhi/ddz (used by the fit) for finite-difference checks
void emu_bouligand_ddz(const double *kh, int nk, const double *p, double *phi, double *dphi) {
    ModelCoef mc = model_coef(p[0], p[1], p[2], p[3]);
    for (int j = 0; j < nk; ++j) phi[j] = bouligand_point_ddz(mc, kh[j], log(kh[j]), dphi[j]);
}

void emu_misfit(const double *kh, int nk, const double *phi, int64_t phi_stride, const double *params, int64_t m,
                const double *prior_mu, const double *prior_sigma, double *out) {
    Priors pr;
    for (int j = 0; j < 4; ++j) { pr.mu[j] = prior_mu ? prior_mu[j] : NAN; pr.sigma[j] = prior_sigma ? prior_sigma[j] : NAN; }
    std::vector<double> sm(warp_slab_doubles(nk, false));
    const std::vector<double> ktab = emu_ktable(kh, nk);
    SerialCtx cx{0, 1, 0, 1, reinterpret_cast<char *>(sm.data())};
    for (int64_t r = 0; r < m; ++r) {
        WarpSlab W = warp_slab_carve(sm.data(), nk, false, kh, ktab.data());
        warp_slab_load(cx, W, phi + r * phi_stride, 0);
        out[r] = misfit_slab(cx, W, params + 4 * r, pr, 0);
    }
}

// The batched fit under emulation: begin, `rounds` rounds of (evaluation, step) over the list of unfinished
// windows, then the per-warp finish -- the device's sequence (model.cu:cb_fit_batch_rows).
// mode 0: SerialCtx (one lane per warp, fast); mode 1: HostCtx (32 pthreads per warp, checks the shuffles).
static int emu_fit_run(FitParams &P, int mode, int rounds) {
    const int64_t n = P.nwin;
    P.handover = n > 4 ? 3 : 0;                      // small enough to exercise both the rounds and the hand-over
    std::vector<double> state(n * CB_FIT_PERSIST_DOUBLES), coef(n * 3 * FC_COUNT), sums(n * 16), tmpl(CB_FIT_STATE_DOUBLES);
    std::vector<int32_t> la(n + 1), lb(n + 1);
    unsigned count[6] = {(unsigned)n, 0, 0, 0, 0, 0};
    P.state = state.data(); P.coef = coef.data(); P.sums = sums.data(); P.list_in = la.data(); P.list_out = lb.data();
    P.count = count; P.stride = n; P.tmpl = tmpl.data();
    fit_template_body(P);
    for (int64_t w = 0; w < n; ++w) fit_begin_window(P, w);
    std::vector<double> sm((fit_finish_smem_bytes(P.nk) + fit_eval_smem_bytes(1, P.nk)) / sizeof(double) + 8);
    for (int r = 0; r < rounds; ++r) {
        P.round = r;
        if (r == 0 && !P.x0_dev) {          // common start point: model values once per CTA, sums per window
            if (mode == 0) {
                std::vector<double> sf(fit_first_smem_bytes(P.nk) / sizeof(double) + 8);
                SerialCtx cx{0, 1, 0, 1, reinterpret_cast<char *>(sf.data())};
                fit_first_body(cx, P);
            } else {
                emu_launch(2, 64, fit_first_smem_bytes(P.nk), [&](const HostCtx &cx) { fit_first_body(cx, P); });
            }
        } else if (mode == 0) {
            SerialCtx cx{0, 1, 0, 1, reinterpret_cast<char *>(sm.data())};
            fit_eval_body(cx, P);
        } else {
            emu_launch(2, 64, fit_eval_smem_bytes(2, P.nk), [&](const HostCtx &cx) { fit_eval_body(cx, P); });
        }
        SerialCtx cx{0, 1, 0, 1, reinterpret_cast<char *>(sm.data())};
        const unsigned cnt = count[r % 3];
        for (unsigned e = 0; e < cnt; ++e) fit_step_entry(cx, P, (int64_t)e);
        std::swap(P.list_in, P.list_out);
    }
    P.round = rounds;
    if (mode == 0) {
        SerialCtx cx{0, 1, 0, 1, reinterpret_cast<char *>(sm.data())};
        fit_finish_body(cx, P);
    } else {
        emu_launch(2, 64, fit_finish_smem_bytes(P.nk), [&](const HostCtx &cx) { fit_finish_body(cx, P); });
    }
    return 0;
}

int emu_fit(const double *kh, int nk, const double *phi, int64_t nwin, const cb_fit_options *opt, const double *x0,
            double *x_out, double *f_out, int32_t *niter, int32_t *status, int32_t *nfev, int mode) {
    FitParams P;
    memset(&P, 0, sizeof(P));
    const std::vector<double> ktab = emu_ktable(kh, nk);
    P.kh = kh; P.ktab = ktab.data(); P.nk = nk; P.phi = phi; P.nwin = nwin; P.opt = *opt; P.x0_dev = x0;
    P.x_out = x_out; P.f_out = f_out; P.niter = niter; P.status = status; P.nfev = nfev;
    return emu_fit_run(P, mode, mode == 0 ? 6 : 3);
}

void emu_tanaka(const double *k, int nk, const double *lnphi, const double *sigma, int64_t nwin, double a0, double a1,
                double b0, double b1, double *out) {
    TanakaParams P;
    P.k = k; P.nk = nk; P.lnphi = lnphi; P.sigma = sigma; P.nwin = nwin;
    P.top_lo = a0; P.top_hi = a1; P.cen_lo = b0; P.cen_hi = b1; P.out = out; P.warps_per_block = 2;
    emu_launch((int)((nwin + 1) / 2), 64, 64, [&](const HostCtx &cx) { tanaka_body(cx, P); });
}

// out = real(ifft2(filter * fft2(data))) through line_fft_body, the four passes of cb_fft2_filter
int emu_fft2_filter(const double *data, int64_t ny, int64_t nx, const double *filter_re_im, double *out) {
    struct Plan { FftPlan fft; std::vector<double2> tw, chirp, filt; std::vector<int32_t> slot; int bl, line_slots; };
    auto make = [](int n, Plan &p) {
        if (cbplan::build_fft_plan(n, p.fft, p.tw)) { p.bl = 0; p.slot = cbplan::digit_reversed_slots(p.fft); }
        else {
            int M = cbplan::bluestein_length(n);
            if (!M || !cbplan::build_fft_plan(M, p.fft, p.tw)) return false;
            p.bl = 1;
            cbplan::build_bluestein_tables(n, p.fft, p.chirp, p.filt);
        }
        p.line_slots = (fft_line_slots(p.fft.n) + 1) & ~1;
        if (p.bl) p.filt.resize(p.line_slots, cplx(0.0, 0.0));
        return true;
    };
    Plan px, py;
    if (!make((int)nx, px) || !make((int)ny, py)) return -1;
    std::vector<double2> A(ny * nx), B(ny * nx);
    auto run = [&](const Plan &pl, LineFftParams P) {
        P.bluestein = pl.bl; P.fft = pl.fft; P.tw = pl.tw.data(); P.chirp = pl.chirp.data(); P.filt = pl.filt.data();
        P.slot = pl.slot.data(); P.line_slots = pl.line_slots;
        emu_launch(2, 64, (size_t)2 * pl.line_slots * 16, [&](const HostCtx &cx) { line_fft_body<GenericFft>(cx, P); });
    };
    LineFftParams P;
    memset(&P, 0, sizeof(P));
    P.in_real = data; P.in_es = 1; P.in_ls = nx; P.out_cplx = A.data(); P.out_es = 1; P.out_ls = nx; P.nlines = ny; P.n = (int)nx; P.scale = 1.0;
    run(px, P);
    memset(&P, 0, sizeof(P));
    P.in_cplx = A.data(); P.in_es = nx; P.in_ls = 1; P.out_cplx = B.data(); P.out_es = nx; P.out_ls = 1; P.nlines = nx; P.n = (int)ny; P.scale = 1.0;
    P.mul = reinterpret_cast<const double2 *>(filter_re_im);
    run(py, P);
    memset(&P, 0, sizeof(P));
    P.in_cplx = B.data(); P.in_es = nx; P.in_ls = 1; P.out_cplx = A.data(); P.out_es = nx; P.out_ls = 1; P.nlines = nx; P.n = (int)ny;
    P.inverse = 1; P.scale = 1.0 / (double)ny;
    run(py, P);
    memset(&P, 0, sizeof(P));
    P.in_cplx = A.data(); P.in_es = 1; P.in_ls = nx; P.out_real = out; P.out_es = 1; P.out_ls = nx; P.nlines = ny; P.n = (int)nx;
    P.inverse = 1; P.scale = 1.0 / (double)nx;
    run(px, P);
    return 0;
}

// the four passes of cb_fft2_filter_windows: nwin windows of edge n gathered from a grid, filtered, stacked
int emu_fft2_filter_windows(const double *grid, int64_t ny, int64_t nx, const int32_t *iy0, const int32_t *ix0,
                            int64_t nwin, int n, const double *filter_re_im, double *out) {
    (void)ny;
    FftPlan fft; std::vector<double2> tw, chirp, filt; std::vector<int32_t> slot; int bl = 0;
    if (cbplan::build_fft_plan(n, fft, tw)) slot = cbplan::digit_reversed_slots(fft);
    else {
        int M = cbplan::bluestein_length(n);
        if (!M || !cbplan::build_fft_plan(M, fft, tw)) return -1;
        bl = 1;
        cbplan::build_bluestein_tables(n, fft, chirp, filt);
    }
    const int line_slots = (fft_line_slots(fft.n) + 1) & ~1;
    if (bl) filt.resize(line_slots, cplx(0.0, 0.0));
    const int64_t half = n / 2 + 1;
    std::vector<double2> A(nwin * n * half), B(nwin * n * half), oph(n * half);
    for (int64_t i = 0; i < (int64_t)oph.size(); ++i) oph[i] = op_half_entry(reinterpret_cast<const double2 *>(filter_re_im), n, i);
    for (int pass = 0; pass < 4; ++pass) {
        if (bl && pass == 2) continue;                  // Bluestein lines: pass 1 does both column transforms
        LineFftParams P;
        fft2_windows_pass(pass, n, nwin, grid, nx, iy0, ix0, A.data(), bl ? nullptr : B.data(), oph.data(), out, P, bl != 0);
        P.bluestein = bl; P.fft = fft; P.tw = tw.data(); P.chirp = chirp.data(); P.filt = filt.data();
        P.slot = slot.data(); P.line_slots = line_slots;
        // same engine selection as cb_fft_plan_create: the static engines stage twiddles + filter in front of the lines
        // and run their pruned passes on lines that fit
        const int stage = (fft.tw_count + line_slots + 1) & ~1;
        auto run_static = [&](auto engine_tag) {
            typedef decltype(engine_tag) E;
            P.stage_slots = stage;
            emu_launch(3, 64, (size_t)(stage + 2 * line_slots) * 16, [&](const HostCtx &cx) { line_fft_body<E>(cx, P); });
        };
        if (bl && PlusOneFft<StaticFft<8, 8, 4>>::serves(fft, n)) run_static(PlusOneFft<StaticFft<8, 8, 4>>());
        else if (bl && PlusOneFft<StaticFft<8, 8, 5>>::serves(fft, n)) run_static(PlusOneFft<StaticFft<8, 8, 5>>());
        else if (bl && PlusOneFft<StaticFft<8, 8, 8>>::serves(fft, n)) run_static(PlusOneFft<StaticFft<8, 8, 8>>());
        else if (bl && StaticFft<8, 8, 8>::matches(fft)) run_static(StaticFft<8, 8, 8>());
        else if (bl && StaticFft<8, 8, 4>::matches(fft)) run_static(StaticFft<8, 8, 4>());
        else if (bl && StaticFft<8, 8, 5>::matches(fft)) run_static(StaticFft<8, 8, 5>());
        else if (bl && StaticFft<8, 10, 8>::matches(fft)) run_static(StaticFft<8, 10, 8>());
        else emu_launch(3, 64, (size_t)2 * line_slots * 16, [&](const HostCtx &cx) { line_fft_body<GenericFft>(cx, P); });
    }
    return 0;
}

// fit with per-row prior means (the inner loop of CurieOptimise.sensitivity)
int emu_fit_rows(const double *kh, int nk, const double *phi, int64_t nwin, const cb_fit_options *opt,
                 const double *prior_mu_rows, double *x_out, double *f_out) {
    FitParams P;
    memset(&P, 0, sizeof(P));
    const std::vector<double> ktab = emu_ktable(kh, nk);
    P.kh = kh; P.ktab = ktab.data(); P.nk = nk; P.phi = phi; P.nwin = nwin; P.opt = *opt; P.prior_mu_dev = prior_mu_rows;
    P.x_out = x_out; P.f_out = f_out;
    return emu_fit_run(P, 0, 4);
}

int emu_mh(const double *kh, int nk, const double *phi, int64_t nwin, const cb_mh_options *opt, const double *x0,
           double *out, int64_t *accepted, int mode) {
    MhParams P;
    memset(&P, 0, sizeof(P));
    const std::vector<double> ktab = emu_ktable(kh, nk);
    P.kh = kh; P.ktab = ktab.data(); P.nk = nk; P.phi = phi; P.nwin = nwin; P.opt = *opt; P.x0_dev = x0; P.out = out; P.accepted = accepted;
    if (mode == 0) {
        P.warps_per_block = 1;
        std::vector<double> sm(mh_slab_doubles(nk));
        for (int64_t w = 0; w < nwin; ++w) {
            SerialCtx cx{0, 1, (int)w, (int)nwin, reinterpret_cast<char *>(sm.data())};
            mh_body(cx, P);
        }
    } else {
        P.warps_per_block = 2;
        emu_launch((int)((nwin + 1) / 2), 64, 2 * mh_slab_doubles(nk) * sizeof(double), [&](const HostCtx &cx) { mh_body(cx, P); });
    }
    return 0;
}

// ---- radon -------------------------------------------------------------------------------------
// count pass (data == NULL) or fill pass; offset has nray+1 entries in the fill pass
int emu_radon_trace(const double *tx, const double *rx, int64_t nray, int nw, int64_t *count,
                    const int64_t *offset, double *data, int64_t *indices) {
    RadonTraceParams P;
    memset(&P, 0, sizeof(P));
    P.tx = tx; P.rx = rx; P.nray = nray; P.nw = nw; P.count = count; P.offset = offset; P.data = data; P.indices = indices;
    emu_launch((int)((nray + 63) / 64), 64, 0, [&](const HostCtx &cx) { radon_trace_body(cx, P); });
    return 0;
}

int emu_radon_apply(const double *grid, int64_t ld, const int32_t *iy0, const int32_t *ix0, int64_t nwin, int nw,
                    int ntheta, const int64_t *offset, const double *data, const int64_t *indices, double *sino,
                    int window_fastest) {
    RadonApplyParams P;
    memset(&P, 0, sizeof(P));
    P.grid = grid; P.ld = ld; P.iy0 = iy0; P.ix0 = ix0; P.nwin = nwin; P.nw = nw; P.ntheta = ntheta;
    P.offset = offset; P.data = data; P.indices = indices; P.sino = sino; P.window_fastest = window_fastest;
    const int64_t total = (int64_t)nw * ntheta * nwin;
    emu_launch((int)((total + 63) / 64), 64, 0, [&](const HostCtx &cx) { radon_apply_body(cx, P); });
    return 0;
}

int emu_az_tail(const double *sino, int nw, int ntheta, int64_t nwin, const double *taper, int nkb, double power,
                double *S) {
    AzTailParams P{sino, nw, ntheta, nkb, nwin, taper, power, S};
    const int nk = 1 + 2 * nkb;
    emu_launch((int)(nwin * ntheta), 32, (size_t)nk * 16 + (size_t)nw * 8, [&](const HostCtx &cx) { az_tail_body(cx, P); });
    return 0;
}

// Warp-cooperative model evaluation (fast path) for direct comparison with bouligand2009: mode 0 serial, 1 = 32 lanes
void emu_model_eval(const double *kh, int nk, const double *x, double *phi, double *ddz, int mode) {
    std::vector<double> zero(nk, 0.0);
    const std::vector<double> ktab = emu_ktable(kh, nk);
    if (mode == 0) {
        std::vector<double> sm(warp_slab_doubles(nk, true));
        SerialCtx cx{0, 1, 0, 1, reinterpret_cast<char *>(sm.data())};
        WarpSlab W = warp_slab_carve(sm.data(), nk, true, kh, ktab.data());
        warp_slab_load(cx, W, zero.data(), 0);
        model_coefs_warp(cx, W, x, 0.0, 0, 1, 0);
        model_eval_warp<false>(cx, W, 0, 0, W.p0, W.dz);
        for (int i = 0; i < nk; ++i) { phi[i] = W.p0[i]; ddz[i] = W.dz[i]; }
    } else {
        emu_launch(1, 32, warp_slab_doubles(nk, true) * sizeof(double), [&](const HostCtx &cx) {
            WarpSlab W = warp_slab_carve(reinterpret_cast<double *>(cx.smem()), nk, true, kh, ktab.data());
            warp_slab_load(cx, W, zero.data(), cx.tid());
            model_coefs_warp(cx, W, x, 0.0, 0, 1, cx.tid());
            model_eval_warp<false>(cx, W, 0, cx.tid(), W.p0, W.dz);
            for (int i = cx.tid(); i < nk; i += 32) { phi[i] = W.p0[i]; ddz[i] = W.dz[i]; }
        });
    }
}

// the spectrum kernels' 0.5 ln(x) (host build of the same algorithm: float-rounded reciprocal seed)
void emu_half_log(const double *x, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = cb_half_log(x[i]);
}

void emu_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
    U32x4 r = philox4x32_10(c0, c1, c2, c3, k0, k1);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}

}  // extern "C"
