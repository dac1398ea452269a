// Host-side plan of the spectrum stage: FFT schedule, Bluestein tables, taper, folded bin CSR,
// shared-memory carve-up.  Pure host code (no CUDA calls) so tests/emu can reuse it.
#pragma once
#include <algorithm>
#include <string>
#include <vector>
#include "fft_plan.h"
#include "spectrum_body.cuh"

struct SpectrumHostPlan {
    int n = 0, nbins = 0, nvh = 0, npad = 0, bluestein = 0;
    FftPlan fft;
    std::vector<double2> tw, chirp, filt;
    std::vector<int32_t> slot;
    std::vector<double> taper_r, taper_c;
    std::vector<double2> plane_t0, plane_t1;   // x-transforms of taper_c and (x - ctr) taper_c, v < nvh
    bool has_taper = false;
    std::vector<int32_t> bin_ptr, bin_idx, col_span;
    int nnz_full = 0;
    // shared memory layout (double2 slots) and launch shape
    int sm_tw = 0, sm_filt = 0, sm_chirp = 0, sm_cw = 0, sm_buf = 0, line_slots = 0;
    int teams = 0, threads = 0;
    int64_t smem_bytes = 0;
    int64_t gt_stride = 0, lh_stride = 0, ws_per_cta = 0;
    std::string error;
};

static const int CB_SPEC_MAX_TEAMS = 16;
static const int CB_SPEC_RED_SLOTS = 48;      // 96 doubles of block-reduction scratch

// Folds the reference-built CSR (flat ids a*n+b of the fftSHIFTED plane) onto the Hermitian
// half plane v <= n/2 that the kernel computes: |F[u,v]| = |F[-u,-v]| for real input.
static inline void fold_bins(int n, int nbins, const int32_t *ptr, const int32_t *idx,
                             std::vector<int32_t> &optr, std::vector<int32_t> &oidx) {
    const int half = n / 2;
    optr.assign(nbins + 1, 0);
    oidx.clear();
    std::vector<int32_t> ids;
    for (int b = 0; b < nbins; ++b) {
        ids.clear();
        for (int e = ptr[b]; e < ptr[b + 1]; ++e) {
            int a = idx[e] / n, c = idx[e] % n;
            int u = (a + (n + 1) / 2) % n;      // np.fft.fftshift: shifted[a] = orig[(a + ceil(n/2)) % n]
            int v = (c + (n + 1) / 2) % n;
            if (v > half) { u = (n - u) % n; v = n - v; }
            // columns v = 0 (and v = n/2 for even n) are their own mirror: fold u as well
            if ((v == 0 || 2 * v == n) && u > half) u = n - u;
            ids.push_back(v * n + u);
        }
        std::sort(ids.begin(), ids.end());
        // mean and population std are unchanged when every member appears the same number of
        // times (for odd n every half-plane point of a bin is hit exactly twice): keep one copy.
        size_t i = 0;
        int mult = 0;
        bool uniform = !ids.empty();
        while (i < ids.size()) {
            size_t j = i;
            while (j < ids.size() && ids[j] == ids[i]) ++j;
            int m = (int)(j - i);
            if (!mult) mult = m;
            else if (m != mult) { uniform = false; break; }
            i = j;
        }
        if (uniform && mult > 1) ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        oidx.insert(oidx.end(), ids.begin(), ids.end());
        optr[b + 1] = (int32_t)oidx.size();
    }
}

static inline bool spectrum_plan_build_host(SpectrumHostPlan &hp, int nw, int nbins, const double *taper_r,
                                            const double *taper_c, const int32_t *bin_ptr,
                                            const int32_t *bin_idx, int64_t smem_limit_bytes,
                                            int max_teams = CB_SPEC_MAX_TEAMS) {
    if (nw < 4 || nw > 16384) { hp.error = "window edge must be in [4, 16384]"; return false; }
    if (nbins < 1) { hp.error = "nbins must be >= 1"; return false; }
    hp.n = nw;
    hp.nbins = nbins;
    hp.nvh = nw / 2 + 1;
    hp.npad = (nw + 1) & ~1;
    // direct mixed radix when nw is 2/3/5-smooth, else Bluestein on a padded smooth length
    if (cbplan::build_fft_plan(nw, hp.fft, hp.tw)) {
        hp.bluestein = 0;
        hp.slot = cbplan::digit_reversed_slots(hp.fft);
    } else {
        int M = cbplan::bluestein_length(nw);
        if (!M || !cbplan::build_fft_plan(M, hp.fft, hp.tw)) { hp.error = "no FFT plan"; return false; }
        hp.bluestein = 1;
        cbplan::build_bluestein_tables(nw, hp.fft, hp.chirp, hp.filt);
    }
    hp.has_taper = taper_r && taper_c;
    if (hp.has_taper) {
        hp.taper_r.assign(taper_r, taper_r + nw);
        hp.taper_c.assign(taper_c, taper_c + nw);
    }
    {   // detrending commutes with the row transform through these two vectors (colspec_body)
        const long double ctr = 0.5L * (long double)(nw - 1);
        hp.plane_t0.resize(hp.nvh);
        hp.plane_t1.resize(hp.nvh);
        for (int v = 0; v < hp.nvh; ++v) {
            cbplan::lcplx a0(0, 0), a1(0, 0);
            for (int x = 0; x < nw; ++x) {
                const long double t = hp.has_taper ? (long double)taper_c[x] : 1.0L;
                const cbplan::lcplx w = cbplan::cis_ratio((long long)v * x, nw);
                a0 += t * w;
                a1 += ((long double)x - ctr) * t * w;
            }
            hp.plane_t0[v] = cplx((double)a0.real(), (double)a0.imag());
            hp.plane_t1[v] = cplx((double)a1.real(), (double)a1.imag());
        }
    }
    for (int b = 0; b < nbins; ++b)
        if (bin_ptr[b + 1] < bin_ptr[b]) { hp.error = "bin_ptr not monotone"; return false; }
    hp.nnz_full = bin_ptr[nbins] - bin_ptr[0];
    for (int e = bin_ptr[0]; e < bin_ptr[nbins]; ++e)
        if (bin_idx[e] < 0 || bin_idx[e] >= nw * nw) { hp.error = "bin_idx out of range"; return false; }
    fold_bins(nw, nbins, bin_ptr, bin_idx, hp.bin_ptr, hp.bin_idx);
    {   // per half-plane column the u that some bin reads: a circular run around u = 0 (low |frequency|),
        // i.e. u <= lo_end or u >= hi_start; the middle of the column lies beyond the last bin edge
        const int halfn = nw / 2;
        std::vector<int> lo_end(hp.nvh, -1), hi_start(hp.nvh, 0xffff);
        for (int32_t id : hp.bin_idx) {
            const int v = id / nw, u = id % nw;
            if (u <= halfn) lo_end[v] = std::max(lo_end[v], u);
            else hi_start[v] = std::min(hi_start[v], u);
        }
        hp.col_span.resize(hp.nvh);
        for (int v = 0; v < hp.nvh; ++v)
            hp.col_span[v] = (int32_t)((uint32_t)(lo_end[v] < 0 ? 0xffff : lo_end[v]) | ((uint32_t)hi_start[v] << 16));
    }

    hp.line_slots = (fft_line_slots(hp.fft.n) + 1) & ~1;
    if (hp.bluestein) hp.filt.resize(hp.line_slots, cplx(0.0, 0.0));   // the kernels stage line_slots entries
    hp.sm_tw = CB_SPEC_RED_SLOTS;
    hp.sm_filt = hp.sm_tw + hp.fft.tw_count;
    hp.sm_chirp = hp.sm_filt + (hp.bluestein ? hp.line_slots : 0);
    hp.sm_cw = hp.sm_chirp + (hp.bluestein ? nw : 0);
    hp.sm_buf = hp.sm_cw + nw;
    int64_t slots_avail = smem_limit_bytes / 16 - hp.sm_buf;
    int teams = (int)std::min<int64_t>(max_teams, slots_avail / hp.line_slots);
    if (teams < 1) { hp.error = "window too large for the shared-memory FFT line"; return false; }
    teams = std::min(teams, std::max(1, hp.nvh));
    // Two CTAs of 8 teams per SM beat one of 16 (r01h: 84.0 vs 86.5 ms): the CTAs drift apart, so one
    // transforms while the other waits on its block barrier or on the bin gathers.  228 KB of shared
    // memory per SM, 1 KB reserved per CTA.
    if (teams == CB_SPEC_MAX_TEAMS && max_teams == CB_SPEC_MAX_TEAMS &&
        16 * ((int64_t)hp.sm_buf + 8 * (int64_t)hp.line_slots) <= (228 / 2 - 1) * 1024)
        teams = 8;
    hp.teams = teams;
    hp.threads = teams * CB_WARP;
    hp.smem_bytes = 16 * ((int64_t)hp.sm_buf + (int64_t)teams * hp.line_slots);
    hp.gt_stride = (((int64_t)hp.nvh * hp.npad) + 15) & ~15LL;
    hp.lh_stride = (((int64_t)hp.nvh * nw) + 31) & ~31LL;
    hp.ws_per_cta = hp.gt_stride * 16 + hp.lh_stride * 8;
    return true;
}

// Fills the plan-dependent part of the kernel parameters from pointers that live where the
// kernel will run (device pointers for the product, host pointers for tests/emu).
static inline void spectrum_fill_params(const SpectrumHostPlan &hp, SpectrumParams &P, const double2 *tw,
                                        const double2 *chirp, const double2 *filt, const int32_t *slot,
                                        const double *taper_r, const double *taper_c,
                                        const int32_t *bin_ptr, const int32_t *bin_idx,
                                        const int32_t *col_span) {
    P.n = hp.n; P.nvh = hp.nvh; P.npad = hp.npad; P.bluestein = hp.bluestein; P.fft = hp.fft;
    P.tw = tw; P.chirp = chirp; P.filt = filt; P.slot = slot;
    P.taper_r = hp.has_taper ? taper_r : nullptr;
    P.taper_c = hp.has_taper ? taper_c : nullptr;
    P.bin_ptr = bin_ptr; P.bin_idx = bin_idx; P.nbins = hp.nbins; P.col_span = col_span;
    P.gt_stride = hp.gt_stride; P.lh_stride = hp.lh_stride;
    P.sm_tw = hp.sm_tw; P.sm_filt = hp.sm_filt; P.sm_chirp = hp.sm_chirp; P.sm_cw = hp.sm_cw; P.sm_buf = hp.sm_buf;
    P.line_slots = hp.line_slots;
}
