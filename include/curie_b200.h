/*
 * curie_b200 -- C-ABI of the B200-native per-window Curie-depth path.
 *
 * Drop-in boundary for brmather/pycurious (reference paths relative to
 * /root/reference).  The reference has no FFI for this path: its boundary is
 * the Python class API (pycurious/grid.py, pycurious/optimise.py), whose hot
 * loops run inside numpy/scipy.  Each entry point below replaces the body of
 * one reference function for a whole batch of windows; the Python classes in
 * pycurious_b200/ keep the reference's signatures and call these through
 * ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - plain C, no torch types; every function returns 0 or a negative cb_status
 *   - "dev" pointers are device (HBM) addresses, "host" pointers are host
 *   - the caller owns every buffer; the library owns only opaque plans
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - all floating point is IEEE binary64; window origins are int32
 *   - calls are asynchronous on `stream` unless stated otherwise
 *   - no CPU fallback: without a CUDA device every compute call returns
 *     CB_ERR_NO_DEVICE
 */
#ifndef CURIE_B200_H
#define CURIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cb_status {
    CB_OK = 0,
    CB_ERR_INVALID = -1,      /* bad argument */
    CB_ERR_NO_DEVICE = -2,    /* no CUDA device / driver */
    CB_ERR_CUDA = -3,         /* CUDA runtime error (see cb_last_error) */
    CB_ERR_UNSUPPORTED = -4,  /* size outside what the kernels support */
    CB_ERR_NOMEM = -5
} cb_status;

#define CB_ABI_VERSION 1
#define CB_MAX_RADIX_PASSES 12

int cb_abi_version(void);
/* Human readable text of the last error raised on the calling thread. */
const char *cb_last_error(void);
/* SM count, compute capability and opt-in shared memory of the current device. */
int cb_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *smem_optin_bytes);

/* Measures the FP64 FMA rate of the current device with a register-resident DFMA chain kernel
 * (roofline denominator for the Bessel / fit / MCMC kernels; MEASURED_PEAKS.json has no FP64 entry).
 * Synchronous; *tflops_out = 2 * DFMA/s / 1e12. */
int cb_fp64_fma_peak(double *tflops_out);

/* Page-lock / release a caller-owned HOST buffer in place (cudaHostRegisterPortable).  The reference reads
 * `self.data` live on every call (pycurious/grid.py:142-155); the drop-in therefore uploads the rows a call
 * needs from the caller's own array each time, and registering that array makes the upload an asynchronous
 * full-speed DMA with no staging copy.  Registering twice is not an error; failure (read-only mapping,
 * pinned-memory limit) is reported and the caller falls back to a pageable copy. */
int cb_host_register(void *host, int64_t bytes);
int cb_host_unregister(void *host);
/* Frees the scratch the library owns per (device, stream): the window-independent k tables of the model kernels
 * and the solver state of cb_fit_batch.  The caller guarantees that no launch of this library is in flight. */
int cb_release_scratch(void);

/* ------------------------------------------------------------------------
 * Spectrum stage:  CurieGrid.subgrid -> [remove_trend_linear] -> radial_spectrum
 * replaces pycurious/grid.py:112-157 (window gather), :199-216 (plane detrend),
 * :218-246 (taper) and :248-288 (fft2, modulus, fftshift, radial binning).
 * ---------------------------------------------------------------------- */
typedef struct cb_spectrum_plan cb_spectrum_plan;

/*
 * Build the plan for nw x nw windows.
 *   taper_r, taper_c   host, nw values each (row / column taper of grid.py:236-239);
 *                      NULL = no taper (grid.py:233-234)
 *   bin_ptr, bin_idx   host CSR with nbins rows: the members of radial bin i are the
 *                      flat indices a*nw+b (row a, column b of the fftSHIFTED modulus
 *                      array) listed in bin_idx[bin_ptr[i] .. bin_ptr[i+1]).  The host
 *                      builds it with the reference's own numpy expressions
 *                      (grid.py:277-282) so membership is bit-identical; the library
 *                      folds it onto the Hermitian half plane it computes.
 * The FFT length is arbitrary: 2/3/5-smooth nw run the mixed-radix engine directly,
 * anything else goes through Bluestein's chirp-z on a padded 2/3/5-smooth length.
 */
int cb_spectrum_plan_create(int nw, int nbins,
                            const double *taper_r, const double *taper_c,
                            const int32_t *bin_ptr, const int32_t *bin_idx,
                            cb_spectrum_plan **out);
int cb_spectrum_plan_destroy(cb_spectrum_plan *plan);

/* Introspection (tests, DESIGN.md numbers). */
typedef struct cb_spectrum_plan_info {
    int32_t nw, nbins, nvh;         /* window edge, bins, half-plane columns nw/2+1 */
    int32_t bluestein;              /* 1 = chirp-z path, 0 = direct mixed radix */
    int32_t fft_len;                /* length of the smooth FFT actually run */
    int32_t npass;
    int32_t radix[CB_MAX_RADIX_PASSES];
    int32_t nnz_full, nnz_folded;   /* CSR entries before / after Hermitian folding */
    int32_t teams, threads, grid_max;   /* launch shape: FFT teams per CTA, threads per CTA, CTAs */
    int64_t smem_bytes;             /* dynamic shared memory per CTA */
    int64_t workspace_bytes_per_cta;
} cb_spectrum_plan_info;
int cb_spectrum_plan_get_info(const cb_spectrum_plan *plan, cb_spectrum_plan_info *info);
/* Copies the folded device CSR back (bin_ptr: nbins+1, bin_idx: nnz_folded half-plane ids v*nw+u). */
int cb_spectrum_plan_get_bins(const cb_spectrum_plan *plan, int32_t *bin_ptr_host, int32_t *bin_idx_host);

/* Scratch the batch call needs (independent of nwin: one slab per resident CTA). */
int64_t cb_spectrum_workspace_bytes(const cb_spectrum_plan *plan);

/*
 * Phi[w, i], sigma[w, i]  (w < nwin, i < nbins) for the windows whose top-left
 * nodes are (iy0[w], ix0[w]) in the row-major grid with leading dimension ld.
 *   detrend != 0  subtracts the least-squares plane first (grid.py:199-216)
 *   power         the `power`/`const` factor of grid.py:283 (2.0 Bouligand, 0.5 Tanaka)
 * The caller validates that every window lies inside the grid (grid.py:148-153).
 */
int cb_spectrum_batch(const cb_spectrum_plan *plan,
                      const double *grid_dev, int64_t ny, int64_t nx, int64_t ld,
                      const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin,
                      int detrend, double power,
                      double *phi_dev, double *sigma_dev,
                      void *workspace_dev, int64_t workspace_bytes,
                      void *stream);

/*
 * Same result as cb_spectrum_batch for windows that sit on a lattice (create_centroid_list,
 * grid.py:159-197): windows that start at the same column ix0 share the x-transforms of the grid
 * rows they overlap (the row taper and the detrending plane commute with that transform), so those
 * are computed once per (distinct ix0, grid row) into `rowspec_dev` and each window only runs its
 * column transforms and the binning.  The host supplies the grouping:
 *   group_of_win_dev[w]   index into group_ix0_dev of window w (ix0[w] == group_ix0[group_of_win[w]])
 *   group_ix0_dev[g]      the distinct ix0 values, ngroups of them
 *   win_order_dev[i]      optional (NULL = 0..nwin-1): the order in which windows are processed; sorting by
 *                         (group, iy0) makes concurrently resident CTAs read overlapping rows (L2 hits)
 *   [row_lo, row_hi)      grid rows to transform for every group; every window needs
 *                         row_lo <= iy0[w] and iy0[w] + nw <= row_hi
 *   rowspec_dev           cb_spectrum_rowspec_bytes(plan, ngroups, row_lo, row_hi) bytes of scratch: the row
 *                         spectra [ngroups][nw/2+1][rows] and, for detrend, the per-row sums (sum d,
 *                         sum (x - xbar) d) [ngroups][rows] from which every window's plane is assembled
 * Worth it when nwin * nw is well above ngroups * (row_hi - row_lo), i.e. when rows are reused.
 */
int64_t cb_spectrum_rowspec_bytes(const cb_spectrum_plan *plan, int ngroups, int row_lo, int row_hi);
int cb_spectrum_batch_shared_rows(const cb_spectrum_plan *plan,
                                  const double *grid_dev, int64_t ny, int64_t nx, int64_t ld,
                                  const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin,
                                  const int32_t *group_of_win_dev, const int32_t *group_ix0_dev, int ngroups,
                                  const int32_t *win_order_dev,
                                  int row_lo, int row_hi,
                                  int detrend, double power,
                                  double *phi_dev, double *sigma_dev,
                                  void *rowspec_dev, int64_t rowspec_bytes,
                                  void *workspace_dev, int64_t workspace_bytes,
                                  void *stream);

/* Detrended windows themselves: out[w] = window - plane  (CurieGrid.remove_trend_linear). */
int cb_detrend_batch(const double *grid_dev, int64_t ny, int64_t nx, int64_t ld, int nw,
                     const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin,
                     double *out_dev /* [nwin, nw, nw] */, void *stream);

/* ------------------------------------------------------------------------
 * Forward model:  scipy.special.kv and bouligand2009 (pycurious/grid.py:552-600)
 * ---------------------------------------------------------------------- */
/* out[i] = K_{nu[i]}(x[i])  (Temme series x<=2, Steed CF2 x>2, upward recurrence). */
int cb_besselk(const double *nu_dev, const double *x_dev, int64_t n, double *out_dev, void *stream);
/* out[m, j] = bouligand2009(kh[j], params[m,0..3]); params rows are (beta, zt, dz, C). */
int cb_bouligand_forward(const double *kh_dev, int nk, const double *params_dev, int64_t m,
                         double *out_dev, void *stream);
/* out[i] = bouligand2009(kh[i], params[i,0..3]): fully element-wise (the caller broadcasts). */
int cb_bouligand_points(const double *kh_dev, const double *params_dev, int64_t n, double *out_dev, void *stream);
/* out[m] = CurieOptimise.min_func(params[m], kh, phi[m or 0]) incl. priors (optimise.py:196-230);
 * phi_stride = 0 shares one spectrum between all rows.  prior_mu[i] = NaN means no prior. */
int cb_misfit(const double *kh_dev, int nk, const double *phi_dev, int64_t phi_stride,
              const double *params_dev, int64_t m,
              const double *prior_mu_host, const double *prior_sigma_host,
              double *out_dev, void *stream);

/* ------------------------------------------------------------------------
 * MAP fit:  CurieOptimise.optimise / optimise_routine (optimise.py:232-301, :446-509)
 * Bounded quasi-Newton (projected BFGS seeded with the Gauss-Newton metric) on
 * min_func, in place of scipy's L-BFGS-B.  One call issues a sequence of
 * asynchronous launches on `stream` (model evaluation: one window per warp;
 * solver step: one window per thread; a straggler kernel) and returns without
 * synchronising.  The solver state between the launches lives in scratch the
 * library owns per (device, stream) -- about 1.5 KB per window, at most 262,144
 * windows at a time (larger inputs go through it in passes); cb_release_scratch
 * frees it.  A window's result does not depend on the batch it arrives in.
 * nk <= 512.
 * ---------------------------------------------------------------------- */
#define CB_FIT_FLAG_LINEAR_DZ 1   /* iterate in dz itself instead of u = exp(-2 k_min dz) */
#define CB_FIT_FLAG_FORWARD_FIRST 2 /* dPhi/dbeta by forward differences far from the minimum, central near it (default: central throughout) */
typedef struct cb_fit_options {
    double x0[4];            /* start (beta, zt, dz, C), used when x0_dev == NULL */
    double lb[4], ub[4];     /* bounds, +-INFINITY = unbounded (self.bounds, optimise.py:111-114) */
    double prior_mu[4];      /* NaN = no prior on that parameter (optimise.py:120-150) */
    double prior_sigma[4];
    int32_t max_iter;        /* 0 -> default 200 */
    int32_t flags;           /* CB_FIT_FLAG_*; 0 = defaults */
    double gtol;             /* projected-gradient inf-norm tolerance, 0 -> 1e-9 */
    double ftol;             /* relative decrease tolerance, 0 -> 1e-15 */
} cb_fit_options;

int cb_fit_batch(const double *kh_dev, int nk,
                 const double *phi_dev /* [nwin, nk] */, int64_t nwin,
                 const cb_fit_options *opt_host,
                 const double *x0_dev /* [nwin,4] or NULL */,
                 double *x_out_dev /* [nwin,4] */, double *f_out_dev /* [nwin] or NULL */,
                 int32_t *niter_dev /* [nwin] or NULL */, int32_t *status_dev /* [nwin] or NULL */,
                 int32_t *nfev_dev /* [nwin] or NULL: model evaluations of nk points each */,
                 void *stream);

/* The launch schedule cb_fit_batch uses for a pass of `nwin` windows: batched rounds at most, and the number of
 * unfinished windows below which the straggler kernel takes over (environment CB_FIT_ROUNDS / CB_FIT_HANDOVER
 * override both; results do not depend on either).  One call launches 3 + 2 * rounds + 1 kernels per pass. */
int cb_fit_schedule(int64_t nwin, int *rounds, int *handover);

/* Same fit with per-row prior means: prior_mu_dev[w, j] (NaN = keep opt->prior_mu[j]) replaces the
 * mean of the prior on parameter j for row w; a prior exists iff opt->prior_mu[j] is not NaN.  This is
 * the inner loop of CurieOptimise.sensitivity (optimise.py:705-716): each row is one simulation with its
 * own perturbed spectrum and its own draw of the prior means. */
int cb_fit_batch_rows(const double *kh_dev, int nk,
                      const double *phi_dev /* [nrow, nk] */, int64_t nrow,
                      const cb_fit_options *opt_host,
                      const double *x0_dev /* [nrow,4] or NULL */,
                      const double *prior_mu_dev /* [nrow,4] or NULL */,
                      double *x_out_dev, double *f_out_dev, int32_t *niter_dev, int32_t *status_dev,
                      int32_t *nfev_dev, void *stream);

/* out[(w*nsim + s), j] = phi[w, j] + sigma[w, j] * z,  z ~ N(0,1) from Philox4x32-10 keyed by
 * (seed, window_offset + w): the spectra perturbation of sensitivity (optimise.py:714), nsim rows per window. */
int cb_perturb_spectra(const double *phi_dev, const double *sigma_dev, int64_t nwin, int nk, int64_t nsim,
                       uint64_t seed, uint64_t window_offset, double *out_dev /* [nwin*nsim, nk] */, void *stream);

/* ------------------------------------------------------------------------
 * Metropolis-Hastings:  CurieOptimise.metropolis_hastings (optimise.py:511-633)
 * One chain per warp, Philox4x32-10 counter-based stream keyed by (seed, chain id).
 * ---------------------------------------------------------------------- */
typedef struct cb_mh_options {
    double x0[4];            /* start, used when x0_dev == NULL */
    double x_scale[4];       /* proposal scale (optimise.py:588-589) */
    double prior_mu[4], prior_sigma[4];
    int64_t nsim, burnin;
    uint64_t seed;
    uint64_t chain_offset;   /* id of window 0 (so shards of one job draw distinct streams) */
    int32_t reduce_mode;     /* 0: full chains out[nwin,4,nsim]; 1: moments out[nwin,10] =
                                mean(beta,zt,dz,C,zt+dz), std(beta,zt,dz,C,zt+dz) (population) */
    int32_t reserved;
} cb_mh_options;

int cb_mh_batch(const double *kh_dev, int nk,
                const double *phi_dev /* [nwin, nk] */, int64_t nwin,
                const cb_mh_options *opt_host,
                const double *x0_dev /* [nwin,4] or NULL */,
                double *out_dev, int64_t *accepted_dev /* [nwin] or NULL */,
                void *stream);

/* ------------------------------------------------------------------------
 * FFT-domain filters of a whole array:  out = real(ifft2(filter * fft2(data)))
 * replaces the numpy.fft.fft2 / ifft2 pairs of CurieGrid.reduce_to_pole (pycurious/grid.py:489-490) and
 * CurieGrid.upward_continuation (grid.py:545-546).  `filter_dev` is a complex128 [ny, nx] array (interleaved
 * re, im) built by the host with the reference's expressions; arbitrary ny, nx (Bluestein for non-smooth
 * lengths) up to the shared-memory line limit (~14,000 points smooth, ~6,000 otherwise).
 * ---------------------------------------------------------------------- */
typedef struct cb_fft_plan cb_fft_plan;
int cb_fft_plan_create(int n, cb_fft_plan **out);
int cb_fft_plan_destroy(cb_fft_plan *plan);
int64_t cb_fft2_filter_workspace_bytes(int64_t ny, int64_t nx);
int cb_fft2_filter(const cb_fft_plan *plan_x /* length nx */, const cb_fft_plan *plan_y /* length ny */,
                   const double *data_dev, int64_t ny, int64_t nx, int64_t ld,
                   const void *filter_dev /* complex128 [ny, nx] */,
                   double *out_dev, int64_t out_ld,
                   void *workspace_dev, int64_t workspace_bytes, void *stream);

/* The same filter applied to a batch of n x n windows cut from a survey grid at (iy0[i], ix0[i]) -- the
 * `process_subgrid` stage of CurieOptimise.optimise (pycurious/optimise.py:285-287) when the callable is
 * reduce_to_pole / upward_continuation (grid.py:405-546).  `plan` is the length-n plan, `filter_dev` the
 * complex128 [n, n] operator, `out_dev` a dense [nwin, n, n] stack that cb_spectrum_batch can read as a grid of
 * nwin*n rows.  The caller validates the origins (0 <= iy0 <= ny-n, 0 <= ix0 <= nx-n).  Real windows have Hermitian
 * spectra and real(ifft2(op * .)) acts through the Hermitian part of `op` only: the library forms that part on the
 * half plane and runs row pairs / half the columns (4 * ceil(n/2) line transforms per window instead of 4 n); on
 * Bluestein lines (n not 2/3/5-smooth) the forward and inverse column transforms share one visit of the line. */
int64_t cb_fft2_filter_windows_workspace_bytes(int n, int64_t nwin);
int cb_fft2_filter_windows(const cb_fft_plan *plan, const double *grid_dev, int64_t ny, int64_t nx, int64_t ld,
                           const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin, const void *filter_dev,
                           double *out_dev /* [nwin, n, n] */, void *workspace_dev, int64_t workspace_bytes,
                           void *stream);

/* ------------------------------------------------------------------------
 * Tanaka et al. (1999):  tanaka1999 + ComputeTanaka (pycurious/grid.py:603-699) for a batch of
 * radial spectra taken with power = 0.5.  out[w, 0..7] = Ztr, btr, dZtr, Zor, bor, dZor, Zb, eZb;
 * (kmin_lo, kmin_hi) / (kmax_lo, kmax_hi) are the reference's kmin_range / kmax_range.
 * ---------------------------------------------------------------------- */
int cb_tanaka_batch(const double *k_dev, int nk,
                    const double *lnphi_dev /* [nwin, nk] */, const double *sigma_dev /* [nwin, nk] */,
                    int64_t nwin, double kmin_lo, double kmin_hi, double kmax_lo, double kmax_hi,
                    double *out_dev /* [nwin, 8] */, void *stream);

/* ------------------------------------------------------------------------
 * Radon transform:  src/cradon.c:4-246 + src/radon.pyx:12-120
 * ---------------------------------------------------------------------- */
typedef struct cb_radon_plan cb_radon_plan;
/*
 * Traces ntheta*nw rays through the nw x nw unit-cell grid on the device, one
 * thread per ray, and keeps the CSR (cradon.c semantics: cell id ix*ny+iy, int64
 * indices/indptr, float64 lengths).  Tx/Rx: host [ntheta, nw, 2] ray end points,
 * generated by the host with the reference's numpy expressions (radon.pyx:30-96).
 * Synchronous.
 */
int cb_radon_plan_create(int nw, int ntheta, const double *tx_host, const double *rx_host,
                         cb_radon_plan **out);
int cb_radon_plan_destroy(cb_radon_plan *plan);
int64_t cb_radon_plan_nnz(const cb_radon_plan *plan);
/* indptr_host: ntheta*(nw+1) entries, per-angle local offsets like the reference's indptr. */
int cb_radon_plan_get_csr(const cb_radon_plan *plan, double *data_host, int64_t *indices_host,
                          int64_t *indptr_host);
/* sinogram[w, r, t] = sum over the CSR row (t, r) of length * window_w.flat[cell]. */
int cb_radon_apply_batch(const cb_radon_plan *plan,
                         const double *grid_dev, int64_t ny, int64_t nx, int64_t ld,
                         const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin,
                         double *sinogram_dev /* [nwin, nw, ntheta] */, void *stream);
/* S[w, t, j] = power * log|DFT_nk(taper * sinogram[w, :, t])[j+1]|, j < nkb, nk = 1 + 2*nkb
 * (the 1-D tail of CurieGrid.azimuthal_spectrum, grid.py:393-401). taper_host NULL = none. */
int cb_azimuthal_tail_batch(const double *sinogram_dev, int nw, int ntheta, int64_t nwin,
                            const double *taper_host, int nkb, double power,
                            double *S_dev /* [nwin, ntheta, nkb] */, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CURIE_B200_H */
