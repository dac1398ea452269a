// Spectrum stage: plan objects, launches and C-ABI entry points (include/curie_b200.h).
// Kernel body: spectrum_body.cuh.  Replaces pycurious/grid.py:112-157, :199-216, :218-288.
#include <cstdlib>
#include <cstring>
#include <new>
#include "spectrum_plan.h"

typedef void (*spectrum_kernel_fn)(const SpectrumParams);

struct cb_spectrum_plan {
    SpectrumHostPlan hp;
    int device = 0, sm_count = 0, occupancy = 1, grid_max = 0;
    double2 *d_tw = nullptr, *d_chirp = nullptr, *d_filt = nullptr;
    int32_t *d_slot = nullptr, *d_bin_ptr = nullptr, *d_bin_idx = nullptr, *d_col_span = nullptr;
    double *d_taper_r = nullptr, *d_taper_c = nullptr;
    double2 *d_plane_t0 = nullptr, *d_plane_t1 = nullptr;
    spectrum_kernel_fn kernel = nullptr, row_kernel = nullptr, col_kernel = nullptr;
    const char *engine = "generic";
};

extern __shared__ double2 cb_dyn_smem[];

template <class Engine>
__global__ void __launch_bounds__(CB_SPEC_MAX_TEAMS * CB_WARP, 1)
spectrum_kernel(const __grid_constant__ SpectrumParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem)};
    spectrum_body<Engine>(cx, P);
}

template <class Engine>
__global__ void __launch_bounds__(CB_SPEC_MAX_TEAMS * CB_WARP, 1)
rowspec_kernel(const __grid_constant__ SpectrumParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem)};
    rowspec_body<Engine>(cx, P);
}

template <class Engine>
__global__ void __launch_bounds__(CB_SPEC_MAX_TEAMS * CB_WARP, 1)
colspec_kernel(const __grid_constant__ SpectrumParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem)};
    colspec_body<Engine>(cx, P);
}

// FFT engines compiled in: static pass lists for the Bluestein lengths of the BASELINE window sizes,
// the generic runtime-radix engine for everything else.
typedef StaticFft<8, 8, 8> Fft512;      // nw 201 (and 162..256; 257 through PlusOneFft)
typedef StaticFft<8, 8, 4> Fft256;      // nw 101 (and 129 through PlusOneFft)
typedef StaticFft<8, 8, 5> Fft320;      // nw 130..160 (and 161 through PlusOneFft)
typedef StaticFft<8, 10, 8> Fft640;     // nw 301

__global__ void __launch_bounds__(256) detrend_kernel(const __grid_constant__ DetrendParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem)};
    detrend_body(cx, P);
}

template <class T>
static int upload(T **dst, const std::vector<T> &src) {
    *dst = nullptr;
    if (src.empty()) return CB_OK;
    CB_CUDA_TRY(cudaMalloc((void **)dst, src.size() * sizeof(T)));
    CB_CUDA_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return CB_OK;
}

template <class Engine>
static void set_kernels(cb_spectrum_plan *pl, const char *name) {
    pl->kernel = spectrum_kernel<Engine>;
    pl->row_kernel = rowspec_kernel<Engine>;
    pl->col_kernel = colspec_kernel<Engine>;
    pl->engine = name;
}

static void pick_spectrum_kernels(cb_spectrum_plan *pl) {
    const SpectrumHostPlan &hp = pl->hp;
    if (hp.bluestein && !getenv("CB_SPEC_GENERIC")) {
        // lines of N/2 + 1 nodes on the length-N engine (Bluestein length 2n - 2, see PlusOneFft)
        if (PlusOneFft<Fft256>::serves(hp.fft, hp.n)) return set_kernels<PlusOneFft<Fft256>>(pl, "static 8.8.4 + 1");
        if (PlusOneFft<Fft320>::serves(hp.fft, hp.n)) return set_kernels<PlusOneFft<Fft320>>(pl, "static 8.8.5 + 1");
        if (PlusOneFft<Fft512>::serves(hp.fft, hp.n)) return set_kernels<PlusOneFft<Fft512>>(pl, "static 8.8.8 + 1");
        if (PlusOneFft<Fft640>::serves(hp.fft, hp.n)) return set_kernels<PlusOneFft<Fft640>>(pl, "static 8.10.8 + 1");
        if (Fft512::matches(hp.fft) && hp.n <= Fft512::FUSED_SPAN) return set_kernels<Fft512>(pl, "static 8.8.8");
        if (Fft256::matches(hp.fft) && hp.n <= Fft256::FUSED_SPAN) return set_kernels<Fft256>(pl, "static 8.8.4");
        if (Fft320::matches(hp.fft) && hp.n <= Fft320::FUSED_SPAN) return set_kernels<Fft320>(pl, "static 8.8.5");
        if (Fft640::matches(hp.fft) && hp.n <= Fft640::FUSED_SPAN) return set_kernels<Fft640>(pl, "static 8.10.8");
    }
    set_kernels<GenericFft>(pl, "generic");
}

extern "C" {

int cb_spectrum_plan_destroy(cb_spectrum_plan *plan) {
    if (!plan) return CB_OK;
    cudaFree(plan->d_tw); cudaFree(plan->d_chirp); cudaFree(plan->d_filt); cudaFree(plan->d_slot);
    cudaFree(plan->d_bin_ptr); cudaFree(plan->d_bin_idx); cudaFree(plan->d_taper_r); cudaFree(plan->d_taper_c);
    cudaFree(plan->d_plane_t0); cudaFree(plan->d_plane_t1); cudaFree(plan->d_col_span);
    delete plan;
    return CB_OK;
}

int cb_spectrum_plan_create(int nw, int nbins, const double *taper_r, const double *taper_c,
                            const int32_t *bin_ptr, const int32_t *bin_idx, cb_spectrum_plan **out) {
    CB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    CB_REQUIRE(bin_ptr && bin_idx, "bin CSR is NULL");
    CB_REQUIRE((taper_r == nullptr) == (taper_c == nullptr), "give both tapers or none");
    int rc = cb_check_device();
    if (rc) return rc;
    cb_spectrum_plan *pl = new (std::nothrow) cb_spectrum_plan();
    if (!pl) return CB_ERR_NOMEM;
    CB_CUDA_TRY(cudaGetDevice(&pl->device));
    cudaDeviceProp prop;
    CB_CUDA_TRY(cudaGetDeviceProperties(&prop, pl->device));
    pl->sm_count = prop.multiProcessorCount;
    // CB_SPEC_TEAMS (tuning knob): FFT teams (warps) per CTA; fewer teams -> more CTAs per SM, each on
    // its own window and in its own phase of the pipeline
    int max_teams = CB_SPEC_MAX_TEAMS;
    if (const char *env = getenv("CB_SPEC_TEAMS")) {
        const int t = atoi(env);
        if (t >= 1 && t <= CB_SPEC_MAX_TEAMS) max_teams = t;
    }
    if (!spectrum_plan_build_host(pl->hp, nw, nbins, taper_r, taper_c, bin_ptr, bin_idx,
                                  (int64_t)prop.sharedMemPerBlockOptin, max_teams)) {
        cb_set_error("spectrum plan for nw=%d: %s", nw, pl->hp.error.c_str());
        delete pl;
        return CB_ERR_UNSUPPORTED;
    }
    SpectrumHostPlan &hp = pl->hp;
#define UP(dst, src) if ((rc = upload(&pl->dst, hp.src)) != CB_OK) { cb_spectrum_plan_destroy(pl); return rc; }
    UP(d_tw, tw) UP(d_chirp, chirp) UP(d_filt, filt) UP(d_slot, slot)
    UP(d_bin_ptr, bin_ptr) UP(d_bin_idx, bin_idx) UP(d_taper_r, taper_r) UP(d_taper_c, taper_c)
    UP(d_plane_t0, plane_t0) UP(d_plane_t1, plane_t1) UP(d_col_span, col_span)
#undef UP
    // the attribute belongs to the kernel, not to the plan: raise it once to the device limit so that
    // plans with different line lengths can coexist
    pick_spectrum_kernels(pl);
    cudaError_t e = cudaSuccess;
    for (spectrum_kernel_fn fn : {pl->kernel, pl->row_kernel, pl->col_kernel}) {
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
        if (e != cudaSuccess) break;
    }
    if (e != cudaSuccess) {
        cb_set_error("cudaFuncSetAttribute(smem=%lld): %s", (long long)prop.sharedMemPerBlockOptin, cudaGetErrorString(e));
        cb_spectrum_plan_destroy(pl);
        return CB_ERR_CUDA;
    }
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pl->kernel, hp.threads, (size_t)hp.smem_bytes);
    if (e != cudaSuccess || occ < 1) {
        cb_set_error("spectrum kernel does not fit an SM (threads=%d smem=%lld): %s", hp.threads,
                     (long long)hp.smem_bytes, cudaGetErrorString(e));
        cb_spectrum_plan_destroy(pl);
        return CB_ERR_UNSUPPORTED;
    }
    pl->occupancy = occ;
    pl->grid_max = pl->sm_count * occ;
    *out = pl;
    return CB_OK;
}

int cb_spectrum_plan_get_info(const cb_spectrum_plan *plan, cb_spectrum_plan_info *info) {
    CB_REQUIRE(plan && info, "NULL argument");
    const SpectrumHostPlan &hp = plan->hp;
    memset(info, 0, sizeof(*info));
    info->nw = hp.n; info->nbins = hp.nbins; info->nvh = hp.nvh; info->bluestein = hp.bluestein;
    info->fft_len = hp.fft.n; info->npass = hp.fft.npass;
    for (int p = 0; p < hp.fft.npass; ++p) info->radix[p] = hp.fft.pass[p].radix;
    info->nnz_full = hp.nnz_full; info->nnz_folded = (int32_t)hp.bin_idx.size();
    info->teams = hp.teams; info->threads = hp.threads; info->grid_max = plan->grid_max;
    info->smem_bytes = hp.smem_bytes; info->workspace_bytes_per_cta = hp.ws_per_cta;
    return CB_OK;
}

int cb_spectrum_plan_get_bins(const cb_spectrum_plan *plan, int32_t *bin_ptr_host, int32_t *bin_idx_host) {
    CB_REQUIRE(plan && bin_ptr_host && bin_idx_host, "NULL argument");
    const SpectrumHostPlan &hp = plan->hp;
    // read back what the kernel actually uses
    CB_CUDA_TRY(cudaMemcpy(bin_ptr_host, plan->d_bin_ptr, hp.bin_ptr.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (!hp.bin_idx.empty())
        CB_CUDA_TRY(cudaMemcpy(bin_idx_host, plan->d_bin_idx, hp.bin_idx.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return CB_OK;
}

int64_t cb_spectrum_workspace_bytes(const cb_spectrum_plan *plan) {
    if (!plan) return 0;
    return (int64_t)plan->grid_max * plan->hp.ws_per_cta + 256;
}

int cb_spectrum_batch(const cb_spectrum_plan *plan, const double *grid_dev, int64_t ny, int64_t nx, int64_t ld,
                      const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin, int detrend, double power,
                      double *phi_dev, double *sigma_dev, void *workspace_dev, int64_t workspace_bytes,
                      void *stream) {
    CB_REQUIRE(plan != nullptr, "plan is NULL");
    CB_REQUIRE(nwin >= 0, "nwin < 0");
    if (nwin == 0) return CB_OK;
    CB_REQUIRE(grid_dev && iy0_dev && ix0_dev && phi_dev && sigma_dev && workspace_dev, "NULL buffer");
    const SpectrumHostPlan &hp = plan->hp;
    CB_REQUIRE(ny >= hp.n && nx >= hp.n && ld >= nx, "grid smaller than the window");
    const int grid = (int)std::min<int64_t>(nwin, plan->grid_max);
    uintptr_t base = ((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255;
    CB_REQUIRE((int64_t)(base - (uintptr_t)workspace_dev) + (int64_t)grid * hp.ws_per_cta <= workspace_bytes,
               "workspace too small (see cb_spectrum_workspace_bytes)");
    SpectrumParams P;
    memset(&P, 0, sizeof(P));
    spectrum_fill_params(hp, P, plan->d_tw, plan->d_chirp, plan->d_filt, plan->d_slot, plan->d_taper_r,
                         plan->d_taper_c, plan->d_bin_ptr, plan->d_bin_idx, plan->d_col_span);
    P.grid = grid_dev; P.ld = ld; P.iy0 = iy0_dev; P.ix0 = ix0_dev; P.nwin = nwin;
    P.detrend = detrend ? 1 : 0; P.power = power;
    P.phi = phi_dev; P.sigma = sigma_dev;
    P.gt = reinterpret_cast<double2 *>(base);
    P.lh = reinterpret_cast<double *>(base + (uintptr_t)grid * (uintptr_t)hp.gt_stride * 16);
    plan->kernel<<<grid, hp.threads, (size_t)hp.smem_bytes, (cudaStream_t)stream>>>(P);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

int64_t cb_spectrum_rowspec_bytes(const cb_spectrum_plan *plan, int ngroups, int row_lo, int row_hi) {
    if (!plan || ngroups < 0 || row_hi < row_lo) return 0;
    const int64_t nrows = ((int64_t)(row_hi - row_lo) + 1) & ~1LL;
    // row spectra [ngroups][nvh][nrows] + per-row sums for the detrend plane [ngroups][nrows], both complex128
    return (int64_t)ngroups * (plan->hp.nvh + 1) * nrows * 16 + 512;
}

int cb_spectrum_batch_shared_rows(const cb_spectrum_plan *plan, const double *grid_dev, int64_t ny, int64_t nx,
                                  int64_t ld, const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin,
                                  const int32_t *group_of_win_dev, const int32_t *group_ix0_dev, int ngroups,
                                  const int32_t *win_order_dev,
                                  int row_lo, int row_hi, int detrend, double power, double *phi_dev,
                                  double *sigma_dev, void *rowspec_dev, int64_t rowspec_bytes, void *workspace_dev,
                                  int64_t workspace_bytes, void *stream) {
    CB_REQUIRE(plan != nullptr, "plan is NULL");
    CB_REQUIRE(nwin >= 0, "nwin < 0");
    if (nwin == 0) return CB_OK;
    CB_REQUIRE(grid_dev && iy0_dev && ix0_dev && phi_dev && sigma_dev && workspace_dev && rowspec_dev, "NULL buffer");
    CB_REQUIRE(group_of_win_dev && group_ix0_dev && ngroups > 0, "missing window groups");
    const SpectrumHostPlan &hp = plan->hp;
    CB_REQUIRE(ny >= hp.n && nx >= hp.n && ld >= nx, "grid smaller than the window");
    CB_REQUIRE(row_lo >= 0 && row_hi <= ny && row_hi - row_lo >= hp.n, "row range must hold a whole window inside the grid");
    CB_REQUIRE(rowspec_bytes >= cb_spectrum_rowspec_bytes(plan, ngroups, row_lo, row_hi),
               "row-spectra buffer too small (see cb_spectrum_rowspec_bytes)");
    const int grid = (int)std::min<int64_t>(nwin, plan->grid_max);
    uintptr_t base = ((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255;
    CB_REQUIRE((int64_t)(base - (uintptr_t)workspace_dev) + (int64_t)grid * hp.ws_per_cta <= workspace_bytes,
               "workspace too small (see cb_spectrum_workspace_bytes)");
    SpectrumParams P;
    memset(&P, 0, sizeof(P));
    spectrum_fill_params(hp, P, plan->d_tw, plan->d_chirp, plan->d_filt, plan->d_slot, plan->d_taper_r,
                         plan->d_taper_c, plan->d_bin_ptr, plan->d_bin_idx, plan->d_col_span);
    P.grid = grid_dev; P.ld = ld; P.iy0 = iy0_dev; P.ix0 = ix0_dev; P.nwin = nwin;
    P.detrend = detrend ? 1 : 0; P.power = power;
    P.phi = phi_dev; P.sigma = sigma_dev;
    P.lh = reinterpret_cast<double *>(base);          // the G^T part of the slab is not used on this path
    P.lh_stride = hp.ws_per_cta / 8;
    P.lh_half = hp.lh_stride;                         // second slab (phase C of w overlaps phase B of w+1)
    P.group_of_win = group_of_win_dev; P.group_ix0 = group_ix0_dev; P.ngroups = ngroups;
    P.win_order = win_order_dev;
    P.row_lo = row_lo; P.nrows_valid = row_hi - row_lo; P.nrows = (P.nrows_valid + 1) & ~1;
    P.rt = reinterpret_cast<double2 *>(((uintptr_t)rowspec_dev + 255) & ~(uintptr_t)255);
    P.plane_t0 = plan->d_plane_t0; P.plane_t1 = plan->d_plane_t1;
    if (P.detrend) {
        const uintptr_t rs = (uintptr_t)(P.rt + (int64_t)ngroups * hp.nvh * P.nrows);
        P.rowsum = reinterpret_cast<double2 *>((rs + 255) & ~(uintptr_t)255);
    }
    const int64_t items = (int64_t)ngroups * (P.nrows / 2);
    const int row_grid = (int)std::min<int64_t>((items + hp.teams - 1) / hp.teams, plan->grid_max);
    plan->row_kernel<<<row_grid, hp.threads, (size_t)hp.smem_bytes, (cudaStream_t)stream>>>(P);
    CB_CUDA_TRY(cudaGetLastError());
    plan->col_kernel<<<grid, hp.threads, (size_t)hp.smem_bytes, (cudaStream_t)stream>>>(P);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

int cb_detrend_batch(const double *grid_dev, int64_t ny, int64_t nx, int64_t ld, int nw,
                     const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin, double *out_dev, void *stream) {
    CB_REQUIRE(nwin >= 0 && nw >= 2, "bad sizes");
    if (nwin == 0) return CB_OK;
    CB_REQUIRE(grid_dev && iy0_dev && ix0_dev && out_dev, "NULL buffer");
    CB_REQUIRE(ny >= nw && nx >= nw && ld >= nx, "grid smaller than the window");
    int rc = cb_check_device();
    if (rc) return rc;
    int dev = 0, sms = 0;
    CB_CUDA_TRY(cudaGetDevice(&dev));
    CB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DetrendParams P{grid_dev, ld, iy0_dev, ix0_dev, nwin, nw, out_dev};
    const int grid = (int)std::min<int64_t>(nwin, (int64_t)sms * 8);
    detrend_kernel<<<grid, 256, 96 * sizeof(double), (cudaStream_t)stream>>>(P);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

}  // extern "C"
