// Forward model hooks, batched MAP fit and Metropolis-Hastings: kernels + C-ABI entry points.
// Bodies: besselk.cuh, bouligand.cuh, fit_body.cuh, mh_body.cuh.
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include "fit_body.cuh"
#include "mh_body.cuh"
#include "tanaka_body.cuh"

extern __shared__ double cb_dyn_smem_d[];

// Misfit and Metropolis-Hastings: one warp (= one spectrum / chain) per CTA -- chain lengths vary between windows,
// and with single-warp CTAs the hardware block scheduler balances them.  These kernels are dependency-latency
// bound (one long serial chain per window): throughput scales with resident warps, so the slab is kept at ~9 KB
// (window-independent tables live in global memory) and registers at 80 for 24 CTAs per SM.
#define CB_FIT_WARPS 1
#ifndef CB_FIT_CTAS_PER_SM
#define CB_FIT_CTAS_PER_SM 24
#endif
#ifndef CB_MH_CTAS_PER_SM
#define CB_MH_CTAS_PER_SM 24
#endif

__global__ void __launch_bounds__(256) besselk_kernel(const double *nu, const double *x, int64_t n, double *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = cb_kv(nu[i], x[i]);
}

__global__ void __launch_bounds__(256)
bouligand_kernel(const double *kh, int nk, const double *params, int64_t m, double *out) {
    const int64_t total = m * (int64_t)nk;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / nk;
        const int j = (int)(i - row * nk);
        const double *p = params + 4 * row;
        const ModelCoef mc = model_coef(p[0], p[1], p[2], p[3]);
        const double k = kh[j];
        out[i] = bouligand_point(mc, k, log(k));
    }
}

__global__ void __launch_bounds__(256)
bouligand_points_kernel(const double *kh, const double *params, int64_t n, double *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double *p = params + 4 * i;
        const ModelCoef mc = model_coef(p[0], p[1], p[2], p[3]);
        const double k = kh[i];
        out[i] = bouligand_point(mc, k, log(k));
    }
}

struct MisfitParams {
    const double *kh;
    const double *ktab;
    int32_t nk;
    const double *phi;
    int64_t phi_stride;
    const double *params;
    int64_t m;
    Priors pr;
    double *out;
};

__global__ void __launch_bounds__(CB_FIT_WARPS * CB_WARP, CB_FIT_CTAS_PER_SM) misfit_kernel(const __grid_constant__ MisfitParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem_d)};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * CB_FIT_WARPS + warp;
    if (row >= P.m) return;
    WarpSlab W = warp_slab_carve(cb_dyn_smem_d + (int64_t)warp * warp_slab_doubles(P.nk, false), P.nk, false, P.kh, P.ktab);
    warp_slab_load(cx, W, P.phi + row * P.phi_stride, lane);
    double x[4];
    for (int j = 0; j < 4; ++j) x[j] = P.params[row * 4 + j];
    const double f = misfit_slab(cx, W, x, P.pr, lane);
    if (lane == 0) P.out[row] = f;
}

// The MAP fit as rounds of two kernels over the list of unfinished windows (fit_body.cuh): the fused model
// evaluation (one warp per window, persistent grid, entries claimed from a counter) and the scalar solver step (one
// thread per window); fit_finish_kernel runs the stragglers to the end, one warp each.
#ifndef CB_FIT_EVAL_WARPS
#define CB_FIT_EVAL_WARPS 8
#endif
#ifndef CB_FIT_EVAL_CTAS_PER_SM
#define CB_FIT_EVAL_CTAS_PER_SM 3
#endif
#ifndef CB_FIT_FINISH_WARPS
#define CB_FIT_FINISH_WARPS 4
#endif
__global__ void fit_template_kernel(const __grid_constant__ FitParams P) { fit_template_body(P); }
__global__ void __launch_bounds__(256, 2) fit_begin_kernel(const __grid_constant__ FitParams P) {
    DevCtx cx{nullptr};
    fit_begin_body(cx, P);
}
__global__ void __launch_bounds__(CB_FIT_EVAL_WARPS * CB_WARP, CB_FIT_EVAL_CTAS_PER_SM) fit_eval_kernel(const __grid_constant__ FitParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem_d)};
    fit_eval_body(cx, P);
}
__global__ void __launch_bounds__(256) fit_first_kernel(const __grid_constant__ FitParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem_d)};
    fit_first_body(cx, P);
}
#ifndef CB_FIT_STEP_BLOCKS
#define CB_FIT_STEP_BLOCKS 4
#endif
__global__ void __launch_bounds__(128, CB_FIT_STEP_BLOCKS) fit_step_kernel(const __grid_constant__ FitParams P) {
    DevCtx cx{nullptr};
    fit_step_body(cx, P);
}
__global__ void __launch_bounds__(CB_FIT_FINISH_WARPS * CB_WARP) fit_finish_kernel(const __grid_constant__ FitParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem_d)};
    fit_finish_body(cx, P);
}

__global__ void __launch_bounds__(CB_FIT_WARPS * CB_WARP, CB_MH_CTAS_PER_SM) mh_kernel(const __grid_constant__ MhParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem_d)};
    mh_body(cx, P);
}

// rPhi = Phi + sigma_Phi * N(0,1) (optimise.py:714): row (w, s) of the output is simulation s of window w.
// Philox4x32-10 keyed by (seed, global window index), counter = (simulation, quad of bins, 0, phase 2).
__global__ void __launch_bounds__(256)
perturb_kernel(const double *phi, const double *sigma, int64_t nwin, int nk, int64_t nsim, uint64_t seed,
               uint64_t window_offset, double *out) {
    const int nq = (nk + 3) / 4;
    const int64_t total = nwin * nsim * nq;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % nq);
        const int64_t row = i / nq, w = row / nsim, sim = row - w * nsim;
        const uint64_t chain = window_offset + (uint64_t)w;
        const uint32_t k0 = (uint32_t)(seed ^ chain), k1 = (uint32_t)((seed >> 32) ^ (chain >> 32));
        double z[4];
        philox_normal4((uint64_t)sim * (uint64_t)nq + (uint64_t)q, 2u, k0, k1, z);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = 4 * q + t;
            if (j < nk) out[row * nk + j] = fma(sigma[w * nk + j], z[t], phi[w * nk + j]);
        }
    }
}

__global__ void __launch_bounds__(128) tanaka_kernel(const __grid_constant__ TanakaParams P) {
    DevCtx cx{nullptr};
    tanaka_body(cx, P);
}

__global__ void __launch_bounds__(256) ktable_kernel(const double *kh, int nk, double *out) {
    const int n = ktable_doubles(nk);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = ktable_entry(kh, nk, i);
}

// Scratch for the window-independent tables: one small buffer per (device, stream), created on first use and
// kept (launches of one stream are ordered, so the fill of the next call cannot overtake the kernel of the
// previous one; different streams never share a buffer).  cudaMallocAsync per call was measured 4x slower
// for back-to-back launches.
#include <map>
#include <mutex>
#define CB_KTABLE_MAX_NK 2048
static std::mutex g_ktab_mutex;
static std::map<std::pair<int, cudaStream_t>, double *> g_ktab;

static int ktable_make(const double *kh_dev, int nk, cudaStream_t st, double **out) {
    *out = nullptr;
    CB_REQUIRE(nk <= CB_KTABLE_MAX_NK, "nk too large");
    int dev = 0;
    CB_CUDA_TRY(cudaGetDevice(&dev));
    double *buf = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_ktab_mutex);
        auto it = g_ktab.find(std::make_pair(dev, st));
        if (it != g_ktab.end()) buf = it->second;
        else {
            CB_CUDA_TRY(cudaMalloc((void **)&buf, (size_t)ktable_doubles(CB_KTABLE_MAX_NK) * sizeof(double)));
            g_ktab[std::make_pair(dev, st)] = buf;
        }
    }
    const int n = ktable_doubles(nk);
    ktable_kernel<<<(n + 255) / 256, 256, 0, st>>>(kh_dev, nk, buf);
    CB_CUDA_TRY(cudaGetLastError());
    *out = buf;
    return CB_OK;
}

// Scratch of the batched fit, one buffer per (device, stream), grown on demand and kept (same ownership rule as the
// k tables above: launches of one stream are ordered, different streams never share a buffer).
#define CB_FIT_PASS_WINDOWS 262144
static std::map<std::pair<int, cudaStream_t>, std::pair<char *, size_t>> g_fit_scratch;
static int fit_scratch(cudaStream_t st, size_t bytes, char **out) {
    int dev = 0;
    CB_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_ktab_mutex);
    auto &slot = g_fit_scratch[std::make_pair(dev, st)];
    if (slot.second < bytes) {
        if (slot.first) {
            CB_CUDA_TRY(cudaStreamSynchronize(st));      // a fit of this stream may still be using the old buffer
            CB_CUDA_TRY(cudaFree(slot.first));
            slot.first = nullptr; slot.second = 0;
        }
        CB_CUDA_TRY(cudaMalloc((void **)&slot.first, bytes));
        slot.second = bytes;
    }
    *out = slot.first;
    return CB_OK;
}

static int elementwise_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

extern "C" {

int cb_release_scratch(void) {
    std::lock_guard<std::mutex> lock(g_ktab_mutex);
    int dev = 0;
    CB_CUDA_TRY(cudaGetDevice(&dev));
    for (auto &kv : g_ktab) {
        CB_CUDA_TRY(cudaSetDevice(kv.first.first));
        cudaFree(kv.second);
    }
    g_ktab.clear();
    for (auto &kv : g_fit_scratch) {
        CB_CUDA_TRY(cudaSetDevice(kv.first.first));
        cudaFree(kv.second.first);
    }
    g_fit_scratch.clear();
    CB_CUDA_TRY(cudaSetDevice(dev));
    return CB_OK;
}

int cb_besselk(const double *nu_dev, const double *x_dev, int64_t n, double *out_dev, void *stream) {
    CB_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return CB_OK;
    CB_REQUIRE(nu_dev && x_dev && out_dev, "NULL buffer");
    int rc = cb_check_device();
    if (rc) return rc;
    besselk_kernel<<<elementwise_grid(n), 256, 0, (cudaStream_t)stream>>>(nu_dev, x_dev, n, out_dev);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

int cb_bouligand_forward(const double *kh_dev, int nk, const double *params_dev, int64_t m, double *out_dev,
                         void *stream) {
    CB_REQUIRE(nk > 0 && m >= 0, "bad sizes");
    if (m == 0) return CB_OK;
    CB_REQUIRE(kh_dev && params_dev && out_dev, "NULL buffer");
    int rc = cb_check_device();
    if (rc) return rc;
    bouligand_kernel<<<elementwise_grid(m * nk), 256, 0, (cudaStream_t)stream>>>(kh_dev, nk, params_dev, m, out_dev);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

int cb_bouligand_points(const double *kh_dev, const double *params_dev, int64_t n, double *out_dev, void *stream) {
    CB_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return CB_OK;
    CB_REQUIRE(kh_dev && params_dev && out_dev, "NULL buffer");
    int rc = cb_check_device();
    if (rc) return rc;
    bouligand_points_kernel<<<elementwise_grid(n), 256, 0, (cudaStream_t)stream>>>(kh_dev, params_dev, n, out_dev);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

enum { SLAB_MISFIT = 0, SLAB_FIT = 1, SLAB_MH = 2 };
static int per_warp_launch_shape(int nk, int64_t nwin, int kind, int *grid, size_t *smem) {
    CB_REQUIRE(nk > 0 && nk <= 512, "nk must be in [1, 512]");
    CB_REQUIRE(kind != SLAB_FIT, "the fit has its own launch shape");
    const int slab = kind == SLAB_MH ? mh_slab_doubles(nk) : warp_slab_doubles(nk, false);
    *smem = (size_t)CB_FIT_WARPS * (size_t)slab * sizeof(double);
    int64_t g = (nwin + CB_FIT_WARPS - 1) / CB_FIT_WARPS;
    CB_REQUIRE(g <= 2147483647LL, "too many windows for one launch");
    *grid = (int)g;
    return CB_OK;
}

int cb_misfit(const double *kh_dev, int nk, const double *phi_dev, int64_t phi_stride, const double *params_dev,
              int64_t m, const double *prior_mu_host, const double *prior_sigma_host, double *out_dev,
              void *stream) {
    CB_REQUIRE(m >= 0, "m < 0");
    if (m == 0) return CB_OK;
    CB_REQUIRE(kh_dev && phi_dev && params_dev && out_dev, "NULL buffer");
    int rc = cb_check_device();
    if (rc) return rc;
    int grid;
    size_t smem;
    if ((rc = per_warp_launch_shape(nk, m, SLAB_MISFIT, &grid, &smem)) != CB_OK) return rc;
    MisfitParams P;
    P.kh = kh_dev; P.nk = nk; P.phi = phi_dev; P.phi_stride = phi_stride; P.params = params_dev; P.m = m; P.out = out_dev;
    for (int j = 0; j < 4; ++j) {
        P.pr.mu[j] = prior_mu_host ? prior_mu_host[j] : NAN;
        P.pr.sigma[j] = prior_sigma_host ? prior_sigma_host[j] : NAN;
    }
    if (smem > 48 * 1024) CB_CUDA_TRY(cudaFuncSetAttribute(misfit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    double *ktab = nullptr;
    if ((rc = ktable_make(kh_dev, nk, (cudaStream_t)stream, &ktab)) != CB_OK) return rc;
    P.ktab = ktab;
    misfit_kernel<<<grid, CB_FIT_WARPS * CB_WARP, smem, (cudaStream_t)stream>>>(P);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

// Batched rounds and hand-over threshold for a pass of `nwin` windows.  A late round costs a launch pair whose
// duration no longer shrinks with the list (the step kernel's ~40 us of dependent solver algebra, ~50 us for an
// evaluation), so a small pass hands its stragglers over sooner: measured on one B200 (tools/gpu_exp_fit_small.sh),
// 19,012 windows 2.68 ms at (12 rounds, 1024) against 2.87 at (20, 2048); 38,025 windows 4.15 at (16, 2048) against
// 4.28; 76,050 and 152,100 windows are fastest at (20, 2048).
int cb_fit_schedule(int64_t nwin, int *rounds, int *handover) {
    int r = nwin <= 24576 ? 12 : (nwin <= 49152 ? 16 : 20), h = nwin <= 24576 ? 1024 : 2048;
    if (const char *e = getenv("CB_FIT_ROUNDS")) r = atoi(e);
    if (const char *e = getenv("CB_FIT_HANDOVER")) h = atoi(e);
    if (r < 1) r = 1;
    if (h < 0) h = 0;
    if (rounds) *rounds = r;
    if (handover) *handover = h;
    return CB_OK;
}

int cb_fit_batch_rows(const double *kh_dev, int nk, const double *phi_dev, int64_t nwin,
                      const cb_fit_options *opt_host, const double *x0_dev, const double *prior_mu_dev,
                      double *x_out_dev, double *f_out_dev, int32_t *niter_dev, int32_t *status_dev,
                      int32_t *nfev_dev, void *stream) {
    CB_REQUIRE(nwin >= 0, "nwin < 0");
    if (nwin == 0) return CB_OK;
    CB_REQUIRE(kh_dev && phi_dev && opt_host && x_out_dev, "NULL buffer");
    for (int j = 0; j < 4; ++j) {
        CB_REQUIRE(!(opt_host->lb[j] > opt_host->ub[j]), "lower bound above upper bound");
        CB_REQUIRE(isnan(opt_host->prior_mu[j]) || opt_host->prior_sigma[j] > 0.0, "prior sigma must be > 0");
    }
    int rc = cb_check_device();
    if (rc) return rc;
    CB_REQUIRE(nk > 0 && nk <= 512, "nk must be in [1, 512]");
    // tuning / test knobs, read per call: CB_FIT_HANDOVER = unfinished windows below which the finish kernel takes
    // over (0: never before the last round), CB_FIT_ROUNDS = batched rounds at most, CB_FIT_PASS = windows per pass
    // of the scratch; every setting gives the same results (tests/test_gpu_parity.py::test_fit_paths_agree)
    int64_t env_pass = CB_FIT_PASS_WINDOWS;
    if (const char *e = getenv("CB_FIT_PASS")) env_pass = atoll(e);
    if (env_pass < 1 || env_pass > CB_FIT_PASS_WINDOWS) env_pass = CB_FIT_PASS_WINDOWS;
    int dev_id = 0, sm_count = 0;
    CB_CUDA_TRY(cudaGetDevice(&dev_id));
    CB_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev_id));
    cudaStream_t st = (cudaStream_t)stream;
    double *ktab = nullptr;
    if ((rc = ktable_make(kh_dev, nk, st, &ktab)) != CB_OK) return rc;
    CB_CUDA_TRY(cudaFuncSetAttribute(fit_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fit_eval_smem_bytes(CB_FIT_EVAL_WARPS, nk)));
    if (fit_finish_smem_bytes(nk) > 48 * 1024)
        CB_CUDA_TRY(cudaFuncSetAttribute(fit_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fit_finish_smem_bytes(nk)));
    // library-owned scratch per (device, stream): state, coefficients, sums, two lists, counters; inputs larger
    // than CB_FIT_PASS_WINDOWS go through it in passes
    const int64_t pass = nwin < env_pass ? nwin : env_pass;
    const size_t per_win = (size_t)(CB_FIT_PERSIST_DOUBLES + 3 * FC_COUNT + 16) * sizeof(double) + 2 * sizeof(int32_t);
    const size_t head = 64 + (size_t)((CB_FIT_STATE_DOUBLES + 7) & ~7) * sizeof(double);
    char *scratch = nullptr;
    if ((rc = fit_scratch(st, (size_t)pass * per_win + head, &scratch)) != CB_OK) return rc;
    for (int64_t off = 0; off < nwin; off += pass) {
        const int64_t n = nwin - off < pass ? nwin - off : pass;
        FitParams P;
        memset(&P, 0, sizeof(P));
        P.kh = kh_dev; P.ktab = ktab; P.nk = nk; P.phi = phi_dev + off * nk; P.nwin = n; P.opt = *opt_host;
        P.x0_dev = x0_dev ? x0_dev + off * 4 : nullptr;
        P.prior_mu_dev = prior_mu_dev ? prior_mu_dev + off * 4 : nullptr;
        P.x_out = x_out_dev + off * 4;
        P.f_out = f_out_dev ? f_out_dev + off : nullptr;
        P.niter = niter_dev ? niter_dev + off : nullptr;
        P.status = status_dev ? status_dev + off : nullptr;
        P.nfev = nfev_dev ? nfev_dev + off : nullptr;
        P.count = reinterpret_cast<unsigned int *>(scratch);
        P.tmpl = reinterpret_cast<double *>(scratch + 64);
        P.state = reinterpret_cast<double *>(scratch + head);
        P.stride = pass;
        P.coef = P.state + (size_t)pass * CB_FIT_PERSIST_DOUBLES;
        P.sums = P.coef + (size_t)pass * 3 * FC_COUNT;
        int32_t *la = reinterpret_cast<int32_t *>(P.sums + (size_t)pass * 16), *lb = la + pass;
        P.list_in = la; P.list_out = lb;
        int env_rounds = 0, env_handover = 0;
        cb_fit_schedule(n, &env_rounds, &env_handover);
        P.handover = env_handover;
        fit_template_kernel<<<1, 1, 0, st>>>(P);
        fit_begin_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(P);
        int64_t g = (n + CB_FIT_EVAL_WARPS - 1) / CB_FIT_EVAL_WARPS;
        const int eval_grid = (int)(g < (int64_t)sm_count * CB_FIT_EVAL_CTAS_PER_SM ? g : (int64_t)sm_count * CB_FIT_EVAL_CTAS_PER_SM);
        g = (n + 127) / 128;
        const int step_grid = (int)(g < (int64_t)sm_count * 16 ? g : (int64_t)sm_count * 16);
        for (int r = 0; r < env_rounds; ++r) {
            P.round = r;
            if (r == 0 && !P.x0_dev) fit_first_kernel<<<eval_grid, 256, fit_first_smem_bytes(nk), st>>>(P);
            else fit_eval_kernel<<<eval_grid, CB_FIT_EVAL_WARPS * CB_WARP, fit_eval_smem_bytes(CB_FIT_EVAL_WARPS, nk), st>>>(P);
            fit_step_kernel<<<step_grid, 128, 0, st>>>(P);
            int32_t *t = P.list_in; P.list_in = P.list_out; P.list_out = t;
        }
        P.round = env_rounds;
        const int fin_grid = (int)(n < (int64_t)sm_count * 8 ? n : (int64_t)sm_count * 8);
        fit_finish_kernel<<<fin_grid, CB_FIT_FINISH_WARPS * CB_WARP, fit_finish_smem_bytes(nk), st>>>(P);
    }
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

int cb_fit_batch(const double *kh_dev, int nk, const double *phi_dev, int64_t nwin, const cb_fit_options *opt_host,
                 const double *x0_dev, double *x_out_dev, double *f_out_dev, int32_t *niter_dev,
                 int32_t *status_dev, int32_t *nfev_dev, void *stream) {
    return cb_fit_batch_rows(kh_dev, nk, phi_dev, nwin, opt_host, x0_dev, nullptr, x_out_dev, f_out_dev, niter_dev,
                             status_dev, nfev_dev, stream);
}

int cb_perturb_spectra(const double *phi_dev, const double *sigma_dev, int64_t nwin, int nk, int64_t nsim,
                       uint64_t seed, uint64_t window_offset, double *out_dev, void *stream) {
    CB_REQUIRE(nwin >= 0 && nsim >= 0 && nk > 0, "bad sizes");
    if (nwin == 0 || nsim == 0) return CB_OK;
    CB_REQUIRE(phi_dev && sigma_dev && out_dev, "NULL buffer");
    int rc = cb_check_device();
    if (rc) return rc;
    const int64_t quads = nwin * nsim * (((int64_t)nk + 3) / 4);
    perturb_kernel<<<elementwise_grid(quads), 256, 0, (cudaStream_t)stream>>>(phi_dev, sigma_dev, nwin, nk, nsim, seed,
                                                                              window_offset, out_dev);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

int cb_tanaka_batch(const double *k_dev, int nk, const double *lnphi_dev, const double *sigma_dev, int64_t nwin,
                    double kmin_lo, double kmin_hi, double kmax_lo, double kmax_hi, double *out_dev, void *stream) {
    CB_REQUIRE(nwin >= 0 && nk > 0, "bad sizes");
    if (nwin == 0) return CB_OK;
    CB_REQUIRE(k_dev && lnphi_dev && sigma_dev && out_dev, "NULL buffer");
    int rc = cb_check_device();
    if (rc) return rc;
    TanakaParams P;
    P.k = k_dev; P.nk = nk; P.lnphi = lnphi_dev; P.sigma = sigma_dev; P.nwin = nwin;
    P.top_lo = kmin_lo; P.top_hi = kmin_hi; P.cen_lo = kmax_lo; P.cen_hi = kmax_hi;
    P.out = out_dev; P.warps_per_block = 4;
    const int64_t g = (nwin + 3) / 4;
    CB_REQUIRE(g <= 2147483647LL, "too many windows for one launch");
    tanaka_kernel<<<(int)g, 128, 0, (cudaStream_t)stream>>>(P);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

int cb_mh_batch(const double *kh_dev, int nk, const double *phi_dev, int64_t nwin, const cb_mh_options *opt_host,
                const double *x0_dev, double *out_dev, int64_t *accepted_dev, void *stream) {
    CB_REQUIRE(nwin >= 0, "nwin < 0");
    if (nwin == 0) return CB_OK;
    CB_REQUIRE(kh_dev && phi_dev && opt_host && out_dev, "NULL buffer");
    CB_REQUIRE(opt_host->nsim >= 0 && opt_host->burnin >= 0, "negative chain length");
    CB_REQUIRE(opt_host->reduce_mode == 0 || opt_host->reduce_mode == 1, "reduce_mode must be 0 or 1");
    int rc = cb_check_device();
    if (rc) return rc;
    int grid;
    size_t smem;
    if ((rc = per_warp_launch_shape(nk, nwin, SLAB_MH, &grid, &smem)) != CB_OK) return rc;
    MhParams P;
    memset(&P, 0, sizeof(P));
    P.kh = kh_dev; P.nk = nk; P.phi = phi_dev; P.nwin = nwin; P.opt = *opt_host; P.x0_dev = x0_dev;
    P.out = out_dev; P.accepted = accepted_dev; P.warps_per_block = CB_FIT_WARPS;
    if (smem > 48 * 1024) CB_CUDA_TRY(cudaFuncSetAttribute(mh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    double *ktab = nullptr;
    if ((rc = ktable_make(kh_dev, nk, (cudaStream_t)stream, &ktab)) != CB_OK) return rc;
    P.ktab = ktab;
    mh_kernel<<<grid, CB_FIT_WARPS * CB_WARP, smem, (cudaStream_t)stream>>>(P);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

}  // extern "C"
