// Radon transform / azimuthal spectrum: plan, launches, C-ABI.  Body: radon_body.cuh.
// Compiled with -fmad=false (see Makefile) for bit-exactness against src/cradon.c.
#include <cstring>
#include <new>
#include <vector>
#include "radon_body.cuh"

struct cb_radon_plan {
    int nw = 0, ntheta = 0;
    int64_t nray = 0, nnz = 0;
    std::vector<int64_t> offset_host;   // [nray+1]
    double *d_tx = nullptr, *d_rx = nullptr, *d_data = nullptr;
    int64_t *d_offset = nullptr, *d_indices = nullptr;
};

extern __shared__ double2 cb_dyn_smem_r[];

__global__ void __launch_bounds__(128) radon_trace_kernel(const __grid_constant__ RadonTraceParams P) {
    DevCtx cx{nullptr};
    radon_trace_body(cx, P);
}

__global__ void __launch_bounds__(128) radon_apply_kernel(const __grid_constant__ RadonApplyParams P) {
    DevCtx cx{nullptr};
    radon_apply_body(cx, P);
}

__global__ void __launch_bounds__(128) az_tail_kernel(const __grid_constant__ AzTailParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_dyn_smem_r)};
    az_tail_body(cx, P);
}

extern "C" {

int cb_radon_plan_destroy(cb_radon_plan *plan) {
    if (!plan) return CB_OK;
    cudaFree(plan->d_tx); cudaFree(plan->d_rx); cudaFree(plan->d_data);
    cudaFree(plan->d_offset); cudaFree(plan->d_indices);
    delete plan;
    return CB_OK;
}

int cb_radon_plan_create(int nw, int ntheta, const double *tx_host, const double *rx_host, cb_radon_plan **out) {
    CB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    CB_REQUIRE(nw >= 2 && ntheta >= 1 && tx_host && rx_host, "bad radon plan arguments");
    int rc = cb_check_device();
    if (rc) return rc;
    cb_radon_plan *pl = new (std::nothrow) cb_radon_plan();
    if (!pl) return CB_ERR_NOMEM;
    pl->nw = nw; pl->ntheta = ntheta; pl->nray = (int64_t)nw * ntheta;
    const size_t eb = (size_t)pl->nray * 2 * sizeof(double);
    int64_t *d_count = nullptr;
#define TRY_OR_FREE(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
        cb_set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); cudaFree(d_count); \
        cb_radon_plan_destroy(pl); return CB_ERR_CUDA; } } while (0)
    TRY_OR_FREE(cudaMalloc((void **)&pl->d_tx, eb));
    TRY_OR_FREE(cudaMalloc((void **)&pl->d_rx, eb));
    TRY_OR_FREE(cudaMemcpy(pl->d_tx, tx_host, eb, cudaMemcpyHostToDevice));
    TRY_OR_FREE(cudaMemcpy(pl->d_rx, rx_host, eb, cudaMemcpyHostToDevice));
    TRY_OR_FREE(cudaMalloc((void **)&d_count, (size_t)pl->nray * sizeof(int64_t)));
    RadonTraceParams P;
    memset(&P, 0, sizeof(P));
    P.tx = pl->d_tx; P.rx = pl->d_rx; P.nray = pl->nray; P.nw = nw; P.count = d_count;
    const int grid = (int)((pl->nray + 127) / 128);
    radon_trace_kernel<<<grid, 128>>>(P);
    TRY_OR_FREE(cudaGetLastError());
    std::vector<int64_t> count((size_t)pl->nray);
    TRY_OR_FREE(cudaMemcpy(count.data(), d_count, (size_t)pl->nray * sizeof(int64_t), cudaMemcpyDeviceToHost));
    pl->offset_host.assign((size_t)pl->nray + 1, 0);
    for (int64_t r = 0; r < pl->nray; ++r) pl->offset_host[r + 1] = pl->offset_host[r] + count[r];
    pl->nnz = pl->offset_host[pl->nray];
    TRY_OR_FREE(cudaMalloc((void **)&pl->d_offset, ((size_t)pl->nray + 1) * sizeof(int64_t)));
    TRY_OR_FREE(cudaMemcpy(pl->d_offset, pl->offset_host.data(), ((size_t)pl->nray + 1) * sizeof(int64_t),
                           cudaMemcpyHostToDevice));
    TRY_OR_FREE(cudaMalloc((void **)&pl->d_data, (size_t)(pl->nnz > 0 ? pl->nnz : 1) * sizeof(double)));
    TRY_OR_FREE(cudaMalloc((void **)&pl->d_indices, (size_t)(pl->nnz > 0 ? pl->nnz : 1) * sizeof(int64_t)));
    P.count = nullptr; P.offset = pl->d_offset; P.data = pl->d_data; P.indices = pl->d_indices;
    radon_trace_kernel<<<grid, 128>>>(P);
    TRY_OR_FREE(cudaGetLastError());
    TRY_OR_FREE(cudaDeviceSynchronize());
    cudaFree(d_count);
#undef TRY_OR_FREE
    *out = pl;
    return CB_OK;
}

int64_t cb_radon_plan_nnz(const cb_radon_plan *plan) { return plan ? plan->nnz : 0; }

int cb_radon_plan_get_csr(const cb_radon_plan *plan, double *data_host, int64_t *indices_host, int64_t *indptr_host) {
    CB_REQUIRE(plan && data_host && indices_host && indptr_host, "NULL argument");
    if (plan->nnz > 0) {
        CB_CUDA_TRY(cudaMemcpy(data_host, plan->d_data, (size_t)plan->nnz * sizeof(double), cudaMemcpyDeviceToHost));
        CB_CUDA_TRY(cudaMemcpy(indices_host, plan->d_indices, (size_t)plan->nnz * sizeof(int64_t), cudaMemcpyDeviceToHost));
    }
    // per-angle local offsets, like the indptr the reference returns for each angle
    for (int t = 0; t < plan->ntheta; ++t) {
        const int64_t base = plan->offset_host[(size_t)t * plan->nw];
        for (int r = 0; r <= plan->nw; ++r)
            indptr_host[(size_t)t * (plan->nw + 1) + r] = plan->offset_host[(size_t)t * plan->nw + r] - base;
    }
    return CB_OK;
}

int cb_radon_apply_batch(const cb_radon_plan *plan, const double *grid_dev, int64_t ny, int64_t nx, int64_t ld,
                         const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin, double *sinogram_dev,
                         void *stream) {
    CB_REQUIRE(plan != nullptr && nwin >= 0, "bad arguments");
    if (nwin == 0) return CB_OK;
    CB_REQUIRE(grid_dev && iy0_dev && ix0_dev && sinogram_dev, "NULL buffer");
    CB_REQUIRE(ny >= plan->nw && nx >= plan->nw && ld >= nx, "grid smaller than the window");
    RadonApplyParams P;
    memset(&P, 0, sizeof(P));
    P.grid = grid_dev; P.ld = ld; P.iy0 = iy0_dev; P.ix0 = ix0_dev; P.nwin = nwin;
    P.nw = plan->nw; P.ntheta = plan->ntheta; P.offset = plan->d_offset; P.data = plan->d_data;
    P.indices = plan->d_indices; P.sino = sinogram_dev; P.window_fastest = nwin >= 32;
    const int64_t total = plan->nray * nwin;
    const int64_t grid = (total + 127) / 128;
    CB_REQUIRE(grid <= 2147483647LL, "batch too large for one launch");
    radon_apply_kernel<<<(int)grid, 128, 0, (cudaStream_t)stream>>>(P);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

int cb_azimuthal_tail_batch(const double *sinogram_dev, int nw, int ntheta, int64_t nwin, const double *taper_host,
                            int nkb, double power, double *S_dev, void *stream) {
    CB_REQUIRE(nw >= 2 && ntheta >= 1 && nkb >= 1 && nwin >= 0, "bad arguments");
    if (nwin == 0) return CB_OK;
    CB_REQUIRE(sinogram_dev && S_dev, "NULL buffer");
    int rc = cb_check_device();
    if (rc) return rc;
    double *d_taper = nullptr;
    if (taper_host) {
        CB_CUDA_TRY(cudaMallocAsync((void **)&d_taper, (size_t)nw * sizeof(double), (cudaStream_t)stream));
        CB_CUDA_TRY_OR(cudaMemcpyAsync(d_taper, taper_host, (size_t)nw * sizeof(double), cudaMemcpyHostToDevice,
                                       (cudaStream_t)stream), cudaFreeAsync(d_taper, (cudaStream_t)stream));
    }
    AzTailParams P{sinogram_dev, nw, ntheta, nkb, nwin, d_taper, power, S_dev};
    const int nk = 1 + 2 * nkb;
    const size_t smem = (size_t)nk * sizeof(double2) + (size_t)nw * sizeof(double);
    auto drop = [&]() { if (d_taper) cudaFreeAsync(d_taper, (cudaStream_t)stream); };
    if (smem > 48 * 1024)
        CB_CUDA_TRY_OR(cudaFuncSetAttribute(az_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), drop());
    const int64_t grid = nwin * ntheta;
    if (grid > 2147483647LL) {
        drop();
        cb_set_error("invalid argument: batch too large for one launch");
        return CB_ERR_INVALID;
    }
    az_tail_kernel<<<(int)grid, 128, smem, (cudaStream_t)stream>>>(P);
    CB_CUDA_TRY_OR(cudaGetLastError(), drop());
    if (d_taper) CB_CUDA_TRY(cudaFreeAsync(d_taper, (cudaStream_t)stream));
    return CB_OK;
}

}  // extern "C"
