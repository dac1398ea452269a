// Error plumbing and device queries of the C-ABI (include/curie_b200.h).
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

static thread_local char g_err[512] = "";

void cb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cb_check_device() {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        cb_set_error("no CUDA device available (%s); curie_b200 has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return CB_ERR_NO_DEVICE;
    }
    return CB_OK;
}

extern "C" {

int cb_abi_version(void) { return CB_ABI_VERSION; }

const char *cb_last_error(void) { return g_err; }

int cb_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *smem_optin_bytes) {
    int rc = cb_check_device();
    if (rc) return rc;
    int dev = 0;
    CB_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CB_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (smem_optin_bytes) *smem_optin_bytes = (int64_t)prop.sharedMemPerBlockOptin;
    return CB_OK;
}

int cb_fp64_fma_peak(double *tflops_out);

// Page-lock a caller-owned host buffer in place (portable: valid for every device of the process) so that
// uploads from it run asynchronously at full PCIe speed without a staging copy.
int cb_host_register(void *host, int64_t bytes) {
    CB_REQUIRE(host != nullptr && bytes > 0, "cb_host_register: NULL buffer or bytes <= 0");
    int rc = cb_check_device();
    if (rc) return rc;
    cudaError_t e = cudaHostRegister(host, (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return CB_OK; }
    if (e != cudaSuccess) {
        cudaGetLastError();
        cb_set_error("cudaHostRegister(%lld bytes): %s", (long long)bytes, cudaGetErrorString(e));
        return CB_ERR_CUDA;
    }
    return CB_OK;
}

int cb_host_unregister(void *host) {
    CB_REQUIRE(host != nullptr, "cb_host_unregister: NULL buffer");
    cudaError_t e = cudaHostUnregister(host);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (e == cudaErrorHostMemoryNotRegistered || e == cudaErrorCudartUnloading || e == cudaErrorContextIsDestroyed) return CB_OK;
        cb_set_error("cudaHostUnregister: %s", cudaGetErrorString(e));
        return CB_ERR_CUDA;
    }
    return CB_OK;
}

}  // extern "C"

// 16 independent DFMA chains per thread, 4096 iterations: 2 * 16 * 4096 flop per thread.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *sink, double a, double b) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123456.789) sink[0] = s;     // never true; keeps the chain alive
}

extern "C" int cb_fp64_fma_peak(double *tflops_out) {
    CB_REQUIRE(tflops_out != nullptr, "NULL output");
    int rc = cb_check_device();
    if (rc) return rc;
    int dev = 0, sms = 0;
    CB_CUDA_TRY(cudaGetDevice(&dev));
    CB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *sink = nullptr;
    CB_CUDA_TRY(cudaMalloc((void **)&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CB_CUDA_TRY(cudaEventCreate(&e0));
    CB_CUDA_TRY(cudaEventCreate(&e1));
    const int grid = sms * 16, block = 256;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CB_CUDA_TRY(cudaEventRecord(e0));
        fp64_peak_kernel<<<grid, block>>>(sink, 0.999999, 1e-9);
        CB_CUDA_TRY(cudaEventRecord(e1));
        CB_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 16.0 * 4096.0 * (double)grid * (double)block;
        const double tf = flop / ((double)ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops_out = best;
    return CB_OK;
}
