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

}  // extern "C"
