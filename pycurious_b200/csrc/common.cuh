// Shared helpers for the curie_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "curie_b200.h"

#define CB_HD __host__ __device__ __forceinline__
#define CB_WARP 32
// kernel bodies: device functions in the product build, plain host functions in tests/emu
#ifdef CB_EMU
#define CB_BODY inline
#else
#define CB_BODY __device__ __forceinline__
#endif

// ---- error plumbing (capi.cu owns the thread-local message) ---------------
void cb_set_error(const char *fmt, ...);
int cb_check_device();   // CB_OK or CB_ERR_NO_DEVICE

#define CB_CUDA_TRY(expr)                                                          \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            cb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                         __FILE__, __LINE__);                                      \
            return CB_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

// CB_CUDA_TRY that runs `cleanup` before returning (owned allocations on an error path)
#define CB_CUDA_TRY_OR(expr, cleanup)                                              \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            cb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                         __FILE__, __LINE__);                                      \
            cleanup;                                                               \
            return CB_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

#define CB_REQUIRE(cond, msg)                                                      \
    do {                                                                           \
        if (!(cond)) {                                                             \
            cb_set_error("invalid argument: %s (%s)", msg, #cond);                 \
            return CB_ERR_INVALID;                                                 \
        }                                                                          \
    } while (0)

// ---- complex arithmetic on double2 ----------------------------------------
CB_HD double2 cplx(double re, double im) { double2 r; r.x = re; r.y = im; return r; }
CB_HD double2 cadd(double2 a, double2 b) { return cplx(a.x + b.x, a.y + b.y); }
CB_HD double2 csub(double2 a, double2 b) { return cplx(a.x - b.x, a.y - b.y); }
CB_HD double2 cmul(double2 a, double2 b) {
    return cplx(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
CB_HD double2 cconj(double2 a) { return cplx(a.x, -a.y); }
CB_HD double2 cscale(double2 a, double s) { return cplx(a.x * s, a.y * s); }
// multiply by -i / +i
CB_HD double2 cmul_mi(double2 a) { return cplx(a.y, -a.x); }
CB_HD double2 cmul_pi(double2 a) { return cplx(-a.y, a.x); }

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif
