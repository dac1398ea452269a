// Execution-context policy: every kernel body in this directory is a template
//     template <class Ctx> __host__ __device__ void xxx_body(Ctx &cx, const Params &p)
// that only talks to the machine through `cx`.  DevCtx maps 1:1 onto CUDA
// intrinsics (zero overhead, everything inlines).  HostCtx (tests/emu only, built
// when CB_EMU is defined) runs the very same body on CPU threads with barriers so
// that kernel LOGIC can be checked in a container without a GPU.  HostCtx is test
// infrastructure; the product library never contains or calls it.
#pragma once
#include "common.cuh"

#if defined(__CUDACC__) && !defined(CB_EMU)
struct DevCtx {
    char *smem_base;
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return blockDim.x; }
    __device__ __forceinline__ int bid() const { return blockIdx.x; }
    __device__ __forceinline__ int nblocks() const { return gridDim.x; }
    __device__ __forceinline__ char *smem() const { return smem_base; }
    __device__ __forceinline__ void sync_block() const { __syncthreads(); }
    __device__ __forceinline__ void sync_warp() const { __syncwarp(); }
    __device__ __forceinline__ double shfl_xor(double v, int o) const { return __shfl_xor_sync(0xffffffffu, v, o); }
    __device__ __forceinline__ double shfl(double v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ int shfl_i(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ double shfl_up(double v, int delta) const { return __shfl_up_sync(0xffffffffu, v, delta); }
    __device__ __forceinline__ int warp_max_i(int v) const { return __reduce_max_sync(0xffffffffu, v); }
    __device__ __forceinline__ int team_lanes() const { return CB_WARP; }
    __device__ __forceinline__ double warp_sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ bool any(bool pred) const { return __any_sync(0xffffffffu, pred); }
    __device__ __forceinline__ bool all(bool pred) const { return __all_sync(0xffffffffu, pred); }
    // fetch-and-increment of a global work counter
    __device__ __forceinline__ unsigned atomic_inc_u32(unsigned *p) const { return atomicAdd(p, 1u); }
    template <class T> __device__ __forceinline__ static T ldg(const T *p) { return __ldg(p); }
    template <class T> __device__ __forceinline__ static T ldcg(const T *p) { return __ldcg(p); }
    template <class T> __device__ __forceinline__ static void stcg(T *p, T v) { __stcg(p, v); }
    // Read-only views of arrays that live in SHARED memory but reach a non-inlined function through a plain
    // pointer: the compiler would issue generic-space loads (LD.E plus descriptor moves) for them; the view
    // converts the base once and reads with ld.shared.
    struct SharedD2 {
        unsigned a;
        __device__ __forceinline__ double2 operator[](int i) const {
            double2 v;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a + 16u * (unsigned)i));
            return v;
        }
    };
    struct SharedD {
        unsigned a;
        __device__ __forceinline__ double operator[](int i) const {
            double v;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a + 8u * (unsigned)i));
            return v;
        }
    };
    __device__ __forceinline__ static SharedD2 shared_view(const double2 *p) { return SharedD2{(unsigned)__cvta_generic_to_shared(p)}; }
    __device__ __forceinline__ static SharedD shared_view(const double *p) { return SharedD{(unsigned)__cvta_generic_to_shared(p)}; }
};
#endif

#ifdef CB_EMU
#include <pthread.h>
#include <string.h>
#include <vector>
struct EmuBlock {
    int nthreads, nblocks, bid;
    pthread_barrier_t block_bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<double> shfl_buf;      // one slot per thread
    std::vector<int> vote_buf;
    std::vector<char> smem;
};
struct HostCtx {
    EmuBlock *blk;
    int tid_;
    int tid() const { return tid_; }
    int nthreads() const { return blk->nthreads; }
    int bid() const { return blk->bid; }
    int nblocks() const { return blk->nblocks; }
    char *smem() const { return blk->smem.data(); }
    void sync_block() const { pthread_barrier_wait(&blk->block_bar); }
    void sync_warp() const { pthread_barrier_wait(&blk->warp_bar[tid_ >> 5]); }
    double shfl_xor(double v, int o) const {
        blk->shfl_buf[tid_] = v;
        sync_warp();
        double r = blk->shfl_buf[tid_ ^ o];
        sync_warp();
        return r;
    }
    double shfl(double v, int src) const {
        blk->shfl_buf[tid_] = v;
        sync_warp();
        double r = blk->shfl_buf[(tid_ & ~31) | src];
        sync_warp();
        return r;
    }
    int shfl_i(int v, int src) const { return (int)shfl((double)v, src); }
    double shfl_up(double v, int delta) const {     // lanes below delta keep their own value
        blk->shfl_buf[tid_] = v;
        sync_warp();
        const int lane = tid_ & 31;
        double r = lane >= delta ? blk->shfl_buf[tid_ - delta] : v;
        sync_warp();
        return r;
    }
    int warp_max_i(int v) const {
        for (int o = 16; o > 0; o >>= 1) { int t = (int)shfl_xor((double)v, o); v = t > v ? t : v; }
        return v;
    }
    int team_lanes() const { return CB_WARP; }
    double warp_sum(double v) const {
        for (int o = 16; o > 0; o >>= 1) v += shfl_xor(v, o);
        return v;
    }
    bool any(bool pred) const {
        blk->vote_buf[tid_] = pred;
        sync_warp();
        bool r = false;
        for (int i = 0; i < 32; ++i) r = r || blk->vote_buf[(tid_ & ~31) | i];
        sync_warp();
        return r;
    }
    bool all(bool pred) const { return !any(!pred); }
    unsigned atomic_inc_u32(unsigned *p) const { return __atomic_fetch_add(p, 1u, __ATOMIC_SEQ_CST); }
    template <class T> static T ldg(const T *p) { return *p; }
    template <class T> static T ldcg(const T *p) { return *p; }
    template <class T> static const T *shared_view(const T *p) { return p; }      // plain memory under emulation
    template <class T> static void stcg(T *p, T v) { *p = v; }
};

// One "warp" = one lane: the per-warp kernels (fit, MCMC) run serially, which is the fast way to
// check their arithmetic on the CPU.  nthreads is the number of emulated warps per block.
struct SerialCtx {
    int warp_, nwarps_, bid_, nblocks_;
    char *smem_;
    int tid() const { return warp_ * CB_WARP; }          // lane 0 of warp `warp_`
    int nthreads() const { return nwarps_ * CB_WARP; }
    int bid() const { return bid_; }
    int nblocks() const { return nblocks_; }
    char *smem() const { return smem_; }
    void sync_block() const {}
    void sync_warp() const {}
    double shfl_xor(double v, int) const { return v; }
    double shfl(double v, int) const { return v; }
    int shfl_i(int v, int) const { return v; }
    double shfl_up(double v, int) const { return v; }
    int warp_max_i(int v) const { return v; }
    int team_lanes() const { return 1; }
    double warp_sum(double v) const { return v; }
    bool any(bool pred) const { return pred; }
    bool all(bool pred) const { return pred; }
    unsigned atomic_inc_u32(unsigned *p) const { return __atomic_fetch_add(p, 1u, __ATOMIC_SEQ_CST); }
    template <class T> static T ldg(const T *p) { return *p; }
    template <class T> static T ldcg(const T *p) { return *p; }
    template <class T> static const T *shared_view(const T *p) { return p; }      // plain memory under emulation
    template <class T> static void stcg(T *p, T v) { *p = v; }
};
#endif
