// FFT-domain filtering of a whole 2-D array (CurieGrid.reduce_to_pole / upward_continuation,
// pycurious/grid.py:405-546): 1-D plans, the line-transform kernel and the C-ABI entry points.
// Kernel body: fft2d_body.cuh.
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <new>
#include <vector>
#include "fft_plan.h"
#include "fft2d_body.cuh"

typedef void (*line_fft_fn)(const LineFftParams);

struct cb_fft_plan {
    int n = 0, bluestein = 0, line_slots = 0;
    FftPlan fft;
    double2 *d_tw = nullptr, *d_chirp = nullptr, *d_filt = nullptr;
    int32_t *d_slot = nullptr;
    int teams = 1, sm_count = 0;
    int64_t smem_bytes = 0;
    line_fft_fn kernel = nullptr;      // generic, or the static engine whose pass list matches (Bluestein lengths 256 / 320 / 512 / 640)
    int stage_slots = 0;               // twiddles + filter staged in shared memory in front of the lines (static engines)
};

extern __shared__ double2 cb_fft2d_smem[];

template <class Engine>
__global__ void __launch_bounds__(512, 1) line_fft_kernel(const __grid_constant__ LineFftParams P) {
    DevCtx cx{reinterpret_cast<char *>(cb_fft2d_smem)};
    line_fft_body<Engine>(cx, P);
}
typedef StaticFft<8, 8, 8> LFft512;
typedef StaticFft<8, 8, 4> LFft256;
typedef StaticFft<8, 8, 5> LFft320;
typedef StaticFft<8, 10, 8> LFft640;

__global__ void __launch_bounds__(256) op_half_kernel(const double2 *op, int n, double2 *op_half) {
    DevCtx cx{nullptr};
    op_half_body(cx, op, n, op_half);
}

template <class T>
static int upload_vec(T **dst, const std::vector<T> &src) {
    *dst = nullptr;
    if (src.empty()) return CB_OK;
    CB_CUDA_TRY(cudaMalloc((void **)dst, src.size() * sizeof(T)));
    CB_CUDA_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return CB_OK;
}

extern "C" {

int cb_fft_plan_destroy(cb_fft_plan *p) {
    if (!p) return CB_OK;
    cudaFree(p->d_tw); cudaFree(p->d_chirp); cudaFree(p->d_filt); cudaFree(p->d_slot);
    delete p;
    return CB_OK;
}

int cb_fft_plan_create(int n, cb_fft_plan **out) {
    CB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    CB_REQUIRE(n >= 2, "length must be >= 2");
    int rc = cb_check_device();
    if (rc) return rc;
    cb_fft_plan *p = new (std::nothrow) cb_fft_plan();
    if (!p) return CB_ERR_NOMEM;
    std::vector<double2> tw, chirp, filt;
    std::vector<int32_t> slot;
    p->n = n;
    if (cbplan::build_fft_plan(n, p->fft, tw)) {
        p->bluestein = 0;
        slot = cbplan::digit_reversed_slots(p->fft);
    } else {
        const int M = cbplan::bluestein_length(n);
        if (!M || !cbplan::build_fft_plan(M, p->fft, tw)) {
            cb_set_error("no FFT plan for length %d", n);
            delete p;
            return CB_ERR_UNSUPPORTED;
        }
        p->bluestein = 1;
        cbplan::build_bluestein_tables(n, p->fft, chirp, filt);
    }
    p->line_slots = (fft_line_slots(p->fft.n) + 1) & ~1;
    if (p->bluestein) filt.resize(p->line_slots, cplx(0.0, 0.0));
    p->kernel = line_fft_kernel<GenericFft>;
    if (p->bluestein && !getenv("CB_SPEC_GENERIC")) {
        // lines of N/2 + 1 nodes on the length-N engine (Bluestein length 2n - 2): the last node goes around the pruned passes
        if (PlusOneFft<LFft256>::serves(p->fft, n)) p->kernel = line_fft_kernel<PlusOneFft<LFft256>>;
        else if (PlusOneFft<LFft320>::serves(p->fft, n)) p->kernel = line_fft_kernel<PlusOneFft<LFft320>>;
        else if (PlusOneFft<LFft512>::serves(p->fft, n)) p->kernel = line_fft_kernel<PlusOneFft<LFft512>>;
        else if (PlusOneFft<LFft640>::serves(p->fft, n)) p->kernel = line_fft_kernel<PlusOneFft<LFft640>>;
        else if (LFft512::matches(p->fft)) p->kernel = line_fft_kernel<LFft512>;
        else if (LFft256::matches(p->fft)) p->kernel = line_fft_kernel<LFft256>;
        else if (LFft320::matches(p->fft)) p->kernel = line_fft_kernel<LFft320>;
        else if (LFft640::matches(p->fft)) p->kernel = line_fft_kernel<LFft640>;
        if (p->kernel != line_fft_kernel<GenericFft>) p->stage_slots = (p->fft.tw_count + p->line_slots + 1) & ~1;
    }
    int dev = 0;
    cudaDeviceProp prop;
    CB_CUDA_TRY_OR(cudaGetDevice(&dev), delete p);
    CB_CUDA_TRY_OR(cudaGetDeviceProperties(&prop, dev), delete p);
    p->sm_count = prop.multiProcessorCount;
    const int64_t line_bytes = (int64_t)p->line_slots * 16;
    const int teams = (int)std::min<int64_t>(16, ((int64_t)prop.sharedMemPerBlockOptin - (int64_t)p->stage_slots * 16) / line_bytes);
    if (teams < 1) {
        cb_set_error("length %d needs a %lld-byte shared-memory line (limit %lld)", n, (long long)line_bytes,
                     (long long)prop.sharedMemPerBlockOptin);
        delete p;
        return CB_ERR_UNSUPPORTED;
    }
    p->teams = teams;
    p->smem_bytes = (int64_t)p->stage_slots * 16 + (int64_t)teams * line_bytes;
    if ((rc = upload_vec(&p->d_tw, tw)) || (rc = upload_vec(&p->d_chirp, chirp)) || (rc = upload_vec(&p->d_filt, filt)) ||
        (rc = upload_vec(&p->d_slot, slot))) {
        cb_fft_plan_destroy(p);
        return rc;
    }
    CB_CUDA_TRY_OR(cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)prop.sharedMemPerBlockOptin), cb_fft_plan_destroy(p));
    *out = p;
    return CB_OK;
}

static int launch_lines(const cb_fft_plan *p, LineFftParams &P, cudaStream_t st) {
    P.n = p->n; P.bluestein = p->bluestein; P.fft = p->fft;
    P.tw = p->d_tw; P.chirp = p->d_chirp; P.filt = p->d_filt; P.slot = p->d_slot; P.line_slots = p->line_slots;
    P.stage_slots = p->stage_slots;
    const int64_t ctas = (P.nlines + p->teams - 1) / p->teams;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)p->sm_count * 2));
    p->kernel<<<grid, p->teams * CB_WARP, (size_t)p->smem_bytes, st>>>(P);
    CB_CUDA_TRY(cudaGetLastError());
    return CB_OK;
}

int64_t cb_fft2_filter_workspace_bytes(int64_t ny, int64_t nx) {
    if (ny <= 0 || nx <= 0) return 0;
    return 2 * ny * nx * 16 + 512;
}

int cb_fft2_filter(const cb_fft_plan *plan_x, const cb_fft_plan *plan_y, const double *data_dev, int64_t ny, int64_t nx,
                   int64_t ld, const void *filter_dev, double *out_dev, int64_t out_ld, void *workspace_dev,
                   int64_t workspace_bytes, void *stream) {
    CB_REQUIRE(plan_x && plan_y, "plan is NULL");
    CB_REQUIRE(data_dev && filter_dev && out_dev && workspace_dev, "NULL buffer");
    CB_REQUIRE(ny >= 2 && nx >= 2 && ld >= nx && out_ld >= nx, "bad sizes");
    CB_REQUIRE(plan_x->n == nx && plan_y->n == ny, "plans do not match the array shape");
    CB_REQUIRE(workspace_bytes >= cb_fft2_filter_workspace_bytes(ny, nx), "workspace too small (see cb_fft2_filter_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    double2 *A = reinterpret_cast<double2 *>(((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255);
    double2 *B = A + ny * nx;
    int rc;
    LineFftParams P;
    // 1. rows forward: data (real) -> A
    memset(&P, 0, sizeof(P));
    P.in_real = data_dev; P.in_es = 1; P.in_ls = ld; P.out_cplx = A; P.out_es = 1; P.out_ls = nx; P.nlines = ny; P.scale = 1.0;
    if ((rc = launch_lines(plan_x, P, st))) return rc;
    // 2. columns forward, times the filter: A -> B
    memset(&P, 0, sizeof(P));
    P.in_cplx = A; P.in_es = nx; P.in_ls = 1; P.out_cplx = B; P.out_es = nx; P.out_ls = 1; P.nlines = nx; P.scale = 1.0;
    P.mul = reinterpret_cast<const double2 *>(filter_dev);
    if ((rc = launch_lines(plan_y, P, st))) return rc;
    // 3. columns inverse: B -> A
    memset(&P, 0, sizeof(P));
    P.in_cplx = B; P.in_es = nx; P.in_ls = 1; P.out_cplx = A; P.out_es = nx; P.out_ls = 1; P.nlines = nx;
    P.inverse = 1; P.scale = 1.0 / (double)ny;
    if ((rc = launch_lines(plan_y, P, st))) return rc;
    // 4. rows inverse, real part: A -> out
    memset(&P, 0, sizeof(P));
    P.in_cplx = A; P.in_es = 1; P.in_ls = nx; P.out_real = out_dev; P.out_es = 1; P.out_ls = out_ld; P.nlines = ny;
    P.inverse = 1; P.scale = 1.0 / (double)nx;
    return launch_lines(plan_x, P, st);
}

int64_t cb_fft2_filter_windows_workspace_bytes(int n, int64_t nwin) {
    if (n <= 0 || nwin <= 0) return 0;
    const int64_t half = n / 2 + 1;
    // A (and B, when the column passes are not fused: 2/3/5-smooth edges) [nwin][n][half] + the operator's Hermitian half
    const int64_t stacks = cbplan::smooth_radices(n).empty() ? 1 : 2;
    return (stacks * nwin + 1) * (int64_t)n * half * 16 + 512;
}

int cb_fft2_filter_windows(const cb_fft_plan *plan, const double *grid_dev, int64_t ny, int64_t nx, int64_t ld,
                           const int32_t *iy0_dev, const int32_t *ix0_dev, int64_t nwin, const void *filter_dev,
                           double *out_dev, void *workspace_dev, int64_t workspace_bytes, void *stream) {
    CB_REQUIRE(plan != nullptr, "plan is NULL");
    if (nwin == 0) return CB_OK;
    CB_REQUIRE(nwin > 0, "nwin < 0");
    CB_REQUIRE(grid_dev && iy0_dev && ix0_dev && filter_dev && out_dev && workspace_dev, "NULL buffer");
    const int64_t n = plan->n;
    CB_REQUIRE(ny >= n && nx >= n && ld >= nx, "windows larger than the grid");
    CB_REQUIRE(nwin * n < ((int64_t)1 << 40), "batch too large");
    CB_REQUIRE(workspace_bytes >= cb_fft2_filter_windows_workspace_bytes((int)n, nwin),
               "workspace too small (see cb_fft2_filter_windows_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    // Real windows have Hermitian spectra and real(ifft2(op * .)) acts through the Hermitian part of op: row PAIRS
    // in one complex transform on the way in and out, the column passes over the half plane v <= n/2 only -- 4 x
    // ceil(n/2) line transforms per window instead of 4 n.  Layout: op_half [n][half], A, B [nwin][n][half].
    const int64_t half = n / 2 + 1;
    double2 *op_half = reinterpret_cast<double2 *>(((uintptr_t)workspace_dev + 255) & ~(uintptr_t)255);
    double2 *A = op_half + n * half;
    double2 *B = A + nwin * n * half;
    op_half_kernel<<<(int)std::min<int64_t>((n * half + 255) / 256, 1024), 256, 0, st>>>(
        reinterpret_cast<const double2 *>(filter_dev), (int)n, op_half);
    CB_CUDA_TRY(cudaGetLastError());
    const bool fused = plan->bluestein != 0;      // column forward x operator x column inverse in one visit of the line
    for (int pass = 0; pass < 4; ++pass) {
        if (fused && pass == 2) continue;
        LineFftParams P;
        fft2_windows_pass(pass, (int)n, nwin, grid_dev, ld, iy0_dev, ix0_dev, A, fused ? nullptr : B, op_half, out_dev, P, fused);
        const int rc = launch_lines(plan, P, st);
        if (rc) return rc;
    }
    return CB_OK;
}

}  // extern "C"
