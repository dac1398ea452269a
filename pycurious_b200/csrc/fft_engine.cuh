// Mixed-radix in-place FFT engine for one warp ("team") working on one
// shared-memory line of N complex doubles, plus the Bluestein chirp-z core.
//
// Replaces numpy.fft.fft2 (pocketfft) behind pycurious/grid.py:270.
//
// Layout of the algorithm
//   forward  = decimation-in-frequency, natural order in, digit-reversed out
//   backward = decimation-in-time with the pass list reversed, digit-reversed in,
//              natural order out
// Both run in place: a radix-R butterfly reads and writes the same R slots, so the
// only synchronisation is one __syncwarp() between passes.  Bluestein's circular
// convolution never needs natural order in the frequency domain (the filter is
// stored digit-reversed), so no reordering pass exists anywhere; the last forward
// pass, the filter multiply and the first backward pass touch the same R
// consecutive slots and are fused in registers (bluestein_mid_pass).
//
// Twiddles: per-pass compact tables T_p[(s-1)*sub + j] = exp(-2 pi i j s / L_p), so
// that consecutive lanes (consecutive j) read consecutive 16-byte words.
//
// Every function is __host__ __device__ and takes (lane, nlanes) so that
// tests can drive the identical code on the CPU, one emulated lane at a time.
#pragma once
#include "common.cuh"

#define CB_MAX_PASS CB_MAX_RADIX_PASSES

// Shared-memory lines are padded by one complex slot per 8 (16 B per 128 B): the unit-stride
// last pass makes lane b touch slots b*R+t, which for R = 8 is a 128-byte lane stride, an
// 8-way bank conflict on LDS.128 without the pad.
#define CB_PAD(i) ((i) + ((i) >> 3))
CB_HD int fft_line_slots(int n) { return CB_PAD(n) + 1; }

struct FftPass {
    int32_t radix;     // R
    int32_t len;       // L: length of the sub-transform this pass splits (DIF) / builds (DIT)
    int32_t sub;       // L / R
    uint32_t magic;    // ceil(2^32 / sub) for b / sub via __umulhi (unused when sub == 1)
    int32_t tw_off;    // offset of this pass' twiddle table (double2 units); -1 when sub == 1
};

struct FftPlan {
    int32_t n;         // transform length N
    int32_t npass;
    int32_t tw_count;  // total twiddles (double2)
    FftPass pass[CB_MAX_PASS];
};

CB_HD uint32_t fastdiv(uint32_t b, uint32_t magic) {
#ifdef __CUDA_ARCH__
    return __umulhi(b, magic);
#else
    return (uint32_t)(((uint64_t)b * (uint64_t)magic) >> 32);
#endif
}

// ---- radix-R forward DFT kernels: y_s = sum_t v_t exp(-2 pi i s t / R) ----
template <int R> struct Dft;

template <> struct Dft<2> {
    CB_HD static void run(double2 *v) {
        double2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <> struct Dft<3> {
    CB_HD static void run(double2 *v) {
        const double S = 0.86602540378443864676;   // sin(2 pi / 3)
        double2 t = cadd(v[1], v[2]);
        double2 m = cplx(fma(-0.5, t.x, v[0].x), fma(-0.5, t.y, v[0].y));
        double2 d = cscale(csub(v[1], v[2]), S);
        v[0] = cadd(v[0], t);
        v[1] = cadd(m, cmul_mi(d));
        v[2] = cadd(m, cmul_pi(d));
    }
};

template <> struct Dft<4> {
    CB_HD static void run(double2 *v) {
        double2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        double2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        v[0] = cadd(a, c);
        v[2] = csub(a, c);
        v[1] = cadd(b, cmul_mi(d));
        v[3] = cadd(b, cmul_pi(d));
    }
};

template <> struct Dft<5> {
    CB_HD static void run(double2 *v) {
        const double C1 = 0.30901699437494742410, C2 = -0.80901699437494742410;
        const double S1 = 0.95105651629515357212, S2 = 0.58778525229247312917;
        double2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
        double2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
        double2 r1 = cplx(fma(C1, a1.x, fma(C2, a2.x, v[0].x)), fma(C1, a1.y, fma(C2, a2.y, v[0].y)));
        double2 r2 = cplx(fma(C2, a1.x, fma(C1, a2.x, v[0].x)), fma(C2, a1.y, fma(C1, a2.y, v[0].y)));
        double2 q1 = cplx(fma(S1, b1.x, S2 * b2.x), fma(S1, b1.y, S2 * b2.y));
        double2 q2 = cplx(fma(S2, b1.x, -S1 * b2.x), fma(S2, b1.y, -S1 * b2.y));
        v[0] = cadd(v[0], cadd(a1, a2));
        v[1] = cadd(r1, cmul_mi(q1));
        v[4] = cadd(r1, cmul_pi(q1));
        v[2] = cadd(r2, cmul_mi(q2));
        v[3] = cadd(r2, cmul_pi(q2));
    }
};

// 10 = 2 x 5 with coprime factors: Good-Thomas index maps n = (5 n1 + 2 n2) mod 10, k = (5 k1 + 6 k2) mod 10 turn
// the 10-point DFT into five 2-point and two 5-point DFTs without any twiddle factor.
template <> struct Dft<10> {
    CB_HD static void run(double2 *v) {
        double2 e[5] = {cadd(v[0], v[5]), cadd(v[2], v[7]), cadd(v[4], v[9]), cadd(v[6], v[1]), cadd(v[8], v[3])};
        double2 o[5] = {csub(v[0], v[5]), csub(v[2], v[7]), csub(v[4], v[9]), csub(v[6], v[1]), csub(v[8], v[3])};
        Dft<5>::run(e);
        Dft<5>::run(o);
        v[0] = e[0]; v[6] = e[1]; v[2] = e[2]; v[8] = e[3]; v[4] = e[4];
        v[5] = o[0]; v[1] = o[1]; v[7] = o[2]; v[3] = o[3]; v[9] = o[4];
    }
};

template <> struct Dft<8> {
    CB_HD static void run(double2 *v) {
        const double H = 0.70710678118654752440;   // sqrt(1/2)
        double2 e[4] = {v[0], v[2], v[4], v[6]};
        double2 o[4] = {v[1], v[3], v[5], v[7]};
        Dft<4>::run(e);
        Dft<4>::run(o);
        // o_k *= exp(-i pi k / 4)
        double2 o1 = cplx(H * (o[1].x + o[1].y), H * (o[1].y - o[1].x));
        double2 o2 = cmul_mi(o[2]);
        double2 o3 = cplx(H * (o[3].y - o[3].x), -H * (o[3].x + o[3].y));
        v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
        v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
        v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
        v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
    }
};

// ---- one DIF pass: butterfly then twiddle ------------------------------------
template <int R>
CB_HD void dif_pass(double2 *buf, const double2 *tw, const FftPass &ps, int n, int lane, int nlanes) {
    const int sub = ps.sub;
    const int nb = n / R;
    for (int b = lane; b < nb; b += nlanes) {
        int blk = (sub == 1) ? b : (int)fastdiv((uint32_t)b, ps.magic);
        int j = b - blk * sub;
        const int i0 = blk * ps.len + j;
        double2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = buf[CB_PAD(i0 + t * sub)];
        Dft<R>::run(v);
        if (sub > 1) {
            const double2 *w = tw + ps.tw_off + j;
#pragma unroll
            for (int s = 1; s < R; ++s) v[s] = cmul(v[s], w[(s - 1) * sub]);
        }
#pragma unroll
        for (int s = 0; s < R; ++s) buf[CB_PAD(i0 + s * sub)] = v[s];
    }
}

// ---- one DIT pass: twiddle then butterfly ------------------------------------
template <int R>
CB_HD void dit_pass(double2 *buf, const double2 *tw, const FftPass &ps, int n, int lane, int nlanes) {
    const int sub = ps.sub;
    const int nb = n / R;
    for (int b = lane; b < nb; b += nlanes) {
        int blk = (sub == 1) ? b : (int)fastdiv((uint32_t)b, ps.magic);
        int j = b - blk * sub;
        const int i0 = blk * ps.len + j;
        double2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = buf[CB_PAD(i0 + t * sub)];
        if (sub > 1) {
            const double2 *w = tw + ps.tw_off + j;
#pragma unroll
            for (int t = 1; t < R; ++t) v[t] = cmul(v[t], w[(t - 1) * sub]);
        }
        Dft<R>::run(v);
#pragma unroll
        for (int s = 0; s < R; ++s) buf[CB_PAD(i0 + s * sub)] = v[s];
    }
}

// ---- fused: last DIF pass (sub == 1) -> * filter -> conj -> first DIT pass ----
// After it, buf holds conj(Y*H) advanced by one backward pass.
template <int R>
CB_HD void bluestein_mid_pass(double2 *buf, const double2 *filt, int n, int lane, int nlanes) {
    const int nb = n / R;
    for (int b = lane; b < nb; b += nlanes) {
        const int i0 = b * R;
        double2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = buf[CB_PAD(i0 + t)];
        Dft<R>::run(v);
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = cconj(cmul(v[t], filt[CB_PAD(i0 + t)]));
        Dft<R>::run(v);
#pragma unroll
        for (int t = 0; t < R; ++t) buf[CB_PAD(i0 + t)] = v[t];
    }
}

#define CB_RADIX_SWITCH(R, CALL)                   \
    switch (R) {                                   \
    case 2: { constexpr int RR = 2; CALL; } break; \
    case 3: { constexpr int RR = 3; CALL; } break; \
    case 4: { constexpr int RR = 4; CALL; } break; \
    case 5: { constexpr int RR = 5; CALL; } break; \
    case 10: { constexpr int RR = 10; CALL; } break; \
    default: { constexpr int RR = 8; CALL; } break; \
    }

// ---- team (= one warp) drivers; `cx` supplies the warp barrier ---------------------
// natural order in -> digit-reversed out
template <class Ctx>
CB_BODY void fft_forward_dif(const Ctx &cx, double2 *buf, const double2 *tw, const FftPlan &pl, int lane) {
    for (int p = 0; p < pl.npass; ++p) {
        cx.sync_warp();
        CB_RADIX_SWITCH(pl.pass[p].radix, dif_pass<RR>(buf, tw, pl.pass[p], pl.n, lane, CB_WARP));
    }
    cx.sync_warp();
}

// buf = a (zero padded to pl.n).  On return buf[k] = conj((a (*) b)[k]) in natural order, where
// `filt` = DIF(b)/N stored digit-reversed (and padded like buf).
template <class Ctx>
CB_BODY void bluestein_convolve(const Ctx &cx, double2 *buf, const double2 *tw, const double2 *filt,
                                const FftPlan &pl, int lane) {
    const int last = pl.npass - 1;
    for (int p = 0; p < last; ++p) {
        cx.sync_warp();
        CB_RADIX_SWITCH(pl.pass[p].radix, dif_pass<RR>(buf, tw, pl.pass[p], pl.n, lane, CB_WARP));
    }
    cx.sync_warp();
    CB_RADIX_SWITCH(pl.pass[last].radix, bluestein_mid_pass<RR>(buf, filt, pl.n, lane, CB_WARP));
    for (int p = last - 1; p >= 0; --p) {
        cx.sync_warp();
        CB_RADIX_SWITCH(pl.pass[p].radix, dit_pass<RR>(buf, tw, pl.pass[p], pl.n, lane, CB_WARP));
    }
    cx.sync_warp();
}

// =====================================================================================
// Static engine: the same passes with every radix / stride a compile-time constant.
// The generic engine above spends as many issue slots on index arithmetic (fastdiv, CB_PAD,
// runtime strides: IMAD + LEA = 24 % of the spectrum kernel's instructions, ncu r01e) as on FP64
// work; with the pass list in the template all of it folds into immediates and the butterfly
// loops unroll.  Selected on the host for the pass lists the BASELINE window sizes produce
// (201 -> 512 = 8.8.8, 101 -> 256 = 8.8.4, 129 -> 320 = 8.8.5, 301 -> 640 = 8.10.8); anything
// else runs the generic engine.  Table layouts (twiddles, filter) are identical.
// =====================================================================================
// w[s] = w1^s, s = 1..R-1, by at most three dependent complex products (<= 4 ulp): the static
// engine loads ONE twiddle per butterfly instead of R-1 -- the kernel is shared-memory-bandwidth
// bound (78 % of the 128 B/clk pipe, ncu r01e), the FP64 pipe has the headroom.
template <int R> CB_HD void tw_powers(double2 w1, double2 *w) {
    w[1] = w1;
#pragma unroll
    for (int s = 2; s < R; ++s) w[s] = (s % 2 == 0) ? cmul(w[s / 2], w[s / 2]) : cmul(w[s - 1], w1);
}

// Radix-8 butterfly whose inputs v[4..7] are zero (first Bluestein pass: the line is zero padded
// beyond n <= N/2):  y_s = E_s + w^s O_s, y_{s+4} = E_s - w^s O_s with E_s = v0 + (-i)^s v2,
// O_s = v1 + (-i)^s v3.
CB_HD void dft8_half_in(double2 *v) {
    const double H = 0.70710678118654752440;
    const double2 e0 = cadd(v[0], v[2]), e2 = csub(v[0], v[2]);
    const double2 e1 = cadd(v[0], cmul_mi(v[2])), e3 = cadd(v[0], cmul_pi(v[2]));
    const double2 o0 = cadd(v[1], v[3]), o2 = csub(v[1], v[3]);
    const double2 o1 = cadd(v[1], cmul_mi(v[3])), o3 = cadd(v[1], cmul_pi(v[3]));
    const double2 t0 = o0;
    const double2 t1 = cplx(H * (o1.x + o1.y), H * (o1.y - o1.x));     // o1 * exp(-i pi/4)
    const double2 t2 = cmul_mi(o2);
    const double2 t3 = cplx(H * (o3.y - o3.x), -H * (o3.x + o3.y));    // o3 * exp(-3i pi/4)
    v[0] = cadd(e0, t0); v[4] = csub(e0, t0);
    v[1] = cadd(e1, t1); v[5] = csub(e1, t1);
    v[2] = cadd(e2, t2); v[6] = csub(e2, t2);
    v[3] = cadd(e3, t3); v[7] = csub(e3, t3);
}

// Radix-8 butterfly of which only the outputs y_0..y_3 are wanted (last Bluestein pass: only the
// first n <= N/2 points of the convolution are read).
CB_HD void dft8_half_out(double2 *v) {
    const double H = 0.70710678118654752440;
    double2 e[4] = {v[0], v[2], v[4], v[6]};
    double2 o[4] = {v[1], v[3], v[5], v[7]};
    Dft<4>::run(e);
    Dft<4>::run(o);
    const double2 o1 = cplx(H * (o[1].x + o[1].y), H * (o[1].y - o[1].x));
    const double2 o2 = cmul_mi(o[2]);
    const double2 o3 = cplx(H * (o[3].y - o[3].x), -H * (o[3].x + o[3].y));
    v[0] = cadd(e[0], o[0]);
    v[1] = cadd(e[1], o1);
    v[2] = cadd(e[2], o2);
    v[3] = cadd(e[3], o3);
}

template <int R0, int R1, int R2, int R3 = 1>
struct StaticFft {
    static constexpr int N = R0 * R1 * R2 * R3;
    static constexpr int NPASS = (R3 > 1) ? 4 : 3;
    template <int P> struct Pass {
        static constexpr int R = P == 0 ? R0 : (P == 1 ? R1 : (P == 2 ? R2 : R3));
        static constexpr int LEN = P == 0 ? N : (P == 1 ? N / R0 : (P == 2 ? N / (R0 * R1) : N / (R0 * R1 * R2)));
        static constexpr int SUB = LEN / R;
        // offset of this pass' twiddle table (layout of cbplan::build_fft_plan: passes with sub > 1 only)
        static constexpr int TW = (P > 0 ? (R0 - 1) * (N / R0) : 0) + (P > 1 ? (R1 - 1) * (N / (R0 * R1)) : 0) +
                                  (P > 2 ? (R2 - 1) * (N / (R0 * R1 * R2)) : 0);
    };
    static bool matches(const FftPlan &pl) {
        if (pl.n != N || pl.npass != NPASS) return false;
        const int r[4] = {R0, R1, R2, R3};
        int off = 0;
        for (int p = 0; p < NPASS; ++p) {
            if (pl.pass[p].radix != r[p]) return false;
            if (pl.pass[p].sub > 1) {
                if (pl.pass[p].tw_off != off) return false;
                off += (r[p] - 1) * pl.pass[p].sub;
            }
        }
        return true;
    }

    // slot of element i0 + t*SUB given i0 (SUB % 8 == 0 keeps the pad linear in t)
    template <int SUB> CB_HD static int slot(int i0, int pad0, int t) {
        if (SUB % 8 == 0) return pad0 + t * (SUB + SUB / 8);
        return CB_PAD(i0 + t * SUB);
    }

    template <int P, bool DIT> CB_HD static void pass(double2 *buf, const double2 *tw, int lane) {
        constexpr int R = Pass<P>::R, LEN = Pass<P>::LEN, SUB = Pass<P>::SUB, NB = N / R;
        const double2 *w0 = tw + Pass<P>::TW;
#pragma unroll
        for (int b0 = 0; b0 < NB; b0 += CB_WARP) {
            const int b = b0 + lane;
            if (NB % CB_WARP != 0 && b >= NB) break;
            const int blk = b / SUB, j = b - blk * SUB;
            const int i0 = blk * LEN + j, pad0 = CB_PAD(i0);
            double2 v[R];
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = buf[slot<SUB>(i0, pad0, t)];
            double2 w[R];
            if (SUB > 1) tw_powers<R>(w0[j], w);
            if (DIT && SUB > 1) {
#pragma unroll
                for (int t = 1; t < R; ++t) v[t] = cmul(v[t], w[t]);
            }
            Dft<R>::run(v);
            if (!DIT && SUB > 1) {
#pragma unroll
                for (int s = 1; s < R; ++s) v[s] = cmul(v[s], w[s]);
            }
#pragma unroll
            for (int s = 0; s < R; ++s) buf[slot<SUB>(i0, pad0, s)] = v[s];
        }
    }

    // fused: last DIF pass (sub == 1) -> * filter -> conj -> first DIT pass
    CB_HD static void mid(double2 *buf, const double2 *filt, int lane) {
        constexpr int R = Pass<NPASS - 1>::R, NB = N / R;
#pragma unroll
        for (int b0 = 0; b0 < NB; b0 += CB_WARP) {
            const int b = b0 + lane;
            if (NB % CB_WARP != 0 && b >= NB) break;
            const int i0 = b * R;
            double2 v[R];
            int sl[R];
#pragma unroll
            for (int t = 0; t < R; ++t) sl[t] = (R == 8) ? 9 * b + t : CB_PAD(i0 + t);
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = buf[sl[t]];
            Dft<R>::run(v);
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = cconj(cmul(v[t], filt[sl[t]]));
            Dft<R>::run(v);
#pragma unroll
            for (int t = 0; t < R; ++t) buf[sl[t]] = v[t];
        }
    }

    // Pruned first / last passes with the line's input and output fused in (no fill pass, no zero
    // padding in memory, no read-back pass):
    //   first: element x of the line comes from ld(x), x < HALF*SUB only (the rest is zero padding)
    //   last : output k of the convolution goes to sk(k, value), k < HALF*SUB only
    static constexpr int HALF = (R0 + 1) / 2;            // inputs / outputs per butterfly that can be live
    static constexpr int FUSED_SPAN = HALF * (N / R0);   // the engine serves windows with n <= FUSED_SPAN

    static constexpr int NIT = (N / R0 + CB_WARP - 1) / CB_WARP;    // first-pass butterflies per lane
    // a second line's inputs are held in registers beside the transform (12 complex values spill a little and still pay)
#ifndef CB_PREFETCH_MAX
#define CB_PREFETCH_MAX 12
#endif
    static constexpr bool PREFETCH = FUSED_SPAN > 0 && NIT * HALF <= CB_PREFETCH_MAX;
    // line index of input (it, t) of lane `lane` in the first pass
    CB_HD static int input_index(int it, int t, int lane) { return it * CB_WARP + lane + t * (N / R0); }

    // The live inputs of this lane's first-pass butterflies, fetched with `raw(x)` (x < FUSED_SPAN).  Split
    // from the pass itself so that a caller can request the NEXT line's inputs before transforming the
    // current one (software prefetch in registers: the loads are in flight during a whole line's passes).
    template <class Raw> CB_HD static void gather(double2 (&in)[NIT][HALF], int lane, Raw raw) {
        constexpr int NB = N / R0;
#pragma unroll
        for (int it = 0; it < NIT; ++it)
#pragma unroll
            for (int t = 0; t < HALF; ++t)
                in[it][t] = (NB % CB_WARP == 0 || it * CB_WARP + lane < NB) ? raw(input_index(it, t, lane)) : cplx(0.0, 0.0);
    }

    CB_HD static void first_pass_pre(double2 *buf, const double2 *tw, int lane, const double2 (&in)[NIT][HALF]) {
        constexpr int R = R0, SUB = N / R0, NB = N / R;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int j = it * CB_WARP + lane;
            if (NB % CB_WARP != 0 && j >= NB) break;
            double2 v[R];
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = t < HALF ? in[it][t] : cplx(0.0, 0.0);
            if (R == 8) dft8_half_in(v);
            else Dft<R>::run(v);
            double2 w[R];
            tw_powers<R>(tw[j], w);
#pragma unroll
            for (int s = 1; s < R; ++s) v[s] = cmul(v[s], w[s]);
            const int pad0 = CB_PAD(j);
#pragma unroll
            for (int s = 0; s < R; ++s) buf[slot<SUB>(j, pad0, s)] = v[s];
        }
    }

    template <class Loader> CB_HD static void first_pass(double2 *buf, const double2 *tw, int lane, Loader ld) {
        constexpr int R = R0, SUB = N / R0, NB = N / R;
        // every input of every butterfly of this lane is requested before the first one is used: the
        // loads come from L2 / HBM (grid rows, shared row spectra) and their latency is paid once per line
        double2 in[NIT][HALF];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int j = it * CB_WARP + lane;
#pragma unroll
            for (int t = 0; t < HALF; ++t)
                in[it][t] = (NB % CB_WARP == 0 || j < NB) ? ld(j + t * SUB) : cplx(0.0, 0.0);
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int j = it * CB_WARP + lane;                     // LEN == N: blk = 0
            if (NB % CB_WARP != 0 && j >= NB) break;
            double2 v[R];
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = t < HALF ? in[it][t] : cplx(0.0, 0.0);
            if (R == 8) dft8_half_in(v);
            else Dft<R>::run(v);
            double2 w[R];
            tw_powers<R>(tw[j], w);
#pragma unroll
            for (int s = 1; s < R; ++s) v[s] = cmul(v[s], w[s]);
            const int pad0 = CB_PAD(j);
#pragma unroll
            for (int s = 0; s < R; ++s) buf[slot<SUB>(j, pad0, s)] = v[s];
        }
    }

    template <class Sink> CB_HD static void last_pass(double2 *buf, const double2 *tw, int lane, Sink sk) {
        constexpr int R = R0, SUB = N / R0, NB = N / R;
#pragma unroll
        for (int b0 = 0; b0 < NB; b0 += CB_WARP) {
            const int j = b0 + lane;
            if (NB % CB_WARP != 0 && j >= NB) break;
            const int pad0 = CB_PAD(j);
            double2 v[R], w[R];
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = buf[slot<SUB>(j, pad0, t)];
            tw_powers<R>(tw[j], w);
#pragma unroll
            for (int t = 1; t < R; ++t) v[t] = cmul(v[t], w[t]);
            if (R == 8) dft8_half_out(v);
            else Dft<R>::run(v);
#pragma unroll
            for (int s = 0; s < HALF; ++s) sk(j + s * SUB, v[s]);
        }
    }

    // conj((a (*) b)[k]) -> sk(k, .) for k < FUSED_SPAN, with a[x] = ld(x) (zero beyond FUSED_SPAN).
    // The caller guarantees that no lane still reads buf (one warp barrier is issued first).  (chirp, n: the
    // Bluestein chirp of the line length n, used by PlusOneFft only.)
    template <class Ctx, class Loader, class Sink>
    CB_BODY static void convolve_io(const Ctx &cx, double2 *buf, const double2 *tw, const double2 *filt, int lane,
                                    const double2 *, int, Loader ld, Sink sk) {
        cx.sync_warp();
        first_pass(buf, tw, lane, ld);
        cx.sync_warp();
        if (NPASS > 2) { pass<1, false>(buf, tw, lane); cx.sync_warp(); }
        if (NPASS > 3) { pass<2, false>(buf, tw, lane); cx.sync_warp(); }
        mid(buf, filt, lane);
        cx.sync_warp();
        if (NPASS > 3) { pass<2, true>(buf, tw, lane); cx.sync_warp(); }
        if (NPASS > 2) { pass<1, true>(buf, tw, lane); cx.sync_warp(); }
        last_pass(buf, tw, lane, sk);
    }

    // convolve_io with the first-pass inputs already in registers (see gather)
    template <class Ctx, class Sink>
    CB_BODY static void convolve_pre(const Ctx &cx, double2 *buf, const double2 *tw, const double2 *filt, int lane,
                                     const double2 (&in)[NIT][HALF], Sink sk) {
        cx.sync_warp();
        first_pass_pre(buf, tw, lane, in);
        cx.sync_warp();
        if (NPASS > 2) { pass<1, false>(buf, tw, lane); cx.sync_warp(); }
        if (NPASS > 3) { pass<2, false>(buf, tw, lane); cx.sync_warp(); }
        mid(buf, filt, lane);
        cx.sync_warp();
        if (NPASS > 3) { pass<2, true>(buf, tw, lane); cx.sync_warp(); }
        if (NPASS > 2) { pass<1, true>(buf, tw, lane); cx.sync_warp(); }
        last_pass(buf, tw, lane, sk);
    }

    template <class Ctx>
    CB_BODY static void convolve(const Ctx &cx, double2 *buf, const double2 *tw, const double2 *filt, const FftPlan &,
                                 int lane) {
        cx.sync_warp();
        pass<0, false>(buf, tw, lane);
        cx.sync_warp();
        if (NPASS > 2) { pass<1, false>(buf, tw, lane); cx.sync_warp(); }
        if (NPASS > 3) { pass<2, false>(buf, tw, lane); cx.sync_warp(); }
        mid(buf, filt, lane);
        cx.sync_warp();
        if (NPASS > 3) { pass<2, true>(buf, tw, lane); cx.sync_warp(); }
        if (NPASS > 2) { pass<1, true>(buf, tw, lane); cx.sync_warp(); }
        pass<0, true>(buf, tw, lane);
        cx.sync_warp();
    }
    template <class Ctx>
    CB_BODY static void forward(const Ctx &cx, double2 *buf, const double2 *tw, const FftPlan &pl, int lane) {
        fft_forward_dif(cx, buf, tw, pl, lane);      // direct (smooth) sizes are not the hot case
    }
};

// A static engine serving ONE more node than its pruned passes hold: lines of n = FUSED_SPAN + 1 = N/2 + 1 nodes
// (129 on 256, 161 on 320, 257 on 512, 321 on 640) with the Bluestein length N = 2n - 2.
//   * The chirp filter is even, b[m] = b[-m] = exp(i pi m^2 / n), so a circular length of 2n - 2 is enough: the only
//     taps that alias are m = +-(n-1), and they are equal (build_bluestein_tables writes that slot twice with the
//     same value; the unpruned engines need nothing more).
//   * The pruned passes take inputs x < n-1 and give outputs k < n-1.  The convolution is linear, so the last input
//     goes around the transform as a rank-1 term, y[k] += a[n-1] b[k-(n-1)], and the last output is one dot product,
//     y[n-1] = sum_j a[j] b[n-1-j], accumulated while the inputs are loaded and summed over the warp.
// Against the next length the planner would have to take (320 for 129, 640 for 257) every pass fills its lanes
// exactly once or twice instead of 1.25 / 2.5 times.
template <class Base>
struct PlusOneFft : Base {
    static constexpr int FUSED_SPAN = Base::FUSED_SPAN + 1;
    static constexpr bool PREFETCH = false;
    static bool serves(const FftPlan &pl, int n) { return Base::matches(pl) && n == FUSED_SPAN && pl.n == 2 * n - 2; }
    template <class Ctx, class Loader, class Sink>
    CB_BODY static void convolve_io(const Ctx &cx, double2 *buf, const double2 *tw, const double2 *filt, int lane,
                                    const double2 *chirp, int n, Loader ld, Sink sk) {
        const double2 a_last = ld(n - 1);                       // the same address on every lane: one broadcast
        const double2 ca = cconj(a_last);
        double2 acc = cplx(0.0, 0.0);
        auto ld1 = [&](int x) -> double2 {                      // x < n - 1 by construction of the pruned first pass
            const double2 z = ld(x);
            const double2 c = chirp[n - 1 - x];                 // b[n-1-x] = conj(chirp[n-1-x])
            acc.x += fma(z.x, c.x, z.y * c.y);
            acc.y += fma(z.y, c.x, -z.x * c.y);
            return z;
        };
        // the sink receives conj(y[k]); conj(a_last b[k-(n-1)]) = conj(a_last) chirp[n-1-k]
        auto sk1 = [&](int k, double2 zc) { sk(k, cadd(zc, cmul(ca, chirp[n - 1 - k]))); };
        Base::convolve_io(cx, buf, tw, filt, lane, chirp, n, ld1, sk1);
        acc.x = cx.warp_sum(acc.x);
        acc.y = cx.warp_sum(acc.y);
        cx.sync_warp();
        if (lane == 0) sk(n - 1, cplx(acc.x + a_last.x, -(acc.y + a_last.y)));      // b[0] = 1
    }
};

struct GenericFft {
    static constexpr int FUSED_SPAN = 0;                  // no fused-I/O path
    static constexpr bool PREFETCH = false;
    static bool matches(const FftPlan &) { return true; }
    template <class Ctx>
    CB_BODY static void convolve(const Ctx &cx, double2 *buf, const double2 *tw, const double2 *filt, const FftPlan &pl,
                                 int lane) {
        bluestein_convolve(cx, buf, tw, filt, pl, lane);
    }
    template <class Ctx>
    CB_BODY static void forward(const Ctx &cx, double2 *buf, const double2 *tw, const FftPlan &pl, int lane) {
        fft_forward_dif(cx, buf, tw, pl, lane);
    }
};
