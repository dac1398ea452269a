// Batched Metropolis-Hastings sampler, one chain (= one window) per warp.
// Replaces CurieOptimise.metropolis_hastings (pycurious/optimise.py:585-633) literally:
//   burn-in  : greedy tempered climb, P = exp(-misfit/1000), accept iff P1 > P0, no uniform drawn
//   sampling : P = exp(-misfit) with the 1e-99 floor on P0, accept iff u <= min(P1/P0, 1)
//   bounds are NOT applied (the reference does not apply them either)
// Random numbers: counter-based Philox4x32-10 (Salmon et al. 2011), key = seed ^ chain id,
// counter = (step, block, phase); the stream is replicated on the host by
// oracle/curie_oracle.py:PhiloxStream so that chains are step-comparable.
#pragma once
#include <stddef.h>
#include "model_eval.cuh"

struct MhParams {
    const double *kh;
    const double *ktab;     // [ktable_doubles(nk)] window-independent tables (ktable_entry)
    int32_t nk;
    const double *phi;      // [nwin, nk]
    int64_t nwin;
    cb_mh_options opt;
    const double *x0_dev;   // [nwin,4] or NULL
    double *out;            // mode 0: [nwin,4,nsim]; mode 1: [nwin,10]
    int64_t *accepted;      // [nwin] or NULL
    int32_t warps_per_block;
};

struct U32x4 { uint32_t v[4]; };

CB_HD void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

CB_HD U32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c0, hi0, lo0);
        mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    U32x4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// 52 random bits -> open interval (0,1): (n + 0.5) * 2^-52, exactly representable
CB_HD double u52(uint32_t hi, uint32_t lo) {
    const uint64_t n = ((uint64_t)hi << 20) ^ (uint64_t)(lo >> 12);
    return ((double)n + 0.5) * (1.0 / 4503599627370496.0);
}

CB_HD void cb_sincos_2pi(double u, double &s, double &c) {
#ifdef __CUDA_ARCH__
    sincospi(2.0 * u, &s, &c);
#else
    s = sin(2.0 * CB_PI * u);
    c = cos(2.0 * CB_PI * u);
#endif
}

// Four standard normals (two Box-Muller pairs) for step t of phase ph.
CB_HD void philox_normal4(uint64_t t, uint32_t ph, uint32_t k0, uint32_t k1, double *z) {
    const uint32_t tl = (uint32_t)t, th = (uint32_t)(t >> 32);
    const U32x4 a = philox4x32_10(tl, th, 0u, ph, k0, k1);
    const U32x4 b = philox4x32_10(tl, th, 1u, ph, k0, k1);
    const double u[4] = {u52(a.v[0], a.v[1]), u52(a.v[2], a.v[3]), u52(b.v[0], b.v[1]), u52(b.v[2], b.v[3])};
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const double r = sqrt(-2.0 * log(u[2 * p]));
        double s, c;
        cb_sincos_2pi(u[2 * p + 1], s, c);
        z[2 * p] = r * c;
        z[2 * p + 1] = r * s;
    }
}

CB_HD double philox_uniform(uint64_t t, uint32_t ph, uint32_t k0, uint32_t k1) {
    const U32x4 c = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), 2u, ph, k0, k1);
    return u52(c.v[0], c.v[1]);
}

// slab of one chain: model arrays, CB_MH_BLOCK coefficient slots, the block's proposals / uniforms / penalties
CB_HD int mh_slab_doubles(int nk) { return warp_slab_doubles(nk, false, 7 * CB_MH_BLOCK + 10, CB_MH_BLOCK); }

#define CB_NEG_LOG_P_FLOOR 227.95592420641054      /* -ln(1e-99), the floor of optimise.py:621 in log space */

template <class Ctx>
CB_BODY void mh_body(const Ctx &cx, const MhParams &P) {
    const int tid = cx.tid();
    const int warp = tid >> 5, lane = tid & 31;
    const int nk = P.nk;
    const int64_t w = (int64_t)cx.bid() * P.warps_per_block + warp;
    if (w >= P.nwin) return;

    WarpSlab W = warp_slab_carve(reinterpret_cast<double *>(cx.smem()) + (int64_t)warp * mh_slab_doubles(nk), nk, false,
                                 P.kh, P.ktab, CB_MH_BLOCK);
    warp_slab_load(cx, W, P.phi + w * nk, lane);
    const int L = cx.team_lanes();
    double *props = W.scratch;                  // [CB_MH_BLOCK][4] proposals of the current block
    double *us = props + 4 * CB_MH_BLOCK;       // [CB_MH_BLOCK] acceptance uniforms
    double *pens = us + CB_MH_BLOCK;            // [CB_MH_BLOCK] prior penalties of the proposals
    double *nlu = pens + CB_MH_BLOCK;           // [CB_MH_BLOCK] -ln(u), for the early-rejection bound
    double *mom = nlu + CB_MH_BLOCK;            // [10] running mean / M2 of (beta, zt, dz, C, zt+dz), lane 0's

    const cb_mh_options &o = P.opt;
    // Priors and proposal scales are read from the kernel parameters where they are used: state that lives in
    // registers across the calls of the evaluation core is saved to and restored from local memory around every
    // call, so the chain carries only x, P0 and f (the moments sit in shared memory, see `fold`).
    const Priors &pr = *reinterpret_cast<const Priors *>(o.prior_mu);
    static_assert(offsetof(cb_mh_options, prior_sigma) == offsetof(cb_mh_options, prior_mu) + 4 * sizeof(double),
                  "Priors mirrors cb_mh_options.prior_mu / prior_sigma");
    double x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = P.x0_dev ? P.x0_dev[w * 4 + j] : o.x0[j];
    const uint64_t chain = o.chain_offset + (uint64_t)w;
    const uint32_t k0 = (uint32_t)(o.seed ^ chain), k1 = (uint32_t)((o.seed >> 32) ^ (chain >> 32));

    // Blocks of proposals.  Until a proposal is accepted every step proposes from the SAME state, so the
    // random draws, the proposals and their order-only coefficients (Gamma functions, 1/Gamma(1+-mu), logs:
    // the scalar third of a step) of the next CB_MH_BLOCK steps are prepared together, one lane each.  The
    // steps are then decided in order; an acceptance invalidates the rest of the block, which is simply
    // prepared again from the new state (acceptance is ~5-10 %).  Draws are keyed by the step index, so the
    // chain is exactly the sequential one.
    // Early rejection: min_func = 0.5 sum r^2 + priors only grows while the points are evaluated, so a
    // proposal whose partial misfit already exceeds what acceptance would need is rejected without
    // finishing the model.  The bound carries a 1e-6 relative margin; inside the margin the evaluation is
    // completed and the reference's own test decides.
    auto prepare = [&](int64_t t0, int n, uint32_t phase) {
        cx.sync_warp();                         // nobody still reads the previous block
        for (int l = lane; l < n; l += L) {
            double z[4], x1[4];
            philox_normal4((uint64_t)(t0 + l), phase, k0, k1, z);
#pragma unroll
            for (int j = 0; j < 4; ++j) { x1[j] = x[j] + z[j] * o.x_scale[j]; props[4 * l + j] = x1[j]; }
            pens[l] = prior_penalty(pr, x1);
            us[l] = phase ? philox_uniform((uint64_t)(t0 + l), phase, k0, k1) : 0.0;
            nlu[l] = phase ? -log(us[l]) : 0.0;
        }
        cx.sync_warp();
        model_coefs_lanes(cx, W, props, n, lane);
    };
    // min_func of proposal l of the block, exact when <= bound (see model_eval_warp)
    auto misfit_of = [&](int l, double bound) -> double {
        const double pen = pens[l];
        const double f = model_eval_warp<true>(cx, W, l, lane, W.p0, nullptr, W.obs, bound - pen);
        return isfinite(f) ? f + pen : CB_BIG_MISFIT;
    };

    // ---- burn-in (optimise.py:598-611): accept iff exp(-f1/1000) > exp(-f0/1000) ----
    double f0 = misfit_slab(cx, W, x, pr, lane);
    double P0 = exp(-f0 / 1000.0);
    for (int64_t t = 0; t < o.burnin;) {
        const int n = (int)(o.burnin - t < CB_MH_BLOCK ? o.burnin - t : CB_MH_BLOCK);
        prepare(t, n, 0u);
        int used = n;
        for (int l = 0; l < n; ++l) {
            const double bound = f0 + 1e-6 * fmax(1.0, fabs(f0));
            const double f1 = misfit_of(l, bound);
            if (f1 > bound) continue;
            const double P1 = exp(-f1 / 1000.0);
            if (P1 > P0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = props[4 * l + j];
                P0 = P1;
                f0 = f1;
                used = l + 1;
                break;
            }
        }
        t += used;
    }

    // ---- sampling (optimise.py:613-631): accept iff u <= min(exp(-f1) / max(P0, 1e-99), 1) ----
    double fcur = misfit_slab(cx, W, x, pr, lane);            // min_func of the current state (= -ln P0 above the floor)
    P0 = exp(-fcur);
    int64_t nacc = 0;
    // Moments (mode 1): the state is recorded once per step but changes on ~10 % of them, so the run of
    // `rep` identical samples is folded in at once (Chan's pairwise update) when the state moves.
    if (lane == 0)
        for (int j = 0; j < 10; ++j) mom[j] = 0.0;
    int64_t nrec = 0, rep = 0;
    auto fold = [&]() {
        if (rep == 0) return;
        if (lane == 0) {
            const double v[5] = {x[0], x[1], x[2], x[3], x[1] + x[2]};
            const double tot = (double)(nrec + rep);
            const double wgt = (double)rep / tot, cross = (double)nrec * wgt;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const double dlt = v[j] - mom[j];
                mom[j] = fma(dlt, wgt, mom[j]);
                mom[5 + j] = fma(dlt * dlt, cross, mom[5 + j]);
            }
        }
        nrec += rep;
        rep = 0;
    };
    double *chain_out = P.out + w * 4 * o.nsim;
    for (int64_t t0 = 0; t0 < o.nsim;) {
        const int n = (int)(o.nsim - t0 < CB_MH_BLOCK ? o.nsim - t0 : CB_MH_BLOCK);
        prepare(t0, n, 1u);
        int used = n;
        for (int l = 0; l < n; ++l) {
            const int64_t t = t0 + l;
            const double u = us[l];
            P0 = fmax(P0, 1e-99);
            // exp(-f1)/P0 >= u  <=>  f1 <= -ln(u P0) = -ln u + min(f_current, -ln 1e-99); the bound only gates
            // the early rejection (1e-6 margin), the reference's own test below decides inside it
            const double need = nlu[l] + fmin(fcur, CB_NEG_LOG_P_FLOOR);
            const double bound = need + 1e-6 * fmax(1.0, fabs(need));
            const double f1 = misfit_of(l, bound);
            bool take = false;
            double P1 = 0.0;
            if (!(f1 > bound)) {
                P1 = exp(-f1);
                take = u <= fmin(P1 / P0, 1.0);
            }
            if (take) {
                if (o.reduce_mode != 0) fold();
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = props[4 * l + j];
                P0 = P1;
                fcur = f1;
                ++nacc;
            }
            if (o.reduce_mode == 0) {
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) chain_out[j * o.nsim + t] = x[j];
                }
            } else {
                ++rep;
            }
            if (take) { used = l + 1; break; }
        }
        t0 += used;
    }
    if (o.reduce_mode != 0) fold();
    if (lane == 0) {
        if (o.reduce_mode != 0) {
            double *mo = P.out + w * 10;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                mo[j] = mom[j];
                mo[5 + j] = o.nsim > 0 ? sqrt(mom[5 + j] / (double)o.nsim) : NAN;
            }
        }
        if (P.accepted) P.accepted[w] = nacc;
    }
}
