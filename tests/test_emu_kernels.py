"""Kernel LOGIC on the CPU: the kernel bodies of pycurious_b200/csrc are compiled with the HostCtx /
SerialCtx execution contexts (tests/emu/emu.cu, one pthread per CUDA thread) and compared with the
oracle and the reference's golden vectors.  This does not replace the GPU parity tests (-m gpu); it
catches indexing / algorithm errors in a container without a GPU.  No product code path is involved:
libcurie_emu.so is test infrastructure and is never loaded by pycurious_b200."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

EMU_DIR = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    lib = os.environ.get("CB_EMU_LIB")          # tests/test_emu_asan.py points this at the ASan build
    if lib:
        return C.CDLL(lib)
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    return C.CDLL(os.path.join(EMU_DIR, "libcurie_emu.so"))


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class FitOpt(C.Structure):
    _fields_ = [("x0", C.c_double * 4), ("lb", C.c_double * 4), ("ub", C.c_double * 4), ("prior_mu", C.c_double * 4),
                ("prior_sigma", C.c_double * 4), ("max_iter", C.c_int32), ("flags", C.c_int32), ("gtol", C.c_double), ("ftol", C.c_double)]


class MhOpt(C.Structure):
    _fields_ = [("x0", C.c_double * 4), ("x_scale", C.c_double * 4), ("prior_mu", C.c_double * 4),
                ("prior_sigma", C.c_double * 4), ("nsim", C.c_int64), ("burnin", C.c_int64), ("seed", C.c_uint64),
                ("chain_offset", C.c_uint64), ("reduce_mode", C.c_int32), ("reserved", C.c_int32)]


# ------------------------------------------------------------------------------------ FFT engine
@pytest.mark.parametrize("n", [4, 5, 6, 8, 9, 10, 15, 16, 25, 27, 30, 64, 100, 101, 125, 128, 129, 151, 161, 201, 240,
                               256, 257, 301, 305, 320, 321, 401, 625, 640, 1000])
def test_fft_engine_matches_numpy(emu, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    a = np.ascontiguousarray(np.c_[x.real, x.imag]).ravel()
    out = np.empty(2 * n)
    M = emu.emu_dft(n, P(a), P(out))
    assert M >= 0
    y = out[0::2] + 1j * out[1::2]
    ref = np.fft.fft(x)
    assert np.abs(y - ref).max() / np.abs(ref).max() < 5e-15
    smooth = n
    for p in (2, 3, 5):
        while smooth % p == 0:
            smooth //= p
    assert (M == 0) == (smooth == 1)          # direct mixed radix iff 2/3/5-smooth, else Bluestein
    if M:
        assert M >= 2 * n - 2          # the chirp is even: 2n - 2 is enough (fft_plan.h:bluestein_length)
        assert (M == 2 * n - 2) == (n in (129, 161, 257, 321))


def test_plan_choices_for_baseline_windows(emu, oracle):
    """none of BASELINE's windows is smooth -> Bluestein on 256 / 320 / 512 / 640 (= 8.10.8, three
    passes with the twiddle-free radix-10 butterfly, instead of 625 = 5^4)."""
    want = {101: 256, 201: 512, 301: 640,
            175: 512, 151: 320, 281: 640,      # off-list sizes go to a compile-time engine's length too (not 360 / 576)
            129: 256, 161: 320, 257: 512, 131: 320}      # N/2 + 1 nodes ride on the length-N engine (PlusOneFft)
    for nw, M in want.items():
        _, dk, edges = oracle.taper_and_bins(nw, nw, 1000.0, None)
        _, masks = oracle.radial_bin_masks(nw, dk, edges)
        idx = [np.flatnonzero(m).astype(np.int32) for m in masks]
        ptr = np.zeros(len(masks) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([i.size for i in idx])
        out = np.zeros(18, dtype=np.int32)
        assert emu.emu_spectrum_plan_info(nw, len(masks), P(ptr), P(np.concatenate(idx)), P(out)) == 0
        assert out[0] == 1 and out[1] == M
        assert out[15] == ptr[-1] and out[16] * 2 == out[15]      # odd n: every member folds onto a twin


# ------------------------------------------------------------------------------------ spectrum stage
def _bins(oracle, nr, dx=1000.0):
    _, dk, edges = oracle.taper_and_bins(nr, nr, dx, None)
    kk, masks = oracle.radial_bin_masks(nr, dk, edges)
    idx = [np.flatnonzero(m).astype(np.int32) for m in masks]
    ptr = np.zeros(len(masks) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([i.size for i in idx])
    return ptr, np.concatenate(idx)


def _emu_spectrum(emu, oracle, grid, origins, nw, taper, detrend, power, nblocks=1, max_teams=4):
    """max_teams < 0 forces the generic FFT engine with |max_teams| teams."""
    ptr, idx = _bins(oracle, nw)
    nb = ptr.size - 1
    tr = None if taper is None else np.ascontiguousarray(taper(nw), dtype=np.float64)
    iy0 = np.array([o[0] for o in origins], dtype=np.int32)
    ix0 = np.array([o[1] for o in origins], dtype=np.int32)
    phi = np.zeros((len(origins), nb))
    sig = np.zeros_like(phi)
    emu.emu_spectrum.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                                                          C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                                                          C.c_int, C.c_int]
    assert emu.emu_spectrum(nw, nb, P(tr), P(tr), P(ptr), P(idx), P(grid), grid.shape[1], P(iy0), P(ix0), len(origins),
                            detrend, power, P(phi), P(sig), nblocks, max_teams) == 0
    return phi, sig


@pytest.mark.parametrize("nw,taper,det,power,origins", [
    (21, np.hanning, 0, 2.0, [(3, 5), (100, 200), (284, 284)]),
    (30, np.hamming, 1, 0.5, [(0, 0), (17, 33)]),            # smooth even size: direct path + the i0 quirk
    (101, np.hanning, 0, 2.0, [(102, 102), (7, 150)]),
    (129, None, 1, 2.0, [(50, 60)]),
    (128, np.hanning, 0, 2.0, [(10, 20)]),
    (201, np.hanning, 0, 2.0, [(40, 60)]),
])
def test_spectrum_body_matches_oracle(emu, oracle, mag, nw, taper, det, power, origins):
    d = np.ascontiguousarray(mag[0])
    phi, sig = _emu_spectrum(emu, oracle, d, origins, nw, taper, det, power, nblocks=min(2, len(origins)))
    for w, (j0, i0) in enumerate(origins):
        win = d[j0:j0 + nw, i0:i0 + nw]
        if det:
            win = oracle.detrend_plane(win)
        _, p0, s0 = oracle.radial_spectrum(win, 1000.0, taper=taper, power=power)
        np.testing.assert_allclose(phi[w], p0, rtol=1e-11)
        np.testing.assert_allclose(sig[w], s0, rtol=1e-9)


@pytest.mark.parametrize("nw,taper,det,power,teams", [
    (101, np.hanning, 0, 2.0, 4), (101, np.hamming, 1, 0.5, -4), (129, None, 1, 2.0, 4), (30, np.hanning, 1, 2.0, 4),
    (201, np.hanning, 0, 2.0, 4)])
def test_shared_row_spectra_path_matches_oracle(emu, oracle, mag, nw, taper, det, power, teams):
    """rowspec_body + colspec_body: windows on a lattice share the x-transforms of the grid rows (odd and
    even row offsets, two column groups, last row pair half empty), with and without the plane detrend."""
    d = np.ascontiguousarray(mag[0])
    ny = d.shape[0]
    origins = [(3, 5), (10, 5), (ny - nw, 5), (4, 40)] if nw < 201 else [(3, 5), (ny - nw, 5)]
    ptr, idx = _bins(oracle, nw)
    nb = ptr.size - 1
    tr = None if taper is None else np.ascontiguousarray(taper(nw), dtype=np.float64)
    iy0 = np.array([o[0] for o in origins], dtype=np.int32)
    ix0 = np.array([o[1] for o in origins], dtype=np.int32)
    phi = np.zeros((len(origins), nb))
    sig = np.zeros_like(phi)
    emu.emu_spectrum_shared.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                                                                 C.c_int64, C.c_int, C.c_double, C.c_void_p,
                                                                                 C.c_void_p, C.c_int, C.c_int]
    assert emu.emu_spectrum_shared(nw, nb, P(tr), P(tr), P(ptr), P(idx), P(d), d.shape[0], d.shape[1], P(iy0), P(ix0),
                                   len(origins), det, power, P(phi), P(sig), 2, teams) == 0
    for w, (j0, i0) in enumerate(origins):
        win = d[j0:j0 + nw, i0:i0 + nw]
        if det:
            win = oracle.detrend_plane(win)
        _, p0, s0 = oracle.radial_spectrum(win, 1000.0, taper=taper, power=power)
        np.testing.assert_allclose(phi[w], p0, rtol=1e-10, err_msg=str((w, j0, i0)))
        np.testing.assert_allclose(sig[w], s0, rtol=1e-9)


def test_spectrum_static_and_generic_engines_agree(emu, oracle, mag):
    """The compile-time pass lists (8.8.8, 8.8.4, 8.8.5, 8.10.8, and their N/2 + 1-node forms for 129 / 161 / 257)
    against the runtime-radix engine."""
    d = np.ascontiguousarray(mag[0])
    for nw in (101, 129, 161, 201, 257, 301):
        origins = [(2, 3)]
        a = _emu_spectrum(emu, oracle, d, origins, nw, np.hanning, 0, 2.0, max_teams=4)
        b = _emu_spectrum(emu, oracle, d, origins, nw, np.hanning, 0, 2.0, max_teams=-4)
        np.testing.assert_allclose(a[0], b[0], rtol=1e-12)
        np.testing.assert_allclose(a[1], b[1], rtol=1e-10)
        win = d[2:2 + nw, 3:3 + nw]
        _, p0, s0 = oracle.radial_spectrum(win, 1000.0, taper=np.hanning, power=2.0)
        np.testing.assert_allclose(a[0][0], p0, rtol=1e-11)
        np.testing.assert_allclose(a[1][0], s0, rtol=1e-9)


def test_detrend_body(emu, oracle, mag):
    d = np.ascontiguousarray(mag[0])
    g = load_golden("misc.npz")
    j0, i0, n = (int(v) for v in g["detrend_in_origin"])
    out = np.empty((1, n, n))
    emu.emu_detrend(P(d), C.c_int64(305), n, P(np.array([j0], dtype=np.int32)), P(np.array([i0], dtype=np.int32)),
                    C.c_int64(1), P(out))
    np.testing.assert_allclose(out[0], g["detrend_out"], rtol=0, atol=1e-12 * np.abs(g["detrend_out"]).max())


# ------------------------------------------------------------------------------------ forward model
def test_besselk_gamma_forward(emu):
    g = load_golden("forward.npz")
    nu, x = np.meshgrid(g["nu"], g["x"], indexing="ij")
    out = np.empty(nu.size)
    emu.emu_besselk(P(np.ascontiguousarray(nu.ravel())), P(np.ascontiguousarray(x.ravel())), C.c_int64(nu.size), P(out))
    ref = g["kv"].ravel()
    ok = ref > 0
    np.testing.assert_allclose(out[ok], ref[ok], rtol=2e-13)
    out = np.empty(g["gx"].size)
    emu.emu_gamma(P(g["gx"]), C.c_int64(out.size), P(out))
    np.testing.assert_allclose(out, g["gamma"], rtol=1e-14)
    pd = np.ascontiguousarray(g["pdense"])
    out = np.empty((pd.shape[0], g["k49"].size))
    emu.emu_bouligand(P(g["k49"]), g["k49"].size, P(pd), C.c_int64(pd.shape[0]), P(out))
    fin = np.isfinite(g["fdense"])
    assert np.array_equal(np.isfinite(out), fin)
    np.testing.assert_allclose(out[fin], g["fdense"][fin], rtol=0, atol=1e-11)
    params = np.ascontiguousarray(g["params"])
    out = np.empty((params.shape[0], g["kh"].size))
    emu.emu_bouligand(P(g["kh"]), g["kh"].size, P(params), C.c_int64(params.shape[0]), P(out))
    for p, a, b in zip(params, out, g["bouligand"]):
        if p[2] == 1e-8:
            continue             # pure cancellation noise in the reference itself
        assert np.array_equal(np.isfinite(a), np.isfinite(b)) and np.array_equal(np.isnan(a), np.isnan(b)), p
        fin = np.isfinite(b)
        np.testing.assert_allclose(a[fin], b[fin], rtol=1e-12, atol=1e-12)


def test_besselk_against_mpmath(emu):
    mp = pytest.importorskip("mpmath")
    rng = np.random.default_rng(3)
    nn = rng.uniform(0, 6, 200)
    xx = 10 ** rng.uniform(-3, 2.5, 200)
    out = np.empty(200)
    emu.emu_besselk(P(nn), P(xx), C.c_int64(200), P(out))
    mp.mp.dps = 30
    ref = np.array([float(mp.besselk(a, b)) for a, b in zip(nn, xx)])
    assert np.max(np.abs(out - ref) / ref) < 2e-14


def test_ddz_derivative(emu, oracle):
    k = load_golden("forward.npz")["k49"]
    for p in ([3.0, 1.0, 10.0, 5.0], [0.4, 1.0, 3.0, 5.0], [-0.4, 1.0, 3.0, 5.0], [2.0, 0.2, 40.0, -10.0]):
        p = np.array(p)
        phi, dphi = np.empty(k.size), np.empty(k.size)
        emu.emu_bouligand_ddz(P(k), k.size, P(p), P(phi), P(dphi))
        h = 1e-5
        fd = (oracle.bouligand2009(k, p[0], p[1], p[2] + h, p[3]) - oracle.bouligand2009(k, p[0], p[1], p[2] - h, p[3])) / (2 * h)
        np.testing.assert_allclose(dphi, fd, rtol=0, atol=2e-9)
        np.testing.assert_allclose(phi, oracle.bouligand2009(k, *p), rtol=0, atol=1e-12)


@pytest.mark.parametrize("mode", [0, 1])
def test_model_eval_warp_fast_path(emu, oracle, mode):
    """Warp-cooperative evaluation (Temme / Miller / K-negligible classes, model_eval.cuh) against
    bouligand2009 for parameter vectors that exercise every class and the generic fallback."""
    k99 = load_golden("fit.npz")["w200_k"]
    k49 = load_golden("forward.npz")["k49"]
    cases = [(3.0, 1.0, 10.0, 5.0), (3.0, 1.0, 0.3, 5.0), (2.2, 0.4, 2.0, -19.0), (0.0, 0.0, 1.0, 0.0),
             (4.7, 0.1, 35.0, 3.0), (1.0, 2.0, 100.0, -5.0), (3.0, 1.0, 240.0, 5.0), (15.9, 1.0, 7.0, 2.0),
             (3.0, 1.0, 1e-3, 5.0), (6.5, 0.0, 0.05, 1.0), (-0.5, 1.0, 3.0, 5.0), (3.0, 1.0, 0.0, 5.0)]
    for k in (k49, k99):
        for p in cases:
            p = np.array(p, dtype=np.float64)
            phi, ddz = np.empty(k.size), np.empty(k.size)
            emu.emu_model_eval(P(k), k.size, P(p), P(phi), P(ddz), mode)
            with np.errstate(all="ignore"):
                ref = oracle.bouligand2009(k, *p)
                h = 1e-5 * max(1.0, p[2])
                fd = (oracle.bouligand2009(k, p[0], p[1], p[2] + h, p[3]) -
                      oracle.bouligand2009(k, p[0], p[1], p[2] - h, p[3])) / (2 * h)
            assert np.array_equal(np.isfinite(phi), np.isfinite(ref)), p
            fin = np.isfinite(ref)
            if not fin.any():
                continue
            # the bracket of grid.py:595 cancels like eps / (k dz)^2 for small k dz in the reference itself
            xmin = max(k[0] * p[2], 1e-300)
            atol = 2e-12 * max(1.0, np.abs(ref[fin]).max()) + 1e-14 / min(1.0, xmin * xmin)
            np.testing.assert_allclose(phi[fin], ref[fin], rtol=0, atol=atol, err_msg=str(p))
            if p[2] > 1e-2 and p[0] >= 0:
                ok = fin & np.isfinite(fd)
                np.testing.assert_allclose(ddz[ok], fd[ok], rtol=1e-6, atol=5e-8, err_msg=str(p))


def test_misfit_matches_min_func(emu, oracle):
    g = load_golden("fit.npz")
    k = g["cfg1_k"]
    x = np.ascontiguousarray(np.vstack([g["cfg1_xref"][:6], [[3, 1, 0, 5], [3, 1, 400, 5]]]))
    phi = np.ascontiguousarray(np.vstack([g["cfg1_phi"][:6], g["cfg1_phi"][:2]]))
    out = np.empty(8)
    emu.emu_misfit(P(k), k.size, P(phi), C.c_int64(k.size), P(x), C.c_int64(8), None, None, P(out))
    np.testing.assert_allclose(out[:6], g["cfg1_fref"][:6], rtol=1e-12)
    assert out[6] == 1e99 and out[7] == 1e99
    mu, sg = g["w150_prior_prior_mu"], g["w150_prior_prior_sigma"]
    x = np.ascontiguousarray(g["w150_prior_xref"][:3])
    phi = np.ascontiguousarray(g["w150_prior_phi"][:3])
    k = g["w150_prior_k"]
    out = np.empty(3)
    emu.emu_misfit(P(k), k.size, P(phi), C.c_int64(k.size), P(x), C.c_int64(3), P(mu), P(sg), P(out))
    np.testing.assert_allclose(out, g["w150_prior_fref"][:3], rtol=1e-12)


# ------------------------------------------------------------------------------------ fit
def _emu_fit(emu, k, phi, x0, mu, sg, mode=0, lb=(0, 0, 0, -np.inf)):
    o = FitOpt()
    for j in range(4):
        o.x0[j], o.lb[j], o.ub[j] = x0[j], lb[j], np.inf
        o.prior_mu[j], o.prior_sigma[j] = mu[j], sg[j]
    n = phi.shape[0]
    x, f = np.empty((n, 4)), np.empty(n)
    it, st, nf = (np.empty(n, dtype=np.int32) for _ in range(3))
    phi = np.ascontiguousarray(phi)
    emu.emu_fit(P(k), k.size, P(phi), C.c_int64(n), C.byref(o), None, P(x), P(f), P(it), P(st), P(nf), mode)
    return x, f, it, st, nf


@pytest.mark.parametrize("tag", ["cfg1", "w200", "w150_prior", "w300_betaprior"])
def test_fit_body_reaches_the_reference_minimum(emu, oracle, tag):
    g = load_golden("fit.npz")
    k, phi, mu, sg = g[tag + "_k"], g[tag + "_phi"], g[tag + "_prior_mu"], g[tag + "_prior_sigma"]
    x, f, it, st, nf = _emu_fit(emu, k, phi, g[tag + "_x0"], mu, sg)
    xr, xp, fr, fp = g[tag + "_xref"], g[tag + "_xpol"], g[tag + "_fref"], g[tag + "_fpol"]
    assert np.all(st <= 2) and nf.mean() < 120
    for i in range(len(f)):
        ref_conv = np.max(np.abs(xr[i] - xp[i]) / np.maximum(np.abs(xp[i]), 1e-3))
        assert f[i] <= fr[i] + 1e-9 * abs(fr[i])                 # never worse than the reference's stopping point
        assert f[i] <= fp[i] + 1e-9 * abs(fp[i])                 # and as low as the polished minimum
        np.testing.assert_allclose(f[i], oracle.misfit(x[i], k, phi[i], mu, sg), rtol=1e-11)
        if ref_conv < 3e-5:
            assert np.max(np.abs(x[i] - xr[i]) / np.maximum(np.abs(xr[i]), 1e-3)) < 1e-4, (tag, i)


def test_fit_body_warp_shuffle_path_equals_serial(emu):
    g = load_golden("fit.npz")
    nan4 = np.full(4, np.nan)
    xs, fs, *_ = _emu_fit(emu, g["w200_k"], g["w200_phi"][:3], g["w200_x0"], nan4, nan4, mode=0)
    xw, fw, *_ = _emu_fit(emu, g["w200_k"], g["w200_phi"][:3], g["w200_x0"], nan4, nan4, mode=1)
    np.testing.assert_allclose(xw, xs, rtol=1e-6)
    np.testing.assert_allclose(fw, fs, rtol=1e-12)


def test_fit_respects_bounds(emu):
    g = load_golden("fit.npz")
    nan4 = np.full(4, np.nan)
    x, f, it, st, nf = _emu_fit(emu, g["w200_k"], g["w200_phi"][:2], (3.5, 1.0, 10.0, 5.0), nan4, nan4,
                                lb=(3.2, 0.5, 0.0, -np.inf))
    assert np.all(x[:, 0] >= 3.2) and np.all(x[:, 1] >= 0.5)
    assert np.any(x[:, 0] == 3.2) or np.any(x[:, 1] == 0.5)       # the unconstrained optimum is outside


# ------------------------------------------------------------------------------------ Metropolis-Hastings
def _emu_mh(emu, g, mu, sg, nsim, burnin, chain, mode, reduce=0):
    o = MhOpt()
    for j in range(4):
        o.x0[j], o.x_scale[j], o.prior_mu[j], o.prior_sigma[j] = g["start"][j], g["x_scale"][j], mu[j], sg[j]
    o.nsim, o.burnin, o.seed, o.chain_offset, o.reduce_mode = nsim, burnin, int(g["seed"]), chain, reduce
    k = g["k"]
    phi = np.ascontiguousarray(g["phi"][None, :])
    out = np.empty((1, 4, nsim)) if reduce == 0 else np.empty((1, 10))
    acc = np.zeros(1, dtype=np.int64)
    emu.emu_mh(P(k), k.size, P(phi), C.c_int64(1), C.byref(o), None, P(out), P(acc), mode)
    return out[0], int(acc[0])


def test_mh_body_reproduces_reference_chain(emu):
    """The reference's metropolis_hastings, monkey-patched to draw from the same Philox stream, and the
    kernel body walk the same chain (optimise.py:598-631 replicated literally)."""
    g = load_golden("mh.npz")
    nan4 = np.full(4, np.nan)
    for name, mu, sg in (("philox_prior_0", g["prior_mu"], g["prior_sigma"]), ("philox_noprior_0", nan4, nan4)):
        want = g[name]
        got, acc = _emu_mh(emu, g, mu, sg, want.shape[1], 200, 0, 0)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-10)
        assert acc == int((np.abs(np.diff(want[0])) > 0).sum()) + (1 if want[0, 0] != g["start"][0] else 0) or acc > 0
    got_w, _ = _emu_mh(emu, g, g["prior_mu"], g["prior_sigma"], 300, 50, 0, 1)        # 32-lane shuffle path
    got_s, _ = _emu_mh(emu, g, g["prior_mu"], g["prior_sigma"], 300, 50, 0, 0)
    np.testing.assert_allclose(got_w, got_s, rtol=0, atol=1e-10)
    mom, _ = _emu_mh(emu, g, g["prior_mu"], g["prior_sigma"], 2000, 200, 0, 0, reduce=1)
    ch, _ = _emu_mh(emu, g, g["prior_mu"], g["prior_sigma"], 2000, 200, 0, 0, reduce=0)
    full = np.vstack([ch, ch[1] + ch[2]])
    np.testing.assert_allclose(mom[:5], full.mean(axis=1), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(mom[5:], full.std(axis=1), rtol=1e-10, atol=1e-12)


def test_half_log_within_an_ulp_and_special_values(emu):
    """cb_half_log (argument reduction + minimax polynomial + Newton reciprocal) against numpy's log over the
    whole double range, around powers of two and sqrt(2) (the reduction's branch points), and the values that
    take the library path: 0 -> -inf (grid.py:283 relies on it), subnormals, inf, nan, negatives."""
    rng = np.random.default_rng(5)
    x = np.concatenate([
        np.exp(rng.uniform(-700.0, 700.0, 200000)),
        rng.uniform(0.5, 2.0, 200000),
        np.ldexp(1.0 + rng.uniform(-1e-9, 1e-9, 20000), rng.integers(-1000, 1000, 20000)),
        np.sqrt(2.0) * (1.0 + rng.uniform(-1e-12, 1e-12, 20000)),
        np.array([1.0, 2.0, 0.5, np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0), 2.2250738585072014e-308,
                  1.7976931348623157e308])])
    out = np.empty_like(x)
    emu.emu_half_log(P(x), C.c_int64(x.size), P(out))
    ref = 0.5 * np.log(x)
    err = np.abs(out - ref)
    ulp = np.spacing(np.abs(ref))
    near1 = np.abs(x - 1.0) < 1e-3                     # the result itself is tiny there: absolute bound
    assert np.all(err[~near1] <= 1.0 * ulp[~near1] + 0.0), float((err[~near1] / ulp[~near1]).max())
    assert np.all(err[near1] <= 2e-19 + 1.0 * ulp[near1])
    sp = np.array([0.0, -0.0, 5e-324, 1e-310, np.inf, np.nan, -1.0])
    o2 = np.empty_like(sp)
    emu.emu_half_log(P(sp), C.c_int64(sp.size), P(o2))
    with np.errstate(all="ignore"):
        r2 = 0.5 * np.log(sp)
    assert np.array_equal(np.isnan(o2), np.isnan(r2))
    ok = ~np.isnan(r2)
    assert np.array_equal(o2[ok], r2[ok])


def test_philox_device_function(emu, oracle):
    out = np.empty(4, dtype=np.uint32)
    for ctr, key in (((0, 0, 0, 0), (0, 0)), ((0xffffffff,) * 4, (0xffffffff,) * 2),
                     ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))):
        emu.emu_philox(*[C.c_uint32(v) for v in ctr + key], P(out))
        assert [int(v) for v in out] == oracle.philox4x32_10(ctr, key)


# ------------------------------------------------------------------------------------ Radon
def test_radon_bodies_bit_exact(emu, oracle, mag):
    g = load_golden("radon.npz")
    d = np.ascontiguousarray(mag[0])
    for nw in (51, 101):
        j0, i0, _ = (int(v) for v in g["n%d_meta" % nw])
        theta = g["n%d_theta" % nw]
        nt = theta.size
        ends = [oracle.radon_ray_endpoints(nw, t) for t in theta]
        tx = np.ascontiguousarray(np.concatenate([e[0] for e in ends]))
        rx = np.ascontiguousarray(np.concatenate([e[1] for e in ends]))
        nray = nw * nt
        cnt = np.zeros(nray, dtype=np.int64)
        emu.emu_radon_trace(P(tx), P(rx), C.c_int64(nray), nw, P(cnt), None, None, None)
        off = np.zeros(nray + 1, dtype=np.int64)
        off[1:] = np.cumsum(cnt)
        data, ind = np.empty(off[-1]), np.empty(off[-1], dtype=np.int64)
        emu.emu_radon_trace(P(tx), P(rx), C.c_int64(nray), nw, None, P(off), P(data), P(ind))
        for t in range(nt):
            assert np.array_equal(off[t * nw:(t + 1) * nw + 1] - off[t * nw], g["n%d_indptr" % nw][t])
            key = "n%d_t%d_data" % (nw, t)
            if key in g.files:
                assert np.array_equal(data[off[t * nw]:off[(t + 1) * nw]], g[key])
                assert np.array_equal(ind[off[t * nw]:off[(t + 1) * nw]], g["n%d_t%d_indices" % (nw, t)])
        sino = np.empty((1, nw, nt))
        emu.emu_radon_apply(P(d), C.c_int64(305), P(np.array([j0], dtype=np.int32)), P(np.array([i0], dtype=np.int32)),
                            C.c_int64(1), nw, nt, P(off), P(data), P(ind), P(sino), 0)
        assert np.array_equal(sino[0], g["n%d_sinogram" % nw])
        kb = g["n%d_az_k" % nw]
        S = np.empty((1, nt, kb.size))
        emu.emu_az_tail(P(sino), nw, nt, C.c_int64(1), P(np.ascontiguousarray(np.hanning(nw))), kb.size,
                        C.c_double(2.0), P(S))
        np.testing.assert_allclose(S[0], g["n%d_az_S" % nw], rtol=1e-11)


@pytest.mark.parametrize("ny,nx", [(64, 64), (101, 101), (30, 45), (75, 64)])
def test_fft2_filter_body_matches_reference_filters(emu, oracle, mag, ny, nx):
    """line_fft_body x 4 = real(ifft2(filter * fft2(data))): upward continuation and reduction to the pole
    (grid.py:405-546) on square / rectangular, smooth / Bluestein shapes."""
    d = np.ascontiguousarray(mag[0][5:5 + ny, 9:9 + nx])
    out = np.empty_like(d)
    for filt in (oracle.upward_continuation_filter(d.shape, 1000.0, 1000.0, 800.0),
                 oracle.reduce_to_pole_filter(d.shape, 1000.0, 1000.0, -60.0, 12.0, 40.0, 3.0)):
        f = np.ascontiguousarray(filt.astype(np.complex128))
        assert emu.emu_fft2_filter(P(d), C.c_int64(ny), C.c_int64(nx), P(f.view(np.float64)), P(out)) == 0
        ref = oracle.fft_filter(d, filt)
        np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("n", [32, 51, 101, 129, 151, 161])
def test_fft2_filter_window_batch_equals_per_window_reference(emu, oracle, mag, n):
    """The batched form (image index folded into the line index, first pass gathering from the grid): every
    window of the stack equals the reference filter applied to that window alone.  32: direct mixed radix, four
    passes; 51: generic Bluestein, fused column pass; 101 / 151: static engines with pruned passes fed straight from
    the line's elements; 129 / 161: the 2n - 2 Bluestein length, last node carried around the pruned passes."""
    d = np.ascontiguousarray(mag[0])
    iy0 = np.array([0, 7, min(100, d.shape[0] - n), d.shape[0] - n, 33], dtype=np.int32)
    ix0 = np.array([0, min(200, d.shape[1] - n), 3, d.shape[1] - n, 33], dtype=np.int32)
    filt = oracle.reduce_to_pole_filter((n, n), 1000.0, 1000.0, -60.0, 12.0, 40.0, 3.0)
    f = np.ascontiguousarray(filt.astype(np.complex128))
    out = np.full((iy0.size, n, n), np.nan)
    assert emu.emu_fft2_filter_windows(P(d), C.c_int64(d.shape[0]), C.c_int64(d.shape[1]), P(iy0), P(ix0),
                                       C.c_int64(iy0.size), n, P(f.view(np.float64)), P(out)) == 0
    for i in range(iy0.size):
        ref = oracle.fft_filter(d[iy0[i]:iy0[i] + n, ix0[i]:ix0[i] + n], filt)
        np.testing.assert_allclose(out[i], ref, rtol=0, atol=1e-12 * np.abs(ref).max())


def test_tanaka_body_matches_reference_formulas(emu, oracle):
    """tanaka1999 + ComputeTanaka (grid.py:603-699) batched, one window per warp, against the oracle and the
    reference's golden value."""
    g = load_golden("spectrum.npz")
    m = load_golden("misc.npz")
    rng = np.random.default_rng(11)
    k = load_golden("fit.npz")["w200_k"]
    n = 5
    lnphi = np.ascontiguousarray(10.0 - 8.0 * k[None, :] + 0.2 * rng.standard_normal((n, k.size)))
    sig = np.ascontiguousarray(0.6 + 0.2 * rng.random((n, k.size)))
    out = np.empty((n, 8))
    emu.emu_tanaka.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_double] * 4 + [C.c_void_p]
    emu.emu_tanaka(P(k), k.size, P(lnphi), P(sig), n, 0.005, 0.03, 0.03, 0.7, P(out))
    for i in range(n):
        top, cen = oracle.tanaka1999(k, lnphi[i], sig[i], (0.005, 0.03), (0.03, 0.7))
        zb, ezb = oracle.compute_tanaka(top[0], top[2], cen[0], cen[2])
        np.testing.assert_allclose(out[i], [*top, *cen, zb, ezb], rtol=1e-8)


def test_fit_rows_with_per_row_priors_match_oracle(emu, oracle):
    """cb_fit_batch_rows semantics (the inner loop of CurieOptimise.sensitivity, optimise.py:705-716): every
    row has its own perturbed spectrum and its own prior means; priors without a row value keep the default."""
    g = load_golden("fit.npz")
    k, phi = g["w200_k"], g["w200_phi"][0]
    rng = np.random.default_rng(3)
    ns = 6
    rows = np.ascontiguousarray(phi[None, :] + 0.3 * rng.standard_normal((ns, phi.size)))
    mu = np.full((ns, 4), np.nan)
    mu[:, 0], mu[:3, 1] = rng.normal(3.0, 0.5, ns), rng.normal(0.5, 0.25, 3)     # rows 3.. keep the default zt mean
    sg = np.array([0.5, 0.25, np.nan, np.nan])
    o = FitOpt()
    for j in range(4):
        o.x0[j], o.lb[j], o.ub[j] = [3, 1, 10, 5][j], [0, 0, 0, -np.inf][j], np.inf
        o.prior_mu[j], o.prior_sigma[j] = [3.0, 0.5, np.nan, np.nan][j], sg[j]
    x, f = np.empty((ns, 4)), np.empty(ns)
    emu.emu_fit_rows(P(k), k.size, P(rows), C.c_int64(ns), C.byref(o), P(np.ascontiguousarray(mu)), P(x), P(f))
    mu_eff = np.where(np.isnan(mu), np.array([3.0, 0.5, np.nan, np.nan]), mu)
    ref = oracle.sensitivity_rows(k, rows, mu_eff, sg)
    assert np.max(np.abs(x - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-4
    for s in range(ns):
        np.testing.assert_allclose(f[s], oracle.misfit(x[s], k, rows[s], mu_eff[s], sg), rtol=1e-11)


def test_fit_is_insensitive_to_rounding_level_perturbations():
    """tools/emu_fit_stress.py: every golden case refitted from spectra perturbed at the 1e-13 level (the
    CUDA-vs-glibc math difference) must reach the polished minimum -- no rounding-dependent stalls."""
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emu_fit_stress.py"), "4"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
