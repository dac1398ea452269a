"""Parity of the CUDA path (called through the C-ABI via pycurious_b200) against the reference's
golden vectors and the CPU oracle.  Needs a B200: run with `pytest -m gpu`."""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden, TAPERS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pc():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import pycurious_b200
    from pycurious_b200 import _lib
    _lib.load()
    return pycurious_b200


@pytest.fixture(scope="module")
def grid(pc, mag):
    d, ext = mag
    return pc.CurieOptimise(d, *ext)


# --------------------------------------------------------------------------- spectrum stage
def test_spectrum_golden_cases(pc, grid, mag):
    """Every golden radial_spectrum case of the reference: k bit-exact, Phi/sigma rtol 1e-10."""
    d, _ = mag
    g = load_golden("spectrum.npz")
    for entry in g["names"]:
        name, tp = str(entry).split(":")
        j0, i0, n, det = (int(v) for v in g[name + "_meta"])
        win = grid.data[j0:j0 + n, i0:i0 + n]
        kw = {"beta": float(g[name + "_kaiser_beta"])} if tp == "kaiser" else {}
        if det:
            win = grid.remove_trend_linear(win)
        k, phi, sig = grid.radial_spectrum(win, taper=TAPERS[tp], power=float(g[name + "_power"]), **kw)
        assert np.array_equal(k, g[name + "_k"]), name
        np.testing.assert_allclose(phi, g[name + "_phi"], rtol=1e-10, err_msg=name)
        np.testing.assert_allclose(sig, g[name + "_sig"], rtol=1e-9, err_msg=name)


def test_fused_detrend_matches_reference(grid):
    g = load_golden("spectrum.npz")
    name = "c200_hann_detrend"
    k, phi, sig = grid.radial_spectrum_batch(200e3, [140.5e3], [160.5e3], detrend=True)
    np.testing.assert_allclose(phi[0], g[name + "_phi"], rtol=1e-10)
    np.testing.assert_allclose(sig[0], g[name + "_sig"], rtol=1e-9)
    m = load_golden("misc.npz")
    j0, i0, n = (int(v) for v in m["detrend_in_origin"])
    out = grid.remove_trend_linear(grid.data[j0:j0 + n, i0:i0 + n])
    np.testing.assert_allclose(out, m["detrend_out"], rtol=0, atol=1e-12 * np.abs(m["detrend_out"]).max())


@pytest.mark.parametrize("nw", [101, 128, 129, 201, 301])
def test_device_bin_map_bit_exact(pc, oracle, nw):
    """The CSR the kernel reduces over is the reference's membership, folded onto the half plane:
    integers compared exactly (gate G1a)."""
    from pycurious_b200 import _plans
    plan = _plans.spectrum_plan(nw, 1000.0, None)
    ptr, idx = plan.folded_bins()
    _, dk, edges = oracle.taper_and_bins(nw, nw, 1000.0, None)
    _, masks = oracle.radial_bin_masks(nw, dk, edges)
    g = load_golden("spectrum.npz")
    half = nw // 2
    for b, m in enumerate(masks):
        a, c = np.nonzero(m)
        u = (a + (nw + 1) // 2) % nw
        v = (c + (nw + 1) // 2) % nw
        fold = v > half
        u = np.where(fold, (nw - u) % nw, u)
        v = np.where(fold, nw - v, v)
        self_mirror = ((v == 0) | (2 * v == nw)) & (u > half)
        u = np.where(self_mirror, nw - u, u)
        want = np.sort(v * nw + u)
        got = idx[ptr[b]:ptr[b + 1]]
        uniq, cnt = np.unique(want, return_counts=True)
        if cnt.size and np.all(cnt == cnt[0]) and cnt[0] > 1:
            want = uniq
            assert got.size * cnt[0] == g["bins%d_counts" % nw][b]
        else:
            assert got.size == g["bins%d_counts" % nw][b]
        assert np.array_equal(got, want), (nw, b)


def test_spectrum_batch_equals_single_and_oracle(grid, oracle, mag):
    d, ext = mag
    spec = oracle.GridSpec(d, *ext)
    xc, yc = grid.create_centroid_list(128e3, 23e3, 31e3)
    k, phi, sig = grid.radial_spectrum_batch(128e3, xc, yc, power=0.5)
    assert phi.shape == (xc.size, k.size)
    for i in range(0, xc.size, 7):
        k0, p0, s0 = oracle.radial_spectrum(oracle.cut_window(spec, 128e3, xc[i], yc[i]), spec.dx, power=0.5)
        assert np.array_equal(k, k0)
        np.testing.assert_allclose(phi[i], p0, rtol=1e-10)
        np.testing.assert_allclose(sig[i], s0, rtol=1e-9)


@pytest.mark.parametrize("detrend", [False, True])
def test_shared_row_spectra_path_equals_direct_path(pc, grid, oracle, mag, detrend, monkeypatch):
    """cb_spectrum_batch_shared_rows (rows transformed once per distinct ix0) against cb_spectrum_batch and
    the oracle on a centroid lattice, in one piece and with the column groups forced into several slices."""
    import torch
    from pycurious_b200 import _plans
    d, ext = mag
    spec = oracle.GridSpec(d, *ext)
    xc, yc = grid.create_centroid_list(100e3, 9e3, 14e3)
    iy0, ix0, n = grid._origins_device(100e3, xc, yc)
    plan = _plans.spectrum_plan(n, grid.dx, _plans.taper_vector(np.hanning, n))
    g = grid.device_grid()
    p_dir, s_dir = plan.run(g, iy0, ix0, detrend, 2.0, share_rows=False)
    assert plan.last_path == "direct"
    plan.last_path = "direct"
    p_sh, s_sh = plan.run(g, iy0, ix0, detrend, 2.0, share_rows=True)
    assert plan.last_path == "shared_rows"
    monkeypatch.setenv("CB_SPEC_ROWSPEC_MAX_GB", "%.6f" % (5.2 * 51 * 306 * 16 / 2 ** 30))   # ~5 groups per slice
    p_ch, s_ch = plan.run(g, iy0, ix0, detrend, 2.0, share_rows=True)
    torch.cuda.synchronize()
    for a, b in ((p_sh, p_dir), (p_ch, p_dir)):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-11, atol=1e-11)
    for a, b in ((s_sh, s_dir), (s_ch, s_dir)):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-9)
    assert np.array_equal(p_sh.cpu().numpy(), p_ch.cpu().numpy())
    ph = p_sh.cpu().numpy()
    for i in range(0, xc.size, 37):
        win = oracle.cut_window(spec, 100e3, xc[i], yc[i])
        if detrend:
            win = oracle.detrend_plane(win)
        _, p0, _ = oracle.radial_spectrum(win, spec.dx)
        np.testing.assert_allclose(ph[i], p0, rtol=1e-10)


@pytest.mark.parametrize("nw", [20, 21, 30, 51, 64, 75, 128, 129, 151, 161, 256, 257, 305, 321])
def test_window_size_sweep_both_paths(pc, oracle, nw):
    """Odd / even / prime / 2-3-5-smooth window edges (Bluestein on static and generic engines, direct
    mixed radix; 129 / 161 / 257 / 321 = N/2 + 1 nodes on the length-N static engines, PlusOneFft), each through cb_spectrum_batch and cb_spectrum_batch_shared_rows, with taper + detrend,
    against the oracle."""
    import torch
    from pycurious_b200 import _plans
    rng = np.random.default_rng(nw)
    ny, nx = 2 * nw + 7, 2 * nw + 12
    data = rng.standard_normal((ny, nx)).cumsum(axis=0).cumsum(axis=1)      # red, trended field
    ext = pc.synth.grid_extent(ny, nx)
    G = pc.CurieGrid(data, *ext)
    origins = [(0, 0), (3, 5), (ny - nw, nx - nw), (3, nx - nw), (8, 5)]
    iy0 = torch.tensor([o[0] for o in origins], dtype=torch.int32, device="cuda")
    ix0 = torch.tensor([o[1] for o in origins], dtype=torch.int32, device="cuda")
    plan = _plans.spectrum_plan(nw, G.dx, _plans.taper_vector(np.hamming, nw))
    g = G.device_grid()
    for share in (False, True):
        phi, sig = plan.run(g, iy0, ix0, True, 2.0, share_rows=share)
        phi, sig = phi.cpu().numpy(), sig.cpu().numpy()
        for w, (j0, i0) in enumerate(origins):
            win = oracle.detrend_plane(data[j0:j0 + nw, i0:i0 + nw])
            k0, p0, s0 = oracle.radial_spectrum(win, G.dx, taper=np.hamming, power=2.0)
            np.testing.assert_allclose(phi[w], p0, rtol=2e-10, atol=1e-10, err_msg=str((nw, share, w)))
            np.testing.assert_allclose(sig[w], s0, rtol=1e-8)


def test_emag2_shaped_grid_far_corner(pc, oracle):
    """BASELINE config 5 geometry: a 21600 x 10800 node grid (1.87 GB).  Windows in the far corner address
    elements beyond 2^31 bytes; all three window sizes (101 / 201 / 301 nodes -> Bluestein 256 / 512 / 640),
    scattered (direct path) and on a small lattice (shared row spectra)."""
    ny, nx = 10800, 21600
    rng = np.random.default_rng(99)
    data = rng.standard_normal((ny, nx))
    ext = pc.synth.grid_extent(ny, nx)
    G = pc.CurieOptimise(data, *ext)
    spec = oracle.GridSpec(data, *ext)
    for window in (100e3, 200e3, 300e3):
        half = (int(round(window / 1000.0)) // 2) * 1000.0
        xs = np.array([ext[1] - half, ext[1] - half - 20e3, half, ext[1] - half])
        ys = np.array([ext[3] - half, ext[3] - half - 20e3, ext[3] - half, half])
        k, phi, sig = G.radial_spectrum_batch(window, xs, ys)
        for i in range(xs.size):
            k0, p0, s0 = oracle.radial_spectrum(oracle.cut_window(spec, window, xs[i], ys[i]), spec.dx)
            assert np.array_equal(k, k0)
            np.testing.assert_allclose(phi[i], p0, rtol=1e-10)
            np.testing.assert_allclose(sig[i], s0, rtol=1e-9)
    # lattice in the corner: 6 x 6 windows of 201 nodes, 20-node stride -> shared path (reuse >= 2)
    xq, yq = np.meshgrid(ext[1] - 100e3 - 20e3 * np.arange(6), ext[3] - 100e3 - 20e3 * np.arange(6))
    from pycurious_b200 import _plans
    k, phi, sig = G.radial_spectrum_batch(200e3, xq.ravel(), yq.ravel())
    assert _plans.spectrum_plan(201, G.dx, _plans.taper_vector(np.hanning, 201)).last_path == "shared_rows"
    for i in (0, 7, 35):
        _, p0, _ = oracle.radial_spectrum(oracle.cut_window(spec, 200e3, xq.ravel()[i], yq.ravel()[i]), spec.dx)
        np.testing.assert_allclose(phi[i], p0, rtol=1e-10)
    x = G.optimise_batch(200e3, xq.ravel()[:4], yq.ravel()[:4])
    assert x.shape == (4, 4) and np.all(np.isfinite(x))


def test_zero_modulus_propagates_like_numpy(pc, oracle):
    """log(0) = -inf (grid.py:283): an all-zero window gives Phi = -inf and sigma = nan in every bin, a window with
    a single non-zero DFT coefficient pattern (constant data, no taper -> only DC) gives -inf wherever |F| is
    exactly 0 -- through the shared-row path and the direct path, like the reference expression does."""
    n = 64
    d = np.zeros((160, 160))
    d[80:, :] = 3.5                                            # lower half constant
    G = pc.CurieGrid(d, 0.0, 159e3, 0.0, 159e3)
    xs = np.array([32e3, 40e3, 48e3, 32e3, 40e3, 48e3])         # nodes 32.., window 64 -> rows 0..63 / 90..153
    ys = np.array([32e3, 32e3, 32e3, 122e3, 122e3, 122e3])
    spec = oracle.GridSpec(d, 0.0, 159e3, 0.0, 159e3)
    for share in (True, False):
        from pycurious_b200 import _plans
        iy0, ix0, nw = G._origins_device(64e3, xs, ys)
        plan = _plans.spectrum_plan(nw, G.dx, None)
        phi, sig = plan.run(G.device_grid(), iy0, ix0, False, 2.0, share_rows=share)
        phi, sig = phi.cpu().numpy(), sig.cpu().numpy()
        assert np.all(np.isneginf(phi[:3])) and np.all(np.isnan(sig[:3]))          # zero windows
        with np.errstate(all="ignore"):
            k_o, phi_o, sig_o = oracle.radial_spectrum(oracle.cut_window(spec, 64e3, xs[3], ys[3]), spec.dx, taper=None)
        # constant window without taper: every non-DC coefficient is exactly 0 in exact arithmetic; in floating
        # point both implementations leave either exact zeros or rounding dust, so only the -inf / finite pattern
        # of bins whose members are ALL exact zeros in the reference is compared
        all_zero = np.isneginf(phi_o) & np.isnan(sig_o)
        assert np.all(np.isneginf(phi[3:][:, all_zero]) | (phi[3:][:, all_zero] < -50.0))


def test_spectrum_properties_full_size(pc):
    """Size-independent properties on a BASELINE-sized window set (4096-wide grid, 201-node windows):
    scaling the data by a shifts Phi by power*ln(a) and leaves sigma; identical windows give identical
    rows; a pure plane is removed by the fused detrend."""
    ny = nx = 1024
    data = pc.synth.fractal_grid(ny, nx, seed=5)
    ext = pc.synth.grid_extent(ny, nx)
    G = pc.CurieGrid(data, *ext)
    xc, yc = G.create_centroid_list(200e3, 50e3, 50e3)
    k, phi, sig = G.radial_spectrum_batch(200e3, xc, yc)
    G2 = pc.CurieGrid(data * 7.5, *ext)
    _, phi2, sig2 = G2.radial_spectrum_batch(200e3, xc, yc)
    np.testing.assert_allclose(phi2, phi + 2.0 * np.log(7.5), rtol=1e-11, atol=1e-10)
    np.testing.assert_allclose(sig2, sig, rtol=1e-9)
    # periodic tiling: windows one period apart see identical data
    tile = np.tile(data[:256, :256], (4, 4))
    G3 = pc.CurieGrid(tile, *ext)
    _, p3, _ = G3.radial_spectrum_batch(200e3, [150e3, 150e3 + 256e3, 150e3], [150e3, 150e3, 150e3 + 512e3])
    assert np.array_equal(p3[0], p3[1]) and np.array_equal(p3[0], p3[2])
    # plane + data with detrend == data's own detrended spectrum
    yy, xx = np.mgrid[0:ny, 0:nx]
    G4 = pc.CurieGrid(data + 3.0 * xx - 2.0 * yy + 500.0, *ext)
    _, p4, s4 = G4.radial_spectrum_batch(200e3, xc[:8], yc[:8], detrend=True)
    _, p5, s5 = G.radial_spectrum_batch(200e3, xc[:8], yc[:8], detrend=True)
    np.testing.assert_allclose(p4, p5, rtol=1e-8, atol=1e-8)


# --------------------------------------------------------------------------- forward model
def test_besselk_and_forward_model(pc):
    import torch
    from pycurious_b200 import _lib
    g = load_golden("forward.npz")
    nu, x = np.meshgrid(g["nu"], g["x"], indexing="ij")
    nu_t = torch.from_numpy(np.ascontiguousarray(nu.ravel())).cuda()
    x_t = torch.from_numpy(np.ascontiguousarray(x.ravel())).cuda()
    out = torch.empty_like(nu_t)
    _lib.check(_lib.load().cb_besselk(nu_t.data_ptr(), x_t.data_ptr(), nu_t.numel(), out.data_ptr(), None))
    got, want = out.cpu().numpy(), g["kv"].ravel()
    ok = want > 0          # scipy flushes kv(2, 700) to 0 earlier than the true value
    np.testing.assert_allclose(got[ok], want[ok], rtol=2e-13)
    with np.errstate(all="ignore"):
        for p, ref in zip(g["params"], g["bouligand"]):
            ours = pc.bouligand2009(g["kh"], *p)
            fin = np.isfinite(ref)
            if p[2] > 1e-6:          # dz = 1e-8 is pure cancellation noise in the reference itself
                assert np.array_equal(np.isfinite(ours), fin), p
            np.testing.assert_allclose(ours[fin & np.isfinite(ours)], ref[fin & np.isfinite(ours)], rtol=1e-12, atol=1e-12)
        dense = np.array([pc.bouligand2009(g["k49"], *p) for p in g["pdense"][:50]])
    ref = g["fdense"][:50]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(dense), fin)
    np.testing.assert_allclose(dense[fin], ref[fin], rtol=0, atol=1e-11)
    # SURVEY.md 8c golden values
    np.testing.assert_allclose(pc.bouligand2009(np.array([0.05, 0.5, 1, 3]), 3, 1, 10, 5),
                               [8.292380801142428, 4.286832827341122, 1.9013876157386207, -4.295836866004333],
                               rtol=1e-13)


def test_min_func_matches_reference(grid):
    g = load_golden("fit.npz")
    k = g["cfg1_k"]
    for i in (0, 5, 12):
        f = grid.min_func(g["cfg1_xref"][i], k, g["cfg1_phi"][i], None)
        np.testing.assert_allclose(f, g["cfg1_fref"][i], rtol=1e-12)
    assert grid.min_func(np.array([3.0, 1.0, 0.0, 5.0]), k, g["cfg1_phi"][0], None) == 1e99
    assert grid.min_func(np.array([3.0, 1.0, 400.0, 5.0]), k, g["cfg1_phi"][0], None) == 1e99


# --------------------------------------------------------------------------- MAP fit
def _check_fit(oracle, g, tag, ours):
    xref, xpol, fref = g[tag + "_xref"], g[tag + "_xpol"], g[tag + "_fref"]
    k, phis = g[tag + "_k"], g[tag + "_phi"]
    mu, sg = g[tag + "_prior_mu"], g[tag + "_prior_sigma"]
    ngate = 0
    for i in range(xref.shape[0]):
        ref_conv = np.max(np.abs(xref[i] - xpol[i]) / np.maximum(np.abs(xpol[i]), 1e-3))
        rel = np.max(np.abs(ours[i] - xref[i]) / np.maximum(np.abs(xref[i]), 1e-3))
        f_ours = oracle.misfit(ours[i], k, phis[i], mu, sg)
        # never worse than the reference's own stopping point
        assert f_ours <= fref[i] + 1e-9 * max(1.0, abs(fref[i])), (tag, i, f_ours, fref[i])
        if ref_conv < 3e-5:       # reference converged: parameters within 1e-4 relative (north_star gate)
            assert rel < 1e-4, (tag, i, rel, ours[i], xref[i])
            ngate += 1
    return ngate


def test_fit_config1_matches_reference(grid, oracle):
    g = load_golden("fit.npz")
    res = grid.optimise_routine(float(g["cfg1_window"]), g["cfg1_xc"], g["cfg1_yc"])
    assert isinstance(res, list) and len(res) == 4 and all(r.shape == (25,) for r in res)
    ngate = _check_fit(oracle, g, "cfg1", np.stack(res, axis=1))
    assert ngate >= 18


def test_fit_200km_priors_detrend(grid, oracle):
    g = load_golden("fit.npz")
    res = grid.optimise_routine(200e3, g["w200_xc"], g["w200_yc"])
    assert _check_fit(oracle, g, "w200", np.stack(res, axis=1)) == 9
    grid.reset_priors()
    grid.add_prior(beta=(3.0, 0.5), zt=(0.5, 0.25))
    try:
        res = grid.optimise_routine(150e3, g["w150_prior_xc"], g["w150_prior_yc"], taper=np.hamming,
                                    process_subgrid=grid.remove_trend_linear)
    finally:
        grid.reset_priors()
    assert _check_fit(oracle, g, "w150_prior", np.stack(res, axis=1)) >= 4
    grid.add_prior(beta=(1.0, 0.1))
    try:
        x = grid.optimise(300e3, float(g["w300_betaprior_xc"][0]), float(g["w300_betaprior_yc"][0]))
    finally:
        grid.reset_priors()
    assert x.shape == (4,)
    _check_fit(oracle, g, "w300_betaprior", x[None, :])


def test_fit_random_truths_bounds_and_priors(grid, oracle):
    """Spectra synthesised from random (beta, zt, dz, C) with noise, fitted under default bounds, a custom
    box (grid.bounds is user-writable, optimise.py:111-114) and priors: never worse than the reference's
    stopping point, parameters within 1e-4 where the reference is itself converged."""
    import torch
    from pycurious_b200 import _lib
    rng = np.random.default_rng(2026)
    k = load_golden("fit.npz")["w200_k"]
    n = 24
    truth = np.c_[rng.uniform(1.5, 4.5, n), rng.uniform(0.0, 3.0, n), rng.uniform(3.0, 60.0, n), rng.uniform(-25, 25, n)]
    phi = np.array([oracle.bouligand2009(k, *t) for t in truth]) + 0.5 * rng.standard_normal((n, k.size))
    phi_t = torch.from_numpy(phi).cuda()
    cases = [("default", [(0.0, None), (0.0, None), (0.0, None), (None, None)], {}),
             ("box", [(2.0, 4.0), (0.1, 2.0), (5.0, 40.0), (None, None)], {}),
             ("priors", [(0.0, None), (0.0, None), (0.0, None), (None, None)], dict(beta=(3.0, 0.3), dz=(20.0, 5.0)))]
    saved = grid.bounds
    try:
        for tag, bounds, priors in cases:
            grid.bounds = bounds
            grid.reset_priors()
            if priors:
                grid.add_prior(**priors)
            mu, sg = grid._prior_arrays()
            x = grid._fit_device(k, phi_t, (3.0, 1.0, 10.0, 5.0)).cpu().numpy()
            ngate = 0
            for i in range(n):
                ref = oracle.map_fit(k, phi[i], bounds=bounds, prior_mu=mu, prior_sigma=sg)
                f_ref = oracle.misfit(ref, k, phi[i], mu, sg)
                f_us = oracle.misfit(x[i], k, phi[i], mu, sg)
                assert f_us <= f_ref * (1 + 1e-9) + 1e-12, (tag, i, f_us, f_ref, x[i], ref)
                lo = np.array([-np.inf if b[0] is None else b[0] for b in bounds])
                hi = np.array([np.inf if b[1] is None else b[1] for b in bounds])
                assert np.all(x[i] >= lo - 1e-12) and np.all(x[i] <= hi + 1e-12), (tag, i, x[i])
                if abs(f_ref - f_us) <= 1e-9 * max(1.0, f_ref):          # the reference reached the same minimum
                    rel = np.max(np.abs(x[i] - ref) / np.maximum(np.abs(ref), 1e-3))
                    if rel < 1e-4:
                        ngate += 1
            assert ngate >= n // 2, (tag, ngate)
    finally:
        grid.bounds = saved
        grid.reset_priors()


def test_fit_custom_process_subgrid_and_single(grid, oracle, mag):
    """An arbitrary callable goes through the host path but is still fitted on the device."""
    d, ext = mag
    spec = oracle.GridSpec(d, *ext)
    xc, yc = 152.5e3, 140.5e3
    x_dev = grid.optimise(150e3, xc, yc, process_subgrid=lambda w: w * 2.0)
    win = oracle.cut_window(spec, 150e3, xc, yc) * 2.0
    k, phi, _ = oracle.radial_spectrum(win, spec.dx)
    xp, fp = oracle.polished_fit(k, phi, oracle.map_fit(k, phi))
    np.testing.assert_allclose(x_dev, xp, rtol=2e-5)


def test_fit_paths_agree(grid, monkeypatch):
    """The batched fit (csrc/fit_body.cuh) reaches a window's result by different launches depending on how many
    windows are unfinished: batched evaluation / step rounds, the CTA-per-window finish kernel after the hand-over,
    several passes over the scratch for large inputs, the shared first evaluation for a common start point or one
    evaluation per window for per-window start points.  Every route must give the same bits."""
    import torch
    g = load_golden("fit.npz")
    k, phi = g["w200_k"], np.ascontiguousarray(np.tile(g["w200_phi"], (40, 1)))          # 1,440 spectra
    phi *= 1.0 + 1e-3 * np.random.default_rng(5).standard_normal(phi.shape)
    phi_t = torch.from_numpy(phi).cuda()
    x0 = (3.0, 1.0, 10.0, 5.0)

    def run(**env):
        for name in ("CB_FIT_HANDOVER", "CB_FIT_ROUNDS", "CB_FIT_PASS"):
            monkeypatch.delenv(name, raising=False)
        for name, v in env.items():
            monkeypatch.setenv(name, str(v))
        x, d = grid._fit_device(k, phi_t, x0, diagnostics=True)
        return x.cpu().numpy(), {kk: v.cpu().numpy() for kk, v in d.items()}

    base, dbase = run()                                       # 1,440 < 2,048: everything in the finish kernel
    assert np.all(np.isfinite(base)) and np.all(dbase["status"] <= 1)
    for env in (dict(CB_FIT_HANDOVER=0), dict(CB_FIT_HANDOVER=0, CB_FIT_ROUNDS=3), dict(CB_FIT_HANDOVER=200),
                dict(CB_FIT_HANDOVER=0, CB_FIT_PASS=500), dict(CB_FIT_PASS=97, CB_FIT_HANDOVER=50)):
        x, d = run(**env)
        assert np.array_equal(x, base), env
        assert np.array_equal(d["nfev"], dbase["nfev"]) and np.array_equal(d["f"], dbase["f"]), env
    # per-window start points (no shared first evaluation): the same start for all must reproduce the common-start run
    monkeypatch.delenv("CB_FIT_PASS", raising=False)
    monkeypatch.setenv("CB_FIT_HANDOVER", "0")
    x0_dev = torch.tensor(x0, dtype=torch.float64, device="cuda").repeat(phi.shape[0], 1).contiguous()
    x, _ = grid._fit_device(k, phi_t, x0, x0_dev=x0_dev, diagnostics=True)
    assert np.array_equal(x.cpu().numpy(), base)


def test_fit_terminates_on_hostile_spectra(grid):
    """Rows of NaN / inf / huge / constant values, mixed with good ones, in one batch: every row finishes (status
    code set, iteration and evaluation counts within the solver's bounds), the good rows are unaffected by their
    neighbours, and a non-finite start gives the reference's 1e99 behaviour (FIT_BAD_START, optimise.py:226-227)."""
    import torch
    g = load_golden("fit.npz")
    k, good = g["w200_k"], g["w200_phi"]
    rng = np.random.default_rng(11)
    rows = [good[i % good.shape[0]].copy() for i in range(64)]
    rows[3][:] = np.nan
    rows[7][10] = np.inf
    rows[11][:] = 1e300
    rows[13][:] = -1e300
    rows[17][:] = 5.0
    rows[19][::2] = -np.inf
    rows[23] = rng.standard_normal(k.size) * 1e6
    rows[29] = rng.standard_normal(k.size) * 1e-300
    phi = np.ascontiguousarray(np.stack(rows))
    x, d = grid._fit_device(k, torch.from_numpy(phi).cuda(), (3.0, 1.0, 10.0, 5.0), diagnostics=True)
    torch.cuda.synchronize()
    x, st, it, nf = x.cpu().numpy(), d["status"].cpu().numpy(), d["niter"].cpu().numpy(), d["nfev"].cpu().numpy()
    assert np.all((st >= 0) & (st <= 4)) and np.all(it <= 200) and np.all(nf <= 3 * 200 * 41 + 3)
    for bad in (3, 7, 19):                                   # non-finite objective at the start
        assert st[bad] == 4 and np.allclose(x[bad], (3.0, 1.0, 10.0, 5.0))
    ref, _ = grid._fit_device(k, torch.from_numpy(np.ascontiguousarray(good)).cuda(), (3.0, 1.0, 10.0, 5.0), diagnostics=True)
    ref = ref.cpu().numpy()
    hostile = {3, 7, 11, 13, 17, 19, 23, 29}
    for i in range(64):
        if i not in hostile:
            assert np.array_equal(x[i], ref[i % good.shape[0]]), i
    assert np.all(np.isfinite(x[[j for j in range(64) if j not in (3, 7, 19)]]))


def test_fft_domain_filters_match_reference(grid, oracle, mag):
    """CurieGrid.upward_continuation / reduce_to_pole (grid.py:405-546): FFTs on the device, golden outputs of
    the reference itself; also as a `process_subgrid` callback of optimise (optimise.py:282-294)."""
    g = load_golden("filters.npz")
    d, _ = mag
    for name in g["names"]:
        j0, i0, n, st = (int(v) for v in g[str(name) + "_meta"])
        win = d[j0:j0 + n, i0:i0 + n]
        for tag, fn in (("up500", lambda w: grid.upward_continuation(w, 500.0)),
                        ("up5000", lambda w: grid.upward_continuation(w, 5000.0)),
                        ("rtp_induced", lambda w: grid.reduce_to_pole(w, -60.0, 12.0)),
                        ("rtp_remanent", lambda w: grid.reduce_to_pole(w, 35.0, -5.0, sinc=50.0, sdec=20.0))):
            ref = g["%s_%s" % (name, tag)]
            out = fn(win)
            assert out.shape == win.shape
            np.testing.assert_allclose(out[::st, ::st], ref, rtol=0, atol=1e-11 * np.abs(ref).max(), err_msg=str((name, tag)))
    # rectangular input warns like the reference and still works
    with pytest.warns(RuntimeWarning):
        out = grid.upward_continuation(d[:64, :96], 1000.0)
    ref = oracle.fft_filter(d[:64, :96], oracle.upward_continuation_filter((64, 96), grid.dx, grid.dy, 1000.0))
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-11 * np.abs(ref).max())
    # as a process_subgrid callback
    x = grid.optimise(150e3, 152.5e3, 140.5e3, process_subgrid=lambda w: grid.upward_continuation(w, 2000.0))
    assert x.shape == (4,) and np.all(np.isfinite(x))


def test_filter_stage_recognised_in_batches(grid, oracle, mag, monkeypatch):
    """process_subgrid=functools.partial(grid.upward_continuation / reduce_to_pole, ...) runs for the whole
    centroid batch on the device (cb_fft2_filter_windows -> spectrum): same spectra as the reference sequence
    subgrid -> filter -> radial_spectrum per window, same fit as the host-callback path, sliced or not."""
    import functools
    d, ext = mag
    spec = oracle.GridSpec(d, *ext)
    xc, yc = grid.create_centroid_list(100e3, 25e3, 30e3)
    stages = [
        (functools.partial(grid.upward_continuation, height=1500.0),
         lambda n: oracle.upward_continuation_filter((n, n), grid.dx, grid.dy, 1500.0)),
        (functools.partial(grid.reduce_to_pole, inc=-60.0, dec=12.0, sdec=20.0, sinc=50.0),
         lambda n: oracle.reduce_to_pole_filter((n, n), grid.dx, grid.dy, -60.0, 12.0, 50.0, 20.0)),
    ]
    for stage, make_filter in stages:
        assert grid._window_filter(stage, 101) is not None
        k, phi, sig = grid._batch_spectra(100e3, xc, yc, np.hanning, stage, {})
        phi, sig = phi.cpu().numpy(), sig.cpu().numpy()
        for i in (0, 7, xc.size - 1):
            win = oracle.fft_filter(oracle.cut_window(spec, 100e3, xc[i], yc[i]), make_filter(101))
            k_o, phi_o, sig_o = oracle.radial_spectrum(win, spec.dx)
            assert np.array_equal(k, k_o)
            np.testing.assert_allclose(phi[i], phi_o, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(sig[i], sig_o, rtol=1e-8, atol=1e-9)
        # the same callable hidden in a lambda takes the host path: identical windows, identical fits
        x_dev = grid.optimise_batch(100e3, xc[:6], yc[:6], process_subgrid=stage)
        x_host = grid.optimise_batch(100e3, xc[:6], yc[:6], process_subgrid=lambda w: stage(w))
        np.testing.assert_allclose(x_dev, x_host, rtol=1e-6, atol=1e-8)
        # slicing by scratch budget does not change anything
        monkeypatch.setattr(type(grid), "FILTER_CHUNK_BYTES", 101 * 101 * 40 * 5)
        _, phi2, _ = grid._batch_spectra(100e3, xc, yc, np.hanning, stage, {})
        monkeypatch.undo()
        assert np.array_equal(phi2.cpu().numpy(), phi)
    m = grid.curie_depth_maps(100e3, 50e3, 50e3, process_subgrid=stages[0][0])
    xq, yq = grid.create_centroid_list(100e3, 50e3, 50e3)
    ref = grid.optimise_batch(100e3, xq, yq, process_subgrid=stages[0][0])
    assert np.array_equal(m["dz"].ravel(), ref[:, 2])
    # partials that are not the window-only form fall back to the host path
    assert grid._window_filter(functools.partial(grid.upward_continuation), 101) is None
    assert grid._window_filter(functools.partial(grid.reduce_to_pole, -60.0, 12.0), 101) is None   # window would be `dec`
    assert grid._window_filter(lambda w: w, 101) is None


@pytest.mark.parametrize("wkm", [50e3, 128e3, 150e3, 160e3])
def test_filter_stage_window_sizes(grid, oracle, mag, wkm):
    """The device filter stage on every line-engine path: 51 nodes (generic Bluestein), 129 and 161 (N/2 + 1 nodes on
    the 256- and 320-point static engines, PlusOneFft), 151 (static engine, pruned passes) -- spectra of the filtered
    windows against the reference sequence subgrid -> upward_continuation -> radial_spectrum."""
    import functools
    d, ext = mag
    spec = oracle.GridSpec(d, *ext)
    xc, yc = grid.create_centroid_list(wkm, 31e3, 37e3)
    n = int(round(wkm / grid.dx)) + 1
    stage = functools.partial(grid.upward_continuation, height=1200.0)
    assert grid._window_filter(stage, n) is not None
    k, phi, sig = grid._batch_spectra(wkm, xc, yc, np.hanning, stage, {})
    phi, sig = phi.cpu().numpy(), sig.cpu().numpy()
    filt = oracle.upward_continuation_filter((n, n), grid.dx, grid.dy, 1200.0)
    for i in (0, xc.size // 2, xc.size - 1):
        win = oracle.fft_filter(oracle.cut_window(spec, wkm, xc[i], yc[i]), filt)
        k_o, phi_o, sig_o = oracle.radial_spectrum(win, spec.dx)
        assert np.array_equal(k, k_o)
        np.testing.assert_allclose(phi[i], phi_o, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(sig[i], sig_o, rtol=1e-8, atol=1e-9)


def test_tanaka_batch_matches_reference(pc, grid, oracle, mag):
    """CurieGrid.tanaka_batch = radial_spectrum(power=0.5) -> tanaka1999 -> ComputeTanaka per centroid."""
    d, ext = mag
    spec = oracle.GridSpec(d, *ext)
    xc, yc = grid.create_centroid_list(200e3, 40e3, 40e3)
    out = grid.tanaka_batch(200e3, xc, yc, kmin_range=(0.005, 0.03), kmax_range=(0.03, 0.7))
    assert out.shape == (xc.size, 8)
    for i in range(0, xc.size, 3):
        k, phi, sig = oracle.radial_spectrum(oracle.cut_window(spec, 200e3, xc[i], yc[i]), spec.dx, power=0.5)
        top, cen = oracle.tanaka1999(k, phi, sig, (0.005, 0.03), (0.03, 0.7))
        zb, ezb = oracle.compute_tanaka(top[0], top[2], cen[0], cen[2])
        np.testing.assert_allclose(out[i], [*top, *cen, zb, ezb], rtol=1e-7)


def test_sensitivity_matches_reference_loop(grid, oracle, mag):
    """CurieOptimise.sensitivity (optimise.py:635-725): the device draws N(Phi, sigma_Phi) with its Philox
    stream; the oracle replays the same stream and runs the reference's loop body (L-BFGS-B) on the same
    rows.  No priors here, so the rows are the only randomness."""
    d, ext = mag
    spec = oracle.GridSpec(d, *ext)
    xc, yc = 152.5e3, 140.5e3
    nsim, seed = 12, 424242
    res = grid.sensitivity(200e3, xc, yc, nsim, seed=seed)
    assert isinstance(res, list) and len(res) == 4 and all(r.shape == (nsim,) for r in res)
    k, phi, sig = oracle.radial_spectrum(oracle.cut_window(spec, 200e3, xc, yc), spec.dx)
    rows = oracle.perturbed_spectra_philox(phi, sig, nsim, seed, 0)
    nan4 = np.full(4, np.nan)
    ref = oracle.sensitivity_rows(k, rows, np.tile(nan4, (nsim, 1)), nan4)
    ours = np.stack(res, axis=1)
    good = 0
    for s in range(nsim):
        xp, fp = oracle.polished_fit(k, rows[s], ref[s])
        conv = np.max(np.abs(ref[s] - xp) / np.maximum(np.abs(xp), 1e-3)) < 3e-5
        f_ours = oracle.misfit(ours[s], k, rows[s])
        assert f_ours <= oracle.misfit(ref[s], k, rows[s]) * (1 + 1e-9) + 1e-12
        if conv:
            good += 1
            assert np.max(np.abs(ours[s] - ref[s]) / np.maximum(np.abs(ref[s]), 1e-3)) < 1e-4, (s, ours[s], ref[s])
    assert good >= nsim // 2


def test_sensitivity_priors_and_batch_containers(grid):
    """Prior means are redrawn per simulation (optimise.py:709-711): with a tight prior on beta the spread of
    the fitted beta follows the prior's width; parallelise_routine(self.sensitivity) keeps the reference's
    nested containers; the batched call streams moments."""
    grid.reset_priors()
    grid.add_prior(beta=(3.0, 0.05))
    try:
        np.random.seed(7)
        b, zt, dz, C = grid.sensitivity(200e3, 152.5e3, 140.5e3, 400)
        assert abs(np.mean(b) - 3.0) < 0.03 and 0.02 < np.std(b) < 0.12
        out = grid.parallelise_routine(200e3, [152.5e3, 150.5e3], [140.5e3, 149.5e3], grid.sensitivity, 16)
        assert len(out) == 4 and len(out[0]) == 2 and out[0][0].shape == (16,)
        mom = grid.sensitivity_batch(200e3, [152.5e3, 150.5e3], [140.5e3, 149.5e3], 64, reduce="moments", seed=3)
        assert mom.shape == (2, 8) and np.all(mom[:, 4:] >= 0)
    finally:
        grid.reset_priors()
    assert grid.prior["beta"] is None


# --------------------------------------------------------------------------- Metropolis-Hastings
def test_mh_chain_matches_reference_stream(grid):
    """Same Philox stream as the monkey-patched reference run: chains coincide step by step."""
    g = load_golden("mh.npz")
    seed = int(g["seed"])
    for ci in range(2):
        xc, yc = float(g["xc"][ci]), float(g["yc"][ci])
        chain = np.array(grid.metropolis_hastings_batch(
            100e3, [xc], [yc], 3000, 200, x_scale=g["x_scale"], beta=3.0, zt=0.5, dz=10.0, C=-19.0, seed=seed,
            reduce="chains", chain_offset=ci))[0]
        want = g["philox_noprior_%d" % ci]
        same = np.all(np.abs(chain - want) < 1e-9, axis=0)
        first_bad = same.size if same.all() else int(np.argmin(same))
        assert first_bad >= 2500, first_bad      # an accept/reject flip from a 1-ulp difference is tolerated late
    grid.add_prior(beta=(3.0, 1.0), zt=(1.0, 1.0), dz=(20.0, 10.0), C=(-19.0, 5.0))
    try:
        res = grid.metropolis_hastings(100e3, float(g["xc"][0]), float(g["yc"][0]), 3000, 200, x_scale=g["x_scale"],
                                       beta=3.0, zt=0.5, dz=10.0, C=-19.0, seed=seed)
    finally:
        grid.reset_priors()
    assert isinstance(res, list) and len(res) == 4 and res[0].shape == (3000,)
    want = g["philox_prior_0"]
    same = np.all(np.abs(np.array(res) - want) < 1e-9, axis=0)
    assert (same.size if same.all() else int(np.argmin(same))) >= 2500


def test_mh_moments_statistical_parity(grid):
    """Posterior means / stds of independent device chains agree with the reference's own chains
    (numpy RNG) within Monte-Carlo error."""
    g = load_golden("mh.npz")
    ref = g["ref_moments"]            # (6 chains, [mean,std], 5)
    grid.add_prior(beta=(3.0, 1.0), zt=(1.0, 1.0), dz=(20.0, 10.0), C=(-19.0, 5.0))
    try:
        n = 24
        xc, yc = np.full(n, float(g["xc"][0])), np.full(n, float(g["yc"][0]))
        mom = grid.metropolis_hastings_batch(100e3, xc, yc, int(g["ref_nsim"]), 1000, x_scale=g["x_scale"],
                                             beta=3.0, zt=0.5, dz=10.0, C=-19.0, seed=99, reduce="moments")
    finally:
        grid.reset_priors()
    dev_mean, dev_std = mom[:, :5], mom[:, 5:]
    for q in range(5):
        for dev, rf in ((dev_mean[:, q], ref[:, 0, q]), (dev_std[:, q], ref[:, 1, q])):
            se = np.sqrt(dev.var(ddof=1) / dev.size + rf.var(ddof=1) / rf.size)
            assert abs(dev.mean() - rf.mean()) < 5.0 * se + 1e-12, (q, dev.mean(), rf.mean(), se)
    # moments mode == moments of the chains mode for the same seed
    ch = grid.metropolis_hastings_batch(100e3, xc[:2], yc[:2], 2000, 100, x_scale=g["x_scale"], beta=3.0, zt=0.5,
                                        dz=10.0, C=-19.0, seed=5, reduce="chains")
    mo = grid.metropolis_hastings_batch(100e3, xc[:2], yc[:2], 2000, 100, x_scale=g["x_scale"], beta=3.0, zt=0.5,
                                        dz=10.0, C=-19.0, seed=5, reduce="moments")
    full = np.concatenate([ch, (ch[:, 1] + ch[:, 2])[:, None, :]], axis=1)
    np.testing.assert_allclose(mo[:, :5], full.mean(axis=2), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(mo[:, 5:], full.std(axis=2), rtol=1e-8, atol=1e-10)
    assert not np.array_equal(ch[0], ch[1])        # distinct chain ids -> distinct streams


def test_parallelise_routine_containers(grid):
    xc, yc = grid.create_centroid_list(100e3, 100e3, 100e3)
    out = grid.parallelise_routine(100e3, xc, yc, grid.metropolis_hastings, 50, 10)
    assert len(out) == 4 and len(out[0]) == xc.size and out[0][0].shape == (50,)
    out = grid.parallelise_routine(100e3, xc[:3], yc[:3], lambda w, x, y: np.array([x, y, w, 0.0]))
    assert len(out) == 4 and np.array_equal(out[0], xc[:3])
    with pytest.raises(ValueError, match="same size"):
        grid.parallelise_routine(100e3, xc, yc[:-1], grid.optimise)
    with pytest.raises(ValueError, match="outside data range"):
        grid.optimise_routine(100e3, [-5.0], [150e3])
    with pytest.raises(ValueError, match="exceeds the data range"):
        grid.optimise_routine(100e3, [10e3], [150e3])


def test_curie_depth_maps_equal_notebook_workflow(grid, monkeypatch):
    """curie_depth_maps (lattice expanded on the device, maps as views of the solver output) against the
    reference's notebook sequence create_centroid_list -> optimise_routine -> reshape -> zt + dz, and its
    two uncertainty modes against the batched sensitivity / Metropolis-Hastings calls with the same streams."""
    xc, yc = grid.create_centroid_list(100e3, 40e3, 50e3)
    beta, zt, dz, C = grid.optimise_routine(100e3, xc, yc)
    m = grid.curie_depth_maps(100e3, 40e3, 50e3)
    rows, cols = m["yc"].size, m["xc"].size
    assert rows * cols == xc.size and m["beta"].shape == (rows, cols)
    for name, ref in (("beta", beta), ("zt", zt), ("dz", dz), ("C", C), ("curie_depth", zt + dz)):
        assert np.array_equal(m[name], ref.reshape(rows, cols)), name
    assert np.array_equal(np.meshgrid(m["xc"], m["yc"])[0].ravel(), xc)

    ms = grid.curie_depth_maps(100e3, 40e3, 50e3, uncertainty="sensitivity", nsim=24, seed=11)
    smp = grid.sensitivity_batch(100e3, xc, yc, 24, seed=11)                 # (n, nsim, 4)
    np.testing.assert_allclose(ms["dz_std"].ravel(), smp[:, :, 2].std(axis=1), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ms["curie_depth_std"].ravel(), (smp[:, :, 1] + smp[:, :, 2]).std(axis=1),
                               rtol=1e-12, atol=1e-14)
    assert np.array_equal(ms["beta"], m["beta"])
    monkeypatch.setattr(type(grid), "SENSITIVITY_ROWS", 24 * 7)              # seven windows per slice: same streams
    ms2 = grid.curie_depth_maps(100e3, 40e3, 50e3, uncertainty="sensitivity", nsim=24, seed=11)
    monkeypatch.undo()
    for name in ("beta_std", "zt_std", "dz_std", "C_std", "curie_depth_std"):
        assert np.array_equal(ms2[name], ms[name]), name

    grid.add_prior(beta=(3.0, 1.0), zt=(1.0, 1.0), dz=(20.0, 10.0), C=(-20.0, 5.0))
    try:
        x_map = grid.optimise_batch(100e3, xc, yc)
        mm = grid.curie_depth_maps(100e3, 40e3, 50e3, uncertainty="mcmc", nsim=400, burnin=50,
                                   x_scale=[0.25, 0.1, 1.0, 0.5], seed=5)
        mom = grid.metropolis_hastings_batch(100e3, xc, yc, 400, 50, x_scale=[0.25, 0.1, 1.0, 0.5], x0=x_map, seed=5)
        assert np.array_equal(mm["curie_depth_mean"].ravel(), mom[:, 4])
        assert np.array_equal(mm["curie_depth_std"].ravel(), mom[:, 9])
        assert np.all(mm["beta_std"] > 0)
    finally:
        grid.reset_priors()
    with pytest.raises(ValueError, match="uncertainty"):
        grid.curie_depth_maps(100e3, 40e3, 50e3, uncertainty="bootstrap")


# --------------------------------------------------------------------------- Radon / azimuthal
def test_radon_csr_bit_exact_and_sinogram(pc, grid):
    g = load_golden("radon.npz")
    for nw in (51, 101):
        j0, i0, _ = (int(v) for v in g["n%d_meta" % nw])
        theta = g["n%d_theta" % nw]
        data, ind, ptr = pc.radon.radon_csr(nw, theta)
        assert ind.dtype == np.int64 and ptr.dtype == np.int64
        assert np.array_equal(ptr, g["n%d_indptr" % nw])
        off = np.concatenate([[0], np.cumsum(ptr[:, -1])])
        for t in range(theta.size):
            key = "n%d_t%d_data" % (nw, t)
            seg = slice(off[t], off[t + 1])
            chk = np.add.reduceat(ind[seg] * (1 + np.arange(ind[seg].size) % 7), ptr[t, :-1])
            assert np.array_equal(chk, g["n%d_index_checksum" % nw][t])
            assert np.array_equal(np.add.reduceat(data[seg], ptr[t, :-1]), g["n%d_length_sum" % nw][t])
            if key in g.files:
                assert np.array_equal(data[seg], g[key]) and np.array_equal(ind[seg], g["n%d_t%d_indices" % (nw, t)])
        win = np.ascontiguousarray(grid.data[j0:j0 + nw, i0:i0 + nw])
        sino = pc.radon.radon2d(win, theta)
        assert np.array_equal(sino, g["n%d_sinogram" % nw])
        kb, S, dth = grid.azimuthal_spectrum(grid.data[j0:j0 + nw, i0:i0 + nw])
        assert np.array_equal(kb, g["n%d_az_k" % nw]) and np.array_equal(dth, g["n%d_az_theta" % nw])
        np.testing.assert_allclose(S, g["n%d_az_S" % nw], rtol=1e-10)
    with pytest.raises(ValueError, match="theta should be within"):
        pc.radon.radon2d(np.zeros((8, 8)), np.array([np.pi]))


# --------------------------------------------------------------------------- the reference's own tests
def test_reference_suite_behaviour(pc, grid, mag):
    """tests/test_grid.py and tests/test_optimise.py of the reference, against the new classes."""
    d, (xmin, xmax, ymin, ymax) = mag
    xc, yc = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax)
    G = pc.CurieGrid(d, xmin, xmax, ymin, ymax)
    sub = G.subgrid(100e3, xc, yc)
    assert sub.shape[0] < G.data.shape[0] and sub.shape[1] < G.data.shape[1]
    k, Phi, sPhi = G.radial_spectrum(G.data, taper=None)               # test_FFT
    i3 = len(k) // 3
    assert Phi[:i3].mean() > Phi[i3:2 * i3].mean() > Phi[2 * i3:].mean()
    assert sPhi[:i3].mean() > sPhi[i3:2 * i3].mean() > sPhi[2 * i3:].mean()
    _, Phi2, _ = G.radial_spectrum(G.data, taper=np.hanning)           # test_taper_functions
    _, Phi3, _ = G.radial_spectrum(G.data, taper=np.hamming)
    assert Phi.mean() > Phi2.mean() and Phi.mean() > Phi3.mean()
    assert np.gradient(Phi, k).mean() > np.gradient(Phi2, k).mean()
    k, P, sP = G.radial_spectrum(G.data, taper=np.hanning, power=0.5)  # test_Tanaka
    (Ztr, btr, dZtr), (Zor, bor, dZor) = pc.tanaka1999(k, P, sP, (0.005, 0.03), (0.03, 0.7))
    Zb, eZb = pc.ComputeTanaka(Ztr, dZtr, Zor, dZor)
    assert abs(Zb - 10.0) < 2.0 and eZb < Zb
    np.testing.assert_allclose([Ztr, btr, dZtr, Zor, bor, dZor, Zb, eZb], load_golden("misc.npz")["tanaka"], rtol=1e-7)
    beta, zt, dz, C = grid.optimise(300e3, xc, yc, taper=np.hanning)   # test_optimisation
    assert abs(beta - 3.0) < 0.3 and abs(zt - 0.305) < 0.1 and abs(dz - 10.305) < 2.0
    grid.add_prior(beta=(1.0, 0.1))                                    # test_priors
    try:
        beta1 = grid.optimise(300e3, xc, yc)[0]
    finally:
        grid.reset_priors()
    assert abs(beta1 - 1.0) < abs(beta - 1.0)


# --------------------------------------------------------------------------- C-ABI error behaviour
def test_c_abi_rejects_bad_arguments(pc, grid):
    """Every entry point returns a negative cb_status and sets cb_last_error instead of faulting: NULL buffers,
    negative counts, a workspace that is too small, inconsistent bounds, a window larger than the grid."""
    import ctypes as C
    import torch
    from pycurious_b200 import _lib, _plans
    lib = _lib.load()
    plan = _plans.spectrum_plan(101, grid.dx, _plans.taper_vector(np.hanning, 101))
    g = grid.device_grid()
    iy0 = torch.zeros(2, dtype=torch.int32, device="cuda")
    phi = torch.empty((2, plan.nbins), dtype=torch.float64, device="cuda")
    ws = plan.workspace
    args = lambda **kw: None
    # workspace too small
    rc = lib.cb_spectrum_batch(plan.handle, g.data_ptr(), g.shape[0], g.shape[1], g.stride(0), iy0.data_ptr(), iy0.data_ptr(),
                               2, 0, 2.0, phi.data_ptr(), phi.data_ptr(), ws.data_ptr(), 1024, None)
    assert rc == -1 and b"workspace" in lib.cb_last_error()
    # NULL output
    rc = lib.cb_spectrum_batch(plan.handle, g.data_ptr(), g.shape[0], g.shape[1], g.stride(0), iy0.data_ptr(), iy0.data_ptr(),
                               2, 0, 2.0, None, phi.data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert rc == -1
    # grid smaller than the window
    rc = lib.cb_spectrum_batch(plan.handle, g.data_ptr(), 50, 50, 50, iy0.data_ptr(), iy0.data_ptr(),
                               2, 0, 2.0, phi.data_ptr(), phi.data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert rc == -1
    # fit: lower bound above upper bound, negative count, NULL spectra
    o = _lib.FitOptions()
    for j in range(4):
        o.x0[j], o.lb[j], o.ub[j], o.prior_mu[j], o.prior_sigma[j] = 1.0, 2.0, 1.0, float("nan"), float("nan")
    k = torch.linspace(0.1, 3.0, plan.nbins, dtype=torch.float64, device="cuda")
    x = torch.empty((2, 4), dtype=torch.float64, device="cuda")
    rc = lib.cb_fit_batch(k.data_ptr(), plan.nbins, phi.data_ptr(), 2, C.byref(o), None, x.data_ptr(), None, None, None, None, None)
    assert rc == -1 and b"bound" in lib.cb_last_error()
    for j in range(4):
        o.lb[j], o.ub[j] = 0.0, float("inf")
    assert lib.cb_fit_batch(k.data_ptr(), plan.nbins, phi.data_ptr(), -1, C.byref(o), None, x.data_ptr(), None, None, None, None, None) == -1
    assert lib.cb_fit_batch(k.data_ptr(), plan.nbins, None, 2, C.byref(o), None, x.data_ptr(), None, None, None, None, None) == -1
    assert lib.cb_fit_batch(k.data_ptr(), 0, phi.data_ptr(), 2, C.byref(o), None, x.data_ptr(), None, None, None, None, None) == -1
    # zero windows is a no-op
    assert lib.cb_fit_batch(k.data_ptr(), plan.nbins, phi.data_ptr(), 0, C.byref(o), None, x.data_ptr(), None, None, None, None, None) == 0
    # MH: bad reduce mode / negative chain length
    m = _lib.MhOptions()
    m.nsim, m.burnin, m.reduce_mode = 10, 1, 7
    out = torch.empty((2, 10), dtype=torch.float64, device="cuda")
    assert lib.cb_mh_batch(k.data_ptr(), plan.nbins, phi.data_ptr(), 2, C.byref(m), None, out.data_ptr(), None, None) == -1
    m.reduce_mode, m.nsim = 1, -5
    assert lib.cb_mh_batch(k.data_ptr(), plan.nbins, phi.data_ptr(), 2, C.byref(m), None, out.data_ptr(), None, None) == -1
    # plan creation: taper given for one axis only, nonsense bins
    h = C.c_void_p()
    t = np.hanning(21)
    ptr = np.array([0, 1], dtype=np.int32)
    idx = np.array([5000], dtype=np.int32)
    assert lib.cb_spectrum_plan_create(21, 1, t.ctypes.data, None, ptr.ctypes.data, idx.ctypes.data, C.byref(h)) == -1
    assert lib.cb_spectrum_plan_create(21, 1, None, None, ptr.ctypes.data, idx.ctypes.data, C.byref(h)) == -4
    torch.cuda.synchronize()
    # the Python face turns statuses into exceptions and keeps the reference's ValueErrors
    with pytest.raises(ValueError, match="outside data range"):
        grid.optimise_routine(100e3, [1e9], [1e9])
    with pytest.raises(ValueError, match="exceeds the data range"):
        grid.optimise_routine(300e3, [grid.xmin + 1e3], [grid.ymin + 1e3])
    with pytest.raises(ValueError, match="same size"):
        grid.optimise_routine(100e3, [150e3, 160e3], [150e3])


# --------------------------------------------------------------------------- the reference's smoke tests
def test_reference_routines_smoke(pc, mag):
    """tests/test_routines.py:22-59 of the reference: every routine once, in one object lifetime."""
    d, (xmin, xmax, ymin, ymax) = mag
    xc, yc = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax)
    max_window = 300e3
    cpd = pc.CurieGrid(d, xmin, xmax, ymin, ymax)
    assert cpd.subgrid(max_window, xc, yc).shape == (301, 301)
    assert cpd.create_centroid_list(0.1 * max_window)[0].size > 0
    assert cpd.remove_trend_linear(cpd.data).shape == d.shape
    k, Phi, sPhi = cpd.radial_spectrum(cpd.data)
    assert k.shape == Phi.shape == sPhi.shape and np.all(np.isfinite(Phi))
    kb, S, th = cpd.azimuthal_spectrum(cpd.data)
    assert S.shape == (th.size, kb.size) and np.all(np.isfinite(S))
    assert np.all(np.isfinite(cpd.reduce_to_pole(cpd.data, 30.0, 30.0)))
    assert np.all(np.isfinite(cpd.upward_continuation(cpd.data, 10e3)))

    cpo = pc.CurieOptimise(d, xmin, xmax, ymin, ymax)
    xc_list, yc_list = cpo.create_centroid_list(0.5 * max_window, spacingX=5e3, spacingY=5e3)
    assert xc_list.size == 961
    cpo.add_prior(beta=(3.0, 0.5))
    x = cpo.optimise(0.5 * max_window, xc_list[0], yc_list[0])
    assert x.shape == (4,) and np.all(np.isfinite(x))
    res = cpo.optimise_routine(0.5 * max_window, xc_list, yc_list)
    assert len(res) == 4 and all(r.shape == (961,) and np.all(np.isfinite(r)) for r in res)
    np.testing.assert_allclose([r[0] for r in res], x, rtol=1e-9)
    mh = cpo.metropolis_hastings(0.5 * max_window, xc_list[0], yc_list[0], 100, 10)
    assert len(mh) == 4 and all(c.shape == (100,) and np.all(np.isfinite(c)) for c in mh)
    sens = cpo.sensitivity(0.5 * max_window, xc_list[0], yc_list[0], 100)
    assert len(sens) == 4 and all(c.shape == (100,) and np.all(np.isfinite(c)) for c in sens)


def test_min_func_under_scipy_tnc(grid, oracle):
    """tests/test_optimise.py:63-208 of the reference: its hard-coded 49-point spectrum, `min_func` driven by
    scipy's TNC from the host (one device evaluation per call).  Goldens: the reference itself on the same
    vectors (tests/golden/make_golden_tnc.py)."""
    from scipy.optimize import minimize
    g = load_golden("tnc.npz")
    k, S, x0 = g["k"], g["S"], g["x0"]
    sigma_S = np.ones_like(S)
    np.testing.assert_allclose(grid.min_func(x0, k, S, sigma_S), float(g["f0"]), rtol=1e-12)
    np.testing.assert_allclose(grid.min_func(g["x_tnc"], k, S, sigma_S), float(g["f_tnc"]), rtol=1e-10)
    np.testing.assert_allclose(grid.min_func(g["x_lbfgsb"], k, S, sigma_S), float(g["f_lbfgsb"]), rtol=1e-10)
    res = minimize(grid.min_func, x0, args=(k, S, sigma_S), method="TNC",
                   bounds=list(zip(np.zeros_like(x0), [None] * 4)))
    assert np.all(np.isfinite(res.x)), res.x
    assert res.fun < 2.0                      # from 3545 at x0 into the valley the reference reaches (0.75 - 0.76)
    # the device solver on the same spectrum: never worse than either reference optimiser
    import torch
    x = grid._fit_device(k, torch.from_numpy(S[None, :].copy()).cuda(), tuple(x0)).cpu().numpy()[0]
    f_us = oracle.misfit(x, k, S)
    assert f_us <= float(g["f_lbfgsb"]) * (1 + 1e-9) and f_us <= float(g["f_tnc"])
    assert x[1] == 0.0                        # zt pinned on its bound, like the reference's result


# --------------------------------------------------------------------------- live data, residency, several GPUs
def test_in_place_edits_of_data_are_seen(pc, mag):
    """The reference reads `grid.data` live (grid.py:142-155): in-place edits between two calls must show."""
    d, ext = mag
    G = pc.CurieOptimise(d.copy(), *ext)
    xc, yc = G.create_centroid_list(100e3, 50e3, 50e3)
    k, phi1, sig1 = G.radial_spectrum_batch(100e3, xc, yc)
    x1 = np.stack(G.optimise_routine(100e3, xc, yc), axis=1)
    keep = G.data
    G.data *= 2.0                                        # same ndarray object
    assert G.data is keep
    _, phi2, sig2 = G.radial_spectrum_batch(100e3, xc, yc)
    np.testing.assert_allclose(phi2, phi1 + 2.0 * np.log(2.0), rtol=1e-12)
    np.testing.assert_allclose(sig2, sig1, rtol=1e-9)
    x2 = np.stack(G.optimise_routine(100e3, xc, yc), axis=1)
    np.testing.assert_allclose(x2[:, 3], x1[:, 3] + 2.0 * np.log(2.0), rtol=1e-5)     # C absorbs the scale
    np.testing.assert_allclose(x2[:, :3], x1[:, :3], rtol=1e-3, atol=1e-6)
    s = G.subgrid(100e3, xc[3], yc[3])                   # a view: edits through it show in both directions
    _, p_a, _ = G.radial_spectrum(s)
    s *= 0.5
    _, p_b, _ = G.radial_spectrum(s)
    np.testing.assert_allclose(p_b, p_a - 2.0 * np.log(2.0), rtol=1e-12)
    # opt-in residency: a held copy is a snapshot until it is invalidated
    G.hold_device()
    _, h1, _ = G.radial_spectrum_batch(100e3, xc, yc)
    G.data += 1000.0
    _, h2, _ = G.radial_spectrum_batch(100e3, xc, yc)
    assert np.array_equal(h1, h2)
    G.invalidate_device()
    _, h3, _ = G.radial_spectrum_batch(100e3, xc, yc)
    assert not np.array_equal(h1, h3)
    G.hold_device(False)
    _, h4, _ = G.radial_spectrum_batch(100e3, xc, yc)
    assert np.array_equal(h3, h4)


def test_region_upload_equals_resident_grid(pc, mag):
    """Row-band / rectangle uploads (the default) give the same bits as the resident whole grid."""
    d, ext = mag
    G = pc.CurieOptimise(d, *ext)
    H = pc.CurieOptimise(d, *ext).hold_device()
    cases = [G.create_centroid_list(100e3, 10e3, 10e3),                      # lattice: whole-width band
             (np.array([60.5e3, 70e3, 81e3]), np.array([250e3, 240e3, 251e3])),   # a corner: rectangle
             (np.array([152e3]), np.array([152e3]))]                              # single window
    for xc, yc in cases:
        a = G.radial_spectrum_batch(100e3, xc, yc)
        b = H.radial_spectrum_batch(100e3, xc, yc)
        assert all(np.array_equal(u, v) for u, v in zip(a, b))
        assert np.array_equal(G.optimise_batch(100e3, xc, yc), H.optimise_batch(100e3, xc, yc))
    m1 = G.curie_depth_maps(100e3, 20e3, 20e3)
    m2 = H.curie_depth_maps(100e3, 20e3, 20e3)
    assert all(np.array_equal(m1[key], m2[key]) for key in m1)


def test_max_processors_fans_out_over_gpus(pc, mag):
    """`CurieOptimise(..., max_processors=N)` (optimise.py:116) spreads a centroid list over min(N, visible GPUs)
    devices, one host thread each; results are identical to the one-GPU run, in reference order."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs at least two visible GPUs")
    d, ext = mag
    one = pc.CurieOptimise(d, *ext, max_processors=1)
    many = pc.CurieOptimise(d, *ext, max_processors=ndev)
    many.MIN_WINDOWS_PER_DEVICE = 8
    xc, yc = one.create_centroid_list(100e3, 10e3, 10e3)
    r1, rn = one.optimise_routine(100e3, xc, yc), many.optimise_routine(100e3, xc, yc)
    assert all(np.array_equal(a, b) for a, b in zip(r1, rn))
    for g in (one, many):
        g.add_prior(beta=(3.0, 1.0), dz=(20.0, 10.0))
    m1 = one.metropolis_hastings_batch(100e3, xc, yc, 300, 50, x_scale=[0.25, 0.1, 1.0, 0.5], seed=7)
    mn = many.metropolis_hastings_batch(100e3, xc, yc, 300, 50, x_scale=[0.25, 0.1, 1.0, 0.5], seed=7)
    assert np.array_equal(m1, mn) and np.array_equal(one.last_accepted, many.last_accepted)
    c1 = one.curie_depth_maps(100e3, 10e3, 10e3, uncertainty="mcmc", nsim=200, burnin=50, seed=3)
    cn = many.curie_depth_maps(100e3, 10e3, 10e3, uncertainty="mcmc", nsim=200, burnin=50, seed=3)
    assert all(np.array_equal(c1[key], cn[key]) for key in c1)
    ch = many.parallelise_routine(100e3, xc[:40], yc[:40], many.metropolis_hastings, 50, 10, seed=5)
    assert len(ch) == 4 and len(ch[0]) == 40 and ch[0][0].shape == (50,)


def test_two_streams_do_not_share_scratch(pc, mag):
    """Spectrum plans are cached per (size, taper, device); their scratch is per CUDA stream, so batches issued
    under two torch.cuda.stream contexts may overlap on the device."""
    import torch
    d, ext = mag
    G = pc.CurieOptimise(d, *ext)
    xa, ya = G.create_centroid_list(100e3, 4e3, 4e3)
    xb, yb = xa[::-1].copy(), ya[::-1].copy()
    _, pa, sa = G.radial_spectrum_batch(100e3, xa, ya)
    _, pb, sb = G.radial_spectrum_batch(100e3, xb, yb, detrend=True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s1):
            _, qa, ta = G.radial_spectrum_batch(100e3, xa, ya, as_numpy=False)
        with torch.cuda.stream(s2):
            _, qb, tb = G.radial_spectrum_batch(100e3, xb, yb, detrend=True, as_numpy=False)
        torch.cuda.synchronize()
        assert np.array_equal(qa.cpu().numpy(), pa) and np.array_equal(ta.cpu().numpy(), sa)
        assert np.array_equal(qb.cpu().numpy(), pb) and np.array_equal(tb.cpu().numpy(), sb)
