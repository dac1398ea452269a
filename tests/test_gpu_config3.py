"""BASELINE config 3 at its real size on the GPU: synthetic 4096 x 4096 grid, 200 km windows, 10 km stride,
all 152,100 windows through the public API (the shared-row-spectra kernels + fit kernel), compared with
tests/golden/cfg3.npz -- outputs of the UNMODIFIED reference on 512 evenly spaced windows of the same list
(tests/golden/make_golden_cfg3.py) -- and with the oracle run in-test.

What the fit gate can and cannot be (pycurious/optimise.py:300): the reference stops L-BFGS-B on a
forward-difference gradient (h = 1e-8) and factr = 1e7.  On about a quarter of the windows of a fractal grid
the misfit valley is flat in dz and that stopping point is NOT a reproducible function of the input: the
fixture holds the reference re-run on the same spectra perturbed by 1e-13 relative (the size of the
difference between two correct FP64 spectra), and on those windows the reference moves by 1e-2 .. 7e-1
relative against itself.  The gates below are therefore
  (1) spectra: k bit-exact, Phi rtol 1e-10, sigma rtol 1e-9 on every sampled window,
  (2) objective: f(ours) <= f(reference) on every sampled window,
  (3) parameters within 1e-4 of the reference wherever the reference is converged (|x_ref - x_polished| < 3e-5),
  (4) a floor on the fraction of ALL sampled windows within 1e-4 of the reference,
  (5) every window where ours differs from the reference by > 1e-4 is a window where the reference is not at
      its own minimum, and on >= 80 % of them the reference differs from ITSELF by > 1e-4 under the 1e-13
      perturbation,
  (6) ours is the deterministic answer: within 1e-4 of the tightly polished minimum on every window whose
      minimum is well conditioned, and the same to 1e-9 on re-run.
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3), axis=-1)


@pytest.fixture(scope="module")
def cfg3():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import pycurious_b200 as pc
    from pycurious_b200 import synth
    g = load_golden("cfg3.npz")
    n = int(g["grid_n"])
    data = synth.fractal_grid(n, n, seed=int(g["seed"]))
    grid = pc.CurieOptimise(data, *synth.grid_extent(n, n))
    xc, yc = grid.create_centroid_list(float(g["window"]), float(g["stride"]), float(g["stride"]))
    assert xc.size == int(g["nwin_total"]) == 152100
    assert np.array_equal(xc[g["idx"]], g["xc"]) and np.array_equal(yc[g["idx"]], g["yc"])
    return grid, g, xc, yc


def test_config3_spectra_match_reference(cfg3):
    grid, g, xc, yc = cfg3
    k, phi, sig = grid.radial_spectrum_batch(float(g["window"]), xc, yc)      # all 152,100 windows, lattice path
    idx = g["idx"]
    assert np.array_equal(k, g["k"])
    np.testing.assert_allclose(phi[idx], g["phi"], rtol=1e-10)
    np.testing.assert_allclose(sig[idx], g["sig"], rtol=1e-9)
    # scattered-window path (no row sharing) on the same sample
    k2, phi2, sig2 = grid.radial_spectrum_batch(float(g["window"]), g["xc"], g["yc"])
    assert np.array_equal(k2, g["k"])
    np.testing.assert_allclose(phi2, g["phi"], rtol=1e-10)
    np.testing.assert_allclose(sig2, g["sig"], rtol=1e-9)


def test_config3_fit_against_reference(cfg3, oracle):
    grid, g, xc, yc = cfg3
    res = grid.optimise_routine(float(g["window"]), xc, yc)
    assert isinstance(res, list) and len(res) == 4 and all(r.shape == (152100,) for r in res)
    full = np.stack(res, axis=1)
    assert np.all(np.isfinite(full))
    ours = full[g["idx"]]
    xref, xpol, fref, fpol, xpert = g["xref"], g["xpol"], g["fref"], g["fpol"], g["xpert"]
    k, phis = g["k"], g["phi"]
    n = xref.shape[0]

    # (2) never worse than the reference's stopping point, and at the polished minimum's objective value
    f_ours = np.array([oracle.misfit(ours[i], k, phis[i]) for i in range(n)])
    assert np.all(f_ours <= fref * (1 + 1e-9)), np.max((f_ours - fref) / fref)
    assert np.all(f_ours <= fpol * (1 + 1e-9)), np.max((f_ours - fpol) / fpol)

    # (3) 1e-4 wherever the reference is converged
    ref_off = _rel(xref, xpol)
    conv = ref_off < 3e-5
    d_ref = _rel(ours, xref)
    assert conv.sum() >= 0.70 * n
    assert np.all(d_ref[conv] < 1e-4), d_ref[conv].max()

    # (4) floor over ALL windows (the reference sits within 1e-4 of its own minimum on ~80 %)
    frac_all = float(np.mean(d_ref < 1e-4))
    assert frac_all >= 0.72, frac_all

    # (5) the rest: reference not at its minimum, and not reproducible against itself
    off = d_ref >= 1e-4
    assert np.all(ref_off[off] >= 3e-5)
    self_spread = np.max(_rel(xpert, xref[:, None, :]), axis=1)
    assert np.mean(self_spread[off] > 1e-4) >= 0.80, np.mean(self_spread[off] > 1e-4)
    # ... by as much as we differ from it: median |ref - ref'| is within a factor 5 of median |ours - ref|
    assert np.median(self_spread[off]) >= 0.2 * np.median(d_ref[off])

    # (6) ours = the polished minimum wherever the reference found it, and close to it elsewhere (flat valleys)
    d_pol = _rel(ours, xpol)
    assert np.all(d_pol[conv] < 1e-4), d_pol[conv].max()
    assert np.mean(d_pol < 1e-3) >= 0.95, np.mean(d_pol < 1e-3)
    again = np.stack(grid.optimise_routine(float(g["window"]), xc, yc), axis=1)
    np.testing.assert_allclose(again, full, rtol=1e-9, atol=1e-12)


def test_config3_oracle_in_test_subset(cfg3, oracle):
    """The oracle (map_fit + polished_fit) run here on the device's own spectra of 32 sampled windows."""
    grid, g, xc, yc = cfg3
    sel = np.arange(0, g["idx"].size, g["idx"].size // 32)[:32]
    k, phi, _ = grid.radial_spectrum_batch(float(g["window"]), g["xc"][sel], g["yc"][sel])
    ours = grid.optimise_batch(float(g["window"]), g["xc"][sel], g["yc"][sel])
    ngate = 0
    for j in range(sel.size):
        ref = oracle.map_fit(k, phi[j])
        xp, fp = oracle.polished_fit(k, phi[j], ref)
        f_us = oracle.misfit(ours[j], k, phi[j])
        assert f_us <= oracle.misfit(ref, k, phi[j]) * (1 + 1e-9)
        assert f_us <= fp * (1 + 1e-9)
        if _rel(ref, xp) < 3e-5:
            assert _rel(ours[j], ref) < 1e-4
            ngate += 1
    assert ngate >= 20


def test_config3_size_independent_properties_all_windows(cfg3):
    """Properties that hold for every one of the 152,100 windows, checked on all of them (no oracle needed):
    (a) the result is finite and inside the reference's box (beta, zt, dz >= 0, optimise.py:111-114);
    (b) the fit of a window does not depend on the batch it arrives in: two half lists == the whole list, bit for bit;
    (c) amplitude scaling: a grid times a shifts every Phi by 2 ln a (grid.py:281, power = 2), which the model absorbs in
        C alone (grid.py:592) -- (beta, zt, dz) stay and C moves by 2 ln a.  The shifted spectra differ from the
        original ones by rounding (1e-16 relative), so this doubles as the all-window measure of how far a
        rounding-level change of the input moves OUR answer (the reference moves by > 1e-4 on a quarter of the
        windows under a 1e-13 change, see the module docstring)."""
    import pycurious_b200 as pc
    from pycurious_b200 import synth
    grid, g, xc, yc = cfg3
    window = float(g["window"])
    full = grid.optimise_batch(window, xc, yc)
    assert full.shape == (xc.size, 4) and np.all(np.isfinite(full))
    assert np.all(full[:, :3] >= 0.0)
    h = xc.size // 2 + 37                                        # not a multiple of the lattice row length
    halves = np.concatenate([grid.optimise_batch(window, xc[:h], yc[:h]), grid.optimise_batch(window, xc[h:], yc[h:])])
    assert np.array_equal(halves, full)
    a = 4.0
    n = grid.data.shape[0]
    scaled = pc.CurieOptimise(grid.data * a, *synth.grid_extent(n, n)).optimise_batch(window, xc, yc)
    expect = full.copy()
    expect[:, 3] += 2.0 * np.log(a)
    d = _rel(scaled, expect)
    # measured (profiles/r02x_scale_property_config3.json): all 152,100 windows within 1e-4 (max 1.1e-5), 99.99 %
    # within 1e-6, median 2e-10
    assert np.mean(d < 1e-4) >= 0.999, (np.mean(d < 1e-4), np.mean(d < 1e-6))
    assert np.mean(d < 1e-6) >= 0.995, np.mean(d < 1e-6)
    assert np.median(d) < 1e-8, np.median(d)
