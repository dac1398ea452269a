"""Pins the CPU oracle (oracle/) to vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, TAPERS


def test_spectrum_matches_reference(oracle, mag):
    d, (xmin, xmax, ymin, ymax) = mag
    g = load_golden("spectrum.npz")
    spec = oracle.GridSpec(d, xmin, xmax, ymin, ymax)
    for entry in g["names"]:
        name, tp = str(entry).split(":")
        j0, i0, n, det = (int(v) for v in g[name + "_meta"])
        win = d[j0:j0 + n, i0:i0 + n]
        if det:
            win = oracle.detrend_plane(win)
        kw = {}
        if tp == "kaiser":
            kw["beta"] = float(g[name + "_kaiser_beta"])
        k, phi, sig = oracle.radial_spectrum(win, spec.dx, taper=TAPERS[tp], power=float(g[name + "_power"]), **kw)
        # same numpy calls in the same order -> bit identical
        assert np.array_equal(k, g[name + "_k"]), name
        assert np.array_equal(phi, g[name + "_phi"]), name
        assert np.array_equal(sig, g[name + "_sig"]), name


@pytest.mark.parametrize("nw", [61, 100, 101, 125, 128, 129, 151, 201, 301, 305])
def test_bin_membership_bit_exact(oracle, nw):
    g = load_golden("spectrum.npz")
    _, dk, edges = oracle.taper_and_bins(nw, nw, 1000.0, None)
    _, masks = oracle.radial_bin_masks(nw, dk, edges)
    counts = np.array([m.sum() for m in masks])
    assert np.array_equal(counts, g["bins%d_counts" % nw])
    nmemb = np.sum(masks, axis=0).astype(np.int8)
    assert np.array_equal(nmemb, g["bins%d_nmemb" % nw])
    first = np.full((nw, nw), -1, dtype=np.int16)
    for i, m in enumerate(masks):
        first[np.logical_and(m, first < 0)] = i
    assert np.array_equal(first, g["bins%d_first" % nw])


def test_bin_golden_integers_from_survey(oracle):
    # sum(counts), points in 0/1/2 bins
    expect = {101: (8076, 2357, 7612, 232), 129: (13180, 3789, 12524, 328),
              201: (31944, 8997, 30864, 540), 301: (71688, 19921, 69672, 1008)}
    for nw, (tot, n0, n1, n2) in expect.items():
        _, dk, edges = oracle.taper_and_bins(nw, nw, 1000.0, None)
        _, masks = oracle.radial_bin_masks(nw, dk, edges)
        nm = np.sum(masks, axis=0)
        assert (nm.sum(), (nm == 0).sum(), (nm == 1).sum(), (nm == 2).sum()) == (tot, n0, n1, n2)


def test_forward_model_matches_reference(oracle):
    g = load_golden("forward.npz")
    with np.errstate(all="ignore"):
        for p, want in zip(g["params"], g["bouligand"]):
            got = oracle.bouligand2009(g["kh"], *p)
            assert np.array_equal(got, want, equal_nan=True), p
        for p, want in zip(g["pdense"], g["fdense"]):
            assert np.array_equal(oracle.bouligand2009(g["k49"], *p), want, equal_nan=True)
        for p, want in zip(g["params"], g["maus"]):
            assert np.array_equal(oracle.maus1995(p[0], p[1], g["kh"], p[3]), want, equal_nan=True)
    # SURVEY.md 8c golden values
    got = oracle.bouligand2009(np.array([0.05, 0.5, 1, 3]), 3, 1, 10, 5)
    np.testing.assert_allclose(got, [8.292380801142428, 4.286832827341122, 1.9013876157386207,
                                     -4.295836866004333], rtol=1e-14)


def test_fit_matches_reference(oracle):
    g = load_golden("fit.npz")
    for tag in ("cfg1", "w200", "w150_prior", "w300_betaprior"):
        k = g[tag + "_k"]
        mu, sg = g[tag + "_prior_mu"], g[tag + "_prior_sigma"]
        n = g[tag + "_phi"].shape[0]
        for i in range(0, n, max(1, n // 5)):
            x = oracle.map_fit(k, g[tag + "_phi"][i], g[tag + "_x0"], prior_mu=mu, prior_sigma=sg)
            assert np.array_equal(x, g[tag + "_xref"][i]), (tag, i)
            f = oracle.misfit(x, k, g[tag + "_phi"][i], mu, sg)
            assert f == g[tag + "_fref"][i]


def test_fit_end_to_end_config1(oracle, mag):
    d, ext = mag
    g = load_golden("fit.npz")
    spec = oracle.GridSpec(d, *ext)
    for i in (0, 12, 24):
        x = oracle.window_fit(spec, float(g["cfg1_window"]), g["cfg1_xc"][i], g["cfg1_yc"][i])
        assert np.array_equal(x, g["cfg1_xref"][i])
    # row 0 and the centre row
    np.testing.assert_allclose(g["cfg1_xref"][0], (2.569498, 0.399673, 11.427695, -19.911127), atol=2e-6)
    np.testing.assert_allclose(g["cfg1_xref"][12], (2.800522, 0.374367, 5.261940, -19.847312), atol=2e-6)


def test_mh_philox_chain_matches_reference(oracle):
    g = load_golden("mh.npz")
    k, seed = g["k"], int(g["seed"])
    # chain 0 uses the spectrum stored in the file (centre window)
    want = g["philox_prior_0"]
    got = np.array(oracle.mh_chain_philox(k, g["phi"], want.shape[1], 200, seed, 0, x0=g["start"],
                                          x_scale=g["x_scale"], prior_mu=g["prior_mu"],
                                          prior_sigma=g["prior_sigma"]))
    assert np.array_equal(got, want)
    want = g["philox_noprior_0"]
    got = np.array(oracle.mh_chain_philox(k, g["phi"], want.shape[1], 200, seed, 0, x0=g["start"],
                                          x_scale=g["x_scale"]))
    assert np.array_equal(got, want)


def test_philox_known_answer(oracle):
    # Random123 kat_vectors: philox4x32-10
    assert oracle.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_radon_matches_reference_cradon(oracle):
    g = load_golden("radon.npz")
    mag = load_golden("mag_grid.npz")["d"]
    for nw in (51, 101):
        j0, i0, _ = (int(v) for v in g["n%d_meta" % nw])
        win = np.ascontiguousarray(mag[j0:j0 + nw, i0:i0 + nw])
        theta = g["n%d_theta" % nw]
        for t, th in enumerate(theta):
            Tx, Rx = oracle.radon_ray_endpoints(nw, th)
            data, ind, ptr = oracle.radon_csr(Tx, Rx, nw, nw)
            assert np.array_equal(ptr, g["n%d_indptr" % nw][t])
            chk = np.add.reduceat(ind * (1 + np.arange(ind.size) % 7), ptr[:-1])
            assert np.array_equal(chk, g["n%d_index_checksum" % nw][t])
            assert np.array_equal(np.add.reduceat(data, ptr[:-1]), g["n%d_length_sum" % nw][t])
            key = "n%d_t%d_data" % (nw, t)
            if key in g.files:
                assert np.array_equal(data, g[key])
                assert np.array_equal(ind, g["n%d_t%d_indices" % (nw, t)])
        assert np.array_equal(oracle.radon2d(win, theta), g["n%d_sinogram" % nw])
        kb, S, dth = oracle.azimuthal_spectrum(win, 1000.0)
        assert np.array_equal(kb, g["n%d_az_k" % nw])
        assert np.array_equal(S, g["n%d_az_S" % nw])


def test_geometry_matches_reference(oracle, mag):
    d, ext = mag
    g = load_golden("misc.npz")
    spec = oracle.GridSpec(d, *ext)
    for tag, args in (("cl100", (100e3,)), ("cl100_s50", (100e3, 50e3, 50e3)), ("cl150_s5", (150e3, 5e3, 5e3)),
                      ("cl200_s7", (200e3, 7e3, 13e3))):
        xq, yq = oracle.centroid_list(spec, *args)
        assert xq.size == int(g[tag + "_n"])
        assert np.array_equal(xq[:400], g[tag + "_x_head"]) and np.array_equal(yq[-400:], g[tag + "_y_tail"])
    for (xc, yc), (j0, i0, n) in zip(g["subgrid_pts"], g["subgrid_origin"]):
        assert oracle.window_origin(spec, 100e3, xc, yc) == (j0, i0, n)
    j0, i0, n = (int(v) for v in g["detrend_in_origin"])
    assert np.array_equal(oracle.detrend_plane(d[j0:j0 + n, i0:i0 + n]), g["detrend_out"])
    with pytest.raises(ValueError, match="outside data range"):
        oracle.window_origin(spec, 100e3, ext[0] - 1.0, 150e3)
    with pytest.raises(ValueError, match="exceeds the data range"):
        oracle.window_origin(spec, 100e3, ext[0] + 10e3, 150e3)


def test_tanaka_matches_reference(oracle, mag):
    d, ext = mag
    g = load_golden("misc.npz")
    spec = oracle.GridSpec(d, *ext)
    k, phi, sig = oracle.radial_spectrum(d, spec.dx, taper=np.hanning, power=0.5)
    (Ztr, btr, dZtr), (Zor, bor, dZor) = oracle.tanaka1999(k, phi, sig, (0.005, 0.03), (0.03, 0.7))
    Zb, eZb = oracle.compute_tanaka(Ztr, dZtr, Zor, dZor)
    np.testing.assert_allclose([Ztr, btr, dZtr, Zor, bor, dZor, Zb, eZb], g["tanaka"], rtol=1e-8)   # ill-conditioned normal equations; grouping differs


def test_fft_filters_match_reference(oracle):
    """upward_continuation / reduce_to_pole restatements against the reference's own outputs."""
    g = load_golden("filters.npz")
    d = load_golden("mag_grid.npz")["d"]
    for name in g["names"]:
        j0, i0, n, st = (int(v) for v in g[str(name) + "_meta"])
        win = d[j0:j0 + n, i0:i0 + n]
        for tag, filt in (("up500", oracle.upward_continuation_filter(win.shape, 1000.0, 1000.0, 500.0)),
                          ("up5000", oracle.upward_continuation_filter(win.shape, 1000.0, 1000.0, 5000.0)),
                          ("rtp_induced", oracle.reduce_to_pole_filter(win.shape, 1000.0, 1000.0, -60.0, 12.0)),
                          ("rtp_remanent", oracle.reduce_to_pole_filter(win.shape, 1000.0, 1000.0, 35.0, -5.0, 50.0, 20.0))):
            ref = g["%s_%s" % (name, tag)]
            np.testing.assert_allclose(oracle.fft_filter(win, filt)[::st, ::st], ref, rtol=0, atol=1e-12 * np.abs(ref).max())


def test_config3_fixture_pins_oracle_fit(oracle):
    """tests/golden/cfg3.npz (the unmodified reference on 512 windows of BASELINE config 3): the oracle's
    L-BFGS-B call reproduces the reference's `optimise` bit for bit on the stored spectra, its objective
    equals `min_func`, and the fixture's reference self-sensitivity shows what the 1e-4 gate can cover."""
    g = load_golden("cfg3.npz")
    k, xref, xpol = g["k"], g["xref"], g["xpol"]
    n = xref.shape[0]
    for i in range(0, n, n // 12):
        assert np.array_equal(oracle.map_fit(k, g["phi"][i]), xref[i]), i
        assert oracle.misfit(xref[i], k, g["phi"][i]) == g["fref"][i]
    rel = lambda a, b: np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3), axis=-1)
    ref_off = rel(xref, xpol)
    spread = np.max(rel(g["xpert"], xref[:, None, :]), axis=1)
    conv = ref_off < 3e-5
    # the reference is converged on ~3/4 of the windows; there it is also stable under a 1e-13 perturbation
    assert 0.70 * n <= conv.sum() <= 0.80 * n
    assert np.mean(spread[conv] < 1e-4) >= 0.98
    # where it is not converged it is not reproducible either: it moves by more than 1e-4 against itself
    assert np.mean(spread[~conv] > 1e-4) >= 0.85
    assert np.median(spread[~conv]) > 1e-2
    assert np.all(g["fpol"] <= g["fref"] * (1 + 1e-12))
