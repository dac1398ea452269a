"""Host-side logic of pycurious_b200 (no GPU): index rules, plans, API surface and errors, the C-ABI
library loads and exports every declared symbol, and the product path refuses to run without CUDA."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


@pytest.fixture(scope="module")
def pc():
    import pycurious_b200
    return pycurious_b200


def test_library_exports_every_declared_symbol(pc):
    from pycurious_b200 import _lib
    lib = _lib.load()                      # loading needs no GPU
    header = open(os.path.join(ROOT, "include", "curie_b200.h")).read()
    declared = set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.cb_abi_version() == 1


def test_structs_match_the_header_layout():
    from pycurious_b200 import _lib
    assert C.sizeof(_lib.FitOptions) == 5 * 32 + 8 + 16
    assert C.sizeof(_lib.MhOptions) == 4 * 32 + 4 * 8 + 8
    assert C.sizeof(_lib.SpectrumPlanInfo) == 4 * 6 + 4 * 12 + 4 * 5 + 4 + 16


def test_no_gpu_fails_loudly(pc, mag):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pycurious_b200._lib import CurieB200Error
    d, ext = mag
    G = pc.CurieOptimise(d, *ext)
    with pytest.raises(CurieB200Error, match="no CPU fallback"):
        G.radial_spectrum(G.subgrid(100e3, 152500.0, 152500.0))
    with pytest.raises(CurieB200Error):
        G.optimise(100e3, 152500.0, 152500.0)
    with pytest.raises(CurieB200Error):
        pc.bouligand2009(np.array([0.1, 1.0]), 3.0, 1.0, 10.0, 5.0)
    # and the C-ABI itself reports the missing device instead of computing on the host
    from pycurious_b200 import _lib
    v = C.c_double()
    assert _lib.load().cb_fp64_fma_peak(C.byref(v)) == -2
    assert b"no CPU fallback" in _lib.load().cb_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pycurious_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "curie_oracle" not in text.replace("oracle/curie_oracle.py:PhiloxStream", ""), f
                assert "libcurie_emu" not in text and "libradon_oracle" not in text, f


def test_api_surface_matches_reference_signatures(pc):
    sig = lambda f: list(inspect.signature(f).parameters)
    G, O = pc.CurieGrid, pc.CurieOptimise
    assert sig(G.__init__) == ["self", "grid", "xmin", "xmax", "ymin", "ymax"]
    assert sig(G.subgrid) == ["self", "window", "xc", "yc"]
    assert sig(G.create_centroid_list) == ["self", "window", "spacingX", "spacingY"]
    assert sig(G.remove_trend_linear) == ["self", "data"]
    assert sig(G.radial_spectrum) == ["self", "subgrid", "taper", "power", "kwargs"]
    assert sig(G.azimuthal_spectrum) == ["self", "subgrid", "taper", "power", "theta", "kwargs"]
    assert sig(O.optimise) == ["self", "window", "xc", "yc", "beta", "zt", "dz", "C", "taper", "process_subgrid", "kwargs"]
    assert sig(O.optimise_routine) == ["self", "window", "xc_list", "yc_list", "beta", "zt", "dz", "C", "taper",
                                       "process_subgrid", "kwargs"]
    assert sig(O.parallelise_routine) == ["self", "window", "xc_list", "yc_list", "func", "args", "kwargs"]
    assert sig(O.metropolis_hastings)[:14] == ["self", "window", "xc", "yc", "nsim", "burnin", "x_scale", "beta", "zt",
                                               "dz", "C", "taper", "process_subgrid", "seed"]
    assert sig(O.sensitivity)[:12] == ["self", "window", "xc", "yc", "nsim", "beta", "zt", "dz", "C", "taper",
                                       "process_subgrid", "seed"]           # optimise.py:635-648 (+ keyword seed)
    assert sig(O.min_func) == ["self", "x", "kh", "Phi", "sigma_Phi"]
    d = inspect.signature(O.optimise).parameters
    assert (d["beta"].default, d["zt"].default, d["dz"].default, d["C"].default) == (3.0, 1.0, 10.0, 5.0)
    assert d["taper"].default is np.hanning
    for name in ("CurieGrid", "bouligand2009", "tanaka1999", "maus1995", "ComputeTanaka", "CurieOptimise"):
        assert hasattr(pc, name)


def test_subgrid_and_centroids_match_reference(pc, mag):
    d, ext = mag
    g = load_golden("misc.npz")
    G = pc.CurieGrid(d, *ext)
    for tag, args in (("cl100", (100e3,)), ("cl100_s50", (100e3, 50e3, 50e3)), ("cl150_s5", (150e3, 5e3, 5e3)),
                      ("cl200_s7", (200e3, 7e3, 13e3))):
        xq, yq = G.create_centroid_list(*args)
        assert xq.size == int(g[tag + "_n"])
        assert np.array_equal(xq[:400], g[tag + "_x_head"]) and np.array_equal(yq[:400], g[tag + "_y_head"])
        assert np.array_equal(xq[-400:], g[tag + "_x_tail"]) and np.array_equal(yq[-400:], g[tag + "_y_tail"])
    from pycurious_b200 import _plans
    pts, org = g["subgrid_pts"], g["subgrid_origin"]
    iy0, ix0, n = _plans.window_origins(G, 100e3, pts[:, 0], pts[:, 1])       # vectorised rule incl. exact ties
    assert n == 101 and np.array_equal(iy0, org[:, 0]) and np.array_equal(ix0, org[:, 1])
    for (xc, yc), (j0, i0, nn) in zip(pts[:50], org[:50]):
        s = G.subgrid(100e3, xc, yc)
        assert s.shape == (nn, nn) and np.shares_memory(s, G.data)
    rng = np.random.default_rng(0)
    q = rng.uniform(ext[0], ext[1], 2000)
    want = np.array([np.abs(G.xcoords - v).argmin() for v in q])
    assert np.array_equal(_plans.nearest_node(G.xcoords, q), want)
    with pytest.raises(ValueError, match="outside data range"):
        G.subgrid(100e3, ext[0] - 1.0, 150e3)
    with pytest.raises(ValueError, match="exceeds the data range"):
        G.subgrid(100e3, ext[0] + 10e3, 150e3)
    with pytest.raises(ValueError, match="node spacing should be identical"):
        pc.CurieGrid(np.zeros((10, 10)), 0.0, 90.0, 0.0, 9.0)      # rtol=1.0: only dx > 2 dy trips it


@pytest.mark.parametrize("nw", [61, 100, 101, 128, 129, 201, 301])
def test_bin_csr_is_the_reference_membership(pc, nw):
    from pycurious_b200 import _plans
    g = load_golden("spectrum.npz")
    kbins, k, ptr, idx = _plans.radial_bins(nw, 1000.0)
    assert np.array_equal(np.diff(ptr), g["bins%d_counts" % nw])
    nmemb = np.bincount(idx, minlength=nw * nw).reshape(nw, nw)
    assert np.array_equal(nmemb, g["bins%d_nmemb" % nw])
    first = np.full(nw * nw, -1, dtype=np.int64)
    for b in range(ptr.size - 1):
        m = idx[ptr[b]:ptr[b + 1]]
        first[m[first[m] < 0]] = b
    assert np.array_equal(first.reshape(nw, nw), g["bins%d_first" % nw])
    if nw == 101:
        assert np.array_equal(k, g["c100_hann_k"])


def test_radon_endpoints_equal_the_scalar_reference_loop(pc, oracle):
    from pycurious_b200 import _plans
    for nw in (11, 51, 100, 101, 201):
        for th in np.pi * np.arange(0.0, 180.0, 2.5) / 180.0:
            Tx, Rx = _plans.radon_endpoints(nw, th)
            Tx0, Rx0 = oracle.radon_ray_endpoints(nw, th)
            assert np.array_equal(Tx, Tx0) and np.array_equal(Rx, Rx0), (nw, th)


def test_priors_bounds_and_dispatch(pc, mag):
    d, ext = mag
    O = pc.CurieOptimise(d, *ext, max_processors=3)
    assert O.max_processors == 3 and O.bounds == [(0.0, None), (0.0, None), (0.0, None), (None, None)]
    O.add_prior(beta=(3.0, 0.5))
    from scipy import stats
    O.add_prior(zt=stats.norm(1.0, 0.25))
    assert O.prior["beta"] == [3.0, 0.5] and O.prior["zt"] == [1.0, 0.25] and O.prior["dz"] is None
    mu, sg = O._prior_arrays()
    assert mu[0] == 3.0 and sg[1] == 0.25 and np.isnan(mu[2]) and np.isnan(mu[3])
    np.testing.assert_allclose(O.objective_routine(beta=2.0, zt=1.5, dz=7.0), 0.5 * (1 / 0.25 + 0.25 / 0.0625))
    with pytest.raises(ValueError, match="prior must be one of"):
        O.add_prior(gamma=(1.0, 1.0))
    with pytest.raises(ValueError, match="scipy.stats"):
        O.add_prior(beta=3.0)
    O.reset_priors()
    assert all(v is None for v in O.prior.values())
    O.bounds = [(0.5, 4.0), (0.0, None), (None, 50.0), (None, None)]
    lb, ub = O._bound_arrays()
    assert list(lb) == [0.5, 0.0, -np.inf, -np.inf] and list(ub) == [4.0, np.inf, 50.0, np.inf]
    assert O._is_own(O.optimise, "optimise") and not O._is_own(O.metropolis_hastings, "optimise")
    assert O._is_detrend(O.remove_trend_linear) and not O._is_detrend(lambda w: w)
    other = pc.CurieOptimise(d, *ext)
    assert not O._is_own(other.optimise, "optimise")
    with pytest.raises(ValueError, match="same size"):
        O.parallelise_routine(100e3, [1.0, 2.0], [1.0], O.optimise)
    O.max_processors = 0
    with pytest.raises(ValueError, match="processors is invalid"):
        O.parallelise_routine(100e3, [150e3], [150e3], lambda w, x, y: np.zeros(4))
    O.max_processors = 1
    out = O.parallelise_routine(100e3, [150e3, 160e3], [150e3, 170e3], lambda w, x, y, s=1.0: np.array([x, y, w, s]), 2.0)
    assert len(out) == 4 and np.array_equal(out[1], [150e3, 170e3]) and np.array_equal(out[3], [2.0, 2.0])
    out = O.parallelise_routine(100e3, [150e3, 160e3], [150e3, 170e3], lambda w, x, y: [np.full(3, x)] * 4)
    assert len(out) == 4 and len(out[0]) == 2 and np.array_equal(out[0][1], np.full(3, 160e3))


def test_host_side_free_functions(pc, oracle, mag):
    d, ext = mag
    k = np.array([0.05, 0.5, 1.0, 3.0])
    assert np.array_equal(pc.maus1995(3.0, 1.0, k, 5.0), oracle.maus1995(3.0, 1.0, k, 5.0))
    g = load_golden("spectrum.npz")
    kk, phi, sig = g["whole305_hann_k"], g["whole305_hann_phi"] * 0.25, g["whole305_hann_sig"] * 0.25
    a = pc.tanaka1999(kk, phi, sig, (0.005, 0.03), (0.03, 0.7))
    b = oracle.tanaka1999(kk, phi, sig, (0.005, 0.03), (0.03, 0.7))
    np.testing.assert_allclose(np.ravel(a), np.ravel(b), rtol=1e-12)
    assert pc.ComputeTanaka(1.0, 0.1, 6.0, 0.2) == (11.0, 0.5)


def test_synthetic_grid_is_deterministic_and_fractal(pc):
    a = pc.synth.fractal_grid(256, 256, seed=3)
    b = pc.synth.fractal_grid(256, 256, seed=3)
    assert a.dtype == np.float64 and np.array_equal(a, b) and abs(a.std() - 100.0) < 1e-9
    assert pc.synth.grid_extent(4096, 4096) == (0.0, 4095000.0, 0.0, 4095000.0)
    G = pc.CurieGrid(pc.synth.fractal_grid(4096, 8, seed=1), *pc.synth.grid_extent(4096, 8)) if False else None
    # BASELINE centroid counts from the reference's own formula
    class H:      # coordinates only
        pass
    h = H()
    h.xcoords, h.dx = np.linspace(0.0, 4095000.0, 4096, retstep=True)
    h.ycoords = h.xcoords
    assert pc.CurieGrid.create_centroid_list(h, 128e3, 8e3, 8e3)[0].size == 246016
    assert pc.CurieGrid.create_centroid_list(h, 200e3, 10e3, 10e3)[0].size == 152100


def test_bench_contract_keys():
    """bench.py: the reference arm runs on the host and prints ONE JSON line with the contract's keys; the GPU
    arm's line (not runnable here) is checked statically for the keys the driver and judge read."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "8", "--grid", "512"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    want = "reference" if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "pycurious", "optimise.py")) else "port"
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == want
    assert d["scaling"] == "strong" and d["config"]["windows_total"] > 0
    src = open(os.path.join(ROOT, "bench.py")).read()
    gpu_line = src[src.index('line = {"metric": METRIC'):src.index("emit(json.dumps(line))")]
    code = "\n".join(l.split("#")[0] for l in gpu_line.splitlines())       # a comment must not swallow a key
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "stages",
                "cpu_baseline", "mcmc", "weak", "check", "config2", "config5"):
        assert '"%s"' % key in code, key
    assert "NCCL_DEBUG" not in "\n".join(l for l in src.splitlines() if "environ" in l)     # NCCL's log stays as the caller set it


def test_lattice_fast_path_equals_generic_rule(pc):
    """window_origins: the O(n) path for create_centroid_list lattices against the per-centroid argmin rule,
    and the fallback for lists that only look like lattices."""
    from pycurious_b200 import _plans
    n = 700
    d = np.zeros((n, n))
    G = pc.CurieGrid(d, 0.0, (n - 1) * 1000.0, 0.0, (n - 1) * 1000.0)
    xc, yc = G.create_centroid_list(100e3, 7300.0, 9100.0)          # spacing not a multiple of dx: ties / rounding
    assert xc.size >= 4096
    assert _plans._lattice_shape(xc, yc) is not None
    iy0, ix0, nw = _plans.window_origins(G, 100e3, xc, yc)
    half = nw // 2
    assert np.array_equal(ix0, _plans.nearest_node(G.xcoords, xc) - half)
    assert np.array_equal(iy0, _plans.nearest_node(G.ycoords, yc) - half)
    xs = xc.copy()
    xs[-7] -= 1000.0                                                  # breaks the lattice in one element only
    assert _plans._lattice_shape(xs, yc) is None
    iy1, ix1, _ = _plans.window_origins(G, 100e3, xs, yc)
    assert ix1[-7] == ix0[-7] - 1 and np.array_equal(np.delete(ix1, -7), np.delete(ix0, -7))
    rng = np.random.default_rng(0)
    perm = rng.permutation(xc.size)
    assert _plans._lattice_shape(xc[perm], yc[perm]) is None
    iy2, ix2, _ = _plans.window_origins(G, 100e3, xc[perm], yc[perm])
    assert np.array_equal(ix2, ix0[perm]) and np.array_equal(iy2, iy0[perm])
    # contiguous blocks of the list (one rank's share): they start and end in the middle of a lattice row
    for lo, hi in ((0, xc.size // 3 + 5), (xc.size // 3 + 5, 2 * xc.size // 3 - 11), (17, xc.size - 3), (40, 40 + 4096)):
        iy3, ix3, _ = _plans.window_origins(G, 100e3, xc[lo:hi], yc[lo:hi])
        assert np.array_equal(ix3, ix0[lo:hi]) and np.array_equal(iy3, iy0[lo:hi])
    # per axis: periodic, runs, one broken element, short lists, constant lists
    for q in (xc, yc, xs, xc[5:5000], yc[5:5000], np.full(5000, xc[3]), np.r_[xc[:4000], xc[:4000][::-1]]):
        assert np.array_equal(_plans.nearest_nodes_of_list(G.xcoords, q), _plans.nearest_node(G.xcoords, q))


def test_lattice_origins_match_centroid_lists(pc):
    """lattice_origins (axes only) against window_origins on the expanded lists, including which centroid
    the reference would name in its two ValueErrors (first offender in x-fastest order)."""
    from pycurious_b200 import _plans
    n = 300
    G = pc.CurieGrid(np.zeros((n, n)), 0.0, (n - 1) * 1000.0, 0.0, (n - 1) * 1000.0)
    xs, ys = G.centroid_lattice(60e3, 7300.0, 9100.0)
    xc, yc = G.create_centroid_list(60e3, 7300.0, 9100.0)
    assert np.array_equal(np.meshgrid(xs, ys)[0].ravel(), xc)
    iyu, ixu, nw = _plans.lattice_origins(G, 60e3, xs, ys)
    iy0, ix0, nw2 = _plans.window_origins(G, 60e3, xc, yc)
    assert nw == nw2
    assert np.array_equal(np.repeat(iyu, xs.size), iy0) and np.array_equal(np.tile(ixu, ys.size), ix0)
    lat = _plans.Lattice(xs, ys)
    assert lat.shape == (ys.size, xs.size) and len(lat) == xc.size

    def message(fn, *a):
        with pytest.raises(ValueError) as e:
            fn(*a)
        return str(e.value)

    for bx, by in [((3, -5e3), None), (None, (2, 4e8)), ((5, 1e3), None), (None, (0, 2e3)), ((4, -1.0), (1, 1e3)),
                   ((4, 1e3), (0, 1e3)), ((2, 5e8), (3, -7.0))]:
        x2, y2 = xs.copy(), ys.copy()
        if bx:
            x2[bx[0]] = bx[1]
        if by:
            y2[by[0]] = by[1]
        xq, yq = np.meshgrid(x2, y2)
        assert message(_plans.lattice_origins, G, 60e3, x2, y2) == \
            message(_plans.window_origins, G, 60e3, xq.ravel(), yq.ravel())


def test_process_subgrid_filter_recognition(pc):
    """Only `functools.partial(<own bound reduce_to_pole / upward_continuation>, ...)` whose sole missing argument
    is the window is turned into a device stage; everything else stays a host callback.  The operator is the
    one the method itself would build."""
    import functools
    n = 64
    G = pc.CurieOptimise(np.zeros((n, n)), 0.0, (n - 1) * 1000.0, 0.0, (n - 1) * 1000.0)
    other = pc.CurieOptimise(np.zeros((n, n)), 0.0, (n - 1) * 1000.0, 0.0, (n - 1) * 1000.0)
    from pycurious_b200.grid import _fftfreqs
    f = G._window_filter(functools.partial(G.upward_continuation, height=750.0), 21)
    kx, ky = _fftfreqs(G.dx, G.dy, (21, 21))
    assert f.shape == (21, 21) and np.array_equal(f, np.exp(-750.0 * np.hypot(kx, ky)))
    r1 = G._window_filter(functools.partial(G.reduce_to_pole, inc=-60.0, dec=12.0), 21)
    r2 = G._window_filter(functools.partial(G.reduce_to_pole, inc=-60.0, dec=12.0, sinc=-60.0, sdec=12.0), 21)
    assert r1.dtype == np.complex128 and np.array_equal(r1, r2) and r1[0, 0] == 0
    for cb in (None, lambda w: w, functools.partial(G.upward_continuation), functools.partial(G.reduce_to_pole, -60.0, 12.0),
               functools.partial(other.upward_continuation, height=5.0), functools.partial(G.remove_trend_linear),
               functools.partial(G.upward_continuation, height=1.0, bogus=2)):
        assert G._window_filter(cb, 21) is None
    with pytest.warns(UserWarning, match="downward continuation"):
        G._window_filter(functools.partial(G.upward_continuation, height=-1.0), 21)


def test_device_region_covers_windows(monkeypatch, mag):
    """CurieGrid._device_region: the uploaded row band / rectangle plus the shifted origins address exactly the
    nodes `subgrid` would return (host logic only: the upload is replaced by a CPU tensor)."""
    import torch
    import pycurious_b200 as pc
    from pycurious_b200 import _lib, _plans
    d, ext = mag
    G = pc.CurieGrid(d, *ext)
    monkeypatch.setattr(_lib, "require_cuda", lambda: torch)
    monkeypatch.setattr(pc.CurieGrid, "_upload", lambda self, block: torch.from_numpy(np.ascontiguousarray(block)))
    monkeypatch.setattr(pc.CurieGrid, "_origins_tensors",
                        lambda self, iy0, ix0, r0, c0: (iy0 - np.int32(r0), ix0 - np.int32(c0)))
    cases = [G.create_centroid_list(100e3, 25e3, 25e3),
             (np.array([60.5e3, 70e3, 81e3]), np.array([250e3, 240e3, 251e3])),
             (np.array([152e3]), np.array([152e3])),
             (np.array([50.5e3, 250.5e3]), np.array([50.5e3, 250.5e3]))]
    shapes = []
    for xc, yc in cases:
        g, iy0, ix0, n = G._windows_device(100e3, xc, yc)
        shapes.append(tuple(g.shape))
        for i in range(len(xc)):
            win = g[int(iy0[i]):int(iy0[i]) + n, int(ix0[i]):int(ix0[i]) + n].numpy()
            assert np.array_equal(win, G.subgrid(100e3, xc[i], yc[i]))
    assert shapes[0] == (301, 305) and shapes[2] == (101, 101) and shapes[3] == (301, 305)
    assert shapes[1][0] < 120 and shapes[1][1] < 130
