"""
Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference and oracle/_ref built by
`make -C oracle`):

    python tests/golden/make_golden.py

The reference package is imported unmodified from /root/reference; its compiled
`radon` extension comes from oracle/_ref/ (built from the reference's own
src/radon.pyx + src/cradon.c).  Nothing of the reference is copied: only its
numerical OUTPUTS (and the 305x305 test anomaly grid, a data file) are stored.

Files written (all small .npz):
  mag_grid.npz    the reference's tests/test_mag_data.txt as a (305,305) array + extent
  spectrum.npz    radial_spectrum outputs + bin counts for several window sizes / tapers
  forward.npz     bouligand2009 / maus1995 / kv / gamma values incl. edge cases
  fit.npz         optimise / optimise_routine outputs (+ tightly polished minima)
  mh.npz          metropolis_hastings chains driven by the Philox stream, and
                  moment statistics of chains driven by numpy's own RNG
  radon.npz       cradon.c CSR integers/lengths, sinograms, azimuthal_spectrum
  misc.npz        centroid lists, subgrid origins, remove_trend_linear, tanaka1999
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PYCURIOUS_REF", "/root/reference")

sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

warnings.simplefilter("ignore")
import pycurious                      # the real reference          # noqa: E402
import radon as _ref_radon            # reference's compiled radon  # noqa: E402
import ref_cradon                     # reference cradon.c shim     # noqa: E402
import curie_oracle as orc            # for the Philox stream + polished_fit only  # noqa: E402

pycurious.radon = _ref_radon
sys.modules["pycurious.radon"] = _ref_radon

from scipy.optimize import minimize   # noqa: E402
from scipy.special import kv, gamma   # noqa: E402


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("wrote", name, "%.1f KB" % (os.path.getsize(path) / 1024))


# ---------------------------------------------------------------- grid
raw = np.loadtxt(os.path.join(REF, "tests", "test_mag_data.txt"))
d = raw[:, 2].reshape(305, 305)
extent = np.array([raw[:, 0].min(), raw[:, 0].max(), raw[:, 1].min(), raw[:, 1].max()])
save("mag_grid.npz", d=d, extent=extent)
xmin, xmax, ymin, ymax = extent
xc0, yc0 = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax)

G = pycurious.CurieGrid(d, xmin, xmax, ymin, ymax)

# ---------------------------------------------------------------- spectrum
TAPERS = {"hanning": np.hanning, "hamming": np.hamming, "none": None,
          "blackman": np.blackman, "kaiser": np.kaiser}
spec_cases = [
    # name, window(m) or None (direct slice), xc, yc, taper, power, detrend, kwargs
    ("c100_hann", 100e3, xc0, yc0, "hanning", 2.0, 0, {}),
    ("c128_hann_p05", 128e3, 120.4e3, 171.7e3, "hanning", 0.5, 0, {}),
    ("c150_hamming", 150e3, 97.0e3, 203.2e3, "hamming", 2.0, 0, {}),
    ("c200_hann", 200e3, 140.5e3, 160.5e3, "hanning", 2.0, 0, {}),
    ("c200_hann_detrend", 200e3, 140.5e3, 160.5e3, "hanning", 2.0, 1, {}),
    ("c300_none", 300e3, xc0, yc0, "none", 2.0, 0, {}),
    ("c100_kaiser", 100e3, 60.5e3, 250.5e3, "kaiser", 2.0, 0, {"beta": 8.6}),
    ("c60_blackman_detrend", 60e3, 230.0e3, 41.0e3, "blackman", 2.0, 1, {}),
]
out = {}
names = []
for name, win, xc, yc, tp, pw, det, kw in spec_cases:
    sub = G.subgrid(win, xc, yc)
    j0 = int(np.abs(G.ycoords - yc).argmin()) - int(round(win / G.dx)) // 2
    i0 = int(np.abs(G.xcoords - xc).argmin()) - int(round(win / G.dx)) // 2
    assert np.array_equal(sub, d[j0:j0 + sub.shape[0], i0:i0 + sub.shape[1]])
    if det:
        sub = G.remove_trend_linear(sub)
    k, phi, sig = G.radial_spectrum(sub, taper=TAPERS[tp], power=pw, **kw)
    out[name + "_k"], out[name + "_phi"], out[name + "_sig"] = k, phi, sig
    out[name + "_meta"] = np.array([j0, i0, sub.shape[0], det], dtype=np.int64)
    out[name + "_power"] = np.array(pw)
    out[name + "_kaiser_beta"] = np.array(kw.get("beta", np.nan))
    names.append(name + ":" + tp)
# whole grid (tests/test_grid.py:35) and direct even-sized slices
for name, sl, tp in [("whole305_none", (slice(0, 305), slice(0, 305)), "none"),
                     ("whole305_hann", (slice(0, 305), slice(0, 305)), "hanning"),
                     ("even128_hann", (slice(10, 138), slice(20, 148)), "hanning"),
                     ("even100_hamming", (slice(200, 300), slice(3, 103)), "hamming"),
                     ("smooth125_hann", (slice(50, 175), slice(60, 185)), "hanning")]:
    sub = d[sl]
    k, phi, sig = G.radial_spectrum(sub, taper=TAPERS[tp])
    out[name + "_k"], out[name + "_phi"], out[name + "_sig"] = k, phi, sig
    out[name + "_meta"] = np.array([sl[0].start, sl[1].start, sub.shape[0], 0], dtype=np.int64)
    out[name + "_power"] = np.array(2.0)
    out[name + "_kaiser_beta"] = np.array(np.nan)
    names.append(name + ":" + tp)
out["names"] = np.array(names)
# bin membership: counts per bin + a flat membership checksum, straight from the reference expressions
for nw in (61, 100, 101, 125, 128, 129, 151, 201, 301, 305):
    sub = np.ones((nw, nw))
    _, dk, kbins = G._taper_spectrum(sub, None)
    i0 = int((nw - 1) // 2)
    ix, iy = np.mgrid[0:nw, 0:nw]
    kk = np.hypot((ix - i0) * dk, (iy - i0) * dk)
    counts = np.empty(kbins.size - 1, dtype=np.int64)
    nmemb = np.zeros((nw, nw), dtype=np.int8)
    first = np.full((nw, nw), -1, dtype=np.int16)
    for i in range(kbins.size - 1):
        m = np.logical_and(kk >= kbins[i], kk <= kbins[i + 1])
        counts[i] = m.sum()
        first[np.logical_and(m, first < 0)] = i
        nmemb += m
    out["bins%d_counts" % nw] = counts
    out["bins%d_nmemb" % nw] = nmemb
    out["bins%d_first" % nw] = first
save("spectrum.npz", **out)

# ---------------------------------------------------------------- forward model
kh = np.array([0.05, 0.5, 1.0, 3.0])
params = np.array([[3, 1, 10, 5], [2.5, 0.3, 20, -18], [0, 0, 1, 0], [1, 0.5, 5, 2], [2, 0.1, 50, -20],
                   [4.2, 2.0, 100, 10], [3, 1, 0, 5], [3, 1, 1e-8, 5], [3, 1, 240, 5], [-1, 1, 10, 5],
                   [-2, 1, 10, 5], [-3, 1, 10, 5], [7.0, 0.0, 3.3, -3.0], [0.37, 1.7, 0.5, 1.0],
                   [3.0, 1.0, 200.0, 5.0], [5.999, 0.2, 12.0, 0.0], [6.0, 0.2, 12.0, 0.0]], dtype=float)
with np.errstate(all="ignore"):
    fw = np.array([pycurious.bouligand2009(kh, *p) for p in params])
    ms = np.array([pycurious.maus1995(p[0], p[1], kh, p[3]) for p in params])
# dense well-conditioned sweep on a realistic k axis (101^2 window)
k49 = out["c100_hann_k"]
rng = np.random.default_rng(7)
pdense = np.c_[rng.uniform(0.2, 6.0, 200), rng.uniform(0.0, 3.0, 200),
               rng.uniform(0.5, 120.0, 200), rng.uniform(-25, 25, 200)]
with np.errstate(all="ignore"):
    fdense = np.array([pycurious.bouligand2009(k49, *p) for p in pdense])
nu = np.array([0.0, 0.3, 0.5, 1.0, 1.37, 1.5, 2.0, 2.5, 2.9999, 3.0, 3.5, 4.0, 5.25, 12.0])
xs = np.array([1e-6, 1e-3, 0.05, 0.5, 1.0, 1.999, 2.0, 2.001, 3.0, 5.0, 10.0, 25.0, 60.0, 150.0, 400.0, 690.0])
KV = kv(nu[:, None], xs[None, :])
gx = np.array([0.5, 1.0, 1.25, 1.5, 2.0, 2.5, 3.3, 4.0, 6.5, 0.1, -0.5, -1.5])
save("forward.npz", kh=kh, params=params, bouligand=fw, maus=ms, k49=k49, pdense=pdense, fdense=fdense,
     nu=nu, x=xs, kv=KV, gx=gx, gamma=gamma(gx))

# ---------------------------------------------------------------- fits
O = pycurious.CurieOptimise(d, xmin, xmax, ymin, ymax, max_processors=1)
fit = {}


def run_fit_set(tag, window, xcs, ycs, priors=None, x0=(3.0, 1.0, 10.0, 5.0), **kw):
    O.reset_priors()
    mu = np.full(4, np.nan)
    sg = np.full(4, np.nan)
    if priors:
        O.add_prior(**priors)
        for j, key in enumerate(("beta", "zt", "dz", "C")):
            if key in priors:
                mu[j], sg[j] = priors[key]
    res = O.optimise_routine(window, xcs, ycs, *x0, **kw)
    xref = np.vstack(res).T
    n = len(xcs)
    ks, phis, fref, xpol, fpol, nfev = [], [], [], [], [], []
    for i in range(n):
        sub = O.subgrid(window, xcs[i], ycs[i])
        if kw.get("process_subgrid") is not None:
            sub = kw["process_subgrid"](sub)
        k, phi, sig = O.radial_spectrum(sub, taper=kw.get("taper", np.hanning))
        ks.append(k)
        phis.append(phi)
        fref.append(O.min_func(xref[i], k, phi, sig))
        r = minimize(O.min_func, np.array(x0, dtype=float), args=(k, phi, sig), bounds=O.bounds)
        assert np.allclose(r.x, xref[i], rtol=0, atol=0), "optimise_routine != direct minimize"
        nfev.append(r.nfev)
        xp, fp = orc.polished_fit(k, phi, xref[i], prior_mu=mu, prior_sigma=sg)
        xpol.append(xp)
        fpol.append(fp)
    fit[tag + "_window"] = np.array(window)
    fit[tag + "_xc"], fit[tag + "_yc"] = np.asarray(xcs), np.asarray(ycs)
    fit[tag + "_k"], fit[tag + "_phi"] = np.array(ks[0]), np.array(phis)
    fit[tag + "_xref"], fit[tag + "_fref"], fit[tag + "_nfev"] = xref, np.array(fref), np.array(nfev)
    fit[tag + "_xpol"], fit[tag + "_fpol"] = np.array(xpol), np.array(fpol)
    fit[tag + "_prior_mu"], fit[tag + "_prior_sigma"] = mu, sg
    fit[tag + "_x0"] = np.array(x0, dtype=float)
    O.reset_priors()
    print(tag, "done", n, "windows; max rel |xref-xpol| per window:",
          np.round(np.max(np.abs(xref - np.array(xpol)) / np.maximum(np.abs(np.array(xpol)), 1e-3), axis=1), 6))


# config 1: 5x5 centroids, 100 km windows (BASELINE.json configs[0])
xcs, ycs = O.create_centroid_list(100e3, 50e3, 50e3)
assert xcs.size == 25
run_fit_set("cfg1", 100e3, xcs, ycs)
# 200 km windows, 3x3
xs3 = np.array([110.5e3, 152.5e3, 200.5e3])
xx, yy = np.meshgrid(xs3, xs3)
run_fit_set("w200", 200e3, xx.ravel(), yy.ravel())
# priors + detrend + hamming, 150 km
xs2 = np.array([90.5e3, 152.5e3, 215.5e3])
xx, yy = np.meshgrid(xs2, xs2[:2])
run_fit_set("w150_prior", 150e3, xx.ravel(), yy.ravel(), priors=dict(beta=(3.0, 0.5), zt=(0.5, 0.25)),
            taper=np.hamming, process_subgrid=O.remove_trend_linear)
run_fit_set("w300_betaprior", 300e3, np.array([xc0]), np.array([yc0]), priors=dict(beta=(1.0, 0.1)))
save("fit.npz", **fit)

# ---------------------------------------------------------------- Metropolis-Hastings
mh = {}
_np_normal, _np_rand = np.random.normal, np.random.rand


class PhiloxPatch:
    """Replaces np.random.normal/rand inside the reference by the device's Philox stream."""

    def __init__(self, seed, chain, burnin):
        self.s = orc.PhiloxStream(seed, chain)
        self.burnin = burnin
        self.n = 0

    def normal(self, size=4):
        assert size == 4
        if self.n == self.burnin and self.s.phase == 0:
            self.s.start_sampling()
        self.n += 1
        return self.s.normal4()

    def rand(self):
        return self.s.uniform()


def ref_chain_philox(window, xc, yc, nsim, burnin, seed, chain, priors=None, **kw):
    O.reset_priors()
    if priors:
        O.add_prior(**priors)
    pp = PhiloxPatch(seed, chain, burnin)
    np.random.normal, np.random.rand = pp.normal, pp.rand
    try:
        res = O.metropolis_hastings(window, xc, yc, nsim, burnin, **kw)
    finally:
        np.random.normal, np.random.rand = _np_normal, _np_rand
        O.reset_priors()
    return np.array(res)


pri = dict(beta=(3.0, 1.0), zt=(1.0, 1.0), dz=(20.0, 10.0), C=(-19.0, 5.0))
xsc = np.array([0.25, 0.1, 1.0, 0.5])
mh["seed"] = np.array(20261019)
mh["x_scale"] = xsc
mh["prior_mu"] = np.array([pri[k][0] for k in ("beta", "zt", "dz", "C")])
mh["prior_sigma"] = np.array([pri[k][1] for k in ("beta", "zt", "dz", "C")])
mh["start"] = np.array([3.0, 0.5, 10.0, -19.0])
mh["window"] = np.array(100e3)
mh["xc"], mh["yc"] = np.array([xc0, 100.5e3]), np.array([yc0, 200.5e3])
for ci in range(2):
    mh["philox_prior_%d" % ci] = ref_chain_philox(100e3, mh["xc"][ci], mh["yc"][ci], 3000, 200, 20261019, ci,
                                                   priors=pri, x_scale=xsc, beta=3.0, zt=0.5, dz=10.0, C=-19.0)
    mh["philox_noprior_%d" % ci] = ref_chain_philox(100e3, mh["xc"][ci], mh["yc"][ci], 3000, 200, 20261019, ci,
                                                     x_scale=xsc, beta=3.0, zt=0.5, dz=10.0, C=-19.0)
# statistical pin: the reference's own RNG, several long chains -> moments and batch-means errors
O.reset_priors()
O.add_prior(**pri)
np.random.seed(12345)
nch, nsim = 6, 40000
mom = np.empty((nch, 2, 5))
for c in range(nch):
    ch = np.array(O.metropolis_hastings(100e3, xc0, yc0, nsim, 1000, x_scale=xsc,
                                        beta=3.0, zt=0.5, dz=10.0, C=-19.0))
    ch = np.vstack([ch, ch[1] + ch[2]])
    mom[c, 0], mom[c, 1] = ch.mean(axis=1), ch.std(axis=1)
O.reset_priors()
mh["ref_moments"] = mom
mh["ref_nsim"] = np.array(nsim)
sub = O.subgrid(100e3, xc0, yc0)
mh["k"], mh["phi"], _ = O.radial_spectrum(sub)
save("mh.npz", **mh)

# ---------------------------------------------------------------- radon / azimuthal
rd = {}
for nw, j0, i0 in ((51, 100, 120), (101, 102, 102)):
    sub = np.ascontiguousarray(d[j0:j0 + nw, i0:i0 + nw])
    theta_deg = np.arange(0.0, 180.0, 5.0)
    theta = np.pi * theta_deg / 180.0
    rd["n%d_meta" % nw] = np.array([j0, i0, nw])
    rd["n%d_theta" % nw] = theta
    rd["n%d_sinogram" % nw] = _ref_radon.radon2d(sub, theta)
    nnz, ptrs, idx_sum, len_sum = [], [], [], []
    for t, th in enumerate(theta):
        Tx, Rx = orc.radon_ray_endpoints(nw, th)
        dat, ind, ptr = ref_cradon.trace(Tx, Rx, nw, nw)
        ptrs.append(ptr)
        if t in (0, 1, 7, 18, 19, 35):
            rd["n%d_t%d_data" % (nw, t)] = dat
            rd["n%d_t%d_indices" % (nw, t)] = ind
        idx_sum.append(np.add.reduceat(ind * (1 + np.arange(ind.size) % 7), ptr[:-1]))
        len_sum.append(np.add.reduceat(dat, ptr[:-1]))
    rd["n%d_indptr" % nw] = np.array(ptrs)
    rd["n%d_index_checksum" % nw] = np.array(idx_sum)
    rd["n%d_length_sum" % nw] = np.array(len_sum)
    kb, S, dth = G.azimuthal_spectrum(sub, taper=np.hanning, theta=5.0)
    rd["n%d_az_k" % nw], rd["n%d_az_S" % nw], rd["n%d_az_theta" % nw] = kb, S, dth
save("radon.npz", **rd)

# ---------------------------------------------------------------- misc
mi = {}
for tag, args in (("cl100", (100e3,)), ("cl100_s50", (100e3, 50e3, 50e3)), ("cl150_s5", (150e3, 5e3, 5e3)),
                  ("cl200_s7", (200e3, 7e3, 13e3))):
    xq, yq = G.create_centroid_list(*args)
    mi[tag + "_n"] = np.array(xq.size)
    mi[tag + "_x_head"], mi[tag + "_y_head"] = xq[:400], yq[:400]
    mi[tag + "_x_tail"], mi[tag + "_y_tail"] = xq[-400:], yq[-400:]
rng = np.random.default_rng(11)
pts = np.c_[rng.uniform(60e3, 245e3, 300), rng.uniform(60e3, 245e3, 300)]
pts[:40] = np.round(pts[:40] / 500.0) * 500.0          # exact ties between two nodes
org = []
for xc, yc in pts:
    s = G.subgrid(100e3, xc, yc)
    off = (s.__array_interface__["data"][0] - G.data.__array_interface__["data"][0]) // 8
    org.append((off // 305, off % 305, s.shape[0]))
mi["subgrid_pts"], mi["subgrid_origin"] = pts, np.array(org)
sub = d[30:131, 150:251]
mi["detrend_in_origin"] = np.array([30, 150, 101])
mi["detrend_out"] = G.remove_trend_linear(sub)
k, phi, sig = G.radial_spectrum(G.data, taper=np.hanning, power=0.5)
(Ztr, btr, dZtr), (Zor, bor, dZor) = pycurious.tanaka1999(k, phi, sig, (0.005, 0.03), (0.03, 0.7))
Zb, eZb = pycurious.ComputeTanaka(Ztr, dZtr, Zor, dZor)
mi["tanaka"] = np.array([Ztr, btr, dZtr, Zor, bor, dZor, Zb, eZb])
save("misc.npz", **mi)
print("done")
