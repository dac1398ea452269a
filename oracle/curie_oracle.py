"""
CPU oracle for the per-window Curie-depth path  (TEST INFRASTRUCTURE ONLY).

This module is a plain numpy/scipy restatement of the algorithm of the
reference (brmather/pycurious v1.1.1).  It exists so that the CUDA path can
be checked on a box where `/root/reference` does not exist.  It is the
*checker*, never the product:

    only  tests/ ,  __graft_entry__.smoke()  and  bench.py's cpu_baseline /
    --impl reference leg  may import this module.

Nothing in `pycurious_b200/` imports it, and the product path raises when the
CUDA library is missing instead of falling back to this code.

Parity pin: every function here is compared against the *real* reference
(imported from /root/reference in the build container) by
`tests/golden/make_golden.py`; the resulting vectors are committed under
`tests/golden/` and `tests/test_oracle_golden.py` re-checks the oracle against
them on every run.  Third-party arithmetic that the reference reaches through
numpy/scipy (pocketfft `fft2`, `scipy.special.kv/gamma`, L-BFGS-B from
`scipy.optimize.minimize`, the legacy MT19937 global RNG) is reached the same
way here (numpy 2.3.5 / scipy 1.18.1 in this image; the reference pins nothing,
`setup.py:76`).

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import ctypes
import os
import warnings

import numpy as np
from scipy.optimize import minimize as _scipy_minimize
from scipy.special import gamma as _gamma, kv as _kv

_HERE = os.path.dirname(os.path.abspath(__file__))

PARAM_NAMES = ("beta", "zt", "dz", "C")


# --------------------------------------------------------------------------
# grid geometry  (pycurious/grid.py:98-197)
# --------------------------------------------------------------------------
class GridSpec:
    """Coordinates of a regular grid, as CurieGrid.__init__ builds them
    (pycurious/grid.py:98-110)."""

    def __init__(self, data, xmin, xmax, ymin, ymax):
        self.data = np.array(data)
        self.ny, self.nx = self.data.shape
        self.xmin, self.xmax, self.ymin, self.ymax = xmin, xmax, ymin, ymax
        self.xcoords, self.dx = np.linspace(xmin, xmax, self.nx, retstep=True)
        self.ycoords, self.dy = np.linspace(ymin, ymax, self.ny, retstep=True)
        # grid.py:109 passes 1.0 as rtol, so this almost never raises
        if not np.allclose(self.dx, self.dy, 1.0):
            raise ValueError("node spacing should be identical {}".format((self.dx, self.dy)))


def window_size_nodes(g: GridSpec, window):
    """nw, n2w of grid.py:138-139."""
    nw = int(round(window / g.dx))
    return nw, nw // 2


def window_origin(g: GridSpec, window, xc, yc):
    """Top-left node (row j0, col i0) and edge length of the window that
    CurieGrid.subgrid cuts (grid.py:130-155).  Raises the same ValueErrors."""
    if xc < g.xmin or xc > g.xmax or yc < g.ymin or yc > g.ymax:
        raise ValueError("Point {} outside data range".format((xc, yc)))
    ic = np.abs(g.xcoords - xc).argmin()      # ties -> lower node (argmin rule)
    jc = np.abs(g.ycoords - yc).argmin()
    _, half = window_size_nodes(g, window)
    i0, i1 = ic - half, ic + half + 1
    j0, j1 = jc - half, jc + half + 1
    if i0 < 0 or i1 > g.nx or j0 < 0 or j1 > g.ny:
        raise ValueError(
            "Window size {} at centroid {} exceeds the data range".format(window, (xc, yc))
        )
    return int(j0), int(i0), int(2 * half + 1)


def cut_window(g: GridSpec, window, xc, yc):
    j0, i0, n = window_origin(g, window, xc, yc)
    return g.data[j0:j0 + n, i0:i0 + n]


def centroid_list(g: GridSpec, window, spacingX=None, spacingY=None):
    """grid.py:179-197."""
    _, half = window_size_nodes(g, window)
    xs = g.xcoords[half:-half]
    ys = g.ycoords[half:-half]
    if spacingX is not None:
        xs = np.arange(xs.min(), xs.max(), spacingX)
    if spacingY is not None:
        ys = np.arange(ys.min(), ys.max(), spacingY)
    xx, yy = np.meshgrid(xs, ys)
    return xx.ravel(), yy.ravel()


# --------------------------------------------------------------------------
# detrend  (pycurious/grid.py:199-216)
# --------------------------------------------------------------------------
def detrend_plane(win):
    """Least-squares plane removal exactly as grid.py:212-216 (LAPACK lstsq on
    the [col-index, row-index, 1] design matrix)."""
    nr, nc = win.shape
    rowq, colq = np.mgrid[0:nc, 0:nr]
    design = np.c_[colq.ravel(), rowq.ravel(), np.ones(colq.size)]
    coef = np.linalg.lstsq(design, win.ravel(), rcond=None)[0]
    return win - np.dot(design, coef).reshape(win.shape)


# --------------------------------------------------------------------------
# spectrum  (pycurious/grid.py:218-339)
# --------------------------------------------------------------------------
def taper_and_bins(nr, nc, dx, taper=np.hanning, scale=0.001, **kw):
    """2-D taper, dk and bin edges of grid.py:226-246."""
    if nr != nc:
        warnings.warn("subgrid is not square {}".format((nr, nc)), RuntimeWarning)
    if taper is None:
        w2d = 1.0
    else:
        w2d = np.outer(taper(nr, **kw), taper(nc, **kw))   # == meshgrid product, grid.py:236-239
    dk = 2.0 * np.pi / (nr - 1) / (dx * scale)
    edges = np.arange(dk, dk * nr / 2, dk)
    return w2d, dk, edges


def wavenumber_plane(nr, dk):
    """kk of grid.py:277-279 (note: uses nr for both axes and i0=(nr-1)//2)."""
    c = int((nr - 1) // 2)
    a, b = np.mgrid[0:nr, 0:nr]
    return np.hypot((a - c) * dk, (b - c) * dk)


def radial_bin_masks(nr, dk, edges):
    """Boolean membership mask per bin, closed on both ends (grid.py:281-282)."""
    kk = wavenumber_plane(nr, dk)
    return kk, [np.logical_and(kk >= edges[i], kk <= edges[i + 1]) for i in range(edges.size - 1)]


def radial_spectrum(win, dx, taper=np.hanning, power=2.0, **kw):
    """(k, Phi, sigma_Phi) of CurieGrid.radial_spectrum (grid.py:335-339 ->
    :264-288): fft2 of the tapered window, modulus, fftshift, then per bin the
    mean / population std of power*log|F| and the mean wavenumber."""
    nr, nc = win.shape
    w2d, dk, edges = taper_and_bins(nr, nc, dx, taper, **kw)
    amp = np.fft.fftshift(np.abs(np.fft.fft2(win * w2d)))
    kk, masks = radial_bin_masks(nr, dk, edges)
    nb = len(masks)
    k = np.empty(nb)
    phi = np.empty(nb)
    sig = np.empty(nb)
    for i, m in enumerate(masks):          # the O(nbins * nw^2) loop, grid.py:281-286
        vals = power * np.log(amp[m])
        phi[i] = vals.mean()
        k[i] = kk[m].mean()
        sig[i] = np.std(vals)
    return k, phi, sig


# --------------------------------------------------------------------------
# forward model and objective  (pycurious/grid.py:552-600, optimise.py:159-230)
# --------------------------------------------------------------------------
def bouligand2009(kh, beta, zt, dz, C):
    """Eq. (4) of Bouligand et al. 2009 as grid.py:586-600 evaluates it."""
    x = kh * dz
    nu = 0.5 * (1.0 + beta)
    lead = C - 2.0 * kh * zt - (beta - 1.0) * np.log(kh) - x
    bracket = 0.5 * np.cosh(x) * _gamma(nu) - _kv(-nu, x) * np.power(0.5 * x, nu)
    return lead + np.log(np.sqrt(np.pi) / _gamma(1.0 + 0.5 * beta) * bracket)


def maus1995(beta, zt, kh, C=0.0):
    """grid.py:735."""
    return C - 2.0 * kh * zt - (beta - 1.0) * np.log(kh)


def prior_penalty(x, prior_mu, prior_sigma):
    """Sum of 0.5*(x-mu)^2/sigma^2 over the parameters that carry a prior
    (optimise.py:171-179, :194).  NaN in prior_mu marks 'no prior'."""
    c = 0.0
    for xi, m, s in zip(x, prior_mu, prior_sigma):
        if not np.isnan(m):
            c += 0.5 * np.sum((xi - m) ** 2 / s ** 2)
    return c


NO_PRIOR = (np.full(4, np.nan), np.full(4, np.nan))


def misfit(x, kh, phi, prior_mu=NO_PRIOR[0], prior_sigma=NO_PRIOR[1]):
    """CurieOptimise.min_func (optimise.py:220-230): sigma_Phi is ignored
    (hard-coded 1.0 at :225), non-finite -> 1e99 without priors."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            syn = bouligand2009(kh, x[0], x[1], x[2], x[3])
            val = 0.5 * np.sum((syn - phi) ** 2 / 1.0 ** 2)
    if not np.isfinite(val):
        return 1e99
    return val + prior_penalty(x, prior_mu, prior_sigma)


DEFAULT_BOUNDS = [(0.0, None), (0.0, None), (0.0, None), (None, None)]   # optimise.py:111-114


def map_fit(kh, phi, x0=(3.0, 1.0, 10.0, 5.0), bounds=DEFAULT_BOUNDS,
            prior_mu=NO_PRIOR[0], prior_sigma=NO_PRIOR[1], full=False, **opts):
    """optimise.py:300: scipy minimize with bounds and no method -> L-BFGS-B,
    forward-difference gradient with absolute step 1e-8."""
    res = _scipy_minimize(misfit, np.array(x0, dtype=float),
                          args=(kh, phi, prior_mu, prior_sigma), bounds=bounds, **opts)
    return res if full else res.x


def polished_fit(kh, phi, x_start, bounds=DEFAULT_BOUNDS,
                 prior_mu=NO_PRIOR[0], prior_sigma=NO_PRIOR[1]):
    """Tightly converged bounded minimum near x_start (used to decide whether the
    reference's own stopping point is converged; SURVEY.md section 7 'hard parts')."""
    x = np.array(x_start, dtype=float)
    f = misfit(x, kh, phi, prior_mu, prior_sigma)
    for _ in range(6):
        r = _scipy_minimize(misfit, x, args=(kh, phi, prior_mu, prior_sigma), bounds=bounds,
                            method="L-BFGS-B",
                            options=dict(ftol=1e-15, gtol=1e-10, maxiter=2000, maxfun=20000))
        lo = np.array([(-np.inf if b[0] is None else b[0]) for b in bounds])
        hi = np.array([(np.inf if b[1] is None else b[1]) for b in bounds])

        def clipped(z):
            return misfit(np.clip(z, lo, hi), kh, phi, prior_mu, prior_sigma)

        r2 = _scipy_minimize(clipped, r.x, method="Nelder-Mead",
                             options=dict(xatol=1e-10, fatol=1e-14, maxiter=4000, maxfev=8000))
        xn = np.clip(r2.x, lo, hi) if r2.fun < r.fun else r.x
        fn = min(r2.fun, r.fun)
        done = abs(f - fn) <= 1e-13 * max(1.0, abs(fn))
        x, f = xn, fn
        if done:
            break
    return x, f


def sensitivity_rows(kh, phi_rows, prior_mu_rows, prior_sigma, x0=(3.0, 1.0, 10.0, 5.0), bounds=DEFAULT_BOUNDS):
    """The loop body of CurieOptimise.sensitivity (optimise.py:705-716) for given draws: row s is one
    simulation with its perturbed spectrum phi_rows[s] and its redrawn prior means prior_mu_rows[s]."""
    return np.array([map_fit(kh, phi_rows[s], x0, bounds, prior_mu_rows[s], prior_sigma)
                     for s in range(len(phi_rows))])


def window_fit(g: GridSpec, window, xc, yc, x0=(3.0, 1.0, 10.0, 5.0), taper=np.hanning,
               detrend=False, bounds=DEFAULT_BOUNDS, prior_mu=NO_PRIOR[0],
               prior_sigma=NO_PRIOR[1], **kw):
    """CurieOptimise.optimise (optimise.py:282-301)."""
    win = cut_window(g, window, xc, yc)
    if detrend:
        win = detrend_plane(win)
    k, phi, _ = radial_spectrum(win, g.dx, taper=taper, **kw)
    return map_fit(k, phi, x0, bounds, prior_mu, prior_sigma)


# --------------------------------------------------------------------------
# Metropolis-Hastings  (pycurious/optimise.py:585-633)
# --------------------------------------------------------------------------
class NumpyGlobalRNG:
    """The reference draws from the legacy global numpy RNG."""

    def normal4(self):
        return np.random.normal(size=4)

    def uniform(self):
        return np.random.rand()


def mh_chain(kh, phi, nsim, burnin, x0=(3.0, 1.0, 10.0, 5.0), x_scale=None,
             prior_mu=NO_PRIOR[0], prior_sigma=NO_PRIOR[1], rng=None):
    """optimise.py:585-633 with the random source made pluggable: `rng` must
    offer normal4() -> 4 standard normals and uniform() -> one U[0,1).
    Burn-in is the reference's greedy tempered climb (no uniform drawn)."""
    rng = rng or NumpyGlobalRNG()
    out = np.empty((nsim, 4))
    cur = np.array(x0, dtype=float)
    step = np.ones(4) if x_scale is None else np.asarray(x_scale, dtype=float)

    def f(x):
        return misfit(x, kh, phi, prior_mu, prior_sigma)

    p_cur = np.exp(-f(cur) / 1000)
    for _ in range(burnin):
        prop = cur + rng.normal4() * step
        p_new = np.exp(-f(prop) / 1000)
        if p_new > p_cur:
            cur, p_cur = prop, p_new
    p_cur = np.exp(-f(cur))
    for i in range(nsim):
        prop = cur + rng.normal4() * step
        p_cur = max(p_cur, 1e-99)
        p_new = np.exp(-f(prop))
        if rng.uniform() <= min(p_new / p_cur, 1.0):
            cur, p_cur = prop, p_new
        out[i] = cur
    return list(out.T)


# --- Philox4x32-10 host replica: the device sampler's random stream ---------
_PH_M0, _PH_M1 = 0xD2511F53, 0xCD9E8D57
_PH_W0, _PH_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """Salmon et al. 2011 Philox-4x32 with 10 rounds; ctr = 4 u32, key = 2 u32."""
    c = [int(v) & 0xFFFFFFFF for v in ctr]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PH_M0 * c[0]
        p1 = _PH_M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0 = (k0 + _PH_W0) & 0xFFFFFFFF
        k1 = (k1 + _PH_W1) & 0xFFFFFFFF
    return c


def _u53(hi, lo):
    """Open-interval (0,1) double from two u32 words: 52 random bits n = (hi<<20) ^ (lo>>12),
    u = (n + 0.5) * 2^-52  (exactly representable, never 0 or 1)."""
    return (((hi << 20) ^ (lo >> 12)) + 0.5) * (1.0 / 4503599627370496.0)


class PhiloxStream:
    """Random stream of the device MH kernel (pycurious_b200/csrc/mh.cu):
    key = (seed_lo ^ chain_lo, seed_hi ^ chain_hi); per step t three blocks
    ctr=(t_lo, t_hi, j, phase) j=0,1,2; blocks 0,1 -> four 52-bit uniforms ->
    two Box-Muller pairs (normals z0..z3), block 2 -> acceptance uniform.
    phase 0 = burn-in, 1 = sampling."""

    def __init__(self, seed, chain):
        s = int(seed) & 0xFFFFFFFFFFFFFFFF
        c = int(chain) & 0xFFFFFFFFFFFFFFFF
        self.key = ((s ^ c) & 0xFFFFFFFF, ((s >> 32) ^ (c >> 32)) & 0xFFFFFFFF)
        self.phase = 0
        self.t = 0
        self._pending_uniform = None

    def start_sampling(self):
        self.phase, self.t = 1, 0

    def _block(self, j):
        return philox4x32_10((self.t & 0xFFFFFFFF, (self.t >> 32) & 0xFFFFFFFF, j, self.phase), self.key)

    def normal4(self):
        a = self._block(0)
        b = self._block(1)
        u = [_u53(a[0], a[1]), _u53(a[2], a[3]), _u53(b[0], b[1]), _u53(b[2], b[3])]
        z = np.empty(4)
        for p in range(2):
            r = np.sqrt(-2.0 * np.log(u[2 * p]))
            ang = 2.0 * np.pi * u[2 * p + 1]
            z[2 * p] = r * np.cos(ang)
            z[2 * p + 1] = r * np.sin(ang)
        if self.phase == 0:
            self.t += 1          # burn-in steps draw no uniform
        else:
            c = self._block(2)
            self._pending_uniform = _u53(c[0], c[1])
            self.t += 1
        return z

    def uniform(self):
        return self._pending_uniform


def perturbed_spectra_philox(phi, sigma, nsim, seed, window=0):
    """Replica of the device's cb_perturb_spectra stream for one window: rows (nsim, nbins) of
    phi + sigma * z with z from Philox4x32-10, key (seed ^ window), counter (sim*nq + q, hi, j, phase 2)."""
    phi, sigma = np.asarray(phi, dtype=float), np.asarray(sigma, dtype=float)
    nk = phi.size
    nq = (nk + 3) // 4
    rng = PhiloxStream(seed, window)
    rng.phase = 2
    out = np.empty((nsim, nk))
    for sim in range(nsim):
        for q in range(nq):
            rng.t = sim * nq + q
            a, b = rng._block(0), rng._block(1)
            u = [_u53(a[0], a[1]), _u53(a[2], a[3]), _u53(b[0], b[1]), _u53(b[2], b[3])]
            for p in range(2):
                r = np.sqrt(-2.0 * np.log(u[2 * p]))
                ang = 2.0 * np.pi * u[2 * p + 1]
                for t, zz in ((2 * p, r * np.cos(ang)), (2 * p + 1, r * np.sin(ang))):
                    j = 4 * q + t
                    if j < nk:
                        out[sim, j] = phi[j] + sigma[j] * zz
    return out


def mh_chain_philox(kh, phi, nsim, burnin, seed, chain, **kw):
    """mh_chain driven by the device's Philox stream, so that device and oracle
    chains are step-comparable."""
    rng = PhiloxStream(seed, chain)

    class _Two:
        def __init__(s):
            s.n = 0

        def normal4(s):
            if s.n == burnin and rng.phase == 0:
                rng.start_sampling()
            s.n += 1
            return rng.normal4()

        def uniform(s):
            return rng.uniform()

    return mh_chain(kh, phi, nsim, burnin, rng=_Two(), **kw)


# --------------------------------------------------------------------------
# Radon transform / azimuthal spectrum  (src/radon.pyx, src/cradon.c, grid.py:384-403)
# --------------------------------------------------------------------------
def radon_ray_endpoints(nw, theta):
    """Tx, Rx (nw,2) for one angle, following src/radon.pyx:30-96 expression by
    expression (numpy scalars, so the rounding is the reference's)."""
    nx = ny = nw
    Tx = np.zeros((nw, 2))
    Rx = np.zeros((nw, 2))
    th = theta
    if th == 0.0:                                         # radon.pyx:32-37
        for n in range(nw):
            Tx[n] = (n, 0)
            Rx[n] = (n, ny - 1)
    elif np.abs(th - np.pi / 2) < 1.e-10:                 # radon.pyx:39-45
        for n in range(nw):
            Tx[n] = (0, n)
            Rx[n] = (nx - 1, n)
    elif th < np.pi / 2:                                  # radon.pyx:47-70
        xs = nw / 2 * (1. - np.cos(th))
        ys = nw / 2 * (1. - np.sin(th))
        for n in range(nw):
            xk = xs + n * np.cos(th)
            yk = ys + n * np.sin(th)
            x0 = 0.0
            y0 = yk + (xk - x0) / np.tan(th)
            if y0 > ny - 1:
                y0 = ny - 1
                x0 = xk - (y0 - yk) * np.tan(th)
            x1 = nx - 1
            y1 = yk - (x1 - xk) / np.tan(th)
            if y1 < 0.0:
                y1 = 0.0
                x1 = xk + (yk - y1) * np.tan(th)
            Tx[n] = (x0, y0)
            Rx[n] = (x1, y1)
    else:                                                 # radon.pyx:73-96
        xs = nw / 2 * (1. - np.cos(th))
        ys = nw / 2 * (1. - np.sin(th))
        tq = np.tan(th - np.pi / 2.0)
        for n in range(nw):
            xk = xs + n * np.cos(th)
            yk = ys + n * np.sin(th)
            x0 = nx - 1
            y0 = yk + (x0 - xk) * tq
            if y0 > ny - 1:
                y0 = ny - 1
                x0 = xk + (y0 - yk) / tq
            x1 = 0.0
            y1 = yk - (xk - x1) * tq
            if y1 < 0.0:
                y1 = 0.0
                x1 = xk - (yk - y1) / tq
            Tx[n] = (x0, y0)
            Rx[n] = (x1, y1)
    return Tx, Rx


_radon_lib = None


def _load_radon_lib():
    global _radon_lib
    if _radon_lib is None:
        path = os.path.join(_HERE, "libradon_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libradon_oracle.so missing: run `make -C oracle`")
        lib = ctypes.CDLL(path)
        lib.oracle_radon_trace.restype = ctypes.c_int64
        lib.oracle_radon_trace.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        _radon_lib = lib
    return _radon_lib


def radon_csr(Tx, Rx, nx, ny):
    """CSR (data f64, indices i64, indptr i64) of the ray/cell intersection
    lengths, computed by oracle/radon_oracle.c (restatement of src/cradon.c)."""
    lib = _load_radon_lib()
    Tx = np.ascontiguousarray(Tx, dtype=np.float64)
    Rx = np.ascontiguousarray(Rx, dtype=np.float64)
    nray = Tx.shape[0]
    indptr = np.empty(nray + 1, dtype=np.int64)
    nnz = lib.oracle_radon_trace(Tx.ctypes.data, Rx.ctypes.data, nray, nx, ny,
                                 None, None, indptr.ctypes.data, 0)
    data = np.empty(nnz, dtype=np.float64)
    indices = np.empty(nnz, dtype=np.int64)
    lib.oracle_radon_trace(Tx.ctypes.data, Rx.ctypes.data, nray, nx, ny,
                           data.ctypes.data, indices.ctypes.data, indptr.ctypes.data, nnz)
    return data, indices, indptr


def radon2d(win, theta):
    """Sinogram (nw, ntheta) as src/radon.pyx:12-120: CSR row . flattened data."""
    theta = np.asarray(theta, dtype=float)
    if np.min(theta) < 0.0 or np.max(theta) >= np.pi:
        raise ValueError('theta should be within [0 pi)')
    nx, ny = win.shape
    if nx != ny:
        raise RuntimeError('data should be a square array')
    from scipy.sparse import csr_matrix
    flat = np.asarray(win, dtype=float).flatten()
    sino = np.zeros((nx, theta.size))
    for t in range(theta.size):
        Tx, Rx = radon_ray_endpoints(nx, theta[t])
        L = csr_matrix(radon_csr(Tx, Rx, nx, ny), shape=(nx, nx * ny))   # radon.pyx:110-115
        sino[:, t] = L.dot(flat)
    return sino


def azimuthal_spectrum(win, dx, taper=np.hanning, power=2.0, theta=5.0, **kw):
    """grid.py:386-403."""
    nr, nc = win.shape
    _, dk, edges = taper_and_bins(nr, nc, dx, taper, **kw)
    dtheta = np.arange(0.0, 180.0, theta)
    sino = radon2d(win, np.pi * dtheta / 180.0)
    S = np.zeros((dtheta.size, edges.size))
    w1d = 1.0 if taper is None else taper(sino.shape[0], **kw)
    nk = 1 + 2 * edges.size
    for i in range(dtheta.size):
        psd = np.abs(np.fft.fft(w1d * sino[:, i], n=nk))
        S[i, :] = power * np.log(psd[1:edges.size + 1])
    return edges, S, dtheta


# --------------------------------------------------------------------------
# Tanaka et al. 1999  (pycurious/grid.py:603-699)  -- section 8(f) "next"
# --------------------------------------------------------------------------
def _weighted_line(X, Y, E):
    """compute_coefficients of grid.py:632-653 (weights are E**2 where E is
    already sigma**2: the reference's sigma^4 quirk)."""
    w = 1.0 / E ** 2
    sx, sy, s1 = np.sum(X * w), np.sum(Y * w), np.sum(w)
    sxx, sxy = np.sum(X ** 2 * w), np.sum(X * Y * w)
    Z = (sx * sy - sxy * s1) / (sx ** 2 - sxx * s1)
    b = (sxy - Z * sxx) / sx
    dZ = np.sqrt(s1 / (sxx * s1 - sx ** 2))
    return Z, b, dZ


def tanaka1999(k, lnPhi, sigma_lnPhi, kmin_range=(0.05, 0.2), kmax_range=(0.05, 0.2)):
    """grid.py:629-674."""
    var = sigma_lnPhi ** 2
    sf = k / (2.0 * np.pi)
    m1 = np.logical_and(sf >= kmin_range[0], sf <= kmin_range[1])
    m2 = np.logical_and(sf >= kmax_range[0], sf <= kmax_range[1])
    X2 = sf[m2]
    top = _weighted_line(sf[m1], lnPhi[m1], var[m1])
    cen = _weighted_line(X2, np.log(np.exp(lnPhi[m2]) / (X2 * 2 * np.pi)),
                         np.log(np.exp(var[m2]) / (X2 * 2 * np.pi)))
    return top, cen


def compute_tanaka(Ztr, dZtr, Zor, dZor):
    """grid.py:697-699."""
    return abs(2.0 * Zor - Ztr), 2.0 * dZor + dZtr


# --------------------------------------------------------------------------
# FFT-domain filters  (pycurious/grid.py:405-546, :738-808)
# --------------------------------------------------------------------------
def dircos(inc, dec):
    """grid.py:786-808: unit vector (x north, y east, z down) from inclination / declination in degrees."""
    d2r = np.pi / 180.0
    return [np.cos(d2r * inc) * np.cos(d2r * dec), np.cos(d2r * inc) * np.sin(d2r * dec), np.sin(d2r * inc)]


def fft_wavenumbers(dx, dy, shape):
    """_fftfreqs (grid.py:738-744): kx varies along axis 0 with spacing dx, ky along axis 1 with dy."""
    fx = 2.0 * np.pi * np.fft.fftfreq(shape[0], dx)
    fy = 2.0 * np.pi * np.fft.fftfreq(shape[1], dy)
    return np.meshgrid(fy, fx)[::-1]


def upward_continuation_filter(shape, dx, dy, height):
    """exp(-height |k|) of grid.py:538-545."""
    kx, ky = fft_wavenumbers(dx, dy, shape)
    return np.exp(-height * np.hypot(kx, ky))


def reduce_to_pole_filter(shape, dx, dy, inc, dec, sinc=None, sdec=None):
    """The RTP operator of grid.py:462-488, literally (including sqrt(|k|) in the imaginary part and the
    zeroed DC term)."""
    fx, fy, fz = dircos(inc, dec)
    if sinc is None or sdec is None:
        mx, my, mz = fx, fy, fz
    else:
        mx, my, mz = dircos(sinc, sdec)
    kx, ky = fft_wavenumbers(dx, dy, shape)
    kz = np.hypot(kx, ky)
    a1 = mz * fz - mx * fx
    a2 = mz * fz - my * fy
    a3 = -my * fx - mx * fy
    b1 = mx * fz + mz * fx
    b2 = my * fz + mz * fy
    with np.errstate(divide="ignore", invalid="ignore"):
        rtp = kz / (a1 * kx ** 2 + a2 * ky ** 2 + a3 * kx * ky + 1j * np.sqrt(kz) * (b1 * kx + b2 * ky))
    rtp[0, 0] = 0
    return rtp


def fft_filter(data, filt):
    """real(ifft2(filt * fft2(data))) (grid.py:489-490, :545-546)."""
    return np.real(np.fft.ifft2(filt * np.fft.fft2(data)))


# --------------------------------------------------------------------------
# synthetic fractal grid  (not in the reference)
# --------------------------------------------------------------------------
def synthetic_grid(ny, nx, seed=20261019, beta=3.0, zt=1.0, dz=20.0, std=100.0):
    """2-D spectral synthesis: white Gaussian field shaped by
    exp(0.5*bouligand2009(kr, beta, zt, dz, 0)), DC removed, rescaled to `std`."""
    rng = np.random.default_rng(seed)
    F = np.fft.fft2(rng.standard_normal((ny, nx)))
    kr = 2.0 * np.pi * np.hypot(np.fft.fftfreq(ny, 1.0)[:, None], np.fft.fftfreq(nx, 1.0)[None, :])
    kr[0, 0] = 1.0
    with np.errstate(all="ignore"):
        amp = np.exp(0.5 * bouligand2009(kr, beta, zt, dz, 0.0))
    amp[0, 0] = 0.0
    out = np.real(np.fft.ifft2(F * amp))
    return np.ascontiguousarray(out * (std / out.std()), dtype=np.float64)
