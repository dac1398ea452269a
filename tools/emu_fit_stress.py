#!/usr/bin/env python
"""Robustness sweep of the fit kernel body under CPU emulation: every golden fit case (and optionally
the synthetic dev fixture) is refitted from `nrep` copies of its spectrum perturbed at the 1e-13 level
-- the size of the differences between the CUDA and the CPU math libraries -- and the outcomes are
compared with the polished minimum.  Catches solver paths that depend on rounding (stalls, traps on
the flat dz tail, iteration-limit exits)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_emu_kernels import FitOpt, P
emu = C.CDLL(os.environ.get("CB_EMU_LIB") or os.path.join(ROOT, "tests/emu/libcurie_emu.so"))
g = np.load(os.path.join(ROOT, "tests/golden/fit.npz"))
nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(1)
worst = 0
for tag in ("cfg1", "w200", "w150_prior", "w300_betaprior"):
    k, phi0 = g[tag + "_k"], g[tag + "_phi"]
    mu, sg, x0 = g[tag + "_prior_mu"], g[tag + "_prior_sigma"], g[tag + "_x0"]
    fpol = g[tag + "_fpol"]
    nw = phi0.shape[0]
    phi = np.repeat(phi0, nrep, axis=0) * (1 + 1e-13 * rng.standard_normal((nw * nrep, phi0.shape[1])))
    phi = np.ascontiguousarray(phi)
    o = FitOpt()
    for j in range(4):
        o.x0[j], o.lb[j], o.ub[j] = x0[j], [0, 0, 0, -np.inf][j], np.inf
        o.prior_mu[j], o.prior_sigma[j] = mu[j], sg[j]
    n = phi.shape[0]
    x, f = np.empty((n, 4)), np.empty(n)
    it, st, nf = (np.empty(n, dtype=np.int32) for _ in range(3))
    emu.emu_fit(P(k), k.size, P(phi), C.c_int64(n), C.byref(o), None, P(x), P(f), P(it), P(st), P(nf), 0)
    ref = np.repeat(fpol, nrep)
    excess = (f - ref) / np.abs(ref)
    bad = np.flatnonzero(excess > 1e-8)
    print("%-16s %4d fits: nfev mean %.1f max %d, status %s, max (f-fpol)/fpol %.2e, bad %d" % (
        tag, n, nf.mean(), nf.max(), np.bincount(st, minlength=5), excess.max(), bad.size))
    for b in bad[:5]:
        print("    window", b // nrep, x[b], f[b], ref[b], it[b], nf[b], st[b])
    worst = max(worst, bad.size)
sys.exit(1 if worst else 0)
