#!/usr/bin/env python
"""Run the fit kernel body under CPU emulation on the dev fixture of tools/make_synth_spectra.py and
report evaluation counts and agreement with the reference L-BFGS-B result."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from test_emu_kernels import FitOpt, P
import curie_oracle as orc
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests/emu/libcurie_emu.so")
emu = C.CDLL(lib)
g = np.load(sys.argv[1])
k, phi = g["k"], np.ascontiguousarray(g["phi"])
o = FitOpt()
for j in range(4):
    o.x0[j] = [3.0, 1.0, 10.0, 5.0][j]; o.lb[j] = [0, 0, 0, -np.inf][j]; o.ub[j] = np.inf
    o.prior_mu[j] = np.nan; o.prior_sigma[j] = np.nan
n = phi.shape[0]
x, f = np.empty((n, 4)), np.empty(n)
it, st, nf = (np.empty(n, dtype=np.int32) for _ in range(3))
emu.emu_fit(P(k), k.size, P(phi), C.c_int64(n), C.byref(o), None, P(x), P(f), P(it), P(st), P(nf), 0)
rel = np.max(np.abs(x - g["xref"]) / np.maximum(np.abs(g["xref"]), 1e-3), axis=1)
print("nfev mean %.1f max %d | iters mean %.1f max %d | status counts %s" % (nf.mean(), nf.max(), it.mean(), it.max(), np.bincount(st, minlength=5)))
print("f <= fref(1+1e-9): %d/%d ; max (f-fref)/fref %.3e ; rel param diff median %.2e, <1e-4: %d" % (
    np.sum(f <= g["fref"] * (1 + 1e-9)), n, np.max((f - g["fref"]) / g["fref"]), np.median(rel), np.sum(rel < 1e-4)))
if "-v" in sys.argv:
    for i in range(n):
        print(i, x[i], f[i], g["fref"][i], it[i], nf[i], st[i], "%.2e" % rel[i])
