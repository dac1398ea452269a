"""
Golden vectors for BASELINE config 3 from the REAL reference (build container only).

    python tests/golden/make_golden_cfg3.py [nsample]

The unmodified reference package is imported from /root/reference and run through its own public API
(`CurieOptimise.subgrid` / `radial_spectrum` / `optimise`) on an evenly spaced sample of the 152,100
windows of config 3 (synthetic 4096 x 4096 grid of pycurious_b200.synth, 200 km windows, 10 km
stride).  For each sampled window the file holds

  phi, sig            radial_spectrum of the reference (k once)
  xref, fref          `optimise` of the reference (scipy L-BFGS-B as pycurious/optimise.py:300 calls
                      it) and `min_func` there
  xpol, fpol          the tightly converged bounded minimum next to xref (oracle.polished_fit):
                      tells whether the reference's own stopping point is converged
  xpert[s]            REFERENCE SELF-SENSITIVITY: the same reference call on the same spectrum after
                      a relative perturbation of 1e-13 * N(0,1) -- the size of the difference
                      between two correct FP64 implementations of the spectrum (this repo's spectra
                      agree with numpy's to < 1e-13) -- for three seeds.  Where |xpert - xref| is
                      not small, the reference's stopping point is not a reproducible function of
                      its input, so no implementation can match it to 1e-4 there.

Nothing of the reference is copied: only numerical outputs are stored (tests/golden/cfg3.npz).
"""
import multiprocessing as mp
import os
import sys
import warnings

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PYCURIOUS_REF", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.simplefilter("ignore")

GRID_N, WINDOW, STRIDE, SEED = 4096, 200e3, 10e3, 20261019
NPERT, PERT_REL = 3, 1e-13
_S = {}


def _init():
    import types
    # the reference imports its compiled radon extension at module level; it is not on this path
    sys.modules.setdefault("pycurious.radon", types.ModuleType("pycurious.radon"))
    import pycurious
    from pycurious_b200 import synth
    data = synth.fractal_grid(GRID_N, GRID_N, seed=SEED)
    _S["G"] = pycurious.CurieOptimise(data, *synth.grid_extent(GRID_N, GRID_N))


def _one(job):
    from scipy.optimize import minimize
    import curie_oracle as orc
    i, xc, yc = job
    G = _S["G"]
    sub = G.subgrid(WINDOW, xc, yc)
    k, phi, sig = G.radial_spectrum(sub)
    xref = np.asarray(G.optimise(WINDOW, xc, yc))
    fref = G.min_func(xref, k, phi, sig)
    xpol, fpol = orc.polished_fit(k, phi, xref)
    rng = np.random.default_rng(1000003 * (i + 1))
    xpert = np.empty((NPERT, 4))
    for s in range(NPERT):
        phi_s = phi * (1.0 + PERT_REL * rng.standard_normal(phi.size))
        # pycurious/optimise.py:298-301 on the perturbed spectrum
        xpert[s] = minimize(G.min_func, np.array([3.0, 1.0, 10.0, 5.0]), args=(k, phi_s, sig), bounds=G.bounds).x
    return i, k, phi, sig, xref, fref, xpol, fpol, xpert


def main():
    nsample = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    _init()
    G = _S["G"]
    xc, yc = G.create_centroid_list(WINDOW, STRIDE, STRIDE)
    idx = np.linspace(0, xc.size - 1, nsample).astype(np.int64)
    jobs = [(int(i), float(xc[i]), float(yc[i])) for i in idx]
    cores = len(os.sched_getaffinity(0))
    with mp.get_context("fork").Pool(cores) as pool:       # workers inherit _S
        res = pool.map(_one, jobs, chunksize=2)
    out = dict(grid_n=GRID_N, window=WINDOW, stride=STRIDE, seed=SEED, nwin_total=xc.size, pert_rel=PERT_REL,
               idx=idx, xc=xc[idx], yc=yc[idx], k=res[0][1],
               phi=np.stack([r[2] for r in res]), sig=np.stack([r[3] for r in res]),
               xref=np.stack([r[4] for r in res]), fref=np.array([r[5] for r in res]),
               xpol=np.stack([r[6] for r in res]), fpol=np.array([r[7] for r in res]),
               xpert=np.stack([r[8] for r in res]))
    path = os.path.join(HERE, "cfg3.npz")
    np.savez_compressed(path, **out)
    print("wrote cfg3.npz %.1f KB, %d windows" % (os.path.getsize(path) / 1024, nsample))


if __name__ == "__main__":
    main()
