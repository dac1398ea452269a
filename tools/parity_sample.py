#!/usr/bin/env python
"""Fit-parity accounting at BASELINE size (config 3): the CUDA path fits all 152,100 windows; the 512 windows
of tests/golden/cfg3.npz (outputs of the UNMODIFIED reference, tests/golden/make_golden_cfg3.py) are compared
parameter by parameter.  Reports, per parameter and for the Curie depth zt + dz, over ALL sampled windows:
the fraction within 1e-4 relative and the p50 / p90 / p99 / max relative deviation of

  ours_vs_ref        this repo's fit against the reference's `optimise`
  ref_vs_polished    the reference against the tightly converged minimum next to it (is it converged?)
  ref_vs_ref_pert    REFERENCE SELF-SENSITIVITY: the reference against itself after a 1e-13 relative
                     perturbation of its input spectrum (three seeds, worst of three)
  ours_vs_polished   this repo's fit against the polished minimum

    python tools/parity_sample.py [extra_oracle_windows] > gpurun_out/r02_parity_config3.json

With `extra_oracle_windows` > 0 an additional evenly spaced sample is re-done by the oracle port on the host
cores (map_fit + polished_fit + one perturbed re-run) and summarised the same way.
"""
import json
import multiprocessing as mp
import os
import sys

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import curie_oracle as orc
from pycurious_b200 import synth

NAMES = ("beta", "zt", "dz", "C", "curie_depth")
_S = {}


def _five(x):
    x = np.asarray(x)
    return np.concatenate([x, (x[..., 1] + x[..., 2])[..., None]], axis=-1)


def _dev(a, b):
    """relative deviation per column (beta, zt, dz, C, zt+dz) of a against b"""
    a5, b5 = _five(a), _five(b)
    return np.abs(a5 - b5) / np.maximum(np.abs(b5), 1e-3)


def _summary(d):
    """d: (n, 5) relative deviations -> per-name stats + the all-parameter gate"""
    out = {}
    for j, nm in enumerate(NAMES):
        c = d[:, j]
        out[nm] = {"within_1e-4": float(np.mean(c < 1e-4)), "p50": float(np.percentile(c, 50)),
                   "p90": float(np.percentile(c, 90)), "p99": float(np.percentile(c, 99)), "max": float(c.max())}
    w = d[:, :4].max(axis=1)
    out["all_four"] = {"within_1e-4": float(np.mean(w < 1e-4)), "p50": float(np.percentile(w, 50)),
                       "p90": float(np.percentile(w, 90)), "p99": float(np.percentile(w, 99)), "max": float(w.max())}
    return out


def _init(data, ext):
    _S["spec"] = orc.GridSpec(data, *ext)


def _one(job):
    i, xc, yc, window = job
    spec = _S["spec"]
    k, phi, sig = orc.radial_spectrum(orc.cut_window(spec, window, xc, yc), spec.dx)
    ref = orc.map_fit(k, phi)
    xp, fp = orc.polished_fit(k, phi, ref)
    rng = np.random.default_rng(1000003 * (i + 1))
    pert = orc.map_fit(k, phi * (1.0 + 1e-13 * rng.standard_normal(phi.size)))
    return phi, ref, orc.misfit(ref, k, phi), xp, fp, pert


def main():
    extra = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg3.npz"))
    N, WINDOW, STRIDE = int(g["grid_n"]), float(g["window"]), float(g["stride"])
    data = synth.fractal_grid(N, N, seed=int(g["seed"]))
    ext = synth.grid_extent(N, N)
    spec = orc.GridSpec(data, *ext)
    xc, yc = orc.centroid_list(spec, WINDOW, STRIDE, STRIDE)
    cpu = None
    if extra > 0:                       # before CUDA is initialised: the workers are forked
        idx2 = np.linspace(0, xc.size - 1, extra).astype(np.int64)
        with mp.get_context("fork").Pool(len(os.sched_getaffinity(0)), initializer=_init, initargs=(data, ext)) as pool:
            cpu = pool.map(_one, [(int(i), float(xc[i]), float(yc[i]), WINDOW) for i in idx2], chunksize=4)

    import pycurious_b200 as pc
    G = pc.CurieOptimise(data, *ext)
    k, phi_g, sig_g = G.radial_spectrum_batch(WINDOW, xc, yc)
    x_g, diag = G.optimise_batch(WINDOW, xc, yc, diagnostics=True)
    idx = g["idx"]
    ours, xref, xpol, xpert = x_g[idx], g["xref"], g["xpol"], g["xpert"]
    f_ours = np.array([orc.misfit(ours[i], g["k"], g["phi"][i]) for i in range(idx.size)])
    d_ref, d_pol, d_refpol = _dev(ours, xref), _dev(ours, xpol), _dev(xref, xpol)
    d_self = np.max(np.stack([_dev(xpert[:, s], xref) for s in range(xpert.shape[1])]), axis=0)
    conv = d_refpol[:, :4].max(axis=1) < 3e-5
    off = d_ref[:, :4].max(axis=1) >= 1e-4
    out = {
        "source": "tests/golden/cfg3.npz: unmodified reference (pycurious 1.1.1, scipy L-BFGS-B) on 512 evenly spaced "
                  "windows of BASELINE config 3; device run over all windows",
        "windows_total": int(xc.size), "sample": int(idx.size),
        "k_bit_exact": bool(np.array_equal(k, g["k"])),
        "phi_max_rel": float(np.max(np.abs(phi_g[idx] - g["phi"]) / np.abs(g["phi"]))),
        "sigma_max_rel": float(np.max(np.abs(sig_g[idx] - g["sig"]) / np.abs(g["sig"]))),
        "device_status_counts": [int((diag["status"] == s).sum()) for s in range(5)],
        "reference_converged": int(conv.sum()),
        "within_1e-4_of_reference_where_converged": int(np.sum(~off & conv)),
        "never_worse_than_reference_objective": int(np.sum(f_ours <= g["fref"] * (1 + 1e-9))),
        "max_(f-f_ref)/f_ref": float(np.max((f_ours - g["fref"]) / np.abs(g["fref"]))),
        "max_(f-f_polished)/f_polished": float(np.max((f_ours - g["fpol"]) / np.abs(g["fpol"]))),
        "ours_vs_ref": _summary(d_ref), "ref_vs_polished": _summary(d_refpol),
        "ref_vs_ref_pert_1e-13": _summary(d_self), "ours_vs_polished": _summary(d_pol),
        "windows_off_by_more_than_1e-4": int(off.sum()),
        "of_those_reference_not_converged": int(np.sum(off & ~conv)),
        "of_those_reference_moves_more_than_1e-4_against_itself": int(np.sum(off & (d_self[:, :4].max(axis=1) > 1e-4))),
        "off_windows_median_dev": {"ours_vs_ref": float(np.median(d_ref[off, :4].max(axis=1))) if off.any() else None,
                                   "ref_vs_ref_pert": float(np.median(d_self[off, :4].max(axis=1))) if off.any() else None},
    }
    if cpu is not None:
        ref2 = np.stack([c[1] for c in cpu]); pol2 = np.stack([c[3] for c in cpu]); pert2 = np.stack([c[5] for c in cpu])
        ours2 = x_g[idx2]
        f2 = np.array([orc.misfit(ours2[j], k, cpu[j][0]) for j in range(idx2.size)])
        fr2 = np.array([c[2] for c in cpu])
        out["oracle_sample"] = {
            "sample": int(idx2.size), "never_worse_than_reference_objective": int(np.sum(f2 <= fr2 * (1 + 1e-9))),
            "ours_vs_ref": _summary(_dev(ours2, ref2)), "ref_vs_polished": _summary(_dev(ref2, pol2)),
            "ref_vs_ref_pert_1e-13": _summary(_dev(pert2, ref2)), "ours_vs_polished": _summary(_dev(ours2, pol2))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
