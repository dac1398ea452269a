#!/usr/bin/env python
"""BASELINE config 4 at its full size: one Metropolis-Hastings chain of nsim = 100,000 samples after 1,000 burn-in
steps for every one of the 152,100 windows of config 3 (1.54e10 steps), priors beta=(3,1), zt=(1,1), dz=(20,10),
C=(mean MAP C, 5), x_scale=[0.25, 0.1, 1.0, 0.5], chains started at the per-window MAP estimate, moments streamed
on the device (CurieOptimise.curie_depth_maps(uncertainty='mcmc')).  Prints one JSON object.
    gpurun -- python tools/measure_config4_full.py > gpurun_out/config4_full.json"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth

n = 4096
nsim, burnin = int(os.environ.get("NSIM", 100000)), 1000
G = pc.CurieOptimise(synth.fractal_grid(n, n, seed=20261019), *synth.grid_extent(n, n))
m0 = G.curie_depth_maps(200e3, 10e3, 10e3, as_numpy=False)
c_mean = float(m0["C"].mean())
G.add_prior(beta=(3.0, 1.0), zt=(1.0, 1.0), dz=(20.0, 10.0), C=(c_mean, 5.0))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
m = G.curie_depth_maps(200e3, 10e3, 10e3, uncertainty="mcmc", nsim=nsim, burnin=burnin, x_scale=[0.25, 0.1, 1.0, 0.5],
                       seed=20261019, as_numpy=False)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
nwin = int(m["beta"].numel())
steps = nwin * (nsim + burnin)
acc = G.last_accepted.double() / nsim
fin = torch.stack([torch.isfinite(v).all() for k, v in m.items() if hasattr(v, "numel")]).all()
print(json.dumps({
    "windows": nwin, "nsim": nsim, "burnin": burnin, "steps": steps, "seconds_spectrum_fit_mcmc": ms * 1e-3,
    "mh_steps_per_s": steps / (ms * 1e-3), "acceptance_mean": float(acc.mean()), "acceptance_min": float(acc.min()),
    "all_finite": bool(fin),
    "curie_depth_posterior_mean_median": float(m["curie_depth_mean"].median()),
    "curie_depth_posterior_std_median": float(m["curie_depth_std"].median()),
    "beta_posterior_mean_median": float(m["beta_mean"].median()), "dz_posterior_mean_median": float(m["dz_mean"].median()),
    "prior_C_mean": c_mean}, indent=1))
