#!/usr/bin/env python
"""Work units behind the FP64 instruction counts (run under ncu, see tools/gpu_visit.sh):

    ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum --csv --log-file counts.csv \
        python tools/count_fp64.py > units.json

One launch of each hot kernel on a 1024^2 synthetic grid (6,889 windows of 201 x 201, the config-3 geometry) and
the number of work units it processed: windows (rowspec + colspec), forward-model points (fit: sum of nfev x
99 bins), Metropolis-Hastings steps (mh).  tools/fp64_counts_from_ncu.py divides the counted thread-level
FP64 instructions (2 x DFMA + DADD + DMUL) by these units -> profiles/fp64_counts.json.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth, _plans

N, WINDOW, STRIDE = 1024, 200e3, 10e3
data = synth.fractal_grid(N, N, seed=20261019)
G = pc.CurieOptimise(data, *synth.grid_extent(N, N)).hold_device()
xc, yc = G.create_centroid_list(WINDOW, STRIDE, STRIDE)
g, iy0, ix0, n = G._windows_device(WINDOW, xc, yc)
plan = _plans.spectrum_plan(n, G.dx, _plans.taper_vector(np.hanning, n))
phi, sig = plan.run(g, iy0, ix0, False, 2.0)
x, diag = G._fit_device(plan.k, phi, (3.0, 1.0, 10.0, 5.0), diagnostics=True)
G.add_prior(beta=(3.0, 1.0), zt=(1.0, 1.0), dz=(20.0, 10.0), C=(float(x[:, 3].mean().item()), 5.0))
nsim, burn = 4000, 40
mom, acc = G._mh_device(plan.k, phi, nsim, burn, [0.25, 0.1, 1.0, 0.5], (3.0, 1.0, 10.0, 5.0), x, 20261019, 1)
torch.cuda.synchronize()
print(json.dumps({"windows": int(xc.size), "window_nodes": int(n), "nbins": int(plan.nbins), "path": plan.last_path,
                  "fit_nfev_sum": int(diag["nfev"].sum().item()), "fit_points": int(diag["nfev"].sum().item()) * int(plan.nbins),
                  "mh_steps": int(xc.size) * (nsim + burn), "mh_acceptance": float(acc.double().mean().item() / nsim)}))
