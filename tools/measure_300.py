#!/usr/bin/env python
"""Spectrum-only and fit-only timings of the 300 km (nw=301, Bluestein 640) windows of config 5 on a 4096^2 grid."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth, _plans
n = 4096
G = pc.CurieOptimise(synth.fractal_grid(n, n, seed=20261019), *synth.grid_extent(n, n))
g = G.device_grid()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out
for window in (100e3, 300e3):
    xc, yc = G.create_centroid_list(window, 20e3, 20e3)
    iy0, ix0, nw = G._origins_device(window, xc, yc)
    plan = _plans.spectrum_plan(nw, G.dx, _plans.taper_vector(np.hanning, nw))
    ms_s, (phi, sig) = timed(lambda: plan.run(g, iy0, ix0, False, 2.0))
    ms_f, _ = timed(lambda: G._fit_device(plan.k, phi, (3.0, 1.0, 10.0, 5.0)))
    print(int(window / 1e3), "km windows", xc.size, "spectrum %.2f ms" % ms_s, "fit %.2f ms" % ms_f, plan.last_path)
