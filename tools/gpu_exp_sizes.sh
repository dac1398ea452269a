#!/bin/bash
# spectrum throughput for window sizes off the BASELINE list (plan choice: static-engine lengths preferred)
python - <<'PY' 2>/dev/null
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth, _plans
n = 2048
G = pc.CurieOptimise(synth.fractal_grid(n, n, seed=20261019), *synth.grid_extent(n, n))
g = G.device_grid()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for nodes in (75, 111, 151, 175, 225, 251, 281, 311):
    window = (nodes - 1) * 1000.0
    xc, yc = G.create_centroid_list(window, 10e3, 10e3)
    iy0, ix0, nw = G._origins_device(window, xc, yc)
    plan = _plans.spectrum_plan(nw, G.dx, _plans.taper_vector(np.hanning, nw))
    ms = timed(lambda: plan.run(g, iy0, ix0, False, 2.0))
    print(nw, "nodes: fft", plan.info().fft_len, "%d windows in %.2f ms = %.0f windows/s" % (xc.size, ms, xc.size / ms * 1e3))
PY
