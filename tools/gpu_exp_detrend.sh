#!/bin/bash
# spectrum-stage timings with and without the fused detrend (config 3 and config 5 / 300 km), one box
python - <<'PY' 2>/dev/null
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
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
    return a.elapsed_time(b) / reps
for window, stride in ((200e3, 10e3), (100e3, 20e3)):
    xc, yc = G.create_centroid_list(window, stride, stride)
    iy0, ix0, nw = G._origins_device(window, xc, yc)
    plan = _plans.spectrum_plan(nw, G.dx, _plans.taper_vector(np.hanning, nw))
    print(int(window / 1e3), "km", xc.size, "windows: plain %.2f ms, detrend %.2f ms" % (
        timed(lambda: plan.run(g, iy0, ix0, False, 2.0)), timed(lambda: plan.run(g, iy0, ix0, True, 2.0))))
PY
