#!/bin/bash
# azimuthal (Radon) spectrum throughput for a centroid batch
python - <<'PY' 2>/dev/null
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth
n = 2048
G = pc.CurieOptimise(synth.fractal_grid(n, n, seed=20261019), *synth.grid_extent(n, n))
G.device_grid()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out
for window in (100e3, 200e3):
    xc, yc = G.create_centroid_list(window, 10e3, 10e3)
    sel = np.linspace(0, xc.size - 1, 4000).astype(int)
    ms, out = timed(lambda: G.azimuthal_spectrum_batch(window, xc[sel], yc[sel], theta=5.0))
    print(int(window / 1e3), "km: 4000 windows x 36 angles in %.2f ms = %.0f windows/s" % (ms, 4000 / ms * 1e3))
PY
