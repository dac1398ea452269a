#!/bin/bash
# filtered-window path: time of the window filter and of the direct spectrum kernel on the stack
python - <<'PY' 2>/dev/null
import os, sys, functools
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
    return a.elapsed_time(b) / reps, out
for window in (200e3, 100e3):
    xc, yc = G.create_centroid_list(window, 10e3, 10e3)
    sel = np.linspace(0, xc.size - 1, 20000).astype(int)
    iy0, ix0, nw = G._origins_device(window, xc[sel], yc[sel])
    filt = G._window_filter(functools.partial(G.upward_continuation, height=1000.0), nw)
    tf, stack = timed(lambda: G._filter_windows_device(filt, g, iy0, ix0, nw))
    plan = _plans.spectrum_plan(nw, G.dx, _plans.taper_vector(np.hanning, nw))
    sy = torch.arange(20000, dtype=torch.int32, device="cuda") * nw
    sx = torch.zeros(20000, dtype=torch.int32, device="cuda")
    ts, _ = timed(lambda: plan.run(stack.view(20000 * nw, nw), sy, sx, False, 2.0, share_rows=False))
    print(int(window / 1e3), "km: filter %.2f ms, direct spectrum %.2f ms for 20000 windows" % (tf, ts))
PY
