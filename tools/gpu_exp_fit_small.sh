#!/bin/bash
# the fit on ONE rank's share of config 3 at 8 GPUs (19,012 windows): CB_FIT_HANDOVER x CB_FIT_ROUNDS sweep, ms per call.
OUT=gpurun_out; mkdir -p $OUT
python - <<'PY' 2>/dev/null | tee $OUT/fit_small.txt
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth, _plans
n = 4096
G = pc.CurieOptimise(synth.fractal_grid(n, n, seed=20261019), *synth.grid_extent(n, n)).hold_device()
g = G.device_grid()
xc, yc = G.create_centroid_list(200e3, 10e3, 10e3)
for share in (8, 4, 2):
    m = xc.size // share
    iy0, ix0, nw = G._origins_device(200e3, xc[:m], yc[:m])
    plan = _plans.spectrum_plan(nw, G.dx, _plans.taper_vector(np.hanning, nw))
    phi, sig = plan.run(g, iy0, ix0, False, 2.0)
    torch.cuda.synchronize()
    def timed(reps=5):
        G._fit_device(plan.k, phi, (3.0, 1.0, 10.0, 5.0)); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): out = G._fit_device(plan.k, phi, (3.0, 1.0, 10.0, 5.0))
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    os.environ.pop("CB_FIT_HANDOVER", None); os.environ.pop("CB_FIT_ROUNDS", None)
    print("windows", m, "default schedule", "fit %.3f ms" % timed(), flush=True)
    for ho, rounds in ((2048, 20), (2048, 16), (1024, 12), (1024, 10), (1024, 14), (2048, 14)):
        os.environ["CB_FIT_HANDOVER"] = str(ho); os.environ["CB_FIT_ROUNDS"] = str(rounds)
        print("windows", m, "handover", ho, "rounds", rounds, "fit %.3f ms" % timed(), flush=True)
PY
