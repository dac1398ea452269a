#!/usr/bin/env python
"""One process, N GPUs through the reference's own knob: CurieOptimise(..., max_processors=N).optimise_routine on
BASELINE config 3 (no torchrun), against max_processors=1.  Prints windows/s end to end and checks equality."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
N, WINDOW, STRIDE = 4096, 200e3, 10e3
data = synth.fractal_grid(N, N, seed=20261019)
ext = synth.grid_extent(N, N)
out = {"gpus_visible": torch.cuda.device_count()}
res = {}
for m in (1, ndev):
    G = pc.CurieOptimise(data, *ext, max_processors=m)
    xc, yc = G.create_centroid_list(WINDOW, STRIDE, STRIDE)
    G.optimise_routine(WINDOW, xc, yc)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        r = G.optimise_routine(WINDOW, xc, yc)
        ts.append(time.perf_counter() - t0)
    res[m] = np.stack(r, axis=1)
    out["max_processors_%d" % m] = {"ms": 1e3 * min(ts), "median_ms": 1e3 * float(np.median(ts)), "windows_per_s": xc.size / min(ts)}
out["equal"] = bool(np.array_equal(res[1], res[ndev]))
out["speedup"] = out["max_processors_1"]["ms"] / out["max_processors_%d" % ndev]["ms"]
print(json.dumps(out, indent=1))
