#!/usr/bin/env python
"""BASELINE config 5 at its real size: an EMAG2-shaped 10800 x 21600 grid (1.87 GB), 100 / 200 / 300 km windows
on a 20 km stride (575,125 / 567,100 / 559,125 centroids), spectrum + MAP fit per window through
CurieOptimise.optimise_batch on one B200.  Prints one JSON object.
    gpurun -- python tools/measure_config5_full.py > gpurun_out/config5_full.json"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth

ny, nx = 10800, 21600
t0 = time.time()
data = synth.fractal_grid(ny, nx, seed=20261019)
t_synth = time.time() - t0
G = pc.CurieOptimise(data, *synth.grid_extent(ny, nx))
t0 = time.time()
G.device_grid(); torch.cuda.synchronize()
res = {"grid": [ny, nx], "synth_s": t_synth, "upload_s": time.time() - t0}
for window in (100e3, 200e3, 300e3):
    xs, ys = G.centroid_lattice(window, 20e3, 20e3)
    run = lambda: G.curie_depth_maps(window, 20e3, 20e3, as_numpy=False)
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m = run(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    n = xs.size * ys.size
    dz = m["dz"]
    res["%dkm" % int(window / 1e3)] = {"centroids": n, "lattice": [int(ys.size), int(xs.size)], "ms": ms,
                                       "windows_per_s": n / ms * 1e3,
                                       "dz_median": float(dz.median()), "beta_median": float(m["beta"].median()),
                                       "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
print(json.dumps(res, indent=1))
