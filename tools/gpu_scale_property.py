"""Amplitude-scaling property of the fit over ALL windows of config 3 (tests/test_gpu_config3.py, property (c)):
how far a rounding-level change of the spectra moves this repo's answer."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import pycurious_b200 as pc
from pycurious_b200 import synth
n = 4096
data = synth.fractal_grid(n, n, seed=20261019)
G = pc.CurieOptimise(data, *synth.grid_extent(n, n))
xc, yc = G.create_centroid_list(200e3, 10e3, 10e3)
full = G.optimise_batch(200e3, xc, yc)
out = {"windows": int(xc.size)}
for a in (4.0, 7.5):
    sc = pc.CurieOptimise(data * a, *synth.grid_extent(n, n)).optimise_batch(200e3, xc, yc)
    ex = full.copy(); ex[:, 3] += 2.0 * np.log(a)
    d = np.max(np.abs(sc - ex) / np.maximum(np.abs(ex), 1e-3), axis=1)
    out["scale_%g" % a] = {"within_1e-4": float(np.mean(d < 1e-4)), "within_1e-6": float(np.mean(d < 1e-6)),
                          "within_1e-8": float(np.mean(d < 1e-8)), "p50": float(np.median(d)), "p99": float(np.quantile(d, 0.99)),
                          "p999": float(np.quantile(d, 0.999)), "max": float(d.max())}
print(json.dumps(out))
