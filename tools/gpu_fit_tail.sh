#!/bin/bash
# the fit's tail on config 3: histogram of solver iterations / evaluations over the 152,100 windows, and the per-launch
# durations of the fit kernels on one rank's share at 8 GPUs (grid 1580 -> 19,044 windows).
OUT=gpurun_out; mkdir -p $OUT
python - <<'PY' 2>/dev/null | tee $OUT/fit_tail.txt
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import pycurious_b200 as pc
from pycurious_b200 import synth
n = 4096
G = pc.CurieOptimise(synth.fractal_grid(n, n, seed=20261019), *synth.grid_extent(n, n))
xc, yc = G.create_centroid_list(200e3, 10e3, 10e3)
x, d = G.optimise_batch(200e3, xc, yc, diagnostics=True)
it, nf = d["niter"], d["nfev"]
print("windows", it.size, "niter mean %.2f max %d" % (it.mean(), it.max()), "nfev mean %.2f max %d" % (nf.mean(), nf.max()))
edges = [0, 4, 6, 8, 10, 12, 14, 16, 20, 24, 32, 48, 64, 100, 201]
print("niter histogram", dict(zip(["<%d" % e for e in edges[1:]], np.histogram(it, bins=edges)[0].tolist())))
edges = [0, 12, 18, 24, 30, 36, 45, 60, 75, 100, 150, 200, 300, 1000]
print("nfev histogram", dict(zip(["<%d" % e for e in edges[1:]], np.histogram(nf, bins=edges)[0].tolist())))
PY
GRID=1580 bash tools/gpu_fit_launches.sh tail1580 | tee -a $OUT/fit_tail.txt
