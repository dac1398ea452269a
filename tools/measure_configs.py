#!/usr/bin/env python
"""Throughput of the BASELINE configurations that are not the bench line (bench.py measures config 3):
config 2 (Tanaka radial_spectrum-only sweep, 128 km windows, 8 km stride, + batched Tanaka fits),
config 3 with the fused detrend, and the three window sizes of config 5 on a 4096^2 grid at 20 km stride.
Prints one JSON object; run on a B200 (`gpurun -- python tools/measure_configs.py > gpurun_out/configs.json`)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth, _plans

n = 4096
data = synth.fractal_grid(n, n, seed=20261019)
ext = synth.grid_extent(n, n)
G = pc.CurieOptimise(data, *ext)
g = G.device_grid()
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


res = {}
# config 2
xc, yc = G.create_centroid_list(128e3, 8e3, 8e3)
iy0, ix0, nw = G._origins_device(128e3, xc, yc)
plan = _plans.spectrum_plan(nw, G.dx, _plans.taper_vector(np.hanning, nw))
ms, _ = timed(lambda: plan.run(g, iy0, ix0, False, 0.5))
res["config2_spectrum"] = {"windows": int(xc.size), "nw": nw, "ms": ms, "windows_per_s": xc.size / ms * 1e3,
                           "fft_len": plan.info().fft_len, "path": plan.last_path}
ms, out = timed(lambda: G.tanaka_batch(128e3, xc, yc, kmin_range=(0.005, 0.03), kmax_range=(0.03, 0.7), as_numpy=False))
res["config2_spectrum_plus_tanaka"] = {"windows": int(xc.size), "ms": ms, "windows_per_s": xc.size / ms * 1e3}
# config 3 with detrend
xc, yc = G.create_centroid_list(200e3, 10e3, 10e3)
iy0, ix0, nw = G._origins_device(200e3, xc, yc)
plan = _plans.spectrum_plan(nw, G.dx, _plans.taper_vector(np.hanning, nw))
ms, _ = timed(lambda: plan.run(g, iy0, ix0, True, 2.0))
res["config3_spectrum_detrend"] = {"windows": int(xc.size), "ms": ms, "windows_per_s": xc.size / ms * 1e3}
# config 5 window sizes
for window in (100e3, 200e3, 300e3):
    xc, yc = G.create_centroid_list(window, 20e3, 20e3)
    ms, x = timed(lambda: G.optimise_batch(window, xc, yc, as_numpy=False))
    res["config5_%dkm_fit" % int(window / 1e3)] = {"windows": int(xc.size), "ms": ms, "windows_per_s": xc.size / ms * 1e3}
# sensitivity: 64 simulations for 2,000 windows
xc, yc = G.create_centroid_list(200e3, 10e3, 10e3)
sel = np.linspace(0, xc.size - 1, 2000).astype(int)
ms, mom = timed(lambda: G.sensitivity_batch(200e3, xc[sel], yc[sel], 64, reduce="moments", seed=1, as_numpy=False), reps=2)
res["sensitivity_2000x64"] = {"fits": 2000 * 64, "ms": ms, "fits_per_s": 2000 * 64 / ms * 1e3}
# process_subgrid stage on the device: upward continuation of every window before its spectrum (20,000 windows)
import functools
sel = np.linspace(0, xc.size - 1, 20000).astype(int)
stage = functools.partial(G.upward_continuation, height=1000.0)
ms, _ = timed(lambda: G._batch_spectra(200e3, xc[sel], yc[sel], np.hanning, stage, {}), reps=2)
res["config3_upward_continuation_spectrum_20000"] = {"windows": 20000, "ms": ms, "windows_per_s": 20000 / ms * 1e3}
# map assembly: lattice on the device, MAP maps
ms, m = timed(lambda: G.curie_depth_maps(200e3, 10e3, 10e3, as_numpy=False), reps=2)
res["config3_curie_depth_maps"] = {"windows": int(m["beta"].numel()), "ms": ms,
                                   "windows_per_s": int(m["beta"].numel()) / ms * 1e3}
print(json.dumps(res, indent=1))
