#!/usr/bin/env python
"""Where the end-to-end time of optimise_routine goes (host pieces timed with a stream sync on both sides)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth, _plans, _lib

N, WINDOW, STRIDE = 4096, 200e3, 10e3
data = synth.fractal_grid(N, N, seed=20261019)
G = pc.CurieOptimise(data, *synth.grid_extent(N, N))
xc, yc = G.create_centroid_list(WINDOW, STRIDE, STRIDE)
SHARE = int(os.environ.get("E2E_SHARE", "1"))          # 8: the block one rank of an 8-GPU run owns
xc, yc = xc[: xc.size // SHARE], yc[: yc.size // SHARE]
out = {}


def T(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    out[name] = [round(min(ts), 3), round(float(np.median(ts)), 3)]
    return r


lib = _lib.load()
t0 = time.perf_counter()
rc = lib.cb_host_register(data.ctypes.data, data.nbytes)
out["host_register_rc_ms"] = [rc, round(1e3 * (time.perf_counter() - t0), 2)]
t = torch.from_numpy(data)
out["is_pinned_after_register"] = bool(t.is_pinned())
T("upload_registered", lambda: t.to("cuda", non_blocking=True))
pinned = torch.from_numpy(data).pin_memory() if not t.is_pinned() else None
if pinned is not None:
    T("upload_torch_pinned", lambda: pinned.to("cuda", non_blocking=True))
T("host_f64", lambda: G._host_f64())
T("window_origins", lambda: _plans.window_origins(G, WINDOW, xc, yc))
iy0, ix0, n = _plans.window_origins(G, WINDOW, xc, yc)
T("device_region", lambda: G._device_region(iy0, ix0, n))
T("origins_tensors", lambda: G._origins_tensors(iy0, ix0, 0, 0))
T("windows_device", lambda: G._windows_device(WINDOW, xc, yc))
g, iy0_t, ix0_t, n = G._windows_device(WINDOW, xc, yc)
plan = _plans.spectrum_plan(n, G.dx, _plans.taper_vector(np.hanning, n))
T("prepare_groups", lambda: plan._prepare_groups(lib, torch, int(iy0_t.numel()), g.shape[1], iy0_t, ix0_t, None))
T("plan_run_cached_groups", lambda: plan.run(g, iy0_t, ix0_t, False, 2.0))
phi, sig = plan.run(g, iy0_t, ix0_t, False, 2.0)
T("fit_device", lambda: G._fit_device(plan.k, phi, (3.0, 1.0, 10.0, 5.0)))
x = G._fit_device(plan.k, phi, (3.0, 1.0, 10.0, 5.0))
T("to_host_T", lambda: G._to_host(x.t()))
T("batch_spectra", lambda: G._batch_spectra(WINDOW, xc, yc, np.hanning, None, {}))
T("optimise_batch", lambda: G.optimise_batch(WINDOW, xc, yc))
T("optimise_routine", lambda: G.optimise_routine(WINDOW, xc, yc))
G.hold_device()
T("optimise_routine_held", lambda: G.optimise_routine(WINDOW, xc, yc))
print(json.dumps(out, indent=1))
