#!/usr/bin/env python
"""Development fixture: radial spectra of a sample of bench.py (config 3) windows, computed with the
oracle on a smaller synthetic grid, plus the reference L-BFGS-B fit of each.  Used to study the
device solver's iteration counts under CPU emulation (tests/emu) without a GPU.

    python tools/make_synth_spectra.py /tmp/synth_spectra.npz [nwin]
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import curie_oracle as orc
from pycurious_b200 import synth

out = sys.argv[1]
nwin = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = 1024
d = synth.fractal_grid(n, n, seed=20261019)
spec = orc.GridSpec(d, *synth.grid_extent(n, n))
xc, yc = orc.centroid_list(spec, 200e3, 10e3, 10e3)
idx = np.linspace(0, xc.size - 1, nwin).astype(int)
phis, xs, fs = [], [], []
for i in idx:
    k, phi, sig = orc.radial_spectrum(orc.cut_window(spec, 200e3, xc[i], yc[i]), spec.dx)
    phis.append(phi)
    r = orc.map_fit(k, phi, full=True)
    xs.append(r.x); fs.append(r.fun)
np.savez(out, k=k, phi=np.array(phis), xref=np.array(xs), fref=np.array(fs))
print("saved", out, np.array(phis).shape)
