"""Golden vectors for the FFT-domain filters, generated from the REAL reference
(imported unmodified from /root/reference, build container only):

    python tests/golden/make_golden_filters.py   ->  tests/golden/filters.npz

CurieGrid.upward_continuation (pycurious/grid.py:492-546) and CurieGrid.reduce_to_pole (grid.py:405-490) on
windows of the reference's test grid: odd, even and non-square-free sizes, induced-only and remanent cases."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYCURIOUS_REF", "/root/reference")
sys.path.insert(0, REF)
warnings.simplefilter("ignore")
import pycurious  # noqa: E402

g = np.load(os.path.join(HERE, "mag_grid.npz"))
d, ext = g["d"], [float(v) for v in g["extent"]]
G = pycurious.CurieGrid(d, *ext)
out = {}
# (name, (row0, col0, edge), stride): outputs are stored subsampled by `stride` to keep the fixture small
cases = [("w64", (10, 20, 64), 1), ("w101", (102, 102, 101), 2), ("w150", (30, 70, 150), 3), ("full", (0, 0, 305), 6)]
for name, (j0, i0, n), st in cases:
    win = d[j0:j0 + n, i0:i0 + n]
    out[name + "_meta"] = np.array([j0, i0, n, st])
    out[name + "_up500"] = G.upward_continuation(win, 500.0)[::st, ::st]
    out[name + "_up5000"] = G.upward_continuation(win, 5000.0)[::st, ::st]
    out[name + "_rtp_induced"] = G.reduce_to_pole(win, -60.0, 12.0)[::st, ::st]
    out[name + "_rtp_remanent"] = G.reduce_to_pole(win, 35.0, -5.0, sinc=50.0, sdec=20.0)[::st, ::st]
out["names"] = np.array([c[0] for c in cases])
np.savez_compressed(os.path.join(HERE, "filters.npz"), **out)
print("wrote filters.npz", os.path.getsize(os.path.join(HERE, "filters.npz")) / 1024, "KB")
