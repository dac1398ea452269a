#!/bin/bash
# the 301-node (300 km) windows of config 5: spectrum / fit split on a 4096^2 lattice at 20 km stride for library
# variants (tools/variants/libcurie_<name>.so; "main" = in-tree), then an ncu --set full capture of the rowspec /
# colspec kernels of the LAST variant.  Usage: bash tools/gpu_prof_301.sh <tag> [variant ...]
TAG=${1:-w301}; shift; OUT=gpurun_out; mkdir -p $OUT
[ $# -eq 0 ] && set -- main
cat > /tmp/w301.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import pycurious_b200 as pc
from pycurious_b200 import synth, _plans
n = 4096
G = pc.CurieOptimise(synth.fractal_grid(n, n, seed=20261019), *synth.grid_extent(n, n)).hold_device()
g = G.device_grid()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out
for window in [float(w) * 1e3 for w in os.environ.get("WINDOWS_KM", "300,200,100").split(",")]:
    xc, yc = G.create_centroid_list(window, 20e3, 20e3)
    iy0, ix0, nw = G._origins_device(window, xc, yc)
    plan = _plans.spectrum_plan(nw, G.dx, _plans.taper_vector(np.hanning, nw))
    ms, (phi, sig) = timed(lambda: plan.run(g, iy0, ix0, False, 2.0))
    i = plan.info()
    print(nw, "nodes fft", i.fft_len, "teams", i.teams, "smem", i.smem_bytes, ":", xc.size, "windows, spectrum %.2f ms (%.3f us/window)" % (ms, 1e3 * ms / xc.size),
          "checksum %.17g" % float(phi.double().sum()))
PY
for v in "$@"; do
  if [ "$v" = main ]; then unset CB_LIB_PATH; else export CB_LIB_PATH=$PWD/tools/variants/libcurie_$v.so; fi
  echo "== $v" | tee -a $OUT/w301_$TAG.txt
  python /tmp/w301.py 2>/dev/null | tee -a $OUT/w301_$TAG.txt
done
if [ -z "$NO_NCU" ]; then
  WINDOWS_KM=300 ncu --set full --clock-control none --import-source on -k regex:'rowspec_kernel|colspec_kernel' -c 2 -f -o $OUT/prof_w301_$TAG python /tmp/w301.py > $OUT/ncu_w301_$TAG.log 2>&1
  echo "ncu rc=$?"
fi
