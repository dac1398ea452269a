#!/usr/bin/env python
"""profiles/fp64_counts.json from an ncu metrics CSV and the unit counts of tools/count_fp64.py:

    python tools/fp64_counts_from_ncu.py gpurun_out/fp64_counts.csv gpurun_out/fp64_units.json > profiles/fp64_counts.json

flop = 2 x DFMA + DADD + DMUL thread-level instructions (predicated-on), per kernel launch; the spectrum figure
adds rowspec_kernel and colspec_kernel.  These per-unit figures replace the 400 flop/point planning convention of
SURVEY.md section 8(d) and are frozen: bench.py reads them for its FP64 roofline."""
import csv
import json
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
iK, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
per = {}
for r in rows:
    name = r[iK].split("(")[0].split("<")[0].replace("void ", "").strip()
    per.setdefault(name, {}).setdefault(r[iM], []).append(float(r[iV].replace(",", "")))
units = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])


def flop(kernel, launch=-1):
    m = per[kernel]
    g = lambda key: m["smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % key][launch]
    return 2.0 * g("dfma") + g("dadd") + g("dmul"), {k: g(k) for k in ("dfma", "dadd", "dmul")}


out = {"source": "ncu thread-level FP64 instruction counts, one launch per kernel, tools/count_fp64.py", "units": units}
f_row, c_row = flop("rowspec_kernel"), flop("colspec_kernel")
out["spectrum_flop_per_window_201"] = (f_row[0] + c_row[0]) / units["windows"]
out["spectrum_instr"] = {"rowspec_kernel": f_row[1], "colspec_kernel": c_row[1]}
# the fit is a sequence of launches (template, begin, first, evaluation / step rounds, finish): all of them, summed;
# a point = one wavenumber of one evaluated parameter vector (nfev counts the vectors, speculative ones included)
fit_kernels = sorted(k for k in per if k.startswith("fit_"))
tot, parts = 0.0, {}
for k in fit_kernels:
    n_launch = len(per[k]["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"])
    fl = sum(flop(k, i)[0] for i in range(n_launch))
    parts[k] = {"launches": n_launch, "flop": fl}
    tot += fl
out["fit_flop_per_point"] = tot / units["fit_points"]
out["fit_kernels"] = parts
f_mh = flop("mh_kernel")
out["mh_flop_per_step_99"] = f_mh[0] / units["mh_steps"]
out["mh_instr"] = f_mh[1]
out["durations_us"] = {k: [round(v / 1e3, 1) for v in m.get("gpu__time_duration.sum", [])] for k, m in per.items()}
print(json.dumps(out, indent=1))
