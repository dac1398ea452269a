#!/usr/bin/env python
"""Per-function share of warp-state samples / executed warp instructions / active lanes for one kernel of an
.ncu-rep (functions = the non-inlined device functions inside the kernel's text section).

    python tools/ncu_funcs.py X.ncu-rep fit_kernel [lib.so]
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "pycurious_b200", "libcurie_b200.so")


def run(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s > /dev/null" % (tmp, lib), shell=True)
mang = "_Z%d%s" % (len(kern), kern)
fmap = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    text = run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)])
    if mang not in text:
        continue
    active, func = False, None
    for ln in text.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            active = mang in m.group(1); func = m.group(1); continue
        if not active:
            continue
        m = re.match(r"^(\S+):\s*$", ln)
        if m and not m.group(1).startswith(".L"):
            func = m.group(1); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            fmap[int(m.group(1), 16)] = func
src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern]))))
h = next(i for i, r in enumerate(src) if len(r) > 5 and r[0] == "Address")
sh = src[h]
ci = {n: sh.index(n) for n in ("# Samples", "Instructions Executed", "Thread Instructions Executed")}
agg = defaultdict(lambda: [0, 0, 0]); tot = [0, 0]; base = None
for r in src[h + 1:]:
    if not r or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            break
        continue
    if base is None:
        base = int(r[0], 16)
    off = int(r[0], 16) - base
    s, ie, te = (int(r[ci[n]] or 0) for n in ("# Samples", "Instructions Executed", "Thread Instructions Executed"))
    a = agg[fmap.get(off, "?")]
    a[0] += s; a[1] += ie; a[2] += te; tot[0] += s; tot[1] += ie
print("| function | samples % | warp-inst % | lanes |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    short = re.sub(r"^\$?_Z\d+\w+?\$", "", k)
    short = re.sub(r"_ZN\d+_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d*?(\d+)", r"\1:", short)[:70]
    print("| %s | %.2f | %.2f | %.1f |" % (short, 100.0 * a[0] / max(1, tot[0]), 100.0 * a[1] / max(1, tot[1]), a[2] / max(1, a[1])))
print("total warp instructions %d, samples %d" % (tot[1], tot[0]))
