#!/usr/bin/env python
"""Turn an .ncu-rep into the text summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [lib.so] > profiles/rNN_X.md

Per kernel: duration, launch shape, occupancy, pipe utilisation, DRAM bytes, stall-reason shares and
the hottest source lines (needs -lineinfo and `ncu --import-source on`)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "pycurious_b200", "libcurie_b200.so")
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
        ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread-inst"),
        ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread-inst"),
        ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread-inst")]


def run(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
print("# ncu summary of `%s`\n" % os.path.basename(rep))
print("Captured with `ncu --set full --clock-control none --import-source on` (each kernel replayed ~40x, cold caches:\n"
      "compare shares, not absolute times with bench.py).\n")
names = []
for r in raw[2:]:
    kn = r[hdr.index("Kernel Name")]
    names.append(kn)
    print("## %s\n" % kn)
    print("| metric | value |\n|---|---|")
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            print("| %s | %s %s |" % (label, r[i], units[i]))
    print()

# source-level: stall reasons + hottest lines
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s > /dev/null" % (tmp, lib), shell=True)
dis = {}
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        dis[f] = run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)])


def norm(name):
    return re.sub(r"\s+", "", name.replace("void ", "")).replace(">>", "> >").replace(" ", "")


def line_map(func, full):
    """PC offset -> (file, line) of ONE kernel: the .text section whose demangled name is `full` (a template has one
    section per instantiation, all starting at offset 0)."""
    out, cur, active = {}, None, False
    func = "_Z%d%s" % (len(func), func)      # mangled prefix: "fit_kernel" must not match "misfit_kernel"
    for text in dis.values():
        if func not in text:
            continue
        for ln in text.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                active = func in m.group(1) and norm(run(["c++filt", m.group(1)]).strip()) == norm(full)
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m and cur:
                out[int(m.group(1), 16)] = cur
    return out


for kn in dict.fromkeys(names):
    short = re.split(r"[(<]", kn.replace("void ", ""))[0].strip()
    src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + short]))))
    try:
        h = next(i for i, r in enumerate(src) if len(r) > 5 and r[0] == "Address")
    except StopIteration:
        continue
    sh = src[h]
    stall_cols = [i for i, n in enumerate(sh) if n.startswith("stall_") and "Not Issued" not in n]
    ci = {n: sh.index(n) for n in ("# Samples", "Instructions Executed", "Thread Instructions Executed")}
    lm = line_map(short, kn)
    base = None
    stalls = defaultdict(int)
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0]
    for r in src[h + 1:]:
        if not r or not r[0].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break                      # only the first launch of this kernel
            continue
        if base is None:
            base = int(r[0], 16)
        off = int(r[0], 16) - base
        for i in stall_cols:
            try:
                stalls[sh[i]] += int(r[i] or 0)
            except ValueError:
                pass
        s, ie, te = (int(r[ci[n]] or 0) for n in ("# Samples", "Instructions Executed", "Thread Instructions Executed"))
        a = agg[lm.get(off, ("?", 0))]
        a[0] += s; a[1] += ie; a[2] += te
        tot[0] += s; tot[1] += ie
    print("## %s: warp-state samples and hottest source lines\n" % short)
    st = sum(stalls.values()) or 1
    print("stall reasons: " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / st)
                                          for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]) + "\n")
    print("| file:line | samples % | warp-inst % | lanes |\n|---|---|---|---|")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][int(os.environ.get("NCU_SORT", "0"))])[:int(os.environ.get("NCU_TOP", "18"))]:
        print("| %s:%d | %.2f | %.2f | %.1f |" % (key[0], key[1], 100.0 * a[0] / max(1, tot[0]), 100.0 * a[1] / max(1, tot[1]),
                                                 a[2] / max(1, a[1])))
    print()
