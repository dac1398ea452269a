#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS table with nvdisasm line info and aggregate executed
instructions / stall samples per source line.

    ncu -i X.ncu-rep --page source --csv -k regex:fit_kernel > sass.csv
    cuobjdump -xelf all lib.so && nvdisasm -g -c model.sm_100a.cubin > dis.txt
    python tools/ncu_lines.py sass.csv dis.txt fit_kernel [topN]
"""
import csv
import re
import sys
from collections import defaultdict


def line_map(dis_path, func_pat):
    """offset -> (file, line) for the function whose name matches func_pat."""
    out, cur, active = {}, None, False
    for ln in open(dis_path, errors="replace"):
        m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*//-+ \.text\.(\S+)", ln)
        if m:
            active = func_pat in m.group(1)
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


def main():
    sass_csv, dis, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lm = line_map(dis, func)
    rows = list(csv.reader(open(sass_csv)))
    h = next(i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address")
    hdr = rows[h]
    ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
    base = int(rows[h + 1][0], 16)
    agg = defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[h + 1:]:
        if len(r) <= ci["Thread Instructions Executed"] or not r[0].startswith("0x"):
            continue
        off = int(r[0], 16) - base
        key = lm.get(off, ("?", 0))
        s, ie, te = int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0)
        a = agg[key]
        a[0] += s; a[1] += ie; a[2] += te; a[3] += 1
        tot[0] += s; tot[1] += ie; tot[2] += te
    print("total samples %d, warp-inst %d, thread-inst %d (avg %.1f lanes)" % (tot[0], tot[1], tot[2], tot[2] / max(1, tot[1])))
    print("%-28s %8s %7s %12s %7s %6s %5s" % ("file:line", "samples", "%", "warp-inst", "%", "lanes", "sass"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s %8d %6.2f%% %12d %6.2f%% %6.1f %5d" % ("%s:%d" % key, a[0], 100.0 * a[0] / max(1, tot[0]), a[1],
                                                          100.0 * a[1] / max(1, tot[1]), a[2] / max(1, a[1]), a[3]))
    byfile = defaultdict(lambda: [0, 0])
    for key, a in agg.items():
        byfile[key[0]][0] += a[0]; byfile[key[0]][1] += a[1]
    print("-- by file")
    for f, a in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("%-28s samples %6.2f%%  warp-inst %6.2f%%" % (f, 100.0 * a[0] / max(1, tot[0]), 100.0 * a[1] / max(1, tot[1])))


if __name__ == "__main__":
    main()
