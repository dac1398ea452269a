#!/usr/bin/env python
"""SASS size report: instructions per function and per source line, from `nvdisasm -g -c <cubin>`.

    cuobjdump -xelf all lib.so && nvdisasm -g -c model.sm_100a.cubin > dis.txt
    python tools/sass_size.py dis.txt fit_kernel [topN]
"""
import re, sys
from collections import defaultdict
dis, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
func, cur = None, None
per_func = defaultdict(int)
per_line = defaultdict(lambda: defaultdict(int))
for ln in open(dis, errors="replace"):
    m = re.match(r"\s*\.type\s+(\S+),@function", ln)
    if m:
        func = m.group(1); continue
    m = re.match(r"^(\S+):\s*$", ln)
    if m and not m.group(1).startswith(".L"):
        func = m.group(1).replace(".text.", ""); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln) and func and pat in func:
        per_func[func] += 1
        per_line[func][cur] += 1
tot = sum(per_func.values())
print("total %d instructions = %.1f KB" % (tot, tot * 16 / 1024))
for f, n in sorted(per_func.items(), key=lambda kv: -kv[1]):
    short = re.sub(r"^\$?_Z\d+\w+?\$", "", f)
    print("%6d  %5.1f KB  %s" % (n, n * 16 / 1024, short[:100]))
    for (fl, c) in sorted(per_line[f].items(), key=lambda kv: -kv[1])[:top if n > 300 else 0]:
        print("          %5d  %s:%s" % (c, fl[0] if fl else "?", fl[1] if fl else 0))
