#!/usr/bin/env python
"""stdin: bench.py output; prints value, ms/step and the stage times on one line (tuning scripts)."""
import json
import sys
for l in sys.stdin:
    if not l.startswith("{"):
        print(l.rstrip()[-200:])
        continue
    d = json.loads(l)
    st = d.get("stages") or {}
    print("value %.0f  ms/step %.2f  spectrum %.2f  fit %.2f  mcmc %s" % (
        d["value"], d["ms_per_step"], st.get("spectrum_ms", -1), st.get("fit_ms", -1),
        (d.get("mcmc") or {}).get("value")))
