#!/bin/bash
# spectrum-stage tuning experiment: FFT teams per CTA
mkdir -p gpurun_out
for t in 16 8 4; do
  echo "== CB_SPEC_TEAMS=$t"
  CB_SPEC_TEAMS=$t python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['stages'], d['plan'])
    else: print(l.rstrip()[-300:])
"
done
