#!/bin/bash
# spectrum-stage tuning experiment: engine / teams knobs; prints value and stage times per setting
mkdir -p gpurun_out
run() {
  echo "== $*"
  env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline --mcmc-steps 0 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['stages'], d['plan'])
    else: print(l.rstrip()[-300:])
"
}
run CB_SPEC_TEAMS=16
run CB_SPEC_TEAMS=8
run CB_SPEC_TEAMS=12
