#!/bin/bash
# A/B of library variants on the MCMC leg of bench.py (config 4 sample).  Usage: bash tools/gpu_ab_mh.sh <name> ...
for v in "$@"; do
  if [ "$v" = main ]; then unset CB_LIB_PATH; else export CB_LIB_PATH=$PWD/tools/variants/libcurie_$v.so; fi
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); m = d['mcmc']; print('$v', 'mcmc %.4g steps/s' % m['value'], 'ms', round(m['ms'], 1), 'acc', round(m['acceptance_mean'], 5))
"
done
