#!/bin/bash
# A/B of library variants (tools/variants/libcurie_<name>.so, selected with CB_LIB_PATH) on the config-3 step.
# Usage: bash tools/gpu_ab.sh <name> [<name> ...]   ("main" = the in-tree library)
OUT=gpurun_out; mkdir -p $OUT
for v in "$@"; do
  if [ "$v" = main ]; then unset CB_LIB_PATH; else export CB_LIB_PATH=$PWD/tools/variants/libcurie_$v.so; fi
  python bench.py --grid ${GRID:-4096} --steps 3 --warmup 2 --no-cpu-baseline --no-extra-configs --mcmc-nsim 0 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$v', d['value'], d['ms_per_step'], d['stages'])
" | tee -a $OUT/ab.txt
done
