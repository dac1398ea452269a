#!/bin/bash
# batched-fit sweep (CB_FIT_ROUNDS = batched rounds before the per-warp finish) on config 3: prints stages.fit_ms
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -q -x -k "fit or config3 or optimise or sensitivity" > $OUT/pytest_fit.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_fit.log
for r in ${ROUNDS:-20 12 16 24 32 48}; do
  CB_FIT_ROUNDS=$r python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extra-configs --mcmc-nsim 0 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('rounds $r', d['value'], d['ms_per_step'], d['stages'])
" | tee -a $OUT/fit_sweep.txt
done
