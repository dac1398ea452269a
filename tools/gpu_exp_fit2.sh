#!/bin/bash
# batched-fit sweep on config 3: CB_FIT_HANDOVER (unfinished windows below which the CTA-per-window finish kernel takes
# over) x CB_FIT_ROUNDS (batched rounds at most); prints stages.fit_ms.  GRID=<edge> for a smaller survey.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -q -x -k "fit or config3 or optimise or sensitivity or max_processors" > $OUT/pytest_fit.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_fit.log
for cfg in ${CONFIGS:-"2048 20" "0 20" "512 20" "8192 20" "2048 32" "4096 24"}; do
  set -- $cfg
  CB_FIT_HANDOVER=$1 CB_FIT_ROUNDS=$2 python bench.py --grid ${GRID:-4096} --steps 3 --warmup 2 --no-cpu-baseline --no-extra-configs --mcmc-nsim 0 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('handover $1 rounds $2 grid ${GRID:-4096}', round(d['value']), round(d['ms_per_step'],2), 'fit_ms', round(d['stages']['fit_ms'],3), 'spec', round(d['stages']['spectrum_ms'],2), 'e2e', round(d['e2e']['ms_per_step'],2))
" | tee -a $OUT/fit_sweep.txt
done
