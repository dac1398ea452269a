#!/bin/bash
# MH steps/s against the chain length (the chain starts at the MAP estimate; early rejection is most effective there)
for n in 2000 10000 50000 100000; do
  python bench.py --grid 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-extra-configs --mcmc-nsim $n 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); m = d['mcmc']; print('nsim $n', 'chains', m['chains_total'], 'mcmc %.4g steps/s' % m['value'], 'ms', round(m['ms'], 1), 'acc', round(m['acceptance_mean'], 5))
"
done
