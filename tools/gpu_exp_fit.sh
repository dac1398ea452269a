#!/bin/bash
# fit / MH occupancy experiment: rebuild is done on the CPU box; here we just run the bench twice
python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['stages'], 'mcmc', d['mcmc']['value'])
"
