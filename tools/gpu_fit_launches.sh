#!/bin/bash
# per-launch durations of the fit kernels of one config-3 step (ncu, gpu__time_duration only).  Usage: <tag>
TAG=${1:-fit}; OUT=gpurun_out; mkdir -p $OUT
ARGS="--grid ${GRID:-4096} --steps 1 --warmup 0 --no-cpu-baseline --no-extra-configs --mcmc-nsim 0"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^fit_' -c 44 --csv --log-file $OUT/fit_launches_$TAG.csv python bench.py $ARGS > $OUT/ncu_l_$TAG.log 2>&1
echo "rc=$?"
python - <<'P'
import csv,sys,glob,os
f=sorted(glob.glob('gpurun_out/fit_launches_*.csv'),key=os.path.getmtime)[-1]
rows=list(csv.reader(open(f))); h=next(i for i,r in enumerate(rows) if r and r[0]=='ID'); hdr=rows[h]
tot={}; seq=[]
for r in rows[h+1:]:
    if len(r)<len(hdr): continue
    d=dict(zip(hdr,r)); k=d['Kernel Name'].split('(')[0]; t=float(d['Metric Value'].replace(',',''))/1e3
    tot[k]=tot.get(k,0)+t; seq.append('%s %.0f'%(k[4:8],t))
print(' '.join(seq)); print({k:round(v) for k,v in tot.items()}, round(sum(tot.values())))
P
