#!/bin/bash
# ncu --set full capture of mh_kernel on a config-4 sample (grid 1024: 6889 chains x 2020 steps).  Usage: <tag>
TAG=${1:-mh}; OUT=gpurun_out; mkdir -p $OUT
ARGS="--grid 1024 --steps 1 --warmup 0 --no-cpu-baseline --no-extra-configs --mcmc-nsim 2000"
python bench.py $ARGS > $OUT/plain_mh_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'mh_kernel' -c 1 -f -o $OUT/prof_mh_$TAG python bench.py $ARGS > $OUT/ncu_mh_$TAG.log 2>&1
echo "ncu rc=$?"
