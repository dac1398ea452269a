#!/bin/bash
# ncu --set full capture of fit_kernel alone (grid 2048: 34,225 windows).  Usage: bash tools/gpu_prof_fit.sh <tag> [warps] [lockstep]
TAG=${1:-fit}; export CB_FIT_WARPS=${2:-24}; export CB_FIT_LOCKSTEP=${3:-1}
OUT=gpurun_out; mkdir -p $OUT
ARGS="--grid 2048 --steps 1 --warmup 1 --no-cpu-baseline --no-extra-configs --mcmc-nsim 0"
python bench.py $ARGS > $OUT/plain_$TAG.log 2>&1 || { tail -5 $OUT/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:fit_kernel -s 1 -c 1 -f -o $OUT/prof_$TAG python bench.py $ARGS > $OUT/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la $OUT | tail -4
