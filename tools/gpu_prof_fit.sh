#!/bin/bash
# fit kernels of config 3 under ncu: per-launch durations (all fit_* launches of one step), then --set full captures of
# fit_eval_kernel (rounds 0 and 8) and fit_step_kernel (round 0).  Usage: bash tools/gpu_prof_fit.sh <tag>
TAG=${1:-fit}
OUT=gpurun_out; mkdir -p $OUT
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-extra-configs --mcmc-nsim 0"
python bench.py $ARGS > $OUT/plain_$TAG.log 2>&1 || { tail -5 $OUT/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'^fit_' -c 60 --csv --log-file $OUT/fit_launches_$TAG.csv python bench.py $ARGS > $OUT/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fit_eval_kernel -c 9 -f -o $OUT/prof_eval_$TAG python bench.py $ARGS > $OUT/ncu_e_$TAG.log 2>&1
echo "eval rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fit_step_kernel -c 1 -f -o $OUT/prof_step_$TAG python bench.py $ARGS > $OUT/ncu_s_$TAG.log 2>&1
echo "step rc=$?"; ls -la $OUT | tail -5
