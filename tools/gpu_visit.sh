#!/bin/bash
# One GPU-box visit (round 2).  Usage under gpurun:  bash tools/gpu_visit.sh <tag> [tests] [parity] [bench] [counts] [full] [big]
TAG=${1:-dev}; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt
nproc >> $OUT/gpu_$TAG.txt
for what in "$@"; do
case $what in
tests)
  python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
  tail -5 $OUT/pytest_gpu_$TAG.log
  python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke_$TAG.log ;;
parity)
  python tools/parity_sample.py ${PARITY_EXTRA:-0} > $OUT/parity_$TAG.json 2> $OUT/parity_$TAG.err; echo "parity rc=$?" ;;
bench)
  python bench.py --steps ${STEPS:-5} --warmup 3 $BENCH_ARGS > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
  tail -c 1500 $OUT/bench_$TAG.json ;;
refbench)
  python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "refbench rc=$?" ;;
counts)
  python tools/count_fp64.py > $OUT/fp64_units_plain_$TAG.json 2> $OUT/fp64_plain_$TAG.err && \
  ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum \
      --clock-control none -k regex:'rowspec_kernel|colspec_kernel|^fit_|mh_kernel' --csv --log-file $OUT/fp64_counts_$TAG.csv \
      python tools/count_fp64.py > $OUT/fp64_units_$TAG.json 2> $OUT/fp64_ncu_$TAG.err; echo "counts rc=$?" ;;
full)
  python bench.py --grid 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs --mcmc-nsim 2000 > $OUT/plain_$TAG.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --grid 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs --mcmc-nsim 2000 > $OUT/ncu_launch_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'rowspec_kernel|colspec_kernel|spectrum_kernel|fit_eval_kernel|fit_step_kernel|fit_finish_kernel|mh_kernel' -c 8 -f -o $OUT/prof_$TAG \
      python bench.py --grid 1024 --steps 1 --warmup 0 --no-cpu-baseline --no-extra-configs --mcmc-nsim 2000 > $OUT/ncu_full_$TAG.log 2>&1 ;;
big)
  ARGS="--steps 1 --warmup 1 --no-cpu-baseline --no-extra-configs --mcmc-nsim 0"
  python bench.py $ARGS > $OUT/plain_big_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'rowspec_kernel|colspec_kernel' -c 2 -f -o $OUT/prof_big_spec_$TAG \
      python bench.py $ARGS > $OUT/ncu_big_spec_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'fit_eval_kernel|fit_step_kernel|fit_finish_kernel|fit_first_kernel' -s 1 -c 4 -f -o $OUT/prof_big_fit_$TAG \
      python bench.py $ARGS > $OUT/ncu_big_fit_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'fit_finish_kernel' -c 1 -f -o $OUT/prof_big_finish_$TAG \
      python bench.py $ARGS > $OUT/ncu_big_finish_$TAG.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_big_$TAG.csv \
      python bench.py $ARGS > $OUT/ncu_big_launch_$TAG.log 2>&1 ;;
esac
done
ls -la $OUT | tail -12
