#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench line, an ncu launch list and one full capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh <tag> [full]
TAG=${1:-dev}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke_$TAG.log
python bench.py --steps 3 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench_$TAG.json
if [ "$2" = "full" ]; then
  # launch list on the same command line as the bench (small grid keeps the serialised replay short)
  python bench.py --grid 1024 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/plain_$TAG.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --grid 1024 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'rowspec_kernel|colspec_kernel|spectrum_kernel|fit_kernel|mh_kernel' -c 4 -f -o $OUT/prof_$TAG \
      python bench.py --grid 1024 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
fi
ls -la $OUT | tail -20
if [ "$2" = "big" ]; then
  # full-size (152,100 windows) captures of the two kernels: DRAM traffic and pipe utilisation per launch
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mcmc-steps 0 > $OUT/plain_big_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'rowspec_kernel|colspec_kernel|spectrum_kernel|fit_kernel' -s 3 -c 3 -f -o $OUT/prof_big_$TAG \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mcmc-steps 0 > $OUT/ncu_big_$TAG.log 2>&1
  ls -la $OUT | tail -5
fi
