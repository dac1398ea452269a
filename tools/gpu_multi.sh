#!/bin/bash
# Multi-GPU visit.  Usage under gpurun --gpus N:  bash tools/gpu_multi.sh <tag> <N> [tests] [bench] [ref]
TAG=${1:-dev}; N=${2:-2}; shift; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus_$TAG.txt; nproc >> $OUT/gpus_$TAG.txt
for what in "$@"; do
case $what in
tests)
  python -m pytest tests -m gpu -q -k "max_processors or fans_out" > $OUT/pytest_multi_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_multi_$TAG.log ;;
bench)
  NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 $BENCH_ARGS > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err; echo "bench rc=$?"
  tail -c 2500 $OUT/bench_${N}gpu_$TAG.json; grep -c "NCCL INFO" $OUT/bench_${N}gpu_$TAG.err; tail -5 $OUT/bench_${N}gpu_$TAG.err ;;
threads)
  python tools/multi_gpu_threads.py $N > $OUT/threads_${N}gpu_$TAG.json 2> $OUT/threads_${N}gpu_$TAG.err; echo "threads rc=$?"; cat $OUT/threads_${N}gpu_$TAG.json ;;
esac
done
