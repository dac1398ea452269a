#!/bin/bash
# run the bench against alternative builds of the library (build_variants/*.so), restoring the default afterwards
cp pycurious_b200/libcurie_b200.so /tmp/default.so
for v in default $(ls build_variants/*.so 2>/dev/null); do
  if [ "$v" != default ]; then cp $v pycurious_b200/libcurie_b200.so; fi
  echo "== $v"
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --mcmc-steps 0 2>&1 | python tools/bench_brief.py
  cp /tmp/default.so pycurious_b200/libcurie_b200.so
done
