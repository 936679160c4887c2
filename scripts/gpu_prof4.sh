#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01v}
for w in c3 c5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fast -s 4 -c 1 -f -o $OUT/prof_${w}fast_$TAG \
      python bench.py --workload $w --arith fast --steps 1 --warmup 3 > $OUT/ncu_${w}fast_$TAG.log 2>&1; tail -1 $OUT/ncu_${w}fast_$TAG.log | cut -c1-200
done
