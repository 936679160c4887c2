#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01y}
for v in NOSYNC NOPHILOX NOMET ALL; do
  JMM_LIBJMMGPU=$PWD/jmmonedmc_b200/variants/libjmmgpu_$v.so timeout 200 python scripts/sweep_grid.py c3 "" "" 2>&1 | grep -v "^c[35]:" | sed "s/^/$v c3 /"
  JMM_LIBJMMGPU=$PWD/jmmonedmc_b200/variants/libjmmgpu_$v.so timeout 200 python scripts/sweep_grid.py c5 "" "" 2>&1 | grep -v "^c[35]:" | sed "s/^/$v c5 /"
done 2>&1 | tee $OUT/ablation_$TAG.log
