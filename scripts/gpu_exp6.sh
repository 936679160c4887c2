#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01r}
for v in A B C D E F; do
  echo "=== variant $v"
  JMM_LIBJMMGPU=$PWD/jmmonedmc_b200/variants/libjmmgpu_$v.so timeout 200 python scripts/sweep_grid.py c3 "" "K=2,WARPS=13" 2>&1 | grep -v "^c[35]:"
  JMM_LIBJMMGPU=$PWD/jmmonedmc_b200/variants/libjmmgpu_$v.so timeout 200 python scripts/sweep_grid.py c5 "" "K=1,WARPS=24,G=8" 2>&1 | grep -v "^c[35]:"
done 2>&1 | tee $OUT/variants_$TAG.log
