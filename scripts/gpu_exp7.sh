#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01s}
run() { v=$1; wl=$2; shift 2; JMM_LIBJMMGPU=$PWD/jmmonedmc_b200/variants/libjmmgpu_$v.so timeout 200 python scripts/sweep_grid.py $wl "$@" 2>&1 | grep -v "^c[35]:" | sed "s/^/$v $wl /"; }
{
run G c3 "" "K=1,WARPS=24" "K=1,WARPS=20" "K=2,WARPS=12"
run H c3 "" "K=1,WARPS=16" "K=1,WARPS=12"
run G c5 "" "K=1,WARPS=24,G=8" "K=1,WARPS=24,G=4"
run H c5 "K=1,WARPS=16,G=8" "K=1,WARPS=16,G=4" "K=1,WARPS=16,G=2"
run I c5 "K=1,WARPS=16,G=8" "K=1,WARPS=16,G=4" "K=1,WARPS=16,G=2"
run J c5 "K=1,WARPS=12,G=8" "K=1,WARPS=12,G=4" "K=1,WARPS=12,G=2"
run K c5 "K=1,WARPS=8,G=8" "K=1,WARPS=8,G=4" "K=1,WARPS=8,G=2" "K=1,WARPS=8,G=4,NSUB=32"
run L c5 "K=1,WARPS=8,G=4" "K=1,WARPS=8,G=2" "K=1,WARPS=8,G=1"
} 2>&1 | tee $OUT/variants_$TAG.log
