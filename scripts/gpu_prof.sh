#!/bin/bash
# ncu --set full of one kernel regex under a given env; usage: gpu_prof.sh <tag> <kernel-regex> [ENV=VAL ...] 
TAG=$1; KRE=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
env "$@" ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 2 --warmup 3 > $OUT/ncu_$TAG.log 2>&1
tail -3 $OUT/ncu_$TAG.log
