#!/bin/bash
# usage: gpu_prof2.sh tag kernel-regex workload arith
ncu --set full --clock-control none --import-source on -k "regex:$2" -s 2 -c 1 -f -o gpurun_out/prof_$1 python bench.py --workload $3 --arith $4 --steps 1 --warmup 3 > gpurun_out/ncu_$1.log 2>&1; tail -1 gpurun_out/ncu_$1.log
