#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -k fast 2>&1 | tail -3
for ar in reference fast; do
ncu --set full --clock-control none --import-source on -k regex:k_chains_step_prod -s 1 -c 1 -f -o gpurun_out/prof_c4_$ar python bench.py --workload c4 --arith $ar --steps 1 --warmup 3 > gpurun_out/ncu_c4_$ar.log 2>&1; tail -1 gpurun_out/ncu_c4_$ar.log
done
