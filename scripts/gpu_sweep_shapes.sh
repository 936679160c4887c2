#!/bin/bash
# Launch-shape experiments for k_sweep_fast (one GPU visit): parity of the checkerboard paths, then scripts/sweep_grid.py
# over JMM_SWEEP_{K,WARPS,G,NSUB} settings for C3 and C5.  Usage (under gpurun): bash scripts/gpu_sweep_shapes.sh [tag]
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-shapes}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 --timeout-method thread -k "checkerboard" > $OUT/pytest_sub_$TAG.log 2>&1
tail -3 $OUT/pytest_sub_$TAG.log
timeout 300 python scripts/sweep_grid.py c3 "" "" "K=2,WARPS=12" "K=1,WARPS=16" "K=1,WARPS=20" "K=1,WARPS=24" 2>&1 | tee $OUT/grid_c3_$TAG.log
timeout 400 python scripts/sweep_grid.py c5 "" "" "G=4" "G=16" "WARPS=24" "WARPS=16" "K=2,WARPS=12" 2>&1 | tee $OUT/grid_c5_$TAG.log
