#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01t}
PT="--timeout 120 --timeout-method thread"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q $PT -k "checkerboard" > $OUT/pytest_sub_$TAG.log 2>&1
tail -3 $OUT/pytest_sub_$TAG.log
grep -n "Timeout\|FAILED\|Error\|assert" $OUT/pytest_sub_$TAG.log | head -10
timeout 300 python scripts/sweep_grid.py c3 "" "K=1,WARPS=24" "K=1,WARPS=20" "K=1,WARPS=16" "K=1,WARPS=12" "K=1,WARPS=8" "K=2,WARPS=12" "K=2,WARPS=8" 2>&1 | tee $OUT/grid_c3_$TAG.log
timeout 400 python scripts/sweep_grid.py c5 "" "K=1,WARPS=24,G=8" "K=1,WARPS=16,G=8" "K=1,WARPS=12,G=8" "K=1,WARPS=24,G=4" "K=1,WARPS=16,G=4" "K=1,WARPS=12,G=4" "K=1,WARPS=24,G=16" "K=1,WARPS=24,G=8,NSUB=32" "K=1,WARPS=16,G=4,NSUB=32" 2>&1 | tee $OUT/grid_c5_$TAG.log
