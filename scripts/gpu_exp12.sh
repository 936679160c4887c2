#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01z}
PT="--timeout 120 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q $PT -k "checkerboard" > $OUT/pytest_sub_$TAG.log 2>&1
tail -3 $OUT/pytest_sub_$TAG.log
grep -n "Timeout\|FAILED\|Error" $OUT/pytest_sub_$TAG.log | head -20
timeout 300 python scripts/sweep_grid.py c3 "" "" 2>&1 | tee $OUT/grid_c3_$TAG.log
timeout 400 python scripts/sweep_grid.py c5 "" "" "G=4" "G=16" "WARPS=24" "WARPS=24,G=16" 2>&1 | tee $OUT/grid_c5_$TAG.log
