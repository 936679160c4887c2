#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01u}
PT="--timeout 120 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q $PT -k "checkerboard" > $OUT/pytest_sub_$TAG.log 2>&1
tail -3 $OUT/pytest_sub_$TAG.log
grep -n "Timeout\|FAILED\|Error" $OUT/pytest_sub_$TAG.log | head -20
