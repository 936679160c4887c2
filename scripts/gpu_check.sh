#!/bin/bash
# whole GPU suite + sanitizer on the C2 kernels of solo.cuh
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-chk}
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method thread > $OUT/pytest_gpu_$TAG.log 2>&1
tail -4 $OUT/pytest_gpu_$TAG.log
SAN_CASES="${SAN_CASES:-solo trio crew bond}" bash scripts/gpu_sanitize.sh > $OUT/sanitizer_$TAG.txt 2>&1; tail -30 $OUT/sanitizer_$TAG.txt
