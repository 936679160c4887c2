#!/bin/bash
# last visit of the round: whole GPU suite + smoke + the C2 bench line and profile with the final kernels
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2z}
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 --timeout-method thread --durations=6 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -10 $OUT/pytest_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
bash scripts/gpu_c2_final.sh $TAG
SAN_CASES="trio crew" bash scripts/gpu_sanitize.sh > $OUT/sanitizer_last_$TAG.txt 2>&1; grep -E "==|SUMMARY" $OUT/sanitizer_last_$TAG.txt
