#!/bin/bash
# compute-sanitizer synccheck (barrier misuse) + initcheck on the named-barrier kernels of solo.cuh
OUT=gpurun_out; mkdir -p $OUT
bash scripts/gpu_sanitize.sh > /dev/null 2>&1 &   # (writes /tmp/san.py; killed right away)
sleep 2; kill %1 2>/dev/null
for c in ${SYNC_CASES:-trio crew}; do
  echo "== synccheck $c"
  SAN_CASE=$c timeout 900 compute-sanitizer --tool synccheck --print-limit 3 python /tmp/san.py 2>&1 | grep -v "^$" | head -${SYNC_LINES:-60}
done
