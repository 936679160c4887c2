#!/bin/bash
# how the histogram path scales with the number of chains (working set = 176 KB per chain)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01h}
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.4g %s  ms/step %.2f  launches %d" % (d["value"], d["unit"], d["ms_per_step"], d["gpu_launches"]))
    else: print(l.rstrip()[:300])
'
for c in 2048 8192 32768; do
  for sl in "" "JMM_NO_SLICE=1"; do
    echo "== c4 fast hist chains=$c $sl"
    env JMM_BENCH_CHAINS=$c JMM_BENCH_PER_STEP=100 $sl timeout 90 python bench.py --workload c4 --steps 2 --warmup 3 --arith fast --hist 2>&1 | tee -a $OUT/bench_c4hist_$TAG.json | python -c "$show"
  done
done
