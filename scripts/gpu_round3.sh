#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01l}
PT="--timeout 90 --timeout-method thread"
timeout 700 python -m pytest tests -m gpu -x -q $PT > $OUT/pytest_gpu_$TAG.log 2>&1
tail -4 $OUT/pytest_gpu_$TAG.log
if ! grep -q " passed" $OUT/pytest_gpu_$TAG.log || grep -q "failed\|Timeout" $OUT/pytest_gpu_$TAG.log; then echo "PARITY NOT GREEN"; grep -n "Timeout\|FAILED\|Error\|assert" $OUT/pytest_gpu_$TAG.log | head -20; fi
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.4g %s  ms/step %.2f  launches %d  fp64 %.3f/%.1f TF = %.4f acc %s e2e %s" % (d["value"], d["unit"], d["ms_per_step"], d["gpu_launches"], r["achieved"], r["peak"], r["frac"], d.get("acceptance"), (d.get("e2e") or {}).get("value")))
    else: print(l.rstrip()[:300])
'
echo "== c2 (default bench, no extras)"
JMM_BENCH_CPU_STEPS=100000 timeout 200 python bench.py --steps 5 --warmup 3 --no-extras 2>&1 | tee -a $OUT/bench_c2_$TAG.json | python -c "$show"
for w in c3 c4 c5; do
  echo "== $w fast"
  timeout 120 python bench.py --workload $w --steps 5 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_${w}_$TAG.json | python -c "$show"
done
for w in c3 c4 c5; do
  echo "== $w reference"
  timeout 120 python bench.py --workload $w --steps 3 --warmup 3 2>&1 | tee -a $OUT/bench_${w}_$TAG.json | python -c "$show"
done
