#!/bin/bash
# GPU visit: fast-arithmetic parity first (under a hard timeout), then lane-count sweeps of the C3/C4/C5 workloads
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01f}
PT="--timeout 90 --timeout-method thread"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q $PT -k "fast or checkerboard or histogram" > $OUT/pytest_fast_$TAG.log 2>&1
tail -8 $OUT/pytest_fast_$TAG.log
if ! grep -q " passed" $OUT/pytest_fast_$TAG.log || grep -q "failed\|Timeout" $OUT/pytest_fast_$TAG.log; then echo "PARITY NOT GREEN: stopping"; grep -n "Timeout\|FAILED\|Error" $OUT/pytest_fast_$TAG.log | head; exit 1; fi
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.4g %s  ms/step %.2f  launches %d  fp64 %.3f/%.1f TF = %.4f acc %s" % (d["value"], d["unit"], d["ms_per_step"], d["gpu_launches"], r["achieved"], r["peak"], r["frac"], d.get("acceptance")))
    else: print(l.rstrip()[:300])
'
for g in 1 2 4; do
  echo "== c4 fast JMM_PROD_G=$g"
  JMM_PROD_G=$g timeout 120 python bench.py --workload c4 --steps 3 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_c4_$TAG.json | python -c "$show"
done
echo "== c4 fast hist"
timeout 200 python bench.py --workload c4 --steps 3 --warmup 3 --arith fast --hist 2>&1 | tee -a $OUT/bench_c4_$TAG.json | python -c "$show"
for g in 1 2; do
  echo "== c3 fast JMM_SWEEP_G=$g"
  JMM_SWEEP_G=$g timeout 120 python bench.py --workload c3 --steps 5 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_c3_$TAG.json | python -c "$show"
done
for g in 2 4 8 16 32; do
  echo "== c5 fast JMM_SWEEP_G=$g"
  JMM_SWEEP_G=$g timeout 120 python bench.py --workload c5 --steps 5 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_c5_$TAG.json | python -c "$show"
done
timeout 600 python -m pytest tests -m gpu -x -q $PT > $OUT/pytest_gpu_$TAG.log 2>&1; tail -5 $OUT/pytest_gpu_$TAG.log
