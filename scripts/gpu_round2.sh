#!/bin/bash
# GPU visit: full parity (hard timeouts), histogram probe, then C3/C4/C5 fast benches with variants
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01k}
PT="--timeout 90 --timeout-method thread"
timeout 700 python -m pytest tests -m gpu -x -q $PT > $OUT/pytest_gpu_$TAG.log 2>&1
tail -4 $OUT/pytest_gpu_$TAG.log
if ! grep -q " passed" $OUT/pytest_gpu_$TAG.log || grep -q "failed\|Timeout" $OUT/pytest_gpu_$TAG.log; then echo "PARITY NOT GREEN"; grep -n "Timeout\|FAILED\|Error\|assert" $OUT/pytest_gpu_$TAG.log | head -20; fi
echo "== hist probe"; PROBE_CHAINS=4096 timeout 100 python scripts/hist_probe.py 2>&1 | tail -8
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.4g %s  ms/step %.2f  launches %d  fp64 %.3f/%.1f TF = %.4f acc %s" % (d["value"], d["unit"], d["ms_per_step"], d["gpu_launches"], r["achieved"], r["peak"], r["frac"], d.get("acceptance")))
    else: print(l.rstrip()[:300])
'
for u in 4 8; do
  echo "== c4 fast JMM_PROD_UNROLL=$u"
  JMM_PROD_UNROLL=$u timeout 120 python bench.py --workload c4 --steps 3 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_c4_$TAG.json | python -c "$show"
done
echo "== c4 fast hist"
timeout 200 python bench.py --workload c4 --steps 2 --warmup 3 --arith fast --hist 2>&1 | tee -a $OUT/bench_c4hist_$TAG.json | python -c "$show"
echo "== c3 fast"
timeout 120 python bench.py --workload c3 --steps 5 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_c3_$TAG.json | python -c "$show"
for g in 4 8 16; do
  echo "== c5 fast JMM_SWEEP_G=$g"
  JMM_SWEEP_G=$g timeout 120 python bench.py --workload c5 --steps 5 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_c5_$TAG.json | python -c "$show"
done
