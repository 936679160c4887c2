#!/bin/bash
# GPU visit: parity of the touched paths, histogram workload, then ncu of the three fast-arithmetic kernels
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01i}
PT="--timeout 90 --timeout-method thread"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compat.py -m gpu -x -q $PT -k "fast or checkerboard or histogram or hist" > $OUT/pytest_fast_$TAG.log 2>&1
tail -4 $OUT/pytest_fast_$TAG.log
if ! grep -q " passed" $OUT/pytest_fast_$TAG.log || grep -q "failed\|Timeout" $OUT/pytest_fast_$TAG.log; then echo "PARITY NOT GREEN: stopping"; grep -n "Timeout\|FAILED\|Error" $OUT/pytest_fast_$TAG.log | head; exit 1; fi
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.4g %s  ms/step %.2f  launches %d  fp64 %.3f/%.1f TF = %.4f acc %s" % (d["value"], d["unit"], d["ms_per_step"], d["gpu_launches"], r["achieved"], r["peak"], r["frac"], d.get("acceptance")))
    else: print(l.rstrip()[:300])
'
echo "== c4 fast hist 8192 chains"
JMM_BENCH_CHAINS=8192 timeout 100 python bench.py --workload c4 --steps 2 --warmup 3 --arith fast --hist 2>&1 | tee -a $OUT/bench_c4hist_$TAG.json | python -c "$show"
echo "== c4 fast hist"
timeout 150 python bench.py --workload c4 --steps 2 --warmup 3 --arith fast --hist 2>&1 | tee -a $OUT/bench_c4hist_$TAG.json | python -c "$show"
for w in c3 c5; do
  echo "== $w fast"
  timeout 100 python bench.py --workload $w --steps 5 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_${w}_$TAG.json | python -c "$show"
done
# launch lists (cold-cache, serialised: shares only)
for w in c3 c5; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_${w}fast_$TAG.csv \
      python bench.py --workload $w --arith fast --steps 1 --warmup 3 > $OUT/ncu_launches_${w}_$TAG.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_prod -s 1 -c 1 -f -o $OUT/prof_c4fast_$TAG \
    python bench.py --workload c4 --arith fast --steps 1 --warmup 3 > $OUT/ncu_c4fast_$TAG.log 2>&1; tail -1 $OUT/ncu_c4fast_$TAG.log
for w in c3 c5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fast -s 2 -c 1 -f -o $OUT/prof_${w}fast_$TAG \
      python bench.py --workload $w --arith fast --steps 1 --warmup 3 > $OUT/ncu_${w}fast_$TAG.log 2>&1; tail -1 $OUT/ncu_${w}fast_$TAG.log
done
ls -la $OUT | tail -12
