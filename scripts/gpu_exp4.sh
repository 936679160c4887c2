#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01p}
PT="--timeout 90 --timeout-method thread"
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q $PT -k "fast or checkerboard" > $OUT/pytest_sub_$TAG.log 2>&1
tail -3 $OUT/pytest_sub_$TAG.log
grep -n "Timeout\|FAILED\|Error\|assert" $OUT/pytest_sub_$TAG.log | head -10
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.4g %s  ms/step %.2f  launches %d  fp64 %.3f/%.1f TF = %.4f acc %s steps %s last %.3f clocks %s" % (d["value"], d["unit"], d["ms_per_step"], d["gpu_launches"], r["achieved"], r["peak"], r["frac"], d.get("acceptance"), d.get("ms_steps"), r["kernel_ms_last_call"], d["clocks"]))
    else: print(l.rstrip()[:300])
'
for ps in 250 2000; do
echo "== c4 fast per_step $ps"; JMM_BENCH_PER_STEP=$ps timeout 150 python bench.py --workload c4 --steps 8 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_c4_$TAG.json | python -c "$show"
done
nvidia-smi --query-gpu=clocks.sm,power.draw,power.limit,temperature.gpu --format=csv
timeout 300 python scripts/sweep_grid.py c3 "" "K=2,WARPS=13" "K=1,WARPS=24" "K=1,WARPS=16" "K=1,WARPS=32" 2>&1 | tee $OUT/grid_c3_$TAG.log
timeout 400 python scripts/sweep_grid.py c5 "" "K=1,WARPS=32,G=8" "K=1,WARPS=32,G=8,NSUB=32" "K=1,WARPS=28,G=8" "K=1,WARPS=32,G=4,NSUB=32" 2>&1 | tee $OUT/grid_c5_$TAG.log
