#!/bin/bash
# Experiment visit: parity of the touched paths, C4 after cooperative scaling, launch-shape grid for C3/C5.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01n}
PT="--timeout 90 --timeout-method thread"
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compat.py -m gpu -x -q $PT -k "fast or checkerboard or production or hist or invariants" > $OUT/pytest_sub_$TAG.log 2>&1
tail -3 $OUT/pytest_sub_$TAG.log
grep -n "Timeout\|FAILED\|Error\|assert" $OUT/pytest_sub_$TAG.log | head -10
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.4g %s  ms/step %.2f  launches %d  fp64 %.3f/%.1f TF = %.4f acc %s" % (d["value"], d["unit"], d["ms_per_step"], d["gpu_launches"], r["achieved"], r["peak"], r["frac"], d.get("acceptance")))
    else: print(l.rstrip()[:300])
'
echo "== c4 fast"; timeout 150 python bench.py --workload c4 --steps 5 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_c4_$TAG.json | python -c "$show"
echo "== c4 reference"; timeout 150 python bench.py --workload c4 --steps 3 --warmup 3 2>&1 | tee -a $OUT/bench_c4_$TAG.json | python -c "$show"
timeout 400 python scripts/sweep_grid.py c3 "" "AUX=0" "K=2,WARPS=13,AUX=0" "K=2,WARPS=13" "K=2,WARPS=16" "K=2,WARPS=12" "K=1,WARPS=24" "K=1,WARPS=28" "K=1,WARPS=32" "K=1,WARPS=16" "K=1,WARPS=20" "K=4,WARPS=8" "K=4,WARPS=4" "K=1,WARPS=24,NSUB=32" 2>&1 | tee $OUT/grid_c3_$TAG.log
JMM_SWEEP_NOUNROLL=1 timeout 100 python scripts/sweep_grid.py c3 "" 2>&1 | tee -a $OUT/grid_c3_$TAG.log
timeout 500 python scripts/sweep_grid.py c5 "" "AUX=0" "K=2,WARPS=9,AUX=0" "K=2,WARPS=9" "K=2,WARPS=8" "K=2,WARPS=16" "K=2,WARPS=12" "K=1,WARPS=32" "K=1,WARPS=16" "K=1,WARPS=24" "K=1,WARPS=32,G=8" "K=2,WARPS=16,G=8" "K=1,WARPS=16,G=2" "K=2,WARPS=8,G=2" "K=1,WARPS=32,NSUB=8" "K=4,WARPS=8" 2>&1 | tee $OUT/grid_c5_$TAG.log
