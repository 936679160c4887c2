#!/bin/bash
# Experiment visit: parity of the touched paths, C4, launch-shape grid for C3/C5, ncu of the sweep kernel.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01o}
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
for w in c3 c5; do echo "== $w fast"; timeout 150 python bench.py --workload $w --steps 5 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_${w}_$TAG.json | python -c "$show"; done
timeout 300 python scripts/sweep_grid.py c3 "" "K=2,WARPS=13" "K=2,WARPS=16" "K=1,WARPS=24" "K=1,WARPS=28" "K=1,WARPS=16" "K=1,WARPS=24,NSUB=32" "K=1,WARPS=24,G=2" "K=1,WARPS=32,G=2" 2>&1 | tee $OUT/grid_c3_$TAG.log
timeout 400 python scripts/sweep_grid.py c5 "" "K=2,WARPS=16" "K=1,WARPS=32" "K=1,WARPS=24" "K=1,WARPS=32,G=8" "K=1,WARPS=32,G=16" "K=2,WARPS=16,G=8" "K=1,WARPS=32,G=8,NSUB=8" "K=1,WARPS=32,G=8,NSUB=32" "K=1,WARPS=28,G=8" 2>&1 | tee $OUT/grid_c5_$TAG.log
for w in c3 c5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fast -s 2 -c 1 -f -o $OUT/prof_${w}fast_$TAG \
      python bench.py --workload $w --arith fast --steps 1 --warmup 3 > $OUT/ncu_${w}fast_$TAG.log 2>&1; tail -1 $OUT/ncu_${w}fast_$TAG.log | cut -c1-200
done
