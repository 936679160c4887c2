#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01q}
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.4g %s  ms/step %.2f  launches %d  fp64 %.3f/%.1f TF = %.4f acc %s steps %s last %.3f" % (d["value"], d["unit"], d["ms_per_step"], d["gpu_launches"], r["achieved"], r["peak"], r["frac"], d.get("acceptance"), d.get("ms_steps"), r["kernel_ms_last_call"]))
    else: print(l.rstrip()[:300])
'
echo "== c4 fast (production phase)"; timeout 150 python bench.py --workload c4 --steps 8 --warmup 3 --arith fast 2>&1 | tee -a $OUT/bench_c4_$TAG.json | python -c "$show"
echo "== c4 reference (production phase)"; timeout 150 python bench.py --workload c4 --steps 3 --warmup 3 2>&1 | tee -a $OUT/bench_c4_$TAG.json | python -c "$show"
for v in A B C D E F; do
  echo "=== variant $v"
  JMM_LIBJMMGPU=$PWD/jmmonedmc_b200/variants/libjmmgpu_$v.so timeout 200 python scripts/sweep_grid.py c3 "" "K=2,WARPS=13" 2>&1 | grep -v "fp64 peak"
  JMM_LIBJMMGPU=$PWD/jmmonedmc_b200/variants/libjmmgpu_$v.so timeout 200 python scripts/sweep_grid.py c5 "" "K=1,WARPS=24,G=8" 2>&1 | grep -v "fp64 peak"
done 2>&1 | tee $OUT/variants_$TAG.log
