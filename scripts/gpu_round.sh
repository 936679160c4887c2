#!/bin/bash
# One GPU visit: full parity suite, smoke, default bench (with C3/C4/C5 extras), reference arm, launch list,
# and one ncu --set full capture per kernel family.  Usage (under gpurun): bash scripts/gpu_round.sh [tag]
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01m}
PT="--timeout 90 --timeout-method thread"
timeout 800 python -m pytest tests -m gpu -x -q $PT > $OUT/pytest_gpu_$TAG.log 2>&1
tail -4 $OUT/pytest_gpu_$TAG.log
if ! grep -q " passed" $OUT/pytest_gpu_$TAG.log || grep -q "failed\|Timeout" $OUT/pytest_gpu_$TAG.log; then echo "PARITY NOT GREEN"; grep -n "Timeout\|FAILED\|Error\|assert" $OUT/pytest_gpu_$TAG.log | head -20; fi
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv | tee $OUT/gpu_$TAG.txt
nproc | tee -a $OUT/gpu_$TAG.txt
timeout 400 python bench.py --steps 10 --warmup 3 2> $OUT/bench_$TAG.err > $OUT/bench_$TAG.json
python - <<EOF
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("C2 value %.4g e2e %.4g frac %.4f cpu %s" % (d["value"], d["e2e"]["value"], r["frac"], d["cpu_baseline"]["value"]))
for k,v in d["other_workloads"].items(): print(k, v.get("value"), v.get("fp64_frac"), v.get("error"))
EOF
tail -3 $OUT/bench_$TAG.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>> $OUT/bench_$TAG.err | tee $OUT/bench_ref_$TAG.json | cut -c1-300
bash scripts/bench_workloads.sh $TAG
# launch list (cold-cache, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/ncu_launches_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step -s 3 -c 1 -f -o $OUT/prof_c2_$TAG \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/ncu_c2_$TAG.log 2>&1; tail -1 $OUT/ncu_c2_$TAG.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_prod -s 1 -c 1 -f -o $OUT/prof_c4fast_$TAG \
    python bench.py --workload c4 --arith fast --steps 1 --warmup 3 --no-cpu > $OUT/ncu_c4fast_$TAG.log 2>&1; tail -1 $OUT/ncu_c4fast_$TAG.log | cut -c1-200
for w in c3 c5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fast -s 4 -c 1 -f -o $OUT/prof_${w}fast_$TAG \
      python bench.py --workload $w --arith fast --steps 1 --warmup 3 --no-cpu > $OUT/ncu_${w}fast_$TAG.log 2>&1; tail -1 $OUT/ncu_${w}fast_$TAG.log | cut -c1-200
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_${w}fast_$TAG.csv \
      python bench.py --workload $w --arith fast --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
done
ls -la $OUT | tail -20
