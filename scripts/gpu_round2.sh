#!/bin/bash
# Round-2 final visit: full parity suite, smoke, the default bench line, the reference arm, launch list, one ncu --set full
# capture per kernel family (with traffic.json), the fp64-peak microbenchmark under ncu, compute-sanitizer on the new kernels.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2z}
export JMM_TRAFFIC_JSON=$PWD/$OUT/traffic_$TAG.json
# gpurun brings back at most 64 MiB: every .ncu-rep is condensed to text on the box (scripts/ncu_summary.py, ncu_lines.py) and removed
condense() {  # rep-basename units traffic-key
  python scripts/ncu_summary.py $OUT/$1.ncu-rep $2 --traffic $3 > $OUT/$1.txt 2>&1
  echo "---- per source line (warp instructions per unit, share of stall samples)" >> $OUT/$1.txt
  python scripts/ncu_lines.py $OUT/$1.ncu-rep $2 40 >> $OUT/$1.txt 2>&1
  rm -f $OUT/$1.ncu-rep
}
PT="--timeout 1500 --timeout-method thread"
timeout 2400 python -m pytest tests -m gpu -q $PT --durations=6 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -12 $OUT/pytest_gpu_$TAG.log
grep -n "FAILED\|Error\|assert " $OUT/pytest_gpu_$TAG.log | head -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv | tee $OUT/gpu_$TAG.txt
nproc | tee -a $OUT/gpu_$TAG.txt
timeout 900 python bench.py --steps 10 --warmup 3 2> $OUT/bench_$TAG.err > $OUT/bench_$TAG.json
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("C2 value %.4g e2e %.4g frac %.4f cpu %s serial %s timed %.2fs clocks %s" % (d["value"], d["e2e"]["value"], r["frac"], d["cpu_baseline"]["value"], d["main_serial"]["value"], d["timed_s"], d["clocks"]))
    for k,v in d["other_workloads"].items(): print(k, "%.4g" % v.get("value", 0), v.get("fp64_frac"), v.get("timed_s"), (v.get("e2e") or {}).get("value"), (v.get("cpu_baseline") or {}).get("value"), v.get("from_step_zero", {}).get("value"), v.get("error"))
    print("strong", d["strong"] and d["strong"]["value"])
except Exception as e:
    print("bench line FAILED", e)
PY
tail -3 $OUT/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>> $OUT/bench_$TAG.err | tee $OUT/bench_ref_$TAG.json | cut -c1-300
for w in c3 c4 c5; do
  timeout 600 python bench.py --workload $w --arith fast --steps 5 --warmup 3 2>> $OUT/bench_$TAG.err > $OUT/bench_${w}_fast_$TAG.json
  python -c "
import json
d = json.loads(open('$OUT/bench_${w}_fast_$TAG.json').read().strip().splitlines()[-1]); c = d['cpu_baseline']
print('$w value %.4g frac %.4f e2e %.4g cpu %s (%s cores, %s) single-core %s serial %s' % (d['value'], d['roofline']['frac'], d['e2e']['value'], c['value'], c['cores'], c['kind'], c.get('single_core_value'), (c.get('main_serial') or {}).get('value')))
"
done
JMM_BENCH_CHAINS=8192 timeout 600 python bench.py --workload c4 --arith fast --steps 5 --warmup 3 --no-cpu 2>> $OUT/bench_$TAG.err > $OUT/bench_c4_8192_fast_$TAG.json
# launch list (cold-cache, serialised: shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/ncu_launches_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_crew -s 3 -c 1 -f -o $OUT/prof_c2_$TAG \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/ncu_c2_$TAG.log 2>&1; tail -1 $OUT/ncu_c2_$TAG.log | cut -c1-200
condense prof_c2_$TAG 51200000 k_chains_step_crew           # unit = one step of a 32-chain CTA (128 CTAs x 400 000 steps)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_prod -s 1 -c 1 -f -o $OUT/prof_c4fast_$TAG \
    python bench.py --workload c4 --arith fast --steps 1 --warmup 3 --no-cpu --no-e2e --min-seconds 0 > $OUT/ncu_c4fast_$TAG.log 2>&1; tail -1 $OUT/ncu_c4fast_$TAG.log | cut -c1-200
condense prof_c4fast_$TAG 40960000 k_chains_step_prod_sliced  # unit = one warp step (2048 tiles x 20 000 steps)
JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=2000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_lanes -s 1 -c 1 -f -o $OUT/prof_c4lanes_$TAG \
    python bench.py --workload c4 --arith fast --steps 1 --warmup 3 --no-cpu --no-e2e --min-seconds 0 > $OUT/ncu_c4lanes_$TAG.log 2>&1; tail -1 $OUT/ncu_c4lanes_$TAG.log | cut -c1-200
condense prof_c4lanes_$TAG 4096000 k_chains_step_lanes      # unit = one warp step (2048 tiles x 2000 steps)
for w in c3 c5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fast -s 4 -c 1 -f -o $OUT/prof_${w}fast_$TAG \
      python bench.py --workload $w --arith fast --steps 1 --warmup 3 --no-cpu --no-e2e --min-seconds 0 > $OUT/ncu_${w}fast_$TAG.log 2>&1; tail -1 $OUT/ncu_${w}fast_$TAG.log | cut -c1-200
  condense prof_${w}fast_$TAG 13421772 k_sweep_fast_$w          # unit = one trial of a 64-half-sweep launch (approx.)
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_${w}fast_$TAG.csv \
      python bench.py --workload $w --arith fast --steps 1 --warmup 3 --no-cpu --no-e2e --min-seconds 0 > /dev/null 2>&1
done
# the roofline denominator itself: the DFMA microbenchmark under ncu (pipe_fp64 counter)
timeout 200 ncu --set full --clock-control none -k regex:k_fp64_peak -s 2 -c 1 -f -o $OUT/prof_fp64peak_$TAG \
    python -c "import jmmonedmc_b200 as J; print(J.lib().jmm_fp64_peak_tflops(0))" > $OUT/ncu_fp64peak_$TAG.log 2>&1; tail -2 $OUT/ncu_fp64peak_$TAG.log | cut -c1-200
condense prof_fp64peak_$TAG 1 k_fp64_peak
SAN_CASES="solo trio crew bond2 lanes lanes80" bash scripts/gpu_sanitize.sh > $OUT/sanitizer_$TAG.txt 2>&1; tail -14 $OUT/sanitizer_$TAG.txt
ls -la $OUT | tail -30
