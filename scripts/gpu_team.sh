#!/bin/bash
# team.cuh (warp-specialised bookkeeping): parity + throughput against lanes.cuh.  Every GPU command under its own timeout:
# a producer/consumer kernel that deadlocks must not take the box down.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2t}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 --timeout-method thread -k "team" > $OUT/pytest_team_$TAG.log 2>&1
tail -6 $OUT/pytest_team_$TAG.log
grep -n "FAILED\|Error\|assert \|Timeout" $OUT/pytest_team_$TAG.log | head -20
b() {  # label, env...
  local label=$1; shift
  env "$@" timeout 120 python bench.py --workload c4 --arith fast --steps 5 --warmup 3 --no-cpu --no-e2e --min-seconds 0.3 2>>$OUT/bench_$TAG.err | tail -1 > $OUT/tmp_line.json
  python - "$label" <<PY
import json,sys
try:
    d=json.loads(open("$OUT/tmp_line.json").read())
    print("%-34s %.4g trials/s  frac %.3f  ms/step %.3f  acc %.3f" % (sys.argv[1], d["value"], d["roofline"]["frac"], d["ms_per_step"], d["acceptance"]))
    open("$OUT/c4_team_$TAG.jsonl","a").write(json.dumps({"label":sys.argv[1], **d})+"\n")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
b "8192 lanes G=8"           JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=5000
b "8192 team"                JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=5000 JMM_TEAM=1
b "8192 team sliced"         JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=5000 JMM_TEAM=1 JMM_FORCE_SLICE=1
b "16384 lanes G=4"          JMM_BENCH_CHAINS=16384 JMM_BENCH_PER_STEP=5000
b "16384 team"               JMM_BENCH_CHAINS=16384 JMM_BENCH_PER_STEP=5000 JMM_TEAM=1 JMM_LANES_G=8
b "8192 team from0"          JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=10000 JMM_BENCH_FROM_ZERO=1 JMM_TEAM=1
JMM_TEAM=1 JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=2000 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_team -s 1 -c 1 -f -o $OUT/prof_c4team_$TAG \
    python bench.py --workload c4 --arith fast --steps 1 --warmup 3 --no-cpu --no-e2e --min-seconds 0 > $OUT/ncu_c4team_$TAG.log 2>&1; tail -1 $OUT/ncu_c4team_$TAG.log | cut -c1-200
python scripts/ncu_summary.py $OUT/prof_c4team_$TAG.ncu-rep 4102000 > $OUT/prof_c4team_$TAG.txt 2>&1
python scripts/ncu_lines.py $OUT/prof_c4team_$TAG.ncu-rep 4102000 60 >> $OUT/prof_c4team_$TAG.txt 2>&1
rm -f $OUT/prof_c4team_$TAG.ncu-rep
tail -5 $OUT/bench_$TAG.err
