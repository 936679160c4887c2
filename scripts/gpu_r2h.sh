#!/bin/bash
# Round-2 visit H: lanes.cuh v5 (two commuting trials per iteration): parity + throughput.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2h}
PT="--timeout 900 --timeout-method thread"
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q $PT -k "sweep_shape or fast_arithmetic or consistent_virial or lanes_paired" > $OUT/pytest_new_$TAG.log 2>&1
tail -5 $OUT/pytest_new_$TAG.log
grep -n "FAILED\|Error\|assert " $OUT/pytest_new_$TAG.log | head -30
b() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --workload c4 --arith fast --steps 5 --warmup 3 --no-cpu --no-e2e --min-seconds 0.3 2>>$OUT/bench_$TAG.err | tail -1 > $OUT/tmp_line.json
  python - "$label" <<PY
import json,sys
try:
    d=json.loads(open("$OUT/tmp_line.json").read())
    print("%-34s %.4g trials/s  frac %.3f  ms/step %.3f  acc %.3f" % (sys.argv[1], d["value"], d["roofline"]["frac"], d["ms_per_step"], d["acceptance"]))
    open("$OUT/c4_lanes_$TAG.jsonl","a").write(json.dumps({"label":sys.argv[1], **d})+"\n")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
b "8192 lanes G=8"           JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=5000
b "8192 lanes G=4"           JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=5000 JMM_LANES_G=4
b "8192 lanes G=16"          JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=5000 JMM_LANES_G=16
b "16384 lanes G=4"          JMM_BENCH_CHAINS=16384 JMM_BENCH_PER_STEP=5000
b "16384 lanes G=8"          JMM_BENCH_CHAINS=16384 JMM_BENCH_PER_STEP=5000 JMM_LANES_G=8
b "32768 lanes G=2"          JMM_BENCH_CHAINS=32768 JMM_BENCH_PER_STEP=5000
b "32768 lanes G=4"          JMM_BENCH_CHAINS=32768 JMM_BENCH_PER_STEP=5000 JMM_LANES_G=4
b "65536 lanes G=2"          JMM_BENCH_CHAINS=65536 JMM_BENCH_PER_STEP=5000 JMM_LANES_G=2
b "65536 prod"               JMM_BENCH_CHAINS=65536 JMM_BENCH_PER_STEP=5000
JMM_BENCH_CHAINS=8192 JMM_BENCH_PER_STEP=2000 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_lanes -s 1 -c 1 -f -o $OUT/prof_c4lanes_$TAG \
    python bench.py --workload c4 --arith fast --steps 1 --warmup 3 --no-cpu --no-e2e --min-seconds 0 > $OUT/ncu_c4lanes_$TAG.log 2>&1; tail -1 $OUT/ncu_c4lanes_$TAG.log | cut -c1-200
python scripts/ncu_summary.py $OUT/prof_c4lanes_$TAG.ncu-rep 4096000 > $OUT/prof_c4lanes_$TAG.txt 2>&1
python scripts/ncu_lines.py $OUT/prof_c4lanes_$TAG.ncu-rep 4096000 50 >> $OUT/prof_c4lanes_$TAG.txt 2>&1
rm -f $OUT/prof_c4lanes_$TAG.ncu-rep
tail -5 $OUT/bench_$TAG.err
