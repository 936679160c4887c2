#!/bin/bash
# Round-2 visit B: lanes.cuh v2 (software-pipelined, thermo ring) — parity + A/B of warps per CTA / lanes per chain.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2b}
PT="--timeout 600 --timeout-method thread"
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q $PT -k "sweep_shape or fast_arithmetic or c2_bench_mode" > $OUT/pytest_new_$TAG.log 2>&1
tail -5 $OUT/pytest_new_$TAG.log
grep -n "FAILED\|Error\|assert " $OUT/pytest_new_$TAG.log | head -30
b() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --workload c4 --arith fast --steps 5 --warmup 3 --no-cpu --no-e2e --min-seconds 0.2 2>>$OUT/bench_$TAG.err | tail -1 > $OUT/tmp_line.json
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
W16=$PWD/jmmonedmc_b200/variants/libjmmgpu_w16.so
b "65536 prod"               JMM_BENCH_CHAINS=65536
b "8192 lanes G=8 w12"       JMM_BENCH_CHAINS=8192
b "8192 lanes G=8 w12 noslice" JMM_BENCH_CHAINS=8192 JMM_NO_SLICE=1
b "8192 lanes G=8 w16"       JMM_BENCH_CHAINS=8192 JMM_LIBJMMGPU=$W16
b "8192 lanes G=8 w16 w14"   JMM_BENCH_CHAINS=8192 JMM_LIBJMMGPU=$W16 JMM_LANES_WARPS=14
b "8192 lanes G=4 w12"       JMM_BENCH_CHAINS=8192 JMM_LANES_G=4
b "8192 lanes G=16 w12"      JMM_BENCH_CHAINS=8192 JMM_LANES_G=16
b "8192 lanes G=16 w16"      JMM_BENCH_CHAINS=8192 JMM_LANES_G=16 JMM_LIBJMMGPU=$W16
b "16384 lanes G=4 w12"      JMM_BENCH_CHAINS=16384
b "16384 lanes G=4 w16"      JMM_BENCH_CHAINS=16384 JMM_LIBJMMGPU=$W16
b "16384 lanes G=8 w12"      JMM_BENCH_CHAINS=16384 JMM_LANES_G=8
b "32768 lanes G=2 w12"      JMM_BENCH_CHAINS=32768
b "32768 lanes G=4 w12"      JMM_BENCH_CHAINS=32768 JMM_LANES_G=4
b "65536 lanes G=2 w12"      JMM_BENCH_CHAINS=65536 JMM_LANES_G=2
b "8192 lanes G=8 from0"     JMM_BENCH_CHAINS=8192 JMM_BENCH_FROM_ZERO=1 JMM_BENCH_PER_STEP=10000
JMM_BENCH_CHAINS=8192 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_lanes -s 1 -c 1 -f -o $OUT/prof_c4lanes_$TAG \
    python bench.py --workload c4 --arith fast --steps 1 --warmup 3 --no-cpu --no-e2e --min-seconds 0 > $OUT/ncu_c4lanes_$TAG.log 2>&1; tail -1 $OUT/ncu_c4lanes_$TAG.log | cut -c1-200
# the new bench line shapes, quickly (full default line comes later)
timeout 600 python bench.py --workload c4 --arith fast --steps 3 --warmup 3 > $OUT/bench_c4_$TAG.json 2>>$OUT/bench_$TAG.err; cut -c1-600 $OUT/bench_c4_$TAG.json
tail -5 $OUT/bench_$TAG.err
