#!/bin/bash
# Round-2 visit A: new parity tests (sweep shape, lanes.cuh) + C4 at 8192..65536 chains on the lanes/prod kernels.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2a}
PT="--timeout 300 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q $PT -k "sweep_shape or fast_arithmetic or c2_bench_mode or philox_many or continue_at" > $OUT/pytest_new_$TAG.log 2>&1
tail -5 $OUT/pytest_new_$TAG.log
grep -n "FAILED\|Error\|assert " $OUT/pytest_new_$TAG.log | head -20
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee $OUT/gpu_$TAG.txt
b() {  # label, env..., -- extra args
  local label=$1; shift
  env "$@" timeout 300 python bench.py --workload c4 --arith fast --steps 5 --warmup 3 --no-cpu 2>>$OUT/bench_$TAG.err | tail -1 > $OUT/tmp_line.json
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
b "65536 prod"            JMM_BENCH_CHAINS=65536
b "8192 prod (G=0)"       JMM_BENCH_CHAINS=8192 JMM_LANES_G=0
b "8192 lanes G=8 w16"    JMM_BENCH_CHAINS=8192
b "8192 lanes G=8 w12"    JMM_BENCH_CHAINS=8192 JMM_LANES_WARPS=12
b "8192 lanes G=8 w8"     JMM_BENCH_CHAINS=8192 JMM_LANES_WARPS=8
b "8192 lanes G=8 noslice" JMM_BENCH_CHAINS=8192 JMM_NO_SLICE=1
b "8192 lanes G=4"        JMM_BENCH_CHAINS=8192 JMM_LANES_G=4
b "8192 lanes G=16"       JMM_BENCH_CHAINS=8192 JMM_LANES_G=16
b "8192 lanes G=8 generic" JMM_BENCH_CHAINS=8192 JMM_LANES_GENERIC=1
b "16384 lanes G=4"       JMM_BENCH_CHAINS=16384
b "16384 lanes G=8"       JMM_BENCH_CHAINS=16384 JMM_LANES_G=8
b "16384 prod"            JMM_BENCH_CHAINS=16384 JMM_LANES_G=0
b "32768 lanes G=2"       JMM_BENCH_CHAINS=32768
b "32768 lanes G=4"       JMM_BENCH_CHAINS=32768 JMM_LANES_G=4
b "32768 prod"            JMM_BENCH_CHAINS=32768 JMM_LANES_G=0
b "8192 lanes G=8 from0"  JMM_BENCH_CHAINS=8192 JMM_BENCH_FROM_ZERO=1 JMM_BENCH_PER_STEP=10000
b "65536 prod from0"      JMM_BENCH_CHAINS=65536 JMM_BENCH_FROM_ZERO=1 JMM_BENCH_PER_STEP=10000
JMM_BENCH_CHAINS=8192 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_lanes -s 1 -c 1 -f -o $OUT/prof_c4lanes_$TAG \
    python bench.py --workload c4 --arith fast --steps 1 --warmup 3 --no-cpu > $OUT/ncu_c4lanes_$TAG.log 2>&1; tail -1 $OUT/ncu_c4lanes_$TAG.log | cut -c1-200
ls -la $OUT | tail
