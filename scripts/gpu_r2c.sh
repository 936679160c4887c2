#!/bin/bash
# Round-2 visit C: lanes.cuh v3 (stage-wise partner batch) A/B, the whole GPU suite, the new default bench line.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2c}
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
b "8192 lanes G=8 w12"       JMM_BENCH_CHAINS=8192
b "8192 lanes G=8 w16"       JMM_BENCH_CHAINS=8192 JMM_LIBJMMGPU=$W16
b "8192 lanes G=8 w16 noslice" JMM_BENCH_CHAINS=8192 JMM_LIBJMMGPU=$W16 JMM_NO_SLICE=1
b "8192 lanes G=4 w12"       JMM_BENCH_CHAINS=8192 JMM_LANES_G=4
b "8192 lanes G=16 w16"      JMM_BENCH_CHAINS=8192 JMM_LANES_G=16 JMM_LIBJMMGPU=$W16
b "16384 lanes G=4 w12"      JMM_BENCH_CHAINS=16384
b "16384 lanes G=8 w16"      JMM_BENCH_CHAINS=16384 JMM_LANES_G=8 JMM_LIBJMMGPU=$W16
b "32768 lanes G=2 w12"      JMM_BENCH_CHAINS=32768
b "32768 lanes G=4 w12"      JMM_BENCH_CHAINS=32768 JMM_LANES_G=4
b "65536 lanes G=2 w12"      JMM_BENCH_CHAINS=65536 JMM_LANES_G=2
b "65536 lanes G=4 w12"      JMM_BENCH_CHAINS=65536 JMM_LANES_G=4
b "65536 prod"               JMM_BENCH_CHAINS=65536
JMM_BENCH_CHAINS=8192 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_lanes -s 1 -c 1 -f -o $OUT/prof_c4lanes_$TAG \
    python bench.py --workload c4 --arith fast --steps 1 --warmup 3 --no-cpu --no-e2e --min-seconds 0 > $OUT/ncu_c4lanes_$TAG.log 2>&1; tail -1 $OUT/ncu_c4lanes_$TAG.log | cut -c1-200
PT="--timeout 1500 --timeout-method thread"
timeout 2400 python -m pytest tests -m gpu -q $PT --durations=8 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -14 $OUT/pytest_gpu_$TAG.log
grep -n "FAILED\|Error\|assert " $OUT/pytest_gpu_$TAG.log | head -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 2> $OUT/bench_$TAG.err2 > $OUT/bench_$TAG.json
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("C2 value %.4g e2e %.4g frac %.4f cpu %s serial %s timed %.2fs" % (d["value"], d["e2e"]["value"], r["frac"], d["cpu_baseline"]["value"], d["main_serial"]["value"], d["timed_s"]))
    for k,v in d["other_workloads"].items(): print(k, v.get("value"), v.get("fp64_frac"), v.get("timed_s"), (v.get("e2e") or {}).get("value"), (v.get("cpu_baseline") or {}).get("value"), v.get("error"))
    print("strong", d["strong"] and d["strong"]["value"])
except Exception as e:
    print("bench line FAILED", e)
PY
tail -5 $OUT/bench_$TAG.err2
