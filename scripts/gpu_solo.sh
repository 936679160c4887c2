#!/bin/bash
# solo.cuh (one chain per thread, a warp per SM) against bond.cuh on the C2 workload: parity, throughput, profile.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2v}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 --timeout-method thread -k "solo or trio or crew or c2_bench" > $OUT/pytest_solo_$TAG.log 2>&1
tail -5 $OUT/pytest_solo_$TAG.log
grep -n "FAILED\|Error\|assert \|Timeout" $OUT/pytest_solo_$TAG.log | head -20
b() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-extras 2>>$OUT/bench_$TAG.err | tail -1 > $OUT/tmp_line.json
  python - "$label" <<PY
import json,sys
try:
    d=json.loads(open("$OUT/tmp_line.json").read())
    print("%-28s %.4g trials/s  e2e %.4g  ms/step %.3f" % (sys.argv[1], d["value"], d["e2e"]["value"], d["ms_per_step"]))
    open("$OUT/c2_solo_$TAG.jsonl","a").write(json.dumps({"label":sys.argv[1], **d})+"\n")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
b "bond (16 lanes/chain)"  JMM_BOND=1
b "solo (1 thread/chain)"  JMM_BOND=3
b "trio (3 warps/32 chains)" JMM_BOND=4
b "crew (5 warps/32 chains)" JMM_BOND=5
JMM_BOND=${PROF_BOND:-5} timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_chains_step_(solo|trio|crew)" -s 1 -c 1 -f -o $OUT/prof_c2solo_$TAG \
    python bench.py --steps 1 --warmup 2 --no-extras > $OUT/ncu_c2solo_$TAG.log 2>&1; tail -1 $OUT/ncu_c2solo_$TAG.log | cut -c1-200
python scripts/ncu_summary.py $OUT/prof_c2solo_$TAG.ncu-rep 25600000 > $OUT/prof_c2solo_$TAG.txt 2>&1
python scripts/ncu_lines.py $OUT/prof_c2solo_$TAG.ncu-rep 25600000 70 >> $OUT/prof_c2solo_$TAG.txt 2>&1
rm -f $OUT/prof_c2solo_$TAG.ncu-rep
tail -5 $OUT/bench_$TAG.err
