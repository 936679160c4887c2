#!/bin/bash
# Round-2 visit E: bond2 v2 (converged, flat) parity + C2 throughput.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2e}
PT="--timeout 900 --timeout-method thread"
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q $PT -k "c2_bench_mode or philox_many or checkpoint or device_adaptation or energy_bookkeeping" > $OUT/pytest_new_$TAG.log 2>&1
tail -5 $OUT/pytest_new_$TAG.log
grep -n "FAILED\|Error\|assert " $OUT/pytest_new_$TAG.log | head -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
c2() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-extras 2>>$OUT/bench_$TAG.err | tail -1 > $OUT/tmp_line.json
  python - "$label" <<PY
import json,sys
try:
    d=json.loads(open("$OUT/tmp_line.json").read())
    print("%-34s %.4g trials/s  e2e %.4g frac %.4f  ms/step %.3f" % (sys.argv[1], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"]))
    open("$OUT/c2_$TAG.jsonl","a").write(json.dumps({"label":sys.argv[1], **d})+"\n")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
c2 "C2 bond2 (default)"
c2 "C2 bond (first kernel)"  JMM_BOND=1
c2 "C2 bond2 block 64"       JMM_BOND_BLOCK=64
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_bond2 -s 3 -c 1 -f -o $OUT/prof_c2_$TAG \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/ncu_c2_$TAG.log 2>&1; tail -1 $OUT/ncu_c2_$TAG.log | cut -c1-200
tail -5 $OUT/bench_$TAG.err
