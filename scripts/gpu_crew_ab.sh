#!/bin/bash
# A/B of k_chains_step_crew's tuning switches (variant libraries under jmmonedmc_b200/variants/, built with -DJMM_CREW_*=0)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-ab}
b() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-extras 2>>$OUT/bench_$TAG.err | tail -1 > $OUT/tmp_line.json
  python - "$label" <<PY
import json,sys
try:
    d=json.loads(open("$OUT/tmp_line.json").read())
    print("%-34s %.4g trials/s  e2e %.4g  ms/step %.3f  %s" % (sys.argv[1], d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel"][:24]))
    open("$OUT/c2_crew_ab_$TAG.jsonl","a").write(json.dumps({"label":sys.argv[1], "value": d["value"], "e2e": d["e2e"]["value"], "ms_per_step": d["ms_per_step"]})+"\n")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
b "tree"
for v in ${AB_VARIANTS:-DV0 FAKE}; do
  b "variant $v" JMM_LIBJMMGPU=$PWD/jmmonedmc_b200/variants/libjmmgpu_$v.so
done
b "tree (again)"
tail -3 $OUT/bench_$TAG.err
