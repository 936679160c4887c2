#!/bin/bash
# the default bench line + launch list + ncu --set full of the C2 kernel (what scripts/gpu_round2.sh does for C2), stand-alone
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2z}
export JMM_TRAFFIC_JSON=$PWD/$OUT/traffic_c2_$TAG.json
condense() {
  python scripts/ncu_summary.py $OUT/$1.ncu-rep $2 --traffic $3 > $OUT/$1.txt 2>&1
  echo "---- per source line (warp instructions per unit, share of stall samples)" >> $OUT/$1.txt
  python scripts/ncu_lines.py $OUT/$1.ncu-rep $2 60 >> $OUT/$1.txt 2>&1
  rm -f $OUT/$1.ncu-rep
}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 --timeout-method thread -k "engine_names or c2_bench" 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_crew -s 3 -c 1 -f -o $OUT/prof_c2_$TAG \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/ncu_c2_$TAG.log 2>&1; tail -1 $OUT/ncu_c2_$TAG.log | cut -c1-200
condense prof_c2_$TAG 51200000 k_chains_step_crew
mkdir -p profiles.tmp && cp profiles/traffic.json profiles.tmp/traffic.json.bak
python - <<PY
import json
t=json.load(open("profiles/traffic.json")); n=json.load(open("$JMM_TRAFFIC_JSON")); t.update(n)
json.dump(t, open("profiles/traffic.json","w"), indent=1, sort_keys=True); json.dump(t, open("$OUT/traffic_merged_$TAG.json","w"), indent=1, sort_keys=True)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/ncu_launches_$TAG.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 2> $OUT/bench_$TAG.err > $OUT/bench_$TAG.json
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
r=d["roofline"]; print("C2 value %.4g e2e %.4g frac %.4f traffic %s kernel %s cpu %s timed %.2fs launches %s" % (d["value"], d["e2e"]["value"], r["frac"], r["traffic"], r["kernel"], d["cpu_baseline"]["value"], d["timed_s"], d["gpu_launches"]))
PY
rm -rf profiles.tmp
