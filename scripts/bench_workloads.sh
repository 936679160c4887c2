#!/bin/bash
# The throughput-bound BASELINE.json workloads (C3, C4, C5) stand-alone, fast arithmetic, each with its CPU
# side-by-side sample.  Usage (under gpurun): bash scripts/bench_workloads.sh [tag]
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-workloads}
for w in c3 c4 c5; do
  timeout 400 python bench.py --workload $w --arith fast --steps 5 --warmup 3 2>> $OUT/bench_$TAG.err > $OUT/bench_${w}_$TAG.json
  python -c "
import json, sys
d = json.loads(open('$OUT/bench_${w}_$TAG.json').read().strip().splitlines()[-1]); c = d['cpu_baseline']
print('$w value %.4g frac %.4f cpu %s (%s cores, %s) single-core %s' % (d['value'], d['roofline']['frac'], c['value'], c['cores'], c['kind'], c.get('single_core_value')))
"
done
tail -3 $OUT/bench_$TAG.err
