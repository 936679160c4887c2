#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01b}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in c2 c3 c4 c5; do
  echo "== workload $w"
  JMM_BENCH_CPU_STEPS=${JMM_BENCH_CPU_STEPS:-100000} python bench.py --workload $w --steps 5 --warmup 3 2>&1 | tee $OUT/bench_${w}_$TAG.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value %.4g %s  ms/step %.2f  launches %d  fp64 %.3f/%.1f TF = %.4f  clocks %s' % (d['value'], d['unit'], d['ms_per_step'], d['gpu_launches'], r['achieved'], r['peak'], r['frac'], d['clocks']))
        if d.get('e2e'): print('   e2e %.4g  cpu %s' % (d['e2e']['value'], d.get('cpu_baseline',{}).get('value')))
        if 'acceptance' in d: print('   acceptance', d['acceptance'])
    else: print(l.rstrip()[:400])
"
done
