#!/bin/bash
# GPU visit: parity tests, then C4 with/without histograms (reference and fast arithmetic)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01e}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_$TAG.log
for extra in "" "--hist" "--arith fast" "--arith fast --hist"; do
  echo "== c4 $extra"
  python bench.py --workload c4 --steps 3 --warmup 3 $extra 2>&1 | tee -a $OUT/bench_c4_$TAG.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value %.4g %s  ms/step %.2f  launches %d  fp64 %.3f/%.1f TF = %.4f' % (d['value'], d['unit'], d['ms_per_step'], d['gpu_launches'], r['achieved'], r['peak'], r['frac']))
    else: print(l.rstrip()[:400])
"
done
