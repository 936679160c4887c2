#!/bin/bash
# quick GPU visit: parity tests + bench under a few JMM_COOP_G settings
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for g in 32 16 8 0; do
  echo "== JMM_COOP_G=$g"
  JMM_COOP_G=$g python bench.py --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g  e2e %.4g  ms/step %.2f  kernel_ms %.2f cpu %.4g' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['cpu_baseline']['value'] or 0))
    else: print(l.rstrip()[:300])
"
done
