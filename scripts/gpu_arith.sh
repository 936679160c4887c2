#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
for w in c3 c4 c5; do for ar in reference fast; do echo "== $w arith=$ar"; python bench.py --workload $w --arith $ar --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('value %.4g  ms/step %.2f  fp64 %.3f TF = %.4f acc %.4f' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], d['acceptance']))
    else: print(l.rstrip()[:300])
"; done; done
