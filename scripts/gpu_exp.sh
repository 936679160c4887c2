#!/bin/bash
# experiment: bench c2 under different library builds / G
run() { echo "== $*"; env "$@" JMM_BENCH_CPU_STEPS=2000 python bench.py --steps 5 --warmup 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g  ms/step %.2f' % (d['value'], d['ms_per_step']))
    else: print(l.rstrip()[:300])
"; }
run JMM_COOP_G=16
run JMM_COOP_G=32
for mb in 6 8; do
  run JMM_LIBJMMGPU=$PWD/build/exp/libjmmgpu_mb$mb.so JMM_COOP_G=16
  run JMM_LIBJMMGPU=$PWD/build/exp/libjmmgpu_mb$mb.so JMM_COOP_G=32
done
