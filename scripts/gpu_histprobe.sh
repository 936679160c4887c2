#!/bin/bash
for g in 1 2; do for c in 4096; do
echo "== G=$g chains=$c"; JMM_PROD_G=$g PROBE_CHAINS=$c timeout 100 python scripts/hist_probe.py 2>&1 | tail -12
done; done
echo "== G=1 no hist"; JMM_PROD_G=1 PROBE_HIST=0 timeout 60 python scripts/hist_probe.py 2>&1 | tail -9
