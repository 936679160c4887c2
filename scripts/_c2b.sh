#!/bin/bash
export JMM_BOND_EARLY_ECHECK=1
timeout 900 python -m pytest tests -m gpu -q --timeout 200 --timeout-method thread > gpurun_out/pytest_gpu_r02r.log 2>&1; tail -3 gpurun_out/pytest_gpu_r02r.log; grep -n "FAILED\|Error" gpurun_out/pytest_gpu_r02r.log | head
for v in 1 "" 1; do echo -n "early=$v: "; if [ -z "$v" ]; then unset JMM_BOND_EARLY_ECHECK; else export JMM_BOND_EARLY_ECHECK=1; fi; timeout 200 python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g e2e %.4g kernel_ms %.3f disc %s' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['config']['echeck_discrepancies']))"; done
