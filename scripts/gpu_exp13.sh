#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02c}
PT="--timeout 120 --timeout-method thread"
timeout 900 python -m pytest tests -m gpu -q $PT > $OUT/pytest_gpu_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log
grep -n "Timeout\|FAILED\|Error" $OUT/pytest_gpu_$TAG.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 value %.4g e2e %.4g frac %.4f kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms']))"
show='
import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.4g %s  ms/step %.2f  fp64 = %.4f acc %s" % (d["value"], d["unit"], d["ms_per_step"], r["frac"], d.get("acceptance")))
'
for w in c4 c3 c5; do timeout 150 python bench.py --workload $w --steps 5 --warmup 3 --arith fast 2>/dev/null | python -c "$show"; done
