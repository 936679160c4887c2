SYNC_LINES=8 bash scripts/gpu_synccheck.sh 2>&1 | tee gpurun_out/synccheck4.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 --timeout-method thread -k "trio or crew or c2_ or one_long" 2>&1 | tail -3
for bnd in 4 5; do JMM_BOND=$bnd timeout 300 python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('JMM_BOND=$bnd', d['roofline']['kernel'][:20], '%.4g' % d['value'], '%.4g' % d['e2e']['value'])"; done
