#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the two-rank C-ABI tests and the strong-scaling bench line at N GPUs.
N=${1:-2}; TAG=${2:-r2f}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_compat.py -m gpu -q --timeout 600 --timeout-method thread -k "two_rank or single_rank_allgather" > $OUT/pytest_multi_$TAG.log 2>&1
  tail -6 $OUT/pytest_multi_$TAG.log
  grep -n "FAILED\|Error\|assert " $OUT/pytest_multi_$TAG.log | head -20
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
   > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
    print("N=%d %s value %.4g e2e %.4g frac %.3f from0 %.4g weak_c2 %.4g cpu %s serial %s" % (d["n_gpus"], d["scaling"], d["value"], d["e2e"]["value"], d["roofline"]["frac"],
          d["from_step_zero"]["value"], d["weak_c2"]["value"], d["cpu_baseline"]["value"], (d["cpu_baseline"].get("main_serial") or {}).get("value")))
    print("config", json.dumps(d["config"])[:400]); print("clocks", d["clocks"], "timed_s", d["timed_s"])
except Exception as e:
    print("bench line FAILED", e)
PY
tail -6 $OUT/bench_${N}gpu_$TAG.err
