#!/bin/bash
# One GPU visit: parity tests, smoke, bench, ncu launch list + one full capture of the dominant kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv | tee $OUT/gpu_$TAG.txt
nproc | tee -a $OUT/gpu_$TAG.txt
python bench.py --steps 10 --warmup 3 2> $OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json
tail -5 $OUT/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 2>> $OUT/bench_$TAG.err | tee $OUT/bench_ref_$TAG.json
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 > $OUT/ncu_launches_$TAG.log 2>&1
# dominant kernel, full set
ncu --set full --clock-control none --import-source on -k regex:k_chains_step -s 3 -c 1 -f -o $OUT/prof_chains_$TAG \
    python bench.py --steps 2 --warmup 3 > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
