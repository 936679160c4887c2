#!/bin/bash
# ncu --set full of the C2 kernel only, condensed (no bench line)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-p}
export JMM_TRAFFIC_JSON=$PWD/$OUT/traffic_c2_$TAG.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chains_step_crew -s 3 -c 1 -f -o $OUT/prof_c2_$TAG \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/ncu_c2_$TAG.log 2>&1; tail -1 $OUT/ncu_c2_$TAG.log | cut -c1-200
python scripts/ncu_summary.py $OUT/prof_c2_$TAG.ncu-rep 51200000 --traffic k_chains_step_crew > $OUT/prof_c2_$TAG.txt 2>&1
echo "---- per source line (warp instructions per unit, share of stall samples)" >> $OUT/prof_c2_$TAG.txt
python scripts/ncu_lines.py $OUT/prof_c2_$TAG.ncu-rep 51200000 90 >> $OUT/prof_c2_$TAG.txt 2>&1
rm -f $OUT/prof_c2_$TAG.ncu-rep
