#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r01x}
timeout 300 python scripts/sweep_grid.py c3 "" "" "STAGGER=200" "STAGGER=500" "STAGGER=1000" "STAGGER=2000" "STAGGER=4000" 2>&1 | tee $OUT/grid_c3_$TAG.log
timeout 400 python scripts/sweep_grid.py c5 "" "" "STAGGER=200" "STAGGER=500" "STAGGER=1000" "STAGGER=2000" "STAGGER=4000" "STAGGER=1000,G=4,WARPS=16" 2>&1 | tee $OUT/grid_c5_$TAG.log
