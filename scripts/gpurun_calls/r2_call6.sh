#!/bin/bash
mkdir -p gpurun_out
BLINKY_FCHUNK=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_ring -c 1 -s 3 -o gpurun_out/prof_r2b_ring_panini -f python scripts/sweep_perf.py panini > gpurun_out/r2_c6_ncu.log 2>&1
BLINKY_FCHUNK=8 BLINKY_LAB=7 timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_ring -c 1 -s 3 -o gpurun_out/prof_r2b_ring_panini_lab7 -f python scripts/sweep_perf.py panini >> gpurun_out/r2_c6_ncu.log 2>&1
tail -2 gpurun_out/r2_c6_ncu.log
