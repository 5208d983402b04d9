#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_c1_smi.txt
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2_c1_tests.log 2>&1
( timeout 300 ./scripts/tma_lab.bin ) > gpurun_out/r2_c1_tma_lab.log 2>&1
for lab in 0 1 2 3 4 6 8 9 15; do
  echo "LAB=$lab" >> gpurun_out/r2_c1_k2lab.log
  BLINKY_LAB=$lab timeout 200 python scripts/quick_perf.py --lens panini --zoom "f_fov 180" --threads 0 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['us_per_frame'], d['min_ms'], d['kernel'][:60])" >> gpurun_out/r2_c1_k2lab.log 2>&1
done
tail -5 gpurun_out/r2_c1_tests.log; cat gpurun_out/r2_c1_k2lab.log
