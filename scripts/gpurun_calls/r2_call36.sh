#!/bin/bash
mkdir -p gpurun_out
{ timeout 500 python scripts/stress_determinism.py 150 --slow-stores 2>&1 | tail -5
  timeout 300 python scripts/stress_determinism.py 100 4k-cube-quincuncial-rubix --slow-stores 2>&1 | tail -5
  timeout 200 python scripts/stress_determinism.py 200 2>&1 | tail -3
} > gpurun_out/r2_c36_stress.log 2>&1
cat gpurun_out/r2_c36_stress.log
