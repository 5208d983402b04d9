#!/bin/bash
# refresh the profiles with the final code: ncu --set full of every workload (16 frames and 1 frame), launch list of the bench
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_c38_all \
  python scripts/ncu_workloads.py --frames 16 --order gpurun_out/r2_c38_order16.json > gpurun_out/r2_c38_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_c38_one \
  python scripts/ncu_workloads.py --frames 1 --order gpurun_out/r2_c38_order1.json 4k-cube-panini 4k-cube-quincuncial-rubix 4k-cube-fisheye1 >> gpurun_out/r2_c38_ncu.log 2>&1
tail -3 gpurun_out/r2_c38_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c38_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2_c38_bench_under_ncu.log 2>&1
tail -c 200 gpurun_out/r2_c38_bench_under_ncu.log
