#!/bin/bash
# final profiles of the round: GPU suite, ncu --set full of every workload (16 frames and 1 frame), launch list, bench (both arms)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_c28_tests.log 2>&1
tail -3 gpurun_out/r2_c28_tests.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_c28_all \
  python scripts/ncu_workloads.py --frames 16 --order gpurun_out/r2_c28_order16.json > gpurun_out/r2_c28_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_c28_one \
  python scripts/ncu_workloads.py --frames 1 --order gpurun_out/r2_c28_order1.json 4k-cube-panini 4k-cube-quincuncial-rubix 4k-cube-fisheye1 >> gpurun_out/r2_c28_ncu.log 2>&1
tail -3 gpurun_out/r2_c28_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c28_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2_c28_bench_under_ncu.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_c28_bench_ref.json 2> gpurun_out/r2_c28_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c28_bench.json 2>> gpurun_out/r2_c28_bench.err
tail -c 400 gpurun_out/r2_c28_bench.json; tail -3 gpurun_out/r2_c28_bench.err
