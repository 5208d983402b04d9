#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_c33_tests.log 2>&1
tail -3 gpurun_out/r2_c33_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c33_bench.json 2> gpurun_out/r2_c33_bench.err
tail -c 300 gpurun_out/r2_c33_bench.json; tail -3 gpurun_out/r2_c33_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
