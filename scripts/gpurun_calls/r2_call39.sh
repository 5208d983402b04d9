#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_transpile.py -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/r2_c39_tests.log 2>&1
cat gpurun_out/r2_c39_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
