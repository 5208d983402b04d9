#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_dropin.py tests/test_capi.py -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2_c35_tests.log 2>&1
tail -3 gpurun_out/r2_c35_tests.log
{
for cfg in "1 3" "2 3" "4 3" "8 3" "4 2" "4 4" "8 2"; do set -- $cfg; BLINKY_HOST_GROUP=$1 BLINKY_HOST_SLOTS=$2 python scripts/pcie_probe.py e2e | sed "s/^/group=$1 slots=$2 /"; done
} > gpurun_out/r2_c35_e2e.log 2>&1
cat gpurun_out/r2_c35_e2e.log | cut -c1-220
