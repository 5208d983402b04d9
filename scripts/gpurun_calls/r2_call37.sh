#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -14 ) > gpurun_out/r2_c37_tests.log 2>&1
cat gpurun_out/r2_c37_tests.log
