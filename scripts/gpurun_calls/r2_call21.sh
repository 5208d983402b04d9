#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest "tests/test_gpu_fullsize.py::test_full_size_device_built_map" -x -q -m gpu -k "C3" 2>&1 | tail -30 ) > gpurun_out/r2_c21_t1.log 2>&1
( BLINKY_SERIAL_GATHER=1 timeout 300 python -m pytest "tests/test_gpu_fullsize.py::test_full_size_device_built_map" -x -q -m gpu -k "C3" 2>&1 | tail -5 ) > gpurun_out/r2_c21_t2.log 2>&1
( BLINKY_RING_BOXES=1 timeout 300 python -m pytest "tests/test_gpu_fullsize.py::test_full_size_device_built_map" -x -q -m gpu -k "C3" 2>&1 | tail -5 ) > gpurun_out/r2_c21_t3.log 2>&1
( BLINKY_RING_CTAS=12 timeout 300 python -m pytest "tests/test_gpu_fullsize.py::test_full_size_device_built_map" -x -q -m gpu -k "C3" 2>&1 | tail -5 ) > gpurun_out/r2_c21_t4.log 2>&1
grep -h "passed\|failed\|Error\|assert" gpurun_out/r2_c21_t*.log | head -20
export BLINKY_SERIAL_GATHER=1
timeout 900 python scripts/sweep_perf.py \
  panini,BLINKY_RING_BOXES=1 panini,BLINKY_RING_BOXES=2 panini,BLINKY_RING_BOXES=3 panini,BLINKY_RING_BOXES=4 \
  panini,BLINKY_RING_BOXES=2,BLINKY_RING_CTAS=16 panini,BLINKY_RING_BOXES=3,BLINKY_RING_CTAS=16 panini,BLINKY_RING_BOXES=2,BLINKY_RING_CTAS=12 panini,BLINKY_RING_BOXES=3,BLINKY_RING_CTAS=12 panini,BLINKY_RING_BOXES=4,BLINKY_RING_CTAS=12 \
  panini,BLINKY_RING_BOXES=2,BLINKY_RING_CTAS=14 panini,BLINKY_RING_BOXES=2,BLINKY_RING_CTAS=18 panini,BLINKY_RING_BOXES=2,BLINKY_FCHUNK=4 panini,BLINKY_RING_BOXES=2,BLINKY_FCHUNK=16 \
  panini:f64,BLINKY_RING_BOXES=2 panini:f1,BLINKY_RING_BOXES=2 panini:f1,BLINKY_RING_BOXES=3 panini:cold,BLINKY_RING_BOXES=3 \
  stereo,BLINKY_RING_BOXES=2 trism,BLINKY_RING_BOXES=2 quinc,BLINKY_RING_BOXES=2 equirect,BLINKY_RING_BOXES=2 \
  > gpurun_out/r2_c21_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c21_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[-80:-40])
PY
