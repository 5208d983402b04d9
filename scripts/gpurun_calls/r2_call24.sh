#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2_c24_tests.log 2>&1
tail -3 gpurun_out/r2_c24_tests.log
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_SERIAL_GATHER=1 panini,BLINKY_RING_CTAS=12 panini,BLINKY_RING_CTAS=14 panini,BLINKY_RING_BOXES=3 \
  panini:f1 panini:f1,BLINKY_RING_BOXES=3 panini:cold panini:f4 panini:f64 \
  trism quinc quinc,BLINKY_SERIAL_GATHER=1 equirect equirect,BLINKY_SERIAL_GATHER=1 hammer hammer,BLINKY_SERIAL_GATHER=1 fisheye1 fisheye1,BLINKY_SERIAL_GATHER=1 panini1080 panini1080:cold stereo \
  > gpurun_out/r2_c24_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c24_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    k=(d.get('kernel') or '')
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), k[k.find('grid='):][:40])
PY
