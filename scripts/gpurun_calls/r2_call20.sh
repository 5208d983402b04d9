#!/bin/bash
# ring kernel: BOX+EMPTY only at 20 (16 with rubix) warps/SM, byte-granular box ring; gather tiles in K3 beside it
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2_c20_tests.log 2>&1
tail -4 gpurun_out/r2_c20_tests.log
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_SERIAL_GATHER=1 panini,BLINKY_RING_CTAS=16 panini,BLINKY_RING_CTAS=12 panini,BLINKY_RING_CTAS=8 panini,BLINKY_FCHUNK=4 panini,BLINKY_FCHUNK=16 \
  panini:f1 panini:cold panini:f4 panini:f64 \
  trism quinc quinc,BLINKY_SERIAL_GATHER=1 equirect equirect,BLINKY_SERIAL_GATHER=1 hammer fisheye1 fisheye1,BLINKY_SERIAL_GATHER=1 panini1080 panini1080:cold stereo \
  > gpurun_out/r2_c20_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c20_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[-95:])
PY
