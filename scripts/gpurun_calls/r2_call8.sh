#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2_c8_tests.log 2>&1
tail -3 gpurun_out/r2_c8_tests.log
timeout 1200 python scripts/sweep_perf.py \
  panini,BLINKY_FCHUNK=8 panini,BLINKY_FCHUNK=8,BLINKY_RING_REGS=128 panini,BLINKY_FCHUNK=8,BLINKY_RING_REGS=128,BLINKY_MAX_BOX=4096 panini,BLINKY_FCHUNK=8,BLINKY_MAX_BOX=4096 \
  panini,BLINKY_FCHUNK=8,BLINKY_RING_STAGES=3 panini,BLINKY_FCHUNK=8,BLINKY_RING_CTAS=8 panini,BLINKY_FCHUNK=8,BLINKY_RING_CTAS=10 \
  panini panini,BLINKY_FCHUNK=4 panini,BLINKY_FCHUNK=16 panini,BLINKY_FCHUNK=8,BLINKY_STATIC_PCT=100 panini,BLINKY_FCHUNK=8,BLINKY_L2_PROMOTION=2 \
  panini:f1 panini:cold panini:f64 \
  trism quinc equirect,BLINKY_MAX_BOX=16384 hammer,BLINKY_MAX_BOX=16384 fisheye1,BLINKY_MAX_BOX=16384 fisheye1 panini1080 stereo \
  > gpurun_out/r2_c8_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c8_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[38:100])
PY
