#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2_c3_tests.log 2>&1
tail -3 gpurun_out/r2_c3_tests.log
timeout 1500 python scripts/sweep_perf.py \
  panini panini:cold panini:f1 panini:f4 panini:f64 \
  panini,BLINKY_RING_CTAS=8 panini,BLINKY_RING_CTAS=12 panini,BLINKY_RING_STAGES=3 panini,BLINKY_RING_STAGES=4 \
  panini,BLINKY_FCHUNK=1 panini,BLINKY_FCHUNK=2 panini,BLINKY_FCHUNK=4 panini,BLINKY_FCHUNK=8 panini,BLINKY_FCHUNK=16 \
  panini,BLINKY_MAX_BOX=4096 panini,BLINKY_MAX_BOX=6144 panini1080 stereo trism \
  quinc,BLINKY_MAX_BOX=4096 quinc quinc,BLINKY_MAX_BOX=16384 \
  equirect equirect,BLINKY_MAX_BOX=16384 \
  hammer hammer,BLINKY_MAX_BOX=16384 fisheye1 fisheye1,BLINKY_MAX_BOX=16384 \
  > gpurun_out/r2_c3_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c3_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[38:120])
PY
