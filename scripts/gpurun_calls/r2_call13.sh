#!/bin/bash
# ring kernel after the instruction diet: parity first, then the sweep
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2_c13_tests.log 2>&1
tail -4 gpurun_out/r2_c13_tests.log
timeout 900 python scripts/sweep_perf.py \
  panini panini:f1 panini:cold panini:f4 panini:f64 panini,BLINKY_RING_STAGES=3 panini,BLINKY_RING_CTAS=8 panini,BLINKY_FCHUNK=16 panini,BLINKY_FCHUNK=4 \
  trism quinc equirect hammer fisheye1 panini1080 panini1080:cold stereo \
  > gpurun_out/r2_c13_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c13_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[38:120])
PY
