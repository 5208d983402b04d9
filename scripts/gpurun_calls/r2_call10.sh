#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_c10_tests.log 2>&1
tail -3 gpurun_out/r2_c10_tests.log
timeout 1200 python scripts/sweep_perf.py \
  panini panini:f1 panini:cold panini:f1,BLINKY_RING_STAGES=2 panini:f1,BLINKY_RING_STAGES=4 panini:f4 panini:f64 \
  trism quinc quinc,BLINKY_SPLIT_PERCENT=100 equirect equirect,BLINKY_MAX_BOX=16384 equirect,BLINKY_MAX_BOX=16384,BLINKY_SPLIT_PERCENT=100 hammer hammer,BLINKY_MAX_BOX=16384 fisheye1,BLINKY_MAX_BOX=16384 fisheye1 fisheye1,BLINKY_MAX_BOX=4096 panini1080 panini1080:cold stereo \
  > gpurun_out/r2_c10_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c10_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[38:86], (d.get('kernel') or '')[-50:])
PY
