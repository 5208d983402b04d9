#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2_c5_tests.log 2>&1
tail -3 gpurun_out/r2_c5_tests.log
timeout 1200 python scripts/sweep_perf.py \
  panini panini,BLINKY_FCHUNK=8 panini,BLINKY_FCHUNK=8,BLINKY_LAB=7 panini,BLINKY_FCHUNK=8,BLINKY_LAB=2 panini,BLINKY_FCHUNK=8,BLINKY_LAB=4 \
  panini,BLINKY_FCHUNK=8,BLINKY_PREFETCH=1 panini,BLINKY_FCHUNK=8,BLINKY_PREFETCH=2 panini,BLINKY_FCHUNK=8,BLINKY_PREFETCH=3 panini,BLINKY_FCHUNK=8,BLINKY_PREFETCH=4 \
  panini,BLINKY_FCHUNK=8,BLINKY_PREFETCH=2,BLINKY_L2_PROMOTION=2 panini,BLINKY_FCHUNK=8,BLINKY_L2_PROMOTION=2 \
  panini,BLINKY_FCHUNK=8,BLINKY_PREFETCH=2,BLINKY_RING_STAGES=3 panini,BLINKY_FCHUNK=4,BLINKY_PREFETCH=2 panini,BLINKY_FCHUNK=16,BLINKY_PREFETCH=2 \
  panini:f1 panini:f1,BLINKY_PREFETCH=2 panini:cold panini:cold,BLINKY_PREFETCH=2 panini:f64,BLINKY_PREFETCH=2 \
  trism,BLINKY_FCHUNK=8,BLINKY_PREFETCH=2 quinc,BLINKY_PREFETCH=2 equirect,BLINKY_MAX_BOX=16384,BLINKY_PREFETCH=2 panini1080,BLINKY_PREFETCH=2 \
  > gpurun_out/r2_c5_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c5_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[38:100])
PY
