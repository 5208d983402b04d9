#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sweep_perf.py \
  panini:f1 panini:f1,BLINKY_STATIC_PCT=100 panini:f1,BLINKY_STATIC_PCT=100,BLINKY_RING_CTAS=16 panini:f1,BLINKY_STATIC_PCT=100,BLINKY_RING_CTAS=16,BLINKY_RING_BOXES=3 \
  panini:f1,BLINKY_STATIC_PCT=100,BLINKY_RING_CTAS=14,BLINKY_RING_BOXES=3 panini:f1,BLINKY_STATIC_PCT=100,BLINKY_RING_BOXES=3 \
  panini:cold,BLINKY_STATIC_PCT=100 panini:cold,BLINKY_STATIC_PCT=100,BLINKY_RING_CTAS=16,BLINKY_RING_BOXES=3 \
  panini:f2,BLINKY_STATIC_PCT=100 panini:f2 panini:f4,BLINKY_STATIC_PCT=100 panini:f4 panini,BLINKY_STATIC_PCT=100 \
  quinc:cold quinc:cold,BLINKY_STATIC_PCT=100 fisheye1:cold fisheye1:cold,BLINKY_STATIC_PCT=100 panini1080:cold panini1080:cold,BLINKY_STATIC_PCT=100 \
  > gpurun_out/r2_c31_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c31_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    k=(d.get('kernel') or '')
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), k[k.find('grid='):][:40])
PY
