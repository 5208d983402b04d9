#!/bin/bash
mkdir -p gpurun_out
export BLINKY_B200_LIB=$PWD/blinky_b200/libblinky_b200_lab.so
export BLINKY_SERIAL_GATHER=1 BLINKY_LAB_NOK3=1
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_LAB=2 panini,BLINKY_LAB=34 panini,BLINKY_LAB=42 panini,BLINKY_LAB=35 \
  panini,BLINKY_LAB=8 panini,BLINKY_LAB=8,BLINKY_RING_CTAS=12 panini,BLINKY_LAB=8,BLINKY_RING_CTAS=8 panini,BLINKY_RING_CTAS=12 panini,BLINKY_RING_CTAS=8 \
  panini,BLINKY_LAB=1 panini,BLINKY_LAB=1,BLINKY_RING_CTAS=12 panini,BLINKY_LAB=1,BLINKY_RING_CTAS=8 panini,BLINKY_LAB=16 panini,BLINKY_LAB=16,BLINKY_RING_CTAS=8 panini,BLINKY_LAB=23 panini,BLINKY_LAB=23,BLINKY_RING_CTAS=8 \
  > gpurun_out/r2_c23_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c23_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    k=(d.get('kernel') or '')
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), k[k.find('grid='):][:30])
PY
