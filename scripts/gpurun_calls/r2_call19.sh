#!/bin/bash
# lab build, ring kernel alone (BOX tiles only), back-to-back launches: what do conflicts / store fragments / flat boxes cost now?
mkdir -p gpurun_out
export BLINKY_B200_LIB=$PWD/blinky_b200/libblinky_b200_lab.so
export BLINKY_SPLIT_PERCENT=0
timeout 900 python scripts/sweep_perf.py panini \
  panini,BLINKY_LAB_NOK3=1 panini,BLINKY_LAB_NOK3=1,BLINKY_LAB=2 panini,BLINKY_LAB_NOK3=1,BLINKY_LAB=4 panini,BLINKY_LAB_NOK3=1,BLINKY_LAB=6 panini,BLINKY_LAB_NOK3=1,BLINKY_LAB=1 \
  panini,BLINKY_LAB_NOK3=1,BLINKY_LAB_FLAT=4 panini,BLINKY_LAB_NOK3=1,BLINKY_LAB_FLAT=4,BLINKY_LAB=6 panini,BLINKY_LAB_NOK3=1,BLINKY_LAB=8 panini,BLINKY_LAB_NOK3=1,BLINKY_LAB=16 panini,BLINKY_LAB_NOK3=1,BLINKY_LAB=23 \
  panini,BLINKY_LAB_NOK3=1,BLINKY_RING_CTAS=6 panini,BLINKY_LAB_NOK3=1,BLINKY_RING_CTAS=9 panini,BLINKY_LAB_NOK3=1,BLINKY_RING_STAGES=3 panini,BLINKY_LAB_NOK3=1,BLINKY_FCHUNK=4 panini,BLINKY_LAB_NOK3=1,BLINKY_FCHUNK=16 \
  panini:f64,BLINKY_LAB_NOK3=1 panini:f4,BLINKY_LAB_NOK3=1 panini:f1,BLINKY_LAB_NOK3=1 \
  stereo,BLINKY_LAB_NOK3=1 stereo,BLINKY_LAB_NOK3=1,BLINKY_LAB=6 quinc,BLINKY_LAB_NOK3=1 quinc,BLINKY_LAB_NOK3=1,BLINKY_LAB=6 \
  > gpurun_out/r2_c19_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c19_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[50:90])
PY
