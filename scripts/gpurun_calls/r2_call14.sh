#!/bin/bash
# lab build: how much do shared-memory bank conflicts, 32-byte store fragments and the stores themselves cost?
mkdir -p gpurun_out
export BLINKY_B200_LIB=$PWD/blinky_b200/libblinky_b200_lab.so
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_LAB=1 panini,BLINKY_LAB=2 panini,BLINKY_LAB=4 panini,BLINKY_LAB=6 panini,BLINKY_LAB=3 panini,BLINKY_LAB=7 \
  panini,BLINKY_RING_CTAS=6 panini,BLINKY_LAB=6,BLINKY_RING_CTAS=6 panini,BLINKY_LAB=6,BLINKY_RING_CTAS=8 panini,BLINKY_LAB=6,BLINKY_RING_CTAS=10 \
  panini,BLINKY_LAB=6,BLINKY_FCHUNK=4 panini,BLINKY_LAB=6,BLINKY_RING_STAGES=3 \
  panini:cold panini:cold,BLINKY_LAB=6 quinc quinc,BLINKY_LAB=6 stereo stereo,BLINKY_LAB=6 \
  > gpurun_out/r2_c14_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c14_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[50:90])
PY
