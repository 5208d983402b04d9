#!/bin/bash
# lab build: is the floor the number of box rows (TMA requests / DRAM pages), not the bytes?
mkdir -p gpurun_out
export BLINKY_B200_LIB=$PWD/blinky_b200/libblinky_b200_lab.so
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_LAB_FLAT=2 panini,BLINKY_LAB_FLAT=4 panini,BLINKY_LAB_FLAT=8 panini,BLINKY_LAB_FLAT=4,BLINKY_LAB=1 panini,BLINKY_LAB_FLAT=4,BLINKY_LAB=7 \
  panini,BLINKY_RING_CTAS=6,BLINKY_LAB=1 panini,BLINKY_RING_CTAS=6,BLINKY_LAB_FLAT=4 panini,BLINKY_RING_STAGES=3,BLINKY_LAB_FLAT=4 \
  stereo stereo,BLINKY_LAB_FLAT=4 quinc quinc,BLINKY_LAB_FLAT=4 trism trism,BLINKY_LAB_FLAT=4 panini:cold,BLINKY_LAB_FLAT=4 \
  > gpurun_out/r2_c15_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c15_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[50:90])
PY
