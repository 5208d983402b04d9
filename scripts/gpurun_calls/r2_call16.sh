#!/bin/bash
# lab build: DRAM-side or L2->SM side?  faces L2-resident (bit 3), no box loads at all (bit 4)
mkdir -p gpurun_out
export BLINKY_B200_LIB=$PWD/blinky_b200/libblinky_b200_lab.so
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_LAB=8 panini,BLINKY_LAB=9 panini,BLINKY_LAB=16 panini,BLINKY_LAB=17 panini,BLINKY_LAB=23 \
  panini,BLINKY_LAB=8,BLINKY_RING_CTAS=6 panini,BLINKY_LAB=16,BLINKY_RING_CTAS=6 panini,BLINKY_LAB=17,BLINKY_RING_CTAS=6 \
  panini,BLINKY_LAB=8,BLINKY_RING_STAGES=3 panini,BLINKY_LAB=8,BLINKY_LAB_FLAT=4 \
  quinc,BLINKY_LAB=8 quinc,BLINKY_LAB=16 stereo,BLINKY_LAB=8 \
  > gpurun_out/r2_c16_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c16_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[50:90])
PY
