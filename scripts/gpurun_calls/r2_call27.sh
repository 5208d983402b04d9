#!/bin/bash
mkdir -p gpurun_out
export BLINKY_B200_LIB=$PWD/blinky_b200/libblinky_b200_lab.so
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_LAB=256 panini,BLINKY_LAB=512 panini,BLINKY_LAB=768 \
  panini,BLINKY_L2_PROMOTION=3 panini,BLINKY_L2_PROMOTION=2 panini,BLINKY_STATIC_PCT=70 panini,BLINKY_STATIC_PCT=95 panini,BLINKY_STATIC_PCT=50 \
  panini,BLINKY_RING_BOXES=3 panini,BLINKY_RING_BYTES=16384 panini,BLINKY_RING_CTAS=13 panini,BLINKY_RING_CTAS=11 \
  panini:f1 panini:f1,BLINKY_RING_CTAS=16 panini:f1,BLINKY_RING_CTAS=16,BLINKY_RING_BOXES=3 panini:f1,BLINKY_STATIC_PCT=50 panini:f1,BLINKY_STATIC_PCT=0 panini:cold panini:cold,BLINKY_RING_CTAS=16,BLINKY_STATIC_PCT=50 \
  quinc quinc,BLINKY_RING_BOXES=3 stereo trism \
  > gpurun_out/r2_c27_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c27_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    k=(d.get('kernel') or '')
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), k[k.find('grid='):][:40], k[k.find('TMA box ring'):][:52])
PY
