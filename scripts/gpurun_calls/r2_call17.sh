#!/bin/bash
# is the launch's fixed cost the GATHER units at the end of the schedule?
mkdir -p gpurun_out
export BLINKY_B200_LIB=$PWD/blinky_b200/libblinky_b200_lab.so
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_SPLIT_PERCENT=0 panini,BLINKY_SPLIT_PERCENT=0,BLINKY_LAB=23 panini,BLINKY_SPLIT_PERCENT=0,BLINKY_LAB=16 panini,BLINKY_SPLIT_PERCENT=0,BLINKY_LAB=1 \
  panini,BLINKY_SPLIT_PERCENT=0,BLINKY_FCHUNK=4 panini,BLINKY_SPLIT_PERCENT=0,BLINKY_FCHUNK=2 panini,BLINKY_SPLIT_PERCENT=0,BLINKY_FCHUNK=16 \
  panini,BLINKY_SPLIT_PERCENT=0,BLINKY_RING_CTAS=6 panini,BLINKY_SPLIT_PERCENT=0,BLINKY_RING_STAGES=3 \
  panini:f64,BLINKY_SPLIT_PERCENT=0 panini:f1,BLINKY_SPLIT_PERCENT=0 panini:cold,BLINKY_SPLIT_PERCENT=0 panini:f4,BLINKY_SPLIT_PERCENT=0 \
  stereo,BLINKY_SPLIT_PERCENT=0 quinc,BLINKY_SPLIT_PERCENT=0 trism,BLINKY_SPLIT_PERCENT=0 panini1080,BLINKY_SPLIT_PERCENT=0 \
  > gpurun_out/r2_c17_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c17_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[50:90])
PY
