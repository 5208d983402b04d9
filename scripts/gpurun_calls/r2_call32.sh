#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sweep_perf.py \
  panini1080:cold panini1080:cold,BLINKY_MERGED_ITEMS=0 panini1080:f1 panini1080:f1,BLINKY_MERGED_ITEMS=0 \
  quinc:cold quinc:cold,BLINKY_MERGED_ITEMS=0 quinc:f1 quinc:f1,BLINKY_MERGED_ITEMS=0 \
  equirect:cold,BLINKY_MERGED_ITEMS=100000 equirect:cold fisheye1:cold,BLINKY_MERGED_ITEMS=100000 fisheye1:cold \
  quinc:f4 quinc:f4,BLINKY_MERGED_ITEMS=100000 panini1080:f4 panini1080:f4,BLINKY_MERGED_ITEMS=100000 panini1080 panini1080,BLINKY_MERGED_ITEMS=100000 \
  > gpurun_out/r2_c32_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c32_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    k=(d.get('kernel') or '')
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), k[k.find('grid='):][:40])
PY
