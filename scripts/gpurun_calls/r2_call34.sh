#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_c34_tests.log 2>&1
tail -3 gpurun_out/r2_c34_tests.log
timeout 600 python scripts/sweep_perf.py \
  panini panini,BLINKY_RING_BOXES=2 panini,BLINKY_RING_BOXES=4 panini:f1 panini:f1,BLINKY_RING_BOXES=4 panini:f1,BLINKY_RING_BOXES=2 panini:f1,BLINKY_RING_CTAS=16 panini:cold panini:f2 panini:f4 panini:f64 \
  trism quinc quinc:cold equirect fisheye1 panini1080 panini1080:cold panini1080:f1 stereo \
  > gpurun_out/r2_c34_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c34_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    k=(d.get('kernel') or '')
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), k[k.find('TMA box ring'):][:48])
PY
