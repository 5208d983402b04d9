#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2_c9_tests.log 2>&1
tail -3 gpurun_out/r2_c9_tests.log
timeout 1200 python scripts/sweep_perf.py \
  panini panini,BLINKY_RING_WARPS=14 panini,BLINKY_RING_WARPS=16 panini,BLINKY_FCHUNK=8 panini,BLINKY_FCHUNK=6 \
  panini:f1 panini:cold panini:f1,BLINKY_RING_WARPS=14 panini:f1,BLINKY_RING_STAGES=3 panini:f4 panini:f64 \
  trism quinc equirect,BLINKY_MAX_BOX=16384 hammer,BLINKY_MAX_BOX=16384 fisheye1,BLINKY_MAX_BOX=16384 fisheye1 fisheye1,BLINKY_MAX_BOX=12288 panini1080 panini1080:cold stereo \
  > gpurun_out/r2_c9_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c9_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[38:110])
PY
