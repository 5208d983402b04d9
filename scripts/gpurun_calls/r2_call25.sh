#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/r2_c25_tests.log 2>&1
tail -3 gpurun_out/r2_c25_tests.log
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_FCHUNK=4 panini,BLINKY_FCHUNK=6 panini,BLINKY_RING_BYTES=12288 panini,BLINKY_RING_CTAS=10 panini,BLINKY_RING_CTAS=14 \
  panini:f1 panini:cold panini:f4 panini:f64 panini:f64,BLINKY_FCHUNK=8 \
  trism quinc equirect hammer fisheye1 panini1080 panini1080:cold stereo stereo,BLINKY_FCHUNK=4 \
  > gpurun_out/r2_c25_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c25_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    k=(d.get('kernel') or '')
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), k[k.find('grid='):][:40], k[-75:-30])
PY
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c25_bench.json 2> gpurun_out/r2_c25_bench.err
tail -c 300 gpurun_out/r2_c25_bench.err
