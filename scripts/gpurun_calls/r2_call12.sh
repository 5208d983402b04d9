#!/bin/bash
# latency-bound or throughput-bound?  ring depth / warps-per-SM sweep on the headline workload
mkdir -p gpurun_out
timeout 900 python scripts/sweep_perf.py \
  panini panini,BLINKY_RING_STAGES=3 panini,BLINKY_RING_STAGES=4 panini,BLINKY_RING_CTAS=6 panini,BLINKY_RING_CTAS=8 panini,BLINKY_RING_CTAS=10 \
  panini,BLINKY_MAX_BOX=4096 panini,BLINKY_MAX_BOX=4096,BLINKY_RING_STAGES=3 panini,BLINKY_MAX_BOX=4096,BLINKY_RING_STAGES=4 panini,BLINKY_MAX_BOX=3072,BLINKY_RING_STAGES=4 \
  panini,BLINKY_FCHUNK=16 panini,BLINKY_FCHUNK=4 panini:f64,BLINKY_FCHUNK=32 panini,BLINKY_L2_PROMOTION=2 panini,BLINKY_L2_PROMOTION=1 panini,BLINKY_STATIC_PCT=100 panini,BLINKY_STATIC_PCT=0 \
  > gpurun_out/r2_c12_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c12_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[38:120], (d.get('plan') or '')[30:100])
PY
