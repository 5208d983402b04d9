#!/bin/bash
mkdir -p gpurun_out
( timeout 300 ./scripts/tma_lab.bin 96 0 0; timeout 300 ./scripts/tma_lab.bin 96 3 0; timeout 300 ./scripts/tma_lab.bin 96 2 0 ) > gpurun_out/r2_c4_tma_lab_dram.log 2>&1
timeout 900 python scripts/sweep_perf.py \
  panini,BLINKY_FCHUNK=8 panini,BLINKY_FCHUNK=8,BLINKY_L2_PROMOTION=1 panini,BLINKY_FCHUNK=8,BLINKY_L2_PROMOTION=2 panini,BLINKY_FCHUNK=8,BLINKY_L2_PROMOTION=3 \
  panini,BLINKY_FCHUNK=8,BLINKY_LAB=1 panini,BLINKY_FCHUNK=8,BLINKY_LAB=2 panini,BLINKY_FCHUNK=8,BLINKY_LAB=3 panini,BLINKY_FCHUNK=8,BLINKY_LAB=4 panini,BLINKY_FCHUNK=8,BLINKY_LAB=5 panini,BLINKY_FCHUNK=8,BLINKY_LAB=6 panini,BLINKY_FCHUNK=8,BLINKY_LAB=7 \
  panini:f1 panini:f1,BLINKY_LAB=2 panini:f1,BLINKY_LAB=7 panini:cold,BLINKY_L2_PROMOTION=3 \
  trism,BLINKY_FCHUNK=8,BLINKY_L2_PROMOTION=3 equirect,BLINKY_MAX_BOX=16384,BLINKY_L2_PROMOTION=3 \
  > gpurun_out/r2_c4_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c4_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[38:120])
PY
BLINKY_FCHUNK=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_ring -c 1 -s 3 -o gpurun_out/prof_r2a_ring_panini -f python scripts/sweep_perf.py panini > gpurun_out/r2_c4_ncu.log 2>&1
tail -3 gpurun_out/r2_c4_ncu.log
grep -E "48x40|64x32 |96x24|176x16|128x16" gpurun_out/r2_c4_tma_lab_dram.log | grep "mode 0" 
