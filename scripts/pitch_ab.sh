#!/bin/bash
# A/B of the box pitch rule (BLINKY_BOX_PITCH=even|odd) on the warp kernel; run under gpurun
for mode in even odd even odd; do
  for cfg in '--lens panini --zoom "f_fov 180" --threads 0' '--lens stereographic --zoom "f_fov 180" --threads 0' '--lens quincuncial --zoom f_cover --rubix --threads 0' '--globe trism --lens stereographic --zoom "f_fov 180" --threads 0' '--w 1920 --h 1080 --ps 1024 --lens panini --zoom "f_fov 170" --threads 0'; do
    BLINKY_BOX_PITCH=$mode eval python scripts/quick_perf.py $cfg 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$mode', d['lens'], d['globe'], d['w'], 'us/frame', d['us_per_frame'], 'frac', d['frac_of_6485'])"
  done
done
