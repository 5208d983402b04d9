#!/bin/bash
# Developer helper (run under gpurun): GPU tests + quick perf table, everything logged to gpurun_out/check.log
mkdir -p gpurun_out
{
  if [ "$1" != "perf" ]; then python -m pytest tests -x -q -m gpu 2>&1 | tail -15; fi
  while read -r cfg; do
    [ -z "$cfg" ] && continue
    eval timeout 300 python scripts/quick_perf.py $cfg > gpurun_out/_qp.log 2>&1
    tail -1 gpurun_out/_qp.log | python -c "
import sys, json
line = sys.stdin.read()
try:
    d = json.loads(line)
    print({k: d[k] for k in ('lens', 'globe', 'w', 'frames', 'cold', 'rubix', 'us_per_frame', 'mpix_s', 'frac_of_6485', 'plan')})
except Exception:
    print('FAILED:', open('gpurun_out/_qp.log').read()[-1500:])
"
  done <<'CFG'
--lens panini --zoom "f_fov 180"
--lens panini --zoom "f_fov 180" --cold
--lens stereographic --zoom "f_fov 180"
--lens quincuncial --zoom f_cover --rubix
--w 1920 --h 1080 --ps 1024 --lens panini --zoom "f_fov 170"
--lens fisheye1 --zoom f_contain
--lens equirect --zoom f_contain
--lens hammer --zoom f_contain
--globe trism --lens stereographic --zoom "f_fov 180"
CFG
} > gpurun_out/check.log 2>&1
tail -40 gpurun_out/check.log
