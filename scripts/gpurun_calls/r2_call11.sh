#!/bin/bash
# re-entry baseline: GPU suite, bench (both arms), launch list, ncu --set full of every workload, knob sweep
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/r2_c11_tests.log 2>&1
tail -3 gpurun_out/r2_c11_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c11_bench.json 2> gpurun_out/r2_c11_bench.err
tail -c 600 gpurun_out/r2_c11_bench.json; tail -3 gpurun_out/r2_c11_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_c11_bench_ref.json 2>> gpurun_out/r2_c11_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c11_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r2_c11_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_c11_all \
  python scripts/ncu_workloads.py --frames 16 > gpurun_out/r2_c11_ncu.log 2>&1
cp gpurun_out/ncu_workloads_order.json gpurun_out/r2_c11_order16.json
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2_c11_one \
  python scripts/ncu_workloads.py --frames 1 --order gpurun_out/r2_c11_order1.json 4k-cube-panini 4k-cube-quincuncial-rubix 4k-cube-fisheye1 >> gpurun_out/r2_c11_ncu.log 2>&1
tail -5 gpurun_out/r2_c11_ncu.log
timeout 900 python scripts/sweep_perf.py \
  panini panini:f1 panini:cold panini:f4 panini:f64 panini,BLINKY_RING_WARPS=14 panini,BLINKY_RING_WARPS=16 \
  trism quinc quinc,BLINKY_SPLIT_PERCENT=100 equirect hammer fisheye1 panini1080 panini1080:cold stereo \
  > gpurun_out/r2_c11_sweep.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_c11_sweep.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d.get('work'), d.get('env'), d.get('frames'), 'cold' if d.get('cold') else '', d.get('us_per_frame'), d.get('min_us'), d.get('error',''), (d.get('kernel') or '')[38:110])
PY
