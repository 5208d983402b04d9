#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_c29_bench_n8.json 2> gpurun_out/r2_c29_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c29_bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','value_warp_only')}, d['e2e']['value'])
print(json.dumps(d['gather'])[:1500])
PY
tail -3 gpurun_out/r2_c29_bench_n8.err
