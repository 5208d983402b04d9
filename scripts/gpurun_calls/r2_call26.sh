#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_c26_bench_n2.json 2> gpurun_out/r2_c26_bench_n2.err
tail -c 2500 gpurun_out/r2_c26_bench_n2.json; tail -5 gpurun_out/r2_c26_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_c26_bench_ref_n2.json 2>> gpurun_out/r2_c26_bench_n2.err
cat gpurun_out/r2_c26_bench_ref_n2.json | cut -c1-300
( timeout 600 python -m pytest tests/test_sharding.py -x -q -m gpu 2>&1 | tail -3 )
