#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 ) > gpurun_out/r2B_bench_n2.json 2> gpurun_out/r2B_bench_n2.err
tail -3 gpurun_out/r2B_bench_n2.err
grep -o '"value": [0-9.e+]*' gpurun_out/r2B_bench_n2.json | head -3
