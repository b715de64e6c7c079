#!/bin/bash
# eight GPUs: config 4 as BASELINE states it (16 384 chains over 8 GPUs), then the headline line without the sub-runs
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
( time timeout 600 $T bench.py --gpus 8 --workload config4 --no-cpu-baseline --no-extras --steps 3 --warmup 3 ) > gpurun_out/r2C_config4_n8.json 2> gpurun_out/r2C_config4_n8.err
grep -o '"value": [0-9.e+]*' gpurun_out/r2C_config4_n8.json | head -2
( time timeout 600 $T bench.py --gpus 8 --no-cpu-baseline --no-extras ) > gpurun_out/r2C_bench_n8.json 2> gpurun_out/r2C_bench_n8.err
grep -o '"value": [0-9.e+]*' gpurun_out/r2C_bench_n8.json | head -2
tail -3 gpurun_out/r2C_bench_n8.err
