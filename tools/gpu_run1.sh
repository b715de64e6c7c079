#!/bin/bash
# GPU session 1 of round 2: full GPU test suite, headline bench, batch-order stride sweep on configs 3 and 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest1.log 2>&1
tail -5 gpurun_out/r2_pytest1.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_c2_a.json 2> gpurun_out/r2_bench_c2_a.err
cat gpurun_out/r2_bench_c2_a.json | cut -c1-400
for s in 0 1 2 4 8 16; do
  BQ_ORDER_STRIDE=$s timeout 300 python tools/probe_dyn.py --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | sed "s/^/stride=$s /" >> gpurun_out/r2_stride_c3.log
done
cat gpurun_out/r2_stride_c3.log | cut -c1-330
for s in 0 1 2 4 8 16; do
  BQ_ORDER_STRIDE=$s timeout 300 python tools/probe_dyn.py --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | sed "s/^/stride=$s /" >> gpurun_out/r2_stride_c4.log
done
cat gpurun_out/r2_stride_c4.log | cut -c1-330
