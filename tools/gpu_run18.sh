#!/bin/bash
# the round's bench lines: N = 1 (default command), then N = 2
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err
tail -3 gpurun_out/r2v_bench_n1.err
cut -c1-200 gpurun_out/r2v_bench_n1.json
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err
cut -c1-400 gpurun_out/r2v_bench_ref.json
