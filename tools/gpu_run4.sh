#!/bin/bash
# round-2 session 2: full GPU suite and the default bench line on the current tree
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2b_pytest.log 2>&1
tail -8 gpurun_out/r2b_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
tail -3 gpurun_out/r2b_bench_n1.err
cut -c1-600 gpurun_out/r2b_bench_n1.json
