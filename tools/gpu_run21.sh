#!/bin/bash
# final tree: GPU suite, smoke, N = 1 bench line, then N = 2
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2z_pytest.log 2>&1
tail -4 gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py ) > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
tail -3 gpurun_out/r2z_bench_n1.err
cut -c1-160 gpurun_out/r2z_bench_n1.json
