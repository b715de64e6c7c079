#!/bin/bash
# final-tree validation: GPU suite, default bench line, DRAM traffic of the config-3 / config-4 launches
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2m_pytest.log 2>&1
tail -6 gpurun_out/r2m_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
tail -3 gpurun_out/r2m_bench_n1.err
cut -c1-300 gpurun_out/r2m_bench_n1.json
for w in config3 config4; do
  CMD="python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
  timeout 600 $CMD > gpurun_out/r2m_$w.json 2> gpurun_out/r2m_$w.err
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sweep_batch -c 2 --csv --log-file gpurun_out/r2m_${w}_dram_traffic.csv $CMD > /dev/null 2>&1
  tail -7 gpurun_out/r2m_${w}_dram_traffic.csv | cut -c1-250
done
