#!/bin/bash
# final measurements of the round: default bench line, DRAM traffic of the config-3 / config-4 launches, launch list
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err
tail -3 gpurun_out/r2o_bench_n1.err
cut -c1-200 gpurun_out/r2o_bench_n1.json
for w in config3 config4; do
  CMD="python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
  timeout 600 $CMD > gpurun_out/r2o_$w.json 2> gpurun_out/r2o_$w.err
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sweep_batch -c 2 --csv --log-file gpurun_out/r2o_${w}_dram_traffic.csv $CMD > /dev/null 2>&1
  tail -3 gpurun_out/r2o_${w}_dram_traffic.csv | cut -c150-330
done
CMD="python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 1"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2o_launches_static.csv $CMD > /dev/null 2>&1
tail -3 gpurun_out/r2o_launches_static.csv | cut -c1-250
