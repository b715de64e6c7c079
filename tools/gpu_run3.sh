#!/bin/bash
# round-2 profiling pass: launch lists and ncu --set full of the two sweep kernels (one plain run right before each)
mkdir -p gpurun_out
CMD="python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 1"
$CMD > gpurun_out/r2_plain_static.log 2> gpurun_out/r2_plain_static.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_static.csv $CMD > /dev/null 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_static -s 2 -c 1 -o gpurun_out/r2_static_prof -f $CMD > gpurun_out/r2_ncu_static.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r2_static_prof.ncu-rep > gpurun_out/r2_static_ncu_full.txt 2>&1
CMD3="python bench.py --workload config3 --no-extras --no-cpu-baseline --steps 1 --warmup 1"
$CMD3 > gpurun_out/r2_plain_c3.log 2> gpurun_out/r2_plain_c3.err && \
ncu --set full --clock-control none --import-source on -k regex:sweep_batch -s 1 -c 1 -o gpurun_out/r2_batch_c3_prof -f $CMD3 > gpurun_out/r2_ncu_c3.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r2_batch_c3_prof.ncu-rep > gpurun_out/r2_batch_c3_ncu_full.txt 2>&1
ls -la gpurun_out/*.ncu-rep
head -30 gpurun_out/r2_static_ncu_full.txt
head -34 gpurun_out/r2_batch_c3_ncu_full.txt
cut -c1-300 gpurun_out/r2_plain_static.log
