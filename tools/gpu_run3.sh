#!/bin/bash
# round-2 profiling pass (static kernel): launch list and ncu --set full, one plain run right before each; the .ncu-rep
# stays on the box (copy-back limit), its text summary comes back
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 1"
$CMD > gpurun_out/r2_plain_static.log 2> gpurun_out/r2_plain_static.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_static.csv $CMD > /dev/null 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_static -s 2 -c 1 -o /tmp/r2_static_prof -f $CMD > gpurun_out/r2_ncu_static.log 2>&1
python profiles/summarize_ncu.py /tmp/r2_static_prof.ncu-rep > gpurun_out/r2_static_ncu_full.txt 2>&1
tail -45 gpurun_out/r2_static_ncu_full.txt
