#!/bin/bash
# ncu --set full of one wave of the two-warp batch kernel (config 3 network at full size), then the N = 2 bench line
mkdir -p gpurun_out
CMD="python tools/probe_dyn.py --net tandem --tasks 20000 --chains 4096 --sweeps 1 --widths 32"
$CMD > gpurun_out/r2A_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_batch -c 1 -o /tmp/r2A_batch -f $CMD > gpurun_out/r2A_ncu.log 2>&1
python profiles/summarize_ncu.py /tmp/r2A_batch.ncu-rep 40 > gpurun_out/r2A_batch_nw_ncu_full.txt 2>&1
head -34 gpurun_out/r2A_batch_nw_ncu_full.txt | cut -c1-160
