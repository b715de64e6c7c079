#!/bin/bash
# ncu --set full of the batch kernel on the PS + delay network (config 4's share of one GPU)
mkdir -p gpurun_out
CMD="python tools/probe_dyn.py --net ps --tasks 50000 --chains 2048 --sweeps 1 --widths 32"
$CMD > gpurun_out/r2p_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_batch -c 1 -o /tmp/r2p_batch -f $CMD > gpurun_out/r2p_ncu.log 2>&1
python profiles/summarize_ncu.py /tmp/r2p_batch.ncu-rep 45 > gpurun_out/r2p_batch_ps_ncu_full.txt 2>&1
head -100 gpurun_out/r2p_batch_ps_ncu_full.txt | cut -c1-190
