#!/bin/bash
# ncu --set full of the batch kernel (tandem at full size) with the source page: where the stall samples are
mkdir -p gpurun_out
CMD="python tools/probe_dyn.py --net tandem --tasks 20000 --chains 4096 --sweeps 1 --widths 32"
$CMD > gpurun_out/r2i_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_batch -c 1 -o /tmp/r2i_batch -f $CMD > gpurun_out/r2i_ncu.log 2>&1
python profiles/summarize_ncu.py /tmp/r2i_batch.ncu-rep 60 > gpurun_out/r2i_batch_ncu_full.txt 2>&1
tail -135 gpurun_out/r2i_batch_ncu_full.txt
