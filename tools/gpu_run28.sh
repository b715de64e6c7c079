#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2G_pf.log
P="python tools/probe_dyn.py"
timeout 300 $P --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/prefetch tandem /" >> gpurun_out/r2G_pf.log
timeout 300 $P --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/prefetch ps /" >> gpurun_out/r2G_pf.log
timeout 300 $P --net tandem --tasks 20000 --chains 1024 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/prefetch tandem /" >> gpurun_out/r2G_pf.log
cut -c1-50,90-140,300-330 gpurun_out/r2G_pf.log
