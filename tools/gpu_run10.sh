#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2l_scan.log
for s in 64 96 128 192; do
  BQ_SUPPORT_SCAN=$s timeout 300 python tools/probe_dyn.py --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/scan=$s /" >> gpurun_out/r2l_scan.log
done
for s in 24 64 128; do
  BQ_SUPPORT_SCAN=$s timeout 300 python tools/probe_dyn.py --net tandem --tasks 2000 --chains 1024 --sweeps 4 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/small scan=$s /" >> gpurun_out/r2l_scan.log
done
cut -c1-140,200-330 gpurun_out/r2l_scan.log
