#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2y_minb.log
P="python tools/probe_dyn.py"
for mb in 3 4; do
  BQ_NW_MINB=$mb timeout 300 $P --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/minb=$mb tandem /" >> gpurun_out/r2y_minb.log
  BQ_NW_MINB=$mb timeout 300 $P --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/minb=$mb ps /" >> gpurun_out/r2y_minb.log
done
BQ_NW_MINB=3 timeout 300 $P --net tandem --tasks 20000 --chains 888 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/minb=3 tandem /" >> gpurun_out/r2y_minb.log
cut -c1-50,90-140,300-330 gpurun_out/r2y_minb.log
