#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2t_wave.log
P="python tools/probe_dyn.py"
timeout 300 $P --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/auto tandem /" >> gpurun_out/r2t_wave.log
timeout 300 $P --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/auto ps /" >> gpurun_out/r2t_wave.log
BQ_BATCH_WAVE=768 timeout 300 $P --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/wave768 tandem /" >> gpurun_out/r2t_wave.log
BQ_BATCH_WAVE=1400 timeout 300 $P --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/wave1400 tandem /" >> gpurun_out/r2t_wave.log
BQ_BATCH_WAVE=700 timeout 300 $P --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/wave700 ps /" >> gpurun_out/r2t_wave.log
BQ_BATCH_NW=1 timeout 300 $P --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/nw1 tandem /" >> gpurun_out/r2t_wave.log
cut -c1-50,90-140,230-330 gpurun_out/r2t_wave.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2t_pytest.log 2>&1
tail -6 gpurun_out/r2t_pytest.log
