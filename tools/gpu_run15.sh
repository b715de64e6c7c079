#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2s_nw.log
for nw in 1 2; do
  BQ_BATCH_NW=$nw timeout 300 python tools/probe_dyn.py --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/nw=$nw ps /" >> gpurun_out/r2s_nw.log
  BQ_BATCH_NW=$nw timeout 300 python tools/probe_dyn.py --net ps --tasks 50000 --chains 1024 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/nw=$nw ps /" >> gpurun_out/r2s_nw.log
done
timeout 300 python tools/probe_dyn.py --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/auto tandem /" >> gpurun_out/r2s_nw.log
timeout 300 python tools/probe_dyn.py --net tandem --tasks 20000 --chains 1024 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/auto tandem /" >> gpurun_out/r2s_nw.log
timeout 300 python tools/probe_dyn.py --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 | sed "s/^/auto ps /" >> gpurun_out/r2s_nw.log
cut -c1-40,90-140,230-330 gpurun_out/r2s_nw.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2s_pytest.log 2>&1
tail -6 gpurun_out/r2s_pytest.log
