#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2n_probe.log
timeout 300 python tools/probe_dyn.py --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 >> gpurun_out/r2n_probe.log
timeout 300 python tools/probe_dyn.py --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 >> gpurun_out/r2n_probe.log
timeout 300 python tools/probe_dyn.py --net tandem --tasks 2000 --chains 1024 --sweeps 4 --widths 32 2>&1 | grep resamples_per_s | cut -c60-420 >> gpurun_out/r2n_probe.log
cut -c1-140,200-330 gpurun_out/r2n_probe.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2n_pytest.log 2>&1
tail -6 gpurun_out/r2n_pytest.log
