#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2j_probe.log
timeout 300 python tools/probe_dyn.py --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c1-420 >> gpurun_out/r2j_probe.log
timeout 300 python tools/probe_dyn.py --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c1-420 >> gpurun_out/r2j_probe.log
timeout 300 python tools/probe_dyn.py --net tandem --tasks 2000 --chains 1024 --sweeps 4 --widths 32 2>&1 | grep resamples_per_s | cut -c1-420 >> gpurun_out/r2j_probe.log
timeout 300 python tools/probe_dyn.py --net tandem --tasks 20000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c1-420 >> gpurun_out/r2j_probe.log
cat gpurun_out/r2j_probe.log
( time timeout 900 python -m pytest tests -m gpu -q -x -k "trajectory or batch or mixed or sequential or big or gg1r" ) > gpurun_out/r2j_pytest.log 2>&1
tail -5 gpurun_out/r2j_pytest.log
