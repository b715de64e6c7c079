#!/bin/bash
# GG1R on the GPU: its tests first, then the whole GPU suite
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "gg1r or mixed or gg1rm" ) > gpurun_out/r2c_gg1r.log 2>&1
tail -15 gpurun_out/r2c_gg1r.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2c_pytest.log 2>&1
tail -8 gpurun_out/r2c_pytest.log
for net in tandem ps; do
  timeout 300 python tools/probe_dyn.py --net $net --tasks $([ $net = tandem ] && echo 20000 || echo 50000) --chains $([ $net = tandem ] && echo 4096 || echo 2048) --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c1-300 >> gpurun_out/r2c_probe.log
done
cat gpurun_out/r2c_probe.log
