#!/bin/bash
# locality experiment: windowed batch order (BQ_ORDER_STRIDE) with FEWER chains resident, so the windows fit in L2
mkdir -p gpurun_out
rm -f gpurun_out/r2g_probe.log
for chains in 256 512 1024 2048; do
  for s in 0 16 32; do
    BQ_ORDER_STRIDE=$s timeout 300 python tools/probe_dyn.py --net tandem --tasks 20000 --chains $chains --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c60-330 | sed "s/^/stride=$s /" >> gpurun_out/r2g_probe.log
  done
done
cat gpurun_out/r2g_probe.log
