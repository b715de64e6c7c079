#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gg1r.py -q -x 2>&1 | tail -15
python tools/probe_dyn.py --net ps --tasks 50000 --chains 2048 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c150-230
