#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_boundary.py -x -q -k list_of_arrivals > gpurun_out/r2_sanitizer.log 2>&1
grep -n "Invalid\|at 0x\|by thread\|Address\|ERROR SUMMARY\|in /" gpurun_out/r2_sanitizer.log | head -30
timeout 300 python tools/probe_rhat.py 5000 256 500 16 1 > gpurun_out/r2_rhat_5k_disp.log 2>&1
timeout 300 python tools/probe_rhat.py 5000 256 500 16 0 > gpurun_out/r2_rhat_5k_nodisp.log 2>&1
timeout 300 python tools/probe_rhat.py 500 256 500 8 1 > gpurun_out/r2_rhat_500_disp.log 2>&1
tail -2 gpurun_out/r2_rhat_5k_disp.log | cut -c1-900
tail -1 gpurun_out/r2_rhat_5k_nodisp.log | cut -c1-900
tail -1 gpurun_out/r2_rhat_500_disp.log | cut -c1-900
