#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_init.py -q -x 2>&1 | tail -12
