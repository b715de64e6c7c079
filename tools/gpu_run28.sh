#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "long_calls or two_warps or checkpoint" 2>&1 | tail -3
