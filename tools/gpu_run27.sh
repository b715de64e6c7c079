#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r2F_pytest.log 2>&1
tail -4 gpurun_out/r2F_pytest.log
