#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/probe_ess_order.py > gpurun_out/r2E_ess_order.jsonl 2> gpurun_out/r2E_ess_order.err
cat gpurun_out/r2E_ess_order.jsonl | cut -c1-600; tail -3 gpurun_out/r2E_ess_order.err
