#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x -k "two_warps or gg1r or trajectory or big" 2>&1 | tail -2
python tools/probe_dyn.py --net tandem --tasks 20000 --chains 4096 --sweeps 2 --widths 32 2>&1 | grep resamples_per_s | cut -c150-230
