#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | cut -c1-160
