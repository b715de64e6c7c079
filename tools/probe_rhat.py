#!/usr/bin/env python3
"""How fast does the ensemble forget its starts?  Split R-hat per summary over blocks of sweeps, config-2 network."""
import json
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bayes_qnet_b200 import GibbsEngine, parallel  # noqa: E402

tasks = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
block = int(sys.argv[3]) if len(sys.argv) > 3 else 500
nblocks = int(sys.argv[4]) if len(sys.argv) > 4 else 20
disperse = int(sys.argv[5]) if len(sys.argv) > 5 else 1
wl = bench.WORKLOADS["config2"]
net, ini, how = bench.make_input(tasks, yaml_text=wl["yaml"], init=wl["init"], fixture=wl["fixture"] if tasks == 5000 else None)
eng = GibbsEngine(net, ini, n_chains=C, seed=777)
if disperse == 2:
    eng.init_chains(0.7)
elif disperse:
    eng.disperse(sweeps=3, scale=0.7, seed=1)
for b in range(nblocks):
    eng.clear_traces()
    eng.sweep(block, "SLICE", "BAYES", trace=True)
    summ = parallel.gather_summaries_device(eng, 1)
    d = parallel.device_diagnostics(summ)
    x = summ.cpu().numpy()
    print(json.dumps({"block": b, "sweeps": (b + 1) * block, "rhat": [round(float(v), 3) for v in d["rhat"]],
                      "ess_per_chain_mean": [round(float(v) / C, 1) for v in d["ess_per_chain_sum"]],
                      "mean": [round(float(v), 4) for v in x.mean(axis=(0, 1))],
                      "between_sd": [round(float(v), 4) for v in x.mean(axis=0).std(axis=0)],
                      "within_sd": [round(float(v), 4) for v in x.std(axis=0).mean(axis=0)]}), flush=True)
