#!/usr/bin/env python3
"""Timing of the cluster path: (a) few chains of the config-2 network, cluster of 1/2/4/8 CTAs per chain; (b) a chain too
large for one CTA's shared memory (E = 40 000): cluster of 2 on chip vs the global-memory build."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bayes_qnet_b200 import GibbsEngine  # noqa: E402

wl = bench.WORKLOADS["config2"]


def timeit(net, ini, chains, ncta, sweeps=20):
    if ncta is None:
        os.environ.pop("BQ_CLUSTER", None)
    else:
        os.environ["BQ_CLUSTER"] = str(ncta)
    with GibbsEngine(net, ini, n_chains=chains, seed=3) as eng:
        n = len(eng.schedule()[1])
        eng.sweep(sweeps, "SLICE", "BAYES")
        best = 1e9
        for _ in range(3):
            eng.sweep(sweeps, "SLICE", "BAYES")
            best = min(best, eng.last_sweep_ms())
        return dict(chains=chains, cluster=eng.kernel_info()["cluster"], state_in_smem=eng.kernel_info()["state_in_smem"],
                    ms=best, resamples_per_s=chains * sweeps * n / (best / 1e3), us_per_sweep=1e3 * best / sweeps)


net, ini, _ = bench.make_input(5000, yaml_text=wl["yaml"], init=wl["init"], fixture=wl["fixture"])
for chains in (1, 4, 16, 32, 64):
    for ncta in (1, 2, 4, 8):
        if chains * ncta > 148:
            continue
        print(json.dumps(dict(case="config2 network, E=20000", **timeit(net, ini, chains, ncta))), flush=True)
net1, ini1, _ = bench.make_input(200, yaml_text=bench.WORKLOADS["config1"]["yaml"], init="DFN2", pct=0.5)
for ncta in (1, 2, 8):
    print(json.dumps(dict(case="config1 M/M/1 200 tasks, 1 chain x 1000 sweeps", **timeit(net1, ini1, 1, ncta, sweeps=1000))), flush=True)
big, inib, _ = bench.make_input(10000, yaml_text=wl["yaml"], init=wl["init"], fixture=None)
for chains in (148, 592):
    for ncta in (1, 2, 4):
        print(json.dumps(dict(case="3-tier network, 10000 tasks, E=40000", **timeit(big, inib, chains, ncta))), flush=True)
