#!/usr/bin/env python3
"""Where the end-to-end step of bench.py spends its host-side time (config 2): per-call wall times."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import bench
from bayes_qnet_b200 import GibbsEngine, parallel

ctx = None
g = numpy.load(os.path.join(bench.ROOT, "tests", "golden", "big_qnet3_5k.npz"), allow_pickle=False)
sys.path.insert(0, os.path.join(bench.ROOT, "tests"))
import helpers
net, ini = helpers.big_state(g, "s0")
C, S = 1024, 20
eng = GibbsEngine(net, ini, n_chains=C, seed=1)
d0 = ini.soa()["d"].copy(); theta0 = numpy.array(net.parameters)
T = {}
def tick(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); T.setdefault(name, []).append(1e3 * (time.perf_counter() - t0)); return r
for it in range(6):
    tick("broadcast_state", lambda: eng.broadcast_state(d0))
    tick("set_params", lambda: eng.set_params(theta0))
    tick("clear_traces", lambda: eng.clear_traces())
    tick("sweep", lambda: eng.sweep(S, "SLICE", "BAYES", trace=True, sync=True))
    summ = tick("gather", lambda: parallel.gather_summaries_device(eng, 1))
    tick("diagnostics", lambda: parallel.device_diagnostics(summ))
    tick("traces_d2h", lambda: eng.traces())
    tick("params_d2h", lambda: eng.params())
print(json.dumps({k: round(sum(v[1:]) / len(v[1:]), 3) for k, v in T.items()}))
