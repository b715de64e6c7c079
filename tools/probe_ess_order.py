#!/usr/bin/env python3
"""ESS per sweep and chain of the chromatic scan order (this package) against the reference's dict order, same
algorithm, same initial state: the reference's long runs in tests/golden/post_<net>.npz (oracle/make_golden.py) vs the
GPU engine started from the state stored with them.  Per summary (service parameters, then per-queue mean waits)."""
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from bayes_qnet_b200 import GibbsEngine, diagnostics  # noqa: E402


def ess_per_sweep(x):
    """x: [n, chains, K] -> mean over chains of the chain's own ESS / n, per summary."""
    n = x.shape[0]
    out = []
    for k in range(x.shape[2]):
        out.append(float(numpy.mean([diagnostics.ess(x[:, c:c + 1, k:k + 1])[0] for c in range(x.shape[1])]) / n))
    return out


for name in sys.argv[1:] or ["mm1", "qnet3", "mm3", "psdelay", "gg1rm"]:
    g = helpers.golden("post_" + name)
    net, arrv = helpers.golden_state(dict((k, g[k]) for k in g.files if k != "theta") | {"theta": g["theta0"]}, "s0")
    ref = numpy.concatenate((numpy.swapaxes(g["theta"], 0, 1), numpy.swapaxes(g["mean_wait"], 0, 1)[:, :, 1:]), axis=2)
    n = ref.shape[0]
    burn = n // 5
    with GibbsEngine(net, arrv, n_chains=64, seed=5) as eng:
        eng.sweep(n, "SLICE", "BAYES", trace=True)
        th, mw, lp = eng.traces()
    ours = numpy.concatenate((th, mw[:, :, 1:]), axis=2)
    a, b = ess_per_sweep(ours[burn:]), ess_per_sweep(ref[burn:])
    print(json.dumps({"net": name, "sweeps": n - burn, "chains_ours": 64, "replicas_ref": int(ref.shape[1]),
                      "ess_per_sweep_ours": [round(v, 4) for v in a], "ess_per_sweep_reference": [round(v, 4) for v in b],
                      "ratio_ours_over_reference": [round(x / max(y, 1e-12), 2) for x, y in zip(a, b)]}))
