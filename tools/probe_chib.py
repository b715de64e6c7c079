#!/usr/bin/env python3
"""Timing probe for bq_gibbs_kernel (GPU): the config-2 network, every chain's transition density into one target."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bayes_qnet_b200 import GibbsEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=5000)
    ap.add_argument("--chains", type=int, default=1024)
    args = ap.parse_args()
    net, ini, _ = bench.make_input(args.tasks)
    with GibbsEngine(net, ini, n_chains=args.chains, seed=3) as eng:
        eng.sweep(20, "SLICE", "BAYES")
        d_star = eng.state(0, 1)[0].copy()
        eng.sweep(1, "SLICE", "NONE")
        n = len(eng.schedule()[1])
        for strict in (False, True):
            eng.gibbs_kernel(d_star, strict=strict)
            t0 = time.time()
            kp = eng.gibbs_kernel(d_star, strict=strict)
            dt = time.time() - t0
            import numpy
            print(json.dumps({"tasks": args.tasks, "chains": args.chains, "events_per_chain": n, "strict": strict,
                              "seconds": dt, "event_normalisers_per_s": n * args.chains / dt,
                              "finite_frac": float(numpy.mean(numpy.isfinite(kp)))}), flush=True)


if __name__ == "__main__":
    main()
