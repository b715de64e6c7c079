#!/usr/bin/env python3
"""Timing probe for the dynamic-order sweep kernel (GPU): tandem source -> G/G/3 -> G/G/3 with LogNormal service
(BASELINE.json configs[2] shape) at a chosen size, for each group width."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayes_qnet_b200 import GibbsEngine, qnetu  # noqa: E402

TANDEM = """
states:
  - name: INITIAL
    queues: [ INITIAL ]
    successors: [ TIER1 ]
    initial: TRUE
  - name: TIER1
    queues: [ Q1 ]
    successors: [ TIER2 ]
  - name: TIER2
    queues: [ Q2 ]
queues:
  - { name: INITIAL, service: [M, 5.0] }
  - { name: Q1, type: GGk, processors: 3, service: [LN, 1.75, 1.0] }
  - { name: Q2, type: GGk, processors: 3, service: [LN, 0.1, 2.0] }
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=2000)
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--sweeps", type=int, default=5)
    ap.add_argument("--widths", default="8,16,32")
    ap.add_argument("--threads", default="128")
    ap.add_argument("--yaml", default=None)
    args = ap.parse_args()
    text = open(args.yaml).read() if args.yaml else TANDEM
    net = qnetu.qnet_from_text(text)
    qnetu.set_seed(13)
    t0 = time.time()
    smp = net.sample(args.tasks)
    sub = smp.subset_by_task(0.2)
    ini = net.gibbs_initialize(sub)
    print("input: %d events, init %.1fs" % (ini.num_events(), time.time() - t0), flush=True)
    for th in args.threads.split(","):
        for w in args.widths.split(","):
            os.environ["BQ_DYN_W"] = w
            os.environ["BQ_DYN_THREADS"] = th
            with GibbsEngine(net, ini, n_chains=args.chains, seed=7) as eng:
                n = len(eng.schedule()[1])
                eng.sweep(1, "SLICE", "BAYES")
                eng.sweep(args.sweeps, "SLICE", "BAYES")
                ms = eng.last_sweep_ms()
                st = eng.stats()
                print(json.dumps({"W": int(w), "threads": int(th), "chains": args.chains, "events": eng.E, "eligible": n,
                                  "ms": ms, "resamples_per_s": args.chains * args.sweeps * n / (ms / 1e3),
                                  "evals_per_resample": st["slice_evals"] / st["nevents"],
                                  "walk_steps_per_resample": st["diff_len"] / st["nevents"],
                                  "stuck": st["n_stuck"]}), flush=True)


if __name__ == "__main__":
    main()
