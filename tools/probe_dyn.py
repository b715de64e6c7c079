#!/usr/bin/env python3
"""Timing probe for the dynamic-order sweep kernels (GPU): tandem source -> G/G/3 -> G/G/3 with LogNormal service
(BASELINE.json configs[2] shape) or source -> PS -> delay station (configs[3] shape) at a chosen size, slice or MH
sweeps, for each group width."""
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


PSDELAY = """
states:
  - name: INITIAL
    queues: [ INITIAL ]
    successors: [ TIER1 ]
    initial: TRUE
  - name: TIER1
    queues: [ QPS ]
    successors: [ TIER2 ]
  - name: TIER2
    queues: [ QDELAY ]
queues:
  - { name: INITIAL, service: [M, 1.0] }
  - { name: QPS, type: PS, service: [M, 0.5] }
  - { name: QDELAY, type: DELAY, service: [M, 2.0] }
"""
NETS = {"tandem": (TANDEM, "DFN2", 32.0), "ps": (PSDELAY, "PS", 80.0 / 3.0)}  # yaml, initializer, SURVEY 8d bytes / resample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=2000)
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--sweeps", type=int, default=5)
    ap.add_argument("--widths", default="8,16,32")
    ap.add_argument("--threads", default="128")
    ap.add_argument("--net", default="tandem", choices=sorted(NETS) + ["qnet3"])
    ap.add_argument("--mode", default="SLICE", choices=["SLICE", "MH"])
    args = ap.parse_args()
    from bayes_qnet_b200 import qnet
    if args.net == "qnet3":
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        import helpers
        text, init, nbytes = helpers.QNET3, "DFN2", 16.0
    else:
        text, init, nbytes = NETS[args.net]
    qnet.set_initialization_type(init)
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
                n = len(eng.sequential_order())
                impl = eng.kernel_info()
                eng.sweep(1, args.mode, "BAYES")
                eng.reset_stats()
                eng.sweep(args.sweeps, args.mode, "BAYES")
                ms = eng.last_sweep_ms()
                st = eng.stats()
                rate = st["nevents"] / (ms / 1e3)
                print(json.dumps({"net": args.net, "mode": args.mode, "W": int(w), "threads": int(th), "chains": args.chains,
                                  "events": eng.E, "eligible": n, "ms": ms, "resamples_per_s": rate,
                                  "algorithmic_GBps": rate * nbytes / 1e9, "reject_frac": st["nreject"] / max(st["nevents"], 1),
                                  "evals_per_resample": st["slice_evals"] / st["nevents"],
                                  "walk_steps_per_resample": st["diff_len"] / st["nevents"],
                                  "redone_frac": st["n_redone"] / max(st["nevents"], 1),
                                  "stuck": st["n_stuck"]}), flush=True)


if __name__ == "__main__":
    main()
