#!/usr/bin/env python3
"""Offline measurement of the REFERENCE at the real size of each BASELINE config ("run a fixed
small number of sweeps at the real size ... do not extrapolate from smaller inputs"), on the build container's CPU,
one pinned single-chain process per figure, from the committed reference-generated states (tests/golden/big_*.npz)
through the reference's own table reader.  Far too slow for a bench.py run (one sweep of config 3 takes the
reference the better part of an hour), so the figures are committed: profiles/r2_reference_fullsize.json, which
bench.py quotes beside its bounded on-box samples.  TEST / MEASUREMENT INFRASTRUCTURE ONLY (executes oracle/_ref).

  python tools/ref_fullsize.py config2 | config3 | config4 | config2_ess [--iters N]
"""
import argparse
import json
import multiprocessing
import os
import platform
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_reference_fullsize.json")
FIX = {"config2": "qnet3_5k", "config3": "lnk_20k", "config4": "psdelay_50k"}

CODE = r"""
import sys, io, os, contextlib, time, json
os.environ.setdefault("BQ_REF_BIG_N", "65536")
sys.path.insert(0, %(ref)r)
import numpy
with contextlib.redirect_stdout(io.StringIO()):
    import qnetu, qnet, sampling, estimation, qstats
    sampling.DEBUG = 0
    estimation.set_sampler(%(sampler)r)
    g = numpy.load(%(fixture)r, allow_pickle=False)
    d, tid = g["s0_d"], g["tid"]
    a = numpy.zeros_like(d); same = tid[1:] == tid[:-1]; a[1:][same] = d[:-1][same]
    lines = ["eid tid qid a d s w state proc obs_a obs_d"]
    for i in range(len(d)):
        lines.append("%%d %%d %%d %%r %%r 0.0 0.0 %%d 0 %%d %%d" %% (g["eid"][i], tid[i], g["qid"][i], float(a[i]), float(d[i]),
                                                               g["state"][i], g["obs_a"][i], g["obs_d"][i]))
    net = qnetu.qnet_from_text(str(g["yaml"]))
    net.parameters = list(g["theta"])
    t0 = time.time()
    ini = qnetu.read_csv_from_lines(net, lines)
    t_load = time.time() - t0
    qnetu.set_seed(%(seed)d)
    qnet.reset_stats()
    waits = []
    t0 = time.time()
    if %(sampler)r == "MH":
        net.estimate(ini)
        arrv, mus = ini, []
        for i in range(%(iters)d):
            arrv = net.gibbs_resample(arrv, burn=0, num_iter=1, return_arrv=False)[-1]
            net.sample_parameters(arrv)
            mus.append(net.parameters[:])
    else:
        mus, _ = estimation.bayes(net, ini, %(iters)d, report_fn=(lambda n, a, i, lp: waits.append(list(qstats.mean_wait(a)))) if %(trace)d else None)
    dt = time.time() - t0
    st = qnet.stats()
print("RESULT " + json.dumps({"nevents": st["nevents"], "nreject": st["nreject"], "seconds": dt, "load_seconds": t_load,
                              "mus": mus if %(trace)d else [], "waits": waits}))
"""


def run_one(args):
    fixture, sampler, iters, seed, trace, core = args
    code = CODE % dict(ref=os.path.join(ROOT, "oracle", "_ref"), fixture=fixture, sampler=sampler, iters=iters, seed=seed,
                       trace=1 if trace else 0)
    cmd = [sys.executable, "-c", code]
    if os.path.exists("/usr/bin/taskset"):
        cmd = ["taskset", "-c", str(core)] + cmd
    out = subprocess.run(cmd, capture_output=True, text=True)
    for line in out.stdout.splitlines():
        if line.startswith("RESULT "):
            return json.loads(line[7:])
    raise RuntimeError(out.stderr[-2000:])


def cpu_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return platform.processor()


def update(key, rec):
    data = {}
    if os.path.exists(OUT):
        data = json.load(open(OUT))
    data[key] = rec
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    print(key, json.dumps(rec)[:600])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["config2", "config3", "config4", "config2_ess"])
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--core", type=int, default=0)
    a = ap.parse_args()
    where = "build container, %s, 1 pinned process of %d vCPUs" % (cpu_name(), os.cpu_count() or 1)
    if a.what == "config2_ess":
        import numpy
        sys.path.insert(0, ROOT)
        from bayes_qnet_b200 import diagnostics
        fixture = os.path.join(ROOT, "tests", "golden", "big_qnet3_5k.npz")
        reps = 4
        with multiprocessing.Pool(reps) as pool:
            res = pool.map(run_one, [(fixture, "SLICE", a.iters, 100 + r, True, a.core + r) for r in range(reps)])
        th = numpy.array([r["mus"] for r in res]).transpose(1, 0, 2)       # [n, reps, P]
        mw = numpy.array([r["waits"] for r in res]).transpose(1, 0, 2)[:, :, 1:]
        x = numpy.concatenate((th, mw), axis=2)
        burn = a.iters // 4
        x = x[burn:]
        n = x.shape[0]
        ess_chain = numpy.array([[diagnostics.ess(x[:, c:c + 1, k:k + 1])[0] for k in range(x.shape[2])] for c in range(reps)])
        rate = numpy.mean([r["nevents"] / r["seconds"] for r in res])
        sweeps_per_s = numpy.mean([a.iters / r["seconds"] for r in res])
        ess_per_sweep = float(ess_chain.mean(axis=0).min() / n)
        numpy.savez_compressed(os.path.join(ROOT, "tests", "golden", "post_qnet3_5k.npz"), theta=th, mean_wait=mw,
                               resamples_per_s=rate)
        update("config2_ess", {"where": where, "replicas": reps, "iterations": a.iters, "burn": burn,
                               "ess_per_sweep_per_chain": ess_per_sweep, "sweeps_per_s_per_core": float(sweeps_per_s),
                               "ess_per_s_per_core": ess_per_sweep * float(sweeps_per_s), "rhat_max": float(numpy.max(diagnostics.rhat(x))),
                               "note": "min over service parameters and per-queue mean waits of the mean per-chain ESS; "
                                       "reference estimation.bayes (slice) on the identical 5000-task initial state; traces in "
                                       "tests/golden/post_qnet3_5k.npz"})
        return
    fixture = os.path.join(ROOT, "tests", "golden", "big_%s.npz" % FIX[a.what])
    rec = {"where": where, "state": "tests/golden/big_%s.npz (reference-generated, identical to the GPU arm's)" % FIX[a.what],
           "iterations": a.iters}
    t0 = time.time()
    r = run_one((fixture, "SLICE", a.iters, 13, False, a.core))
    rec["slice_resamples_per_s_per_core"] = r["nevents"] / r["seconds"]
    rec["slice_seconds"] = r["seconds"]
    rec["slice_resamples"] = r["nevents"]
    rec["load_seconds"] = r["load_seconds"]
    update(a.what, rec)
    if a.what != "config4":  # PS has no arrivalLik: slice only (queues.pyx:1140)
        r = run_one((fixture, "MH", a.iters, 13, False, a.core))
        rec["mh_proposals_per_s_per_core"] = r["nevents"] / r["seconds"]
        rec["mh_reject_frac"] = r["nreject"] / max(r["nevents"], 1)
        rec["mh_seconds"] = r["seconds"]
        update(a.what, rec)
    print("total %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
