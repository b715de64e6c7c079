#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (oracle/_ref, built by oracle/build_ref.py from
/root/reference/src with a mechanical Py2->Py3 patch).  TEST INFRASTRUCTURE ONLY.

The reference cannot travel to the GPU box, so its outputs are committed as small fixtures:

  lpdf_kat.npz        Distribution.lpdf values (distributions.pyx) at fixed points, incl. the reference's own
                      known answers (test_distributions.py:44-46, 62-75).
  cond_<net>.npz      for each network: a state produced by the reference (sample -> subset_by_task ->
                      gibbs_initialize [-> a few slice sweeps]), flat fp64 arrays + links, Qnet.log_prob, and
                      for a set of hidden events the values of GGkGibbs.inner_dfun (qnet.pyx:1340-1357) on a grid
                      -- the test_likdelta.py:43-75 identity turned into fixed vectors.
  mh_<net>.npz        the MH proposal log-density (queues.pair_proposal / departure_proposal, queues.pyx:1239-1283) on
                      grids for the same initial states; postmh_<net>.npz: long estimation.bayes runs with
                      sampling_type MH (Qnet.gibbs_resample, qnet.pyx:337)
  post_<net>.npz      traces of estimation.bayes (estimation.py:104) run long by the reference from a saved
                      initial state: service parameters per iteration and per-queue mean waiting time -- the
                      statistical pin for posterior agreement.

  big_<cfg>.npz       the same kind of vectors on states at BASELINE.json sizes (5 000-task 3-tier network, 20 000-task
                      G/G/3 tandem, 50 000-task PS + delay network), after the reference's initialisation and after one
                      reference sweep, including final events with candidates up to Tmax and far moves.  Run one at a
                      time: --only big_qnet3_5k | big_lnk_20k | big_psdelay_50k | ...

Usage: python oracle/make_golden.py [--only NAME] [--skip-post]
"""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
# the reference holds at most 10 000 events per queue unless told otherwise (arrivals.pyx:237; knob added by
# oracle/build_ref.py): the config-size states need more
os.environ.setdefault("BQ_REF_BIG_N", "65536")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, REF)

_sink = io.StringIO()
with contextlib.redirect_stdout(_sink):
    import qnetu, qnet, sampling, estimation, arrivals, qstats, distributions  # noqa: E402
sampling.DEBUG = 0


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


NETS = {
    # config 1 of BASELINE.json: source M(1.0) -> single-server M(0.7)
    "mm1": """
states:
  - name: INITIAL
    queues: [ INITIAL ]
    successors: [ TIER1 ]
    initial: TRUE
  - name: TIER1
    queues: [ WEB1 ]
queues:
  - { name: INITIAL, service: [M, 1.0] }
  - { name: WEB1, service: [M, 0.7] }
""",
    # the paper's 3-tier model (test_qnet.py:1056-1078) with the parameters of test_qnet.py:589
    "qnet3": """
states:
  - name: INITIAL
    queues: [ INITIAL ]
    successors: [ TIER1 ]
    initial: TRUE
  - name: TIER1
    queues: [ WEB1, WEB2, WEB3 ]
    successors: [ TIER2 ]
  - name: TIER2
    queues: [ APP1, APP2 ]
    successors: [ TIER3 ]
  - name: TIER3
    queues: [ DB ]
queues:
  - { name: INITIAL, service: [M, 1.0] }
  - { name: WEB1, service: [M, 0.2] }
  - { name: WEB2, service: [M, 0.2] }
  - { name: WEB3, service: [M, 0.2] }
  - { name: APP1, service: [M, 0.05] }
  - { name: APP2, service: [M, 0.05] }
  - { name: DB, service: [M, 0.3] }
""",
    # delay station (test_delay_station.py:116-126) behind a FIFO stage
    "delay": """
states:
  - name: I0
    queues: [ I0 ]
    successors: [ S1 ]
  - name: S1
    queues: [ Q1 ]
    successors: [ S2 ]
  - name: S2
    queues: [ Q2 ]
queues:
  - { name: I0, service: [M, 1.0] }
  - { name: Q1, service: [M, 0.5] }
  - { name: Q2, service: [M, 2.0], type: DELAY }
""",
    # lognormal / gamma single-server tandem (test_mh_ggk.py ln1a shape; gamma3 of test_qnet.py)
    "lng1": """
states:
  - name: INITIAL
    queues: [ INITIAL ]
    successors: [ TIER1 ]
    initial: TRUE
  - name: TIER1
    queues: [ WEB1 ]
    successors: [ TIER2 ]
  - name: TIER2
    queues: [ APP1 ]
queues:
  - { name: INITIAL, service: [M, 10.0] }
  - { name: WEB1, processors: 1, type: GG1, service: [LN, 0.75, 0.8] }
  - { name: APP1, processors: 1, type: GG1, service: [G, 2.0, 1.5] }
""",
    # multi-server FIFO (examples/models/mm3.yml)
    "mm3": """
states:
  - name: INITIAL
    queues: [ INITIAL ]
    successors: [ TIER1 ]
    initial: TRUE
  - name: TIER1
    queues: [ WEB1 ]
    successors: [ TIER2 ]
  - name: TIER2
    queues: [ APP1 ]
queues:
  - { name: INITIAL, service: [M, 10.0]  }
  - { name: WEB1, processors: 3, service: [M, 30.0] }
  - { name: APP1, processors: 4, service: [M, 30.0] }
""",
    # lognormal G/G/3 tandem (test_mh_ggk.py:924-939), source scaled so the stations are stable
    "lnk": """
states:
  - name: INITIAL
    queues: [ INITIAL ]
    successors: [ TIER1 ]
    initial: TRUE
  - name: TIER1
    queues: [ WEB1 ]
    successors: [ TIER2 ]
  - name: TIER2
    queues: [ APP1 ]
queues:
  - { name: INITIAL, service: [M, 5.0]  }
  - { name: WEB1, processors: 3, type: GGk, service: [LN, 1.75, 1.0] }
  - { name: APP1, processors: 3, type: GGk, service: [LN, 0.1, 2.0] }
""",
    # mixture service templates (test_mixture_template.py:41, test_qnet.py:1165-1199) on single-server queues
    "mix": """
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
  - { name: INITIAL, service: [M, 1.0] }
  - { name: Q1, service: [ MIX, 0.6, [M, 0.3], 0.4, [LN, 0.0, 0.7] ] }
  - { name: Q2, service: [ MIX, 0.5, [G, 2.0, 0.2], 0.5, [M, 1.0] ] }
""",
    # processor sharing + delay station (test_ps.py:266-292 + test_delay_station.py:116-126): config 4's shape
    "psdelay": """
states:
  - name: I0
    queues: [ I0 ]
    successors: [ S1 ]
  - name: S1
    queues: [ Q1 ]
    successors: [ S2 ]
  - name: S2
    queues: [ Q2 ]
queues:
  - { name: I0, service: [M, 1.0] }
  - { name: Q1, service: [M, 0.5], type: PS }
  - { name: Q2, service: [M, 2.0], type: DELAY }
""",
    # random-order single-server queues (QueueGG1R; test_gg1r.py:436-451 m2one, test_likdelta.py:119-137 mmrss)
    "gg1r": """
states:
 - name: I0
   queues: [I0]
   successors: [S1]
 - name: S1
   queues: [Q1,Q2]
   successors: [S2]
 - name: S2
   queues: [Q11]
queues:
  - { name: I0, service: [M, 1.0] }
  - { name: Q1, service: [M, 1.2], type: GG1R }
  - { name: Q2, service: [G, 2.0, 0.6], type: GG1R }
  - { name: Q11, service: [M, 0.75], type: GG1R }
""",
    # the same with exponential service everywhere (test_gg1r.py:436-451): long posterior runs
    "gg1rm": """
states:
 - name: I0
   queues: [I0]
   successors: [S1]
 - name: S1
   queues: [Q1,Q2]
   successors: [S2]
 - name: S2
   queues: [Q11]
queues:
  - { name: I0, service: [M, 1.0] }
  - { name: Q1, service: [M, 1.2], type: GG1R }
  - { name: Q2, service: [M, 1.5], type: GG1R }
  - { name: Q11, service: [M, 0.75], type: GG1R }
""",
}

# name -> (tasks, pct observed, initializer, seed, sweeps before the second snapshot)
COND = {
    "mm1": (60, 0.5, "DFN2", 13, 3),
    "qnet3": (80, 0.2, "DFN2", 14, 3),
    "delay": (60, 0.3, "NONE", 15, 3),   # DelayStation has no departure_of_service: DFN cannot initialise it
    "lng1": (60, 0.3, "DFN2", 16, 3),
    "mm3": (80, 0.2, "DFN2", 17, 5),
    "lnk": (60, 0.2, "DFN2", 18, 5),
    "psdelay": (50, 0.2, "PS", 19, 3),
    "mix": (70, 0.3, "DFN2", 20, 3),
    # the reference's initialisers do not pass its own validate() on GG1R networks (validate_caches, queues.pyx:843-853,
    # on the partially filled queue); its tests start from the sampled state with the hidden flags set
    # (test_likdelta.py:43-48), and so does this: initializer "TRUE"
    "gg1r": (70, 0.3, "TRUE", 31, 3),
}

# name -> (tasks, pct, initializer, seed, iterations, replicas)
POST = {
    "mm1": (200, 0.5, "DFN2", 13, 1500, 4),      # BASELINE.json configs[0]
    "qnet3": (150, 0.2, "DFN2", 21, 1200, 3),
    "lng1": (100, 0.3, "DFN2", 22, 3000, 8),     # slow-mixing Gamma scale: needs long runs for a usable MCSE
    "mm3": (100, 0.3, "DFN2", 23, 2500, 8),       # heavily loaded 3-server queue: waits mix slowly
    "psdelay": (80, 0.3, "PS", 24, 1500, 8),
    "gg1rm": (80, 0.3, "TRUE", 32, 1500, 8),
}


def flatten(arrv):
    """Rows = events, each task's events contiguous and in visit order; links as row indices."""
    evts = []
    for tid in sorted(arrv.all_tasks()):
        evts.extend(arrv.events_of_task(tid))
    row = dict((e.eid, i) for i, e in enumerate(evts))
    lk = lambda e: row[e.eid] if e is not None else -1
    out = dict(
        eid=numpy.array([e.eid for e in evts], dtype=numpy.int32),
        tid=numpy.array([e.tid for e in evts], dtype=numpy.int32),
        qid=numpy.array([e.qid for e in evts], dtype=numpy.int32),
        state=numpy.array([e.state for e in evts], dtype=numpy.int32),
        a=numpy.array([e.a for e in evts], dtype=numpy.float64),
        d=numpy.array([e.d for e in evts], dtype=numpy.float64),
        s=numpy.array([e.s for e in evts], dtype=numpy.float64),
        proc=numpy.array([e.proc for e in evts], dtype=numpy.int32),
        obs_a=numpy.array([e.obs_a for e in evts], dtype=numpy.uint8),
        obs_d=numpy.array([e.obs_d for e in evts], dtype=numpy.uint8),
        prev_byt=numpy.array([lk(e.previous_by_task()) for e in evts], dtype=numpy.int32),
        next_byt=numpy.array([lk(e.next_by_task()) for e in evts], dtype=numpy.int32),
        prev_byq=numpy.array([lk(e.previous_by_queue()) for e in evts], dtype=numpy.int32),
        next_byq=numpy.array([lk(e.next_by_queue()) for e in evts], dtype=numpy.int32),
        aux=numpy.array([_component(e) for e in evts], dtype=numpy.uint8),
    )
    return out, evts


def _component(e):
    """Mixture component of the event (Event.variable of its queue's template), 0 for plain queues."""
    name = "COMPONENT_%s" % e.queue().name
    try:
        v = e.variable(name)
        return int(v) if v is not None else 0
    except Exception:
        return 0


def make_state(name, ntasks, pct, init, seed):
    net = qnetu.qnet_from_text(NETS[name])
    qnetu.set_seed(seed)
    with quiet():
        if init == "TRUE":  # the sampled state itself, hidden flags set
            smp = net.sample(ntasks)
            ini = smp.subset_by_task(pct)
            # as after loading the state from a file: s = d - max(a, previous departure) in floating point, not the
            # simulator's draw, so that the commencement max(a, d - s) (arrivals.pyx:190-192), at which QueueGG1R counts
            # the jobs present, is the number any holder of (a, d) alone computes
            ini.recompute_all_service()
            ini.regenerate_all_caches()
        else:
            qnet.set_initialization_type(init)
            smp = net.sample(ntasks)
            sub = smp.subset_by_task(pct)
            ini = net.gibbs_initialize(sub)
            qnet.set_initialization_type("DFN2")
    ini.validate()
    return net, ini


def cond_grid(net, arrv, max_events, rng):
    """inner_dfun on a grid for up to max_events resample-eligible events (qnet.pyx:688)."""
    flat, evts = flatten(arrv)
    elig = [i for i, e in enumerate(evts)
            if not e.obs_d and (e.next_by_task() is None or not e.next_by_task().obs_a)]
    rng.shuffle(elig)
    elig = sorted(elig[:max_events])
    tmax = 2 * max(e.d for e in evts)
    rows, xs, vals = [], [], []
    for i in elig:
        e = evts[i]
        nxt = e.next_by_task()
        lo, hi = e.a, (nxt.d if nxt is not None else min(tmax, e.d + 3 * (e.d - e.a) + 1.0))
        g = numpy.concatenate(([e.d], lo + (hi - lo) * numpy.array([0.03, 0.2, 0.4, 0.5, 0.6, 0.8, 0.97]),
                               lo + (hi - lo) * rng.rand(4),
                               e.d + 0.02 * min(e.d - lo, hi - e.d) * numpy.array([-1.0, -0.3, 0.3, 1.0]),
                               e.d + 0.3 * min(e.d - lo, hi - e.d) * numpy.array([-1.0, 1.0]),
                               [lo - 0.1, hi + 0.1]))
        if any(type(q).__name__ == "QueuePS" for q in net.queues):
            g = g[:-2]  # QueuePS asserts (queues.pyx:1068) when evaluated outside [a_e, upper]
        dfn = qnet.GGkGibbs(net, arrv, e, 0.0).dfn()
        with quiet():
            v = numpy.array([dfn(float(x)) for x in g])
        rows.append(i); xs.append(g); vals.append(v)
    return flat, numpy.array(rows, dtype=numpy.int32), numpy.array(xs), numpy.array(vals)


def do_cond(name):
    ntasks, pct, init, seed, nsweeps = COND[name]
    net, arrv = make_state(name, ntasks, pct, init, seed)
    rng = numpy.random.RandomState(seed + 1000)
    out = dict(yaml=NETS[name], theta=numpy.array(net.parameters, dtype=numpy.float64))
    flat, rows, xs, vals = cond_grid(net, arrv, 40, rng)
    for k, v in flat.items():
        out["s0_" + k] = v
    out.update(s0_rows=rows, s0_x=xs, s0_g=vals, s0_logprob=net.log_prob(arrv))
    with quiet():
        net.slice_resample(arrv, nsweeps)
    arrv.validate()
    flat, rows, xs, vals = cond_grid(net, arrv, 40, rng)
    for k, v in flat.items():
        out["s1_" + k] = v
    out.update(s1_rows=rows, s1_x=xs, s1_g=vals, s1_logprob=net.log_prob(arrv))
    # the same state under different parameters (exercises the lpdf constants)
    theta2 = [p * 1.3 + 0.05 for p in net.parameters]
    net.parameters = theta2
    flat, rows, xs, vals = cond_grid(net, arrv, 15, rng)
    out.update(s2_theta=numpy.array(theta2), s2_rows=rows, s2_x=xs, s2_g=vals, s2_logprob=net.log_prob(arrv))
    numpy.savez_compressed(os.path.join(OUT, "cond_%s.npz" % name), **out)
    print("cond_%s: E=%d, %d+%d+%d events on grids" % (name, len(flat["d"]), len(out["s0_rows"]), len(out["s1_rows"]), len(rows)))


def do_mh(name):
    """MH-within-Gibbs pins: the TWOS proposal log-density (queues.pair_proposal / departure_proposal,
    queues.pyx:1239-1283) on grids for resample-eligible events of the cond_<name> initial state."""
    import queues
    ntasks, pct, init, seed, nsweeps = COND[name]
    net, arrv = make_state(name, ntasks, pct, init, seed)
    rng = numpy.random.RandomState(seed + 2000)
    flat, evts = flatten(arrv)
    elig = [i for i, e in enumerate(evts)
            if not e.obs_d and (e.next_by_task() is None or not e.next_by_task().obs_a)]
    rng.shuffle(elig)
    elig = sorted(elig[:40])
    rows, xs, vals, lo_hi = [], [], [], []
    for i in elig:
        e = evts[i]
        nxt = e.next_by_task()
        with quiet():
            fn = queues.pair_proposal(arrv, e, nxt) if nxt is not None else queues.departure_proposal(arrv, e)
        L, U = fn.range()
        hi = U if numpy.isfinite(U) else L + 3.0 * (e.d - L) + 1.0
        g = numpy.concatenate(([e.d], L + (hi - L) * numpy.array([0.001, 0.05, 0.3, 0.5, 0.7, 0.95, 0.999]),
                               L + (hi - L) * rng.rand(4), [L - 0.1, hi + 0.1]))
        with quiet():
            v = numpy.array([float(fn(float(x))) for x in g])
        rows.append(i); xs.append(g); vals.append(v); lo_hi.append((L, U))
    out = dict(yaml=NETS[name], theta=numpy.array(net.parameters, dtype=numpy.float64),
               rows=numpy.array(rows, dtype=numpy.int32), x=numpy.array(xs), h=numpy.array(vals), range=numpy.array(lo_hi))
    for k, v in flat.items():
        out["s0_" + k] = v
    numpy.savez_compressed(os.path.join(OUT, "mh_%s.npz" % name), **out)
    print("mh_%s: %d events on grids" % (name, len(rows)))


def do_kern(name):
    """Qnet.gibbs_kernel (qnet.pyx:741-809) and Qnet.parameter_kernel (qnet.pyx:811-837): the transition densities
    estimation.chib_evidence averages, from a state one slice sweep away from the target."""
    ntasks, pct, init, seed, nsweeps = COND[name]
    net, arrv = make_state(name, ntasks, pct, init, seed)
    qnetu.set_seed(seed + 3000)
    with quiet():
        star = net.slice_resample(arrv.duplicate(), 3)[-1]
        frm = net.slice_resample(star.duplicate(), 1)[-1]
        order_evts = frm.duplicate().all_events()
        kp = net.gibbs_kernel(frm, star)
        theta_from = list(net.parameters)
        theta_to = [1.1 * t if i % 2 == 0 else 0.9 * t for i, t in enumerate(theta_from)]
        pk = net.parameter_kernel(theta_from, theta_to, frm)
    flat, evts = flatten(frm)
    flat_to, _ = flatten(star)
    row = dict((e.eid, i) for i, e in enumerate(evts))
    out = dict(yaml=NETS[name], theta=numpy.array(net.parameters, dtype=numpy.float64), kernel=float(kp),
               order=numpy.array([row[e.eid] for e in order_evts], dtype=numpy.int32), to_d=flat_to["d"],
               theta_to=numpy.array(theta_to), param_kernel=float(pk))
    for k, v in flat.items():
        out["s0_" + k] = v
    numpy.savez_compressed(os.path.join(OUT, "kern_%s.npz" % name), **out)
    print("kern_%s: gibbs_kernel %.6f parameter_kernel %.6f" % (name, kp, pk))


def do_qstats(name):
    """Every exported statistic of the reference's qstats.py (qstats.py:174-184, plus the task-level ones) on the
    cond_<name> initial state."""
    ntasks, pct, init, seed, nsweeps = COND[name]
    net, arrv = make_state(name, ntasks, pct, init, seed)
    flat, _ = flatten(arrv)
    out = dict(yaml=NETS[name], theta=numpy.array(net.parameters, dtype=numpy.float64))
    for k, v in flat.items():
        out["s0_" + k] = v
    names = [fn.__name__ for fn in qstats.STATS] + ["max_service", "mean_task_service", "mean_task_length", "total_service"]
    for n in names:
        out["stat_" + n] = numpy.atleast_1d(numpy.array(getattr(qstats, n)(arrv), dtype=numpy.float64))
    out["stat_names"] = numpy.array(names)
    numpy.savez_compressed(os.path.join(OUT, "qstats_%s.npz" % name), **out)
    print("qstats_%s: %d statistics" % (name, len(names)))


def _post_rep(args):
    """One replica of the reference's estimation.bayes from the common initial state (own process: the reference
    keeps its sampler state in module globals)."""
    name, r = args
    sampler = "SLICE"
    if name.endswith("@MH"):
        name, sampler = name[:-3], "MH"
    ntasks, pct, init, seed, niter, nrep = POST[name]
    net, ini = make_state(name, ntasks, pct, init, seed)
    with quiet():
        estimation.set_sampler(sampler)
    arrv = ini.duplicate()
    qnetu.set_seed(1000 * seed + r)
    w = []
    qnet.reset_stats()
    t0 = time.time()
    with quiet():
        if sampler == "SLICE":
            mus, _ = estimation.bayes(net, arrv, niter, report_fn=lambda n, a, i, lp: w.append(qstats.mean_wait(a)))
        else:
            # estimation.bayes with sampling_type MH cannot take a report_fn (estimation.py:141 passes `lp`, which
            # only the slice branch defines), so its loop (estimation.py:123-151) is spelled out here
            net.estimate(arrv)
            mus = []
            for i in range(niter):
                arrv = net.gibbs_resample(arrv, burn=0, num_iter=1, return_arrv=False)[-1]
                w.append(qstats.mean_wait(arrv))
                net.sample_parameters(arrv)
                mus.append(net.parameters[:])
    dt = time.time() - t0
    return numpy.array(mus), numpy.array(w), qnet.stats()["nevents"] / dt


def do_post(name, sampler="SLICE"):
    import multiprocessing
    key = name + ("@MH" if sampler == "MH" else "")
    ntasks, pct, init, seed, niter, nrep = POST[name]
    net, ini = make_state(name, ntasks, pct, init, seed)
    theta0 = list(net.parameters)
    flat, _ = flatten(ini)
    with multiprocessing.Pool(min(nrep, os.cpu_count() or 1)) as pool:
        res = pool.map(_post_rep, [(key, r) for r in range(nrep)])
    thetas, waits, rate = [x[0] for x in res], [x[1] for x in res], [x[2] for x in res]
    out = dict(yaml=NETS[name], theta0=numpy.array(theta0), theta=numpy.array(thetas), mean_wait=numpy.array(waits),
               ref_resamples_per_s=numpy.array(rate))
    for k, v in flat.items():
        out["s0_" + k] = v
    fname = "post_%s.npz" % name if sampler == "SLICE" else "postmh_%s.npz" % name
    numpy.savez_compressed(os.path.join(OUT, fname), **out)
    print("%s: E=%d iters=%d reps=%d  ref %.0f resamples/s  theta mean %s" % (
        fname, len(flat["d"]), niter, nrep, numpy.mean(rate), numpy.array(thetas)[:, niter // 5:].mean(axis=(0, 1))))


def do_lpdf():
    pts = numpy.array([1e-6, 0.01, 0.1, 0.5, 1.0, 2.0, 3.7, 10.0, 55.0])
    cases = []
    for p in (0.05, 0.7, 1.0, 30.0):
        f = distributions.Exponential(p)
        cases.append((0, p, 0.0, [f.lpdf(x) for x in pts]))
    for sh, sc in ((3.0, 0.5), (1.0, 2.0), (0.7, 1.5), (12.0, 0.1)):
        f = distributions.Gamma(sh, sc)
        cases.append((1, sh, sc, [f.lpdf(x) for x in pts]))
    for mu, sd in ((0.0, 1.0), (1.75, 1.0), (0.1, 2.0), (-1.0, 0.3)):
        f = distributions.LogNormal(mu, sd)
        cases.append((2, mu, sd, [f.lpdf(x) for x in pts]))
    numpy.savez_compressed(os.path.join(OUT, "lpdf_kat.npz"), x=pts,
                           dist=numpy.array([c[0] for c in cases], dtype=numpy.int32),
                           p0=numpy.array([c[1] for c in cases]), p1=numpy.array([c[2] for c in cases]),
                           lpdf=numpy.array([c[3] for c in cases]),
                           ln_at_zero=distributions.LogNormal(0.0, 1.0).lpdf(0.0))
    print("lpdf_kat: %d cases; Gamma(3,0.5).lpdf(2) = %.6f (reference test pins -1.227411)" % (
        len(cases), distributions.Gamma(3.0, 0.5).lpdf(2.0)))


# ---- states at BASELINE.json sizes ------------------------------------------------------------------------------
# name -> (network, tasks, pct observed, initializer, seed, events on grids after init, after one reference sweep)
BIG = {
    "qnet3_5k": ("qnet3", 5000, 0.2, "DFN2", 13, 220, 120),      # BASELINE configs[1]
    "lnk_5k": ("lnk", 5000, 0.2, "DFN2", 13, 220, 120),          # configs[2] network at a quarter of its size
    "lnk_20k": ("lnk", 20000, 0.2, "DFN2", 13, 260, 140),        # configs[2]
    "psdelay_10k": ("psdelay", 10000, 0.2, "PS", 13, 220, 120),  # configs[3] network at a fifth of its size
    "psdelay_50k": ("psdelay", 50000, 0.2, "PS", 13, 260, 140),  # configs[3]
}


def big_grid(net, arrv, evts, n_events, rng):
    """inner_dfun on grids for n_events resample-eligible events of a config-size state, chosen so that what only
    happens at scale is covered: final events (bracket up to Tmax = 2 max d, qnet.pyx:672, 698-699), candidates far
    from the current value (long diff lists, re-ranking of the next event by many positions), as well as near ones."""
    elig = [i for i, e in enumerate(evts)
            if not e.obs_d and (e.next_by_task() is None or not e.next_by_task().obs_a)]
    finals = [i for i in elig if evts[i].next_by_task() is None]
    others = [i for i in elig if evts[i].next_by_task() is not None]
    rng.shuffle(finals)
    rng.shuffle(others)
    n_fin = min(len(finals), n_events // 3)
    pick = sorted(finals[:n_fin] + others[:n_events - n_fin])
    tmax = 2 * max(e.d for e in evts)
    is_ps = any(type(q).__name__ == "QueuePS" for q in net.queues)
    rows, xs, vals = [], [], []
    for i in pick:
        e = evts[i]
        nxt = e.next_by_task()
        lo = e.a
        hi = nxt.d if nxt is not None else tmax
        span = hi - lo
        near = min(e.d - lo, hi - e.d)
        pts = [e.d, lo + 0.03 * span, lo + 0.5 * span, lo + 0.97 * span, lo + span * rng.rand(), lo + span * rng.rand(),
               e.d - 0.3 * near, e.d + 0.3 * near, e.d - 0.02 * near, e.d + 0.02 * near]
        if nxt is None:  # final event: the reference's bracket reaches Tmax; also moderately far points
            scale = max(e.d - e.a, 1e-3)
            pts += [e.d + 5.0 * scale, e.d + 50.0 * scale, e.d + 500.0 * scale, 0.5 * (e.d + tmax), tmax * (1.0 - 1e-9)]
        if not is_ps:  # QueuePS asserts (queues.pyx:1068) when evaluated outside [a_e, upper]
            pts += [lo - 0.1, hi + 0.1]
        g = numpy.array(pts)
        dfn = qnet.GGkGibbs(net, arrv, e, 0.0).dfn()
        with quiet():
            v = numpy.array([dfn(float(x)) for x in g])
        pad = 17 - len(g)
        rows.append(i)
        xs.append(numpy.concatenate((g, numpy.full(pad, numpy.nan))))
        vals.append(numpy.concatenate((v, numpy.full(pad, numpy.nan))))
    return numpy.array(rows, dtype=numpy.int32), numpy.array(xs), numpy.array(vals)


def do_big(name):
    """A state of BASELINE size produced by the reference (sample -> subset_by_task -> gibbs_initialize), inner_dfun
    grids on it, one reference slice sweep, grids again.  Stored compactly: rows are task-major (each task's events in
    visit order), a / s / proc / by-task links are derivable, the by-queue order is the order by arrival (queue 0 by
    departure), which tests/helpers.py re-derives (checked here before saving)."""
    netname, ntasks, pct, init, seed, n0, n1 = BIG[name]
    t0 = time.time()
    net = qnetu.qnet_from_text(NETS[netname])
    qnetu.set_seed(seed)
    with quiet():
        if init == "TRUE":  # the sampled state itself, hidden flags set
            smp = net.sample(ntasks)
            ini = smp.subset_by_task(pct)
            # as after loading the state from a file: s = d - max(a, previous departure) in floating point, not the
            # simulator's draw, so that the commencement max(a, d - s) (arrivals.pyx:190-192), at which QueueGG1R counts
            # the jobs present, is the number any holder of (a, d) alone computes
            ini.recompute_all_service()
            ini.regenerate_all_caches()
        else:
            qnet.set_initialization_type(init)
            smp = net.sample(ntasks)
            sub = smp.subset_by_task(pct)
            ini = net.gibbs_initialize(sub)
            qnet.set_initialization_type("DFN2")
    ini.validate()
    print("big_%s: reference init of %d tasks took %.0f s" % (name, ntasks, time.time() - t0), flush=True)
    rng = numpy.random.RandomState(seed + 7000)
    out = dict(yaml=NETS[netname], theta=numpy.array(net.parameters, dtype=numpy.float64))
    arrv = ini
    for tag, n_ev in (("s0", n0), ("s1", n1)):
        flat, evts = flatten(arrv)
        if tag == "s0":
            for k in ("tid", "qid", "state", "obs_a", "obs_d", "eid"):
                out[k] = flat[k]
        # the by-queue order must be derivable from the times (no ties inside a queue but queue 0's zero arrivals)
        for q in range(len(net.queues)):
            rws = numpy.flatnonzero(flat["qid"] == q)
            if type(net.queues[q]).__name__ in ("QueuePS", "DelayStation") and tag == "s1":
                continue  # the reference never relinks a PS queue or a delay station (queues.pyx:1149-1183, 745-784)
            key = flat["d"][rws] if q == 0 else flat["a"][rws]
            order = rws[numpy.argsort(key, kind="stable")]
            assert numpy.all(flat["prev_byq"][order[1:]] == order[:-1]), "queue %d is not in arrival order" % q
        t1 = time.time()
        rows, xs, vals = big_grid(net, arrv, evts, n_ev, rng)
        out[tag + "_d"] = flat["d"]
        out.update({tag + "_rows": rows, tag + "_x": xs, tag + "_g": vals, tag + "_logprob": net.log_prob(arrv)})
        print("big_%s %s: %d events on grids (%d final) in %.0f s" % (
            name, tag, len(rows), sum(1 for i in rows if evts[i].next_by_task() is None), time.time() - t1), flush=True)
        if tag == "s0":
            t1 = time.time()
            qnet.reset_stats()
            with quiet():
                net.slice_resample(arrv, 1)
            dt = time.time() - t1
            arrv.validate()
            out["ref_sweep_seconds"] = dt
            out["ref_sweep_resamples"] = qnet.stats()["nevents"]
            print("big_%s: one reference sweep: %d resamples in %.0f s" % (name, qnet.stats()["nevents"], dt), flush=True)
    numpy.savez_compressed(os.path.join(OUT, "big_%s.npz" % name), **out)
    print("big_%s: E=%d written (%.1f MB)" % (name, len(out["s0_d"]), os.path.getsize(os.path.join(OUT, "big_%s.npz" % name)) / 1e6))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--skip-post", action="store_true")
    ap.add_argument("--skip-cond", action="store_true")
    ap.add_argument("--skip-mh", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if args.only and args.only.startswith("big_"):
        do_big(args.only[4:])
        sys.exit(0)
    if args.only in (None, "lpdf"):
        do_lpdf()
    if not args.skip_cond:
        for n in COND:
            if args.only in (None, n):
                do_cond(n)
    for n in ("qnet3", "mm3", "psdelay"):
        if args.only in (None, n) and not args.skip_cond:
            do_qstats(n)
    for n in ("mm1", "qnet3", "lng1", "delay", "mm3"):  # transition densities (Chib evidence)
        if args.only in (None, n, "kern"):
            do_kern(n)
    if args.only == "kern":
        sys.exit(0)
    if not args.skip_mh:  # MH-within-Gibbs (gibbs_resample): proposal grids and long reference runs; no PS (queues.pyx:1140)
        for n in COND:
            # no MH proposal for PS (queues.pyx:1140), for mixture templates (DepartureProposal wants a
            # DistributionTemplate) nor for random-order queues (QueueGG1R.arrivalLik raises, queues.pyx:856-857)
            if n not in ("psdelay", "mix", "gg1r") and args.only in (None, n):
                do_mh(n)
        for n in ("mm1", "qnet3", "mm3"):
            if args.only in (None, n) and not args.skip_post:
                do_post(n, "MH")
    if not args.skip_post:
        for n in POST:
            if args.only in (None, n):
                do_post(n)
