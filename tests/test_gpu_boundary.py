"""Boundary behaviour of the reference-facing Python API on the GPU path: callback order of estimation.bayes / sem
(estimation.py:29-169), sem momentum, parameter updates from a LIST of Arrivals (qnet.pyx:844-896), engine-cache
staleness, networks with more queues than threads in the CTA, and the gibbs.py front end in its default StEM mode."""
import os
import subprocess
import sys

import numpy
import pytest

import helpers
from bayes_qnet_b200 import GibbsEngine, estimation, qnet, qnetu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bayes_report_fn_sees_the_parameters_the_sweep_used():
    """estimation.bayes calls report_fn BEFORE sample_parameters (estimation.py:148-155): at iteration i the net holds
    the parameters of iteration i - 1 (the ML initialisation at i = 0), and the state is the one just swept."""
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 60, 0.3, seed=3)
    seen = []

    def report(net_, arrv_, i, lp):
        arrv_.validate()
        seen.append((i, list(net_.parameters), float(lp), arrv_.soa()["d"].copy()))

    work = ini.duplicate()
    mus, arrvl = estimation.bayes(net, work, 6, report_fn=report, seed=11)
    assert [s[0] for s in seen] == list(range(6)) and len(mus) == 6
    net0 = qnetu.qnet_from_text(helpers.QNET3)
    theta0 = net0.estimate(ini.duplicate())
    assert numpy.allclose(seen[0][1], theta0, rtol=1e-12)
    for i in range(1, 6):
        assert numpy.allclose(seen[i][1], mus[i - 1], rtol=1e-12), "callback %d saw other parameters" % i
    assert all(numpy.isfinite(s[2]) for s in seen)
    assert not numpy.array_equal(seen[0][3], seen[1][3])
    assert numpy.allclose(net.parameters, mus[-1])


def test_sem_callbacks_momentum_and_burn():
    """sem(): burn_report_fn for i < burn, report_fn after, both after the ML update; momentum mixes the old and the
    new estimate (estimation.py:64-66); mus holds the post-burn iterations only."""
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 60, 0.3, seed=4)
    burned, kept = [], []
    mus, arrvl = estimation.sem(net, ini.duplicate(), 3, 5, report_fn=lambda n, a, i, lp: kept.append((i, list(n.parameters))),
                                burn_report_fn=lambda n, a, i: burned.append(i), seed=5)
    assert burned == [0, 1, 2] and [k[0] for k in kept] == [3, 4, 5, 6, 7] and len(mus) == 5
    for (i, th), mu in zip(kept, mus):
        assert numpy.allclose(th, mu, rtol=1e-12)
    # momentum: same seed, damped updates -> theta_i = m theta_{i-1} + (1 - m) ML_i; with m = 1 - 1e-12 nothing moves
    net2 = qnetu.qnet_from_text(helpers.QNET3)
    theta_init = net2.estimate(ini.duplicate())
    mus_m, _ = estimation.sem(net2, ini.duplicate(), 0, 4, momentum=0.5, seed=5)
    net3 = qnetu.qnet_from_text(helpers.QNET3)
    mus_0, _ = estimation.sem(net3, ini.duplicate(), 0, 4, momentum=0.0, seed=5)
    assert numpy.allclose(mus_m[0], mus_0[0], rtol=1e-12)  # first iteration: no previous estimate to mix with
    assert not numpy.allclose(mus_m[1], mus_0[1], rtol=1e-6)
    net4 = qnetu.qnet_from_text(helpers.QNET3)
    mus_1, _ = estimation.sem(net4, ini.duplicate(), 0, 4, momentum=1.0 - 1e-13, seed=5)
    assert numpy.allclose(mus_1[3], mus_1[0], rtol=1e-9)


def test_parameter_updates_from_a_list_of_arrivals():
    """Qnet.estimate / sample_parameters given a list pool the events of all Arrivals per queue (qnet.pyx:844-896)."""
    yaml_text = helpers.MIXED["delay_gg1_gg4"]
    net, smp, sub, ini = helpers.synthetic(yaml_text, 80, 0.25, seed=21)
    qnetu.set_seed(22)
    ini2 = net.gibbs_initialize(net.sample(50).subset_by_task(0.5))
    theta = net.estimate([ini, ini2])
    off = 0
    for qi, q in enumerate(net.queues):
        s = numpy.concatenate([numpy.array([e.s for e in a.events_of_queue(q)]) for a in (ini, ini2)])
        s = s[s > 0]
        if q.service.code == 0:
            assert abs(theta[off] - s.mean()) < 1e-12 * max(1.0, s.mean())
        elif q.service.code == 2:
            ls = numpy.log(s[s >= 1e-50])
            assert abs(theta[off] - ls.mean()) < 1e-10 and abs(theta[off + 1] - ls.std()) < 1e-10
        else:  # Gamma (Minka fixed point): mean is matched exactly by construction
            assert abs(theta[off] * theta[off + 1] - s.mean()) < 1e-9 * s.mean()
        off += q.service.numParameters()
    single = qnetu.qnet_from_text(yaml_text).estimate(ini)
    assert not numpy.allclose(single, theta)
    # posterior draws from the pooled data: finite, positive scales, and centred on the pooled estimate
    draws = numpy.array([net.sample_parameters([ini, ini2]) for _ in range(40)])
    assert numpy.all(numpy.isfinite(draws))
    assert abs(draws[:, 0].mean() / theta[0] - 1.0) < 0.2  # source queue: Exponential scale


def test_engine_cache_notices_structure_changes():
    """The cached single-chain engine behind the in-place API is rebuilt when observation flags change."""
    net, smp, sub, ini = helpers.synthetic(helpers.MM1, 40, 0.5, seed=9)
    net.slice_resample(ini, 1)
    n0 = qnet.stats()["nevents"]
    qnet.reset_stats()
    net.slice_resample(ini, 1)
    per_sweep = qnet.stats()["nevents"]
    # observe one more task: fewer events are eligible afterwards
    hidden_tasks = sorted({e.tid for e in ini if not e.obs_d})
    for e in ini.events_of_task(hidden_tasks[0]):
        e.obs_a = 1
        e.obs_d = 1
    qnet.reset_stats()
    net.slice_resample(ini, 1)
    assert qnet.stats()["nevents"] < per_sweep
    assert n0 > 0


def test_more_queues_than_threads_in_the_cta():
    """Q = 40 queues with BQ_THREADS = 32: every queue's constants, traces and parameter updates are handled by
    strided loops (the static kernels used to cover only the first blockDim.x queues)."""
    n_web = 38
    names = ["W%d" % i for i in range(n_web)]
    yaml_text = ("states:\n  - {name: INITIAL, queues: [INITIAL], successors: [T1], initial: TRUE}\n"
                 "  - {name: T1, queues: [%s], successors: [T2]}\n  - {name: T2, queues: [DB]}\nqueues:\n"
                 "  - { name: INITIAL, service: [M, 0.05] }\n" % ", ".join(names))
    for i, n in enumerate(names):
        yaml_text += "  - { name: %s, service: [%s] }\n" % (n, "M, %.2f" % (0.5 + 0.01 * i) if i % 2 else "G, 2.0, %.2f" % (0.2 + 0.01 * i))
    yaml_text += "  - { name: DB, service: [LN, -2.0, 0.5] }\n"
    net, smp, sub, ini = helpers.synthetic(yaml_text, 400, 0.3, seed=31)
    old = os.environ.get("BQ_THREADS")
    os.environ["BQ_THREADS"] = "32"
    try:
        with GibbsEngine(net, ini, n_chains=2, seed=77) as eng:
            assert eng.kernel_info()["threads"] == 32 and eng.Q == 40
            off, order = eng.schedule()
            eng.sweep(2, "SLICE", "BAYES", trace=True)
            th, mw, lp = eng.traces()
            oc = helpers.oracle_chain(net, ini)
            for s in range(2):
                oc.slice_sweep(order, eng.seed, 1, s)
                oc.update_params("BAYES", eng.seed, 1, s)
            assert helpers.relerr(eng.state(1, 2)[0], oc.d).max() < 1e-9
            assert helpers.relerr(eng.params(1, 2)[0], oc.theta).max() < 1e-9
            assert numpy.all(numpy.isfinite(mw)) and numpy.all(mw[:, :, 1:].max(axis=(0, 1)) >= 0)
            assert abs(eng.log_prob()[1] - oc.log_prob()) < 1e-9 * abs(oc.log_prob())
            # transition density of a sweep (the Chib kernel had the same guard)
            kp = eng.gibbs_kernel(ini.soa()["d"])[1]
            assert numpy.isfinite(kp) or kp == -numpy.inf
    finally:
        if old is None:
            del os.environ["BQ_THREADS"]
        else:
            os.environ["BQ_THREADS"] = old


def test_fifo_queue_with_more_servers_than_events_is_a_delay_station():
    """The reference's DELAY == GGk(k = 100) check (test_mh_ggk.py:778-784) and more: same log_prob, and the sweep
    follows the oracle's path on the k = 100 network (the oracle keeps all 100 servers)."""
    delay = """
states:
  - {name: INITIAL, queues: [INITIAL], successors: [T1], initial: TRUE}
  - {name: T1, queues: [Q1], successors: [T2]}
  - {name: T2, queues: [Q2]}
queues:
  - { name: INITIAL, service: [M, 1.0] }
  - { name: Q1, type: DELAY, service: [G, 2.0, 0.4] }
  - { name: Q2, service: [M, 0.5] }
"""
    k100 = delay.replace("type: DELAY", "type: GGk, processors: 100")
    netd, smp, sub, inid = helpers.synthetic(delay, 50, 0.3, seed=8)
    netk = qnetu.qnet_from_text(k100)
    from bayes_qnet_b200.arrivals import Arrivals
    soa = inid.soa()
    inik = Arrivals.from_arrays(netk, soa["tid"], soa["qid"], soa["state"], soa["a"], soa["d"], soa["obs_a"], soa["obs_d"])
    assert abs(netd.log_prob(inid) - netk.log_prob(inik)) < 1e-9
    with GibbsEngine(netk, inik, n_chains=2, seed=3) as eng:
        off, order = eng.schedule()
        eng.sweep(3, "SLICE", "STEM")
        oc = helpers.oracle_chain(netk, inik)
        for s in range(3):
            oc.slice_sweep(order, eng.seed, 1, s)
            oc.update_params("STEM", eng.seed, 1, s)
        full = eng.state(1, 2, full=True)
        assert helpers.relerr(full["d"][0], oc.d).max() < 1e-9
        assert helpers.relerr(full["s"][0], oc.s).max() < 1e-9
        assert helpers.relerr(eng.params(1, 2)[0], oc.theta).max() < 1e-9
        q1 = numpy.flatnonzero(soa["qid"] == 1)
        assert sorted(full["proc"][0][q1]) == list(range(len(q1)))


def test_gibbs_cli_stem_mode_writes_statistics(tmp_path):
    """python -m bayes_qnet_b200.gibbs in its default StEM mode with --statistics: the per-iteration files are filled
    (gibbs.py:156 relies on sem() calling report_fn)."""
    net, smp, sub, ini = helpers.synthetic(helpers.MM1, 60, 0.5, seed=2)
    netf, arrf = tmp_path / "net.yml", tmp_path / "arrv.txt"
    netf.write_text(helpers.MM1)
    qnetu.write_to_table(str(arrf), ini)
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "bayes_qnet_b200.gibbs", "--network", str(netf), "--arrvtbl", str(arrf),
                        "--gibbs-iter", "6", "--burn", "2", "--statistics", "mean_wait,mean_service", "--do-init", "1",
                        "--output-mu", "mu.txt", "--seed", "5"], cwd=str(tmp_path), env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for name in ("train_mean_wait.txt", "train_mean_service.txt"):
        lines = (tmp_path / name).read_text().strip().splitlines()
        assert len(lines) == 6 and lines[0].split()[0] == "2", (name, lines)
    assert len((tmp_path / "mu.txt").read_text().strip().splitlines()) == 6
