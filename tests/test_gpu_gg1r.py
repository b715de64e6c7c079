"""Random-order single-server queues (QueueGG1R, queues.pyx:788-1027) on the GPU.  The conditional against the
reference's inner_dfun vectors, sweep trajectories against the oracle and the posterior against the reference's long
runs are the "gg1r" / "gg1rm" / "gg1r_ps_gg2" cases of test_gpu_parity.py; here: every kernel walks the same path, the
state the C-ABI returns (links in departure order), the reference-API entry points, and what is refused."""
import numpy
import pytest

import helpers
from bayes_qnet_b200 import GibbsEngine, capi, estimation, qnetu

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _true_start(yaml_text, ntasks, pct, seed):
    """The sampled state with the hidden flags set -- how the reference's own GG1R tests start (test_likdelta.py:43-48)."""
    net = qnetu.qnet_from_text(yaml_text)
    qnetu.set_seed(seed)
    return net, net.sample(ntasks).subset_by_task(pct)


@pytest.mark.parametrize("kernel", ["batch", "4", "8", "32"])
def test_every_dynamic_kernel_walks_the_oracles_path(kernel, monkeypatch):
    """Batch kernel (32 events of a chain at once, range claims over arrival positions and departure slots) and the
    group kernel at several widths: sweeps with posterior draws, traces, suffstats and the derived state."""
    if kernel == "batch":
        monkeypatch.setenv("BQ_DYN_BATCH", "1")
    else:
        monkeypatch.setenv("BQ_DYN_BATCH", "0")
        monkeypatch.setenv("BQ_DYN_W", kernel)
    g = helpers.golden("cond_gg1r")
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=5, seed=911, chain_offset=2) as eng:
        off, order = eng.schedule()
        eng.sweep(4, "SLICE", "BAYES", trace=True)
        d, th = eng.state(), eng.params()
        tr_th, tr_mw, tr_lp = eng.traces()
        for c in (0, 4):
            oc = helpers.oracle_chain(net, arrv)
            for s in range(4):
                oc.slice_sweep(order, eng.seed, 2 + c, s)
                assert abs(tr_lp[s, c] - oc.log_prob()) < 1e-8 * max(1.0, abs(oc.log_prob()))
                oc.update_params("BAYES", eng.seed, 2 + c, s)
                assert helpers.relerr(tr_th[s, c], oc.theta).max() < 1e-7
            assert helpers.relerr(d[c], oc.d).max() < 1e-7
            full = eng.state(c, c + 1, full=True)
            assert helpers.relerr(full["s"][0], oc.s).max() < 1e-7
            assert numpy.array_equal(full["next_byq"][0], oc.next_byq) and numpy.array_equal(full["prev_byq"][0], oc.prev_byq)
        assert eng.stats()["nevents"] == 5 * 4 * len(order)


def test_links_are_in_departure_order_and_state_stays_feasible():
    """After many sweeps of a larger network: by-queue links of the random-order queues follow the departures
    (queue_order ARRV_BY_D, queues.pyx:796), services are d - max(a, previous departure) >= 0, the log-probability --
    which carries the start-on-arrival rule -- stays finite, observed times never move."""
    net, sub = _true_start(helpers.MIXED["gg1r_ps_gg2"], 400, 0.2, seed=8)
    soa = sub.soa()
    with GibbsEngine(net, sub, n_chains=4, seed=5) as eng:
        assert numpy.all(numpy.isfinite(eng.log_prob()))
        eng.sweep(30, "SLICE", "BAYES")
        assert numpy.all(numpy.isfinite(eng.log_prob()))
        full = eng.state(full=True)
        for c in range(4):
            d, a, s = full["d"][c], full["a"][c], full["s"][c]
            assert numpy.all(s >= 0)
            obs = soa["obs_d"] == 1
            assert numpy.array_equal(d[obs], soa["d"][obs])
            for q, queue in enumerate(net.queues):
                if queue.code != 3:
                    continue
                rows = numpy.nonzero(soa["qid"] == q)[0]
                head = [r for r in rows if full["prev_byq"][c][r] < 0]
                assert len(head) == 1
                lst, r = [], head[0]
                while r >= 0:
                    lst.append(r)
                    r = full["next_byq"][c][r]
                assert len(lst) == len(rows)
                dd = d[lst]
                assert numpy.all(numpy.diff(dd) >= 0)
                want = dd - numpy.maximum(a[lst], numpy.concatenate(([0.0], dd[:-1])))
                assert numpy.abs(want - s[lst]).max() < 1e-12
        assert len(set(full["d"][:, eng.schedule()[1][0]].tolist())) == 4


def test_reference_api_entry_points_run_on_gg1r_networks():
    """estimation.bayes / sem and Qnet.slice_resample on a random-order network, as test_gg1r.py:281-293 drives them."""
    g = helpers.golden("cond_gg1r")
    net, arrv = helpers.golden_state(g, "s0")
    lp0 = net.log_prob(arrv)
    assert abs(lp0 - float(g["s0_logprob"])) < TOL * abs(lp0)
    out = net.slice_resample(arrv, 3, return_arrv=True)
    out[-1].validate()
    assert numpy.isfinite(net.log_prob(out[-1]))
    mus, arrvl = estimation.bayes(net, arrv, 20, seed=3)
    assert len(mus) == 20 and numpy.all(numpy.isfinite(mus))
    mus, arrvl = estimation.sem(net, arrv, 2, 10)
    assert numpy.all(numpy.isfinite(mus))
    assert "type: GG1R" in net.as_yaml()


def test_what_the_reference_cannot_do_is_refused():
    g = helpers.golden("cond_gg1r")
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=1, seed=1) as eng:
        with pytest.raises(capi.BqError) as ei:  # QueueGG1R.arrivalLik raises: "Use slice-only" (queues.pyx:856-857)
            eng.sweep(1, "MH", "NONE")
        assert ei.value.code == capi.BQ_ERR_UNSUPPORTED
        with pytest.raises(capi.BqError):
            eng.mh_proposal(0, int(eng.schedule()[1][0]), numpy.array([1.0]))
        with pytest.raises(capi.BqError):
            eng.gibbs_kernel(arrv.soa()["d"])


def test_repair_sweeps_walk_an_initialised_state_into_the_support():
    """gibbs_initialize on a random-order network leaves a few observed jobs that start on arrival with others present
    (log_prob = -inf; the reference's initialisers stop there with an assertion).  GibbsEngine.repair_random_order runs
    slice sweeps with the rule softened to a penalty until every chain is inside the support, then restores the rule:
    log_prob finite, observed times untouched, ordinary sweeps stay inside the support afterwards."""
    from bayes_qnet_b200 import initialize
    g = helpers.golden("cond_gg1r")
    net = qnetu.qnet_from_text(str(g["yaml"]))
    net.parameters = list(g["theta"])
    qnetu.set_seed(1)
    sub = net.sample(150).subset_by_task(0.2)
    ini = net.gibbs_initialize(sub)
    assert len(initialize.random_order_violations(net, ini)) > 0
    soa = ini.soa()
    obs = soa["obs_d"] == 1
    with GibbsEngine(net, ini, n_chains=4, seed=4) as eng:
        assert not numpy.any(numpy.isfinite(eng.log_prob()))
        used = eng.repair_random_order(max_sweeps=600, check_every=20)
        assert 0 < used <= 600
        assert numpy.all(numpy.isfinite(eng.log_prob()))
        for c in range(4):
            assert initialize.random_order_violations(net, eng.to_arrivals(c)) == []
        eng.sweep(10, "SLICE", "BAYES")
        assert numpy.all(numpy.isfinite(eng.log_prob()))
        d = eng.state()
        assert numpy.array_equal(d[:, obs], numpy.broadcast_to(soa["d"][obs], (4, int(obs.sum()))))
        assert eng.repair_random_order() == 0
    net2, arrv2 = helpers.golden_state(helpers.golden("cond_mm3"), "s0")
    with GibbsEngine(net2, arrv2, n_chains=1, seed=1) as eng:
        with pytest.raises(capi.BqError):
            eng._L.bq_set_relaxed  # symbol exists
            capi.check(eng._L.bq_set_relaxed(eng._h, 1))
