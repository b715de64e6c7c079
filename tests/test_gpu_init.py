"""Per-chain starting states on the device (bq_init_chains): validate()-style invariants of every
chain's state, distinct chains, sweeps that continue from them exactly like the oracle, and posterior agreement with the
reference's long runs when the ensemble starts from them."""
import numpy
import pytest

import helpers
from bayes_qnet_b200 import GibbsEngine, capi, diagnostics, qnetu

pytestmark = pytest.mark.gpu


def _check_state(net, soa, full, d_init):
    d, a, s = full["d"], full["a"], full["s"]
    n = d.shape[0]
    elig = (soa["obs_d"] == 0) & ((soa["next_byt"] < 0) | (soa["obs_a"][numpy.maximum(soa["next_byt"], 0)] == 0))
    assert numpy.array_equal(d[:, ~elig], numpy.broadcast_to(d_init[~elig], (n, int((~elig).sum()))))  # fixed times untouched
    assert numpy.all(s >= 0)
    has = soa["prev_byt"] >= 0
    assert numpy.array_equal(a[:, has], d[:, soa["prev_byt"][has]])
    qtype = numpy.array([q.code for q in net.queues])[soa["qid"]]
    nq = soa["next_byq"]
    fifo = (nq >= 0) & (qtype == 0)
    assert numpy.all(a[:, nq[fifo]] >= a[:, fifo])     # the handle's queue order is still the arrival order
    assert numpy.all(d[:, nq[fifo]] >= d[:, fifo])     # single-server FIFO: departures in the same order
    return elig


@pytest.mark.parametrize("name", ["qnet3", "delay", "lng1", "mm1"])
def test_every_chain_gets_a_valid_start_of_its_own(name):
    g = helpers.golden("cond_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    soa = arrv.soa()
    with GibbsEngine(net, arrv, n_chains=48, seed=5, chain_offset=7) as eng:
        eng.init_chains(0.7)
        full = eng.state(full=True)
        elig = _check_state(net, soa, full, soa["d"])
        moved = numpy.mean(full["d"][:, elig] != soa["d"][elig])
        assert moved > 0.99
        assert len({tuple(row) for row in numpy.round(full["d"][:, elig][:, :8], 12)}) == 48   # all chains differ
        assert numpy.all(numpy.isfinite(eng.log_prob()))
        # a second call draws new states; sigma = 0 is allowed
        eng.init_chains(0.0)
        full2 = eng.state(full=True)
        _check_state(net, soa, full2, soa["d"])
        assert not numpy.array_equal(full2["d"], full["d"])
        # sweeps continue from such a state exactly as the oracle does
        c = 5
        start = eng.to_arrivals(c)
        oc = helpers.oracle_chain(net, start)
        off, order = eng.schedule()
        eng.sweep(2, "SLICE", "NONE")
        for sw in range(2):
            oc.slice_sweep(order, eng.seed, 7 + c, sw)
        assert helpers.relerr(eng.state(c, c + 1)[0], oc.d).max() < 1e-9
        # shard invariance: the same global chain id gives the same start
    with GibbsEngine(net, arrv, n_chains=4, seed=5, chain_offset=7 + 20) as part:
        part.init_chains(0.7)
        with GibbsEngine(net, arrv, n_chains=48, seed=5, chain_offset=7) as whole:
            whole.init_chains(0.7)
            assert numpy.array_equal(part.state(), whole.state(20, 24))


def test_posterior_from_per_chain_starts_agrees_with_reference_runs():
    g = helpers.golden("post_qnet3")
    net, arrv = helpers.golden_state(dict((k, g[k]) for k in g.files if k != "theta") | {"theta": g["theta0"]}, "s0")
    ref_th, ref_mw = g["theta"], g["mean_wait"]
    n = ref_th.shape[1]
    with GibbsEngine(net, arrv, n_chains=32, seed=11) as eng:
        eng.sweep(0, "SLICE", "NONE")
        from bayes_qnet_b200 import qnet as _q
        _q.capi_update_only(eng, "STEM")          # the ML start of estimation.bayes (estimation.py:124)
        eng.init_chains(0.7)
        eng.sweep(300, "SLICE", "BAYES")
        eng.sweep(1200, "SLICE", "BAYES", trace=True)
        th, mw, _ = eng.traces()
    burn = n // 5
    ours = numpy.concatenate((th, mw[:, :, 1:]), axis=2)
    ref = numpy.concatenate((ref_th[:, burn:], ref_mw[:, burn:, 1:]), axis=2).transpose(1, 0, 2)
    se = numpy.sqrt(diagnostics.mcse(ours) ** 2 + diagnostics.mcse(ref) ** 2)
    z = numpy.abs(ours.mean(axis=(0, 1)) - ref.mean(axis=(0, 1))) / se
    assert numpy.all(z < 4.0), z
    # most summaries have forgotten their starts by now; the source queue's scale has not (the hidden tasks after the last
    # observed one are only bounded below, and a stretched tail contracts diffusively under single-site updates -- in
    # the reference's sampler as here), which the Monte-Carlo error above accounts for through the between-chain term
    assert numpy.nanmedian(diagnostics.rhat(ours)) < 1.2


def test_dynamic_order_networks_refuse():
    g = helpers.golden("cond_mm3")
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=2, seed=1) as eng:
        with pytest.raises(capi.BqError) as ei:
            eng.init_chains(0.5)
        assert ei.value.code == capi.BQ_ERR_UNSUPPORTED


@pytest.mark.parametrize("name", ["lnk", "psdelay", "delay_gg1_gg4"])
def test_host_starts_for_dynamic_order_networks(name):
    """GibbsEngine.init_chains_host: independent runs of the host initialiser as per-chain starts of a multi-server /
    processor-sharing network (where the device initialiser does not apply): observed times kept, every start valid
    (finite log_prob, s >= 0), starts differ in times and in queue order, chain c uses start c mod n_starts, sweeps go
    on from there and the chains of one start separate."""
    if name in helpers.MIXED:
        net, smp, sub, ini = helpers.mixed(name, ntasks=120, pct=0.25, seed=7)
    else:
        g = helpers.golden("cond_" + name)
        net = qnetu.qnet_from_text(str(g["yaml"]))
        net.parameters = list(g["theta"])
        from bayes_qnet_b200 import qnet as _q
        old = _q.get_initializer()
        _q.set_initialization_type("PS" if name == "psdelay" else "DFN2")
        try:
            qnetu.set_seed(3)
            ini = net.gibbs_initialize(net.sample(120).subset_by_task(0.25))
        finally:
            _q.set_initialization_type(old)
    soa = ini.soa()
    obs = soa["obs_d"] == 1
    with GibbsEngine(net, ini, n_chains=12, seed=5) as eng:
        starts = eng.init_chains_host(n_starts=4, seed=11)
        assert starts.shape == (4, eng.E)
        d = eng.state()
        for c in range(12):
            assert numpy.array_equal(d[c], starts[c % 4])
            assert numpy.array_equal(d[c][obs], soa["d"][obs])
        assert len(set(map(bytes, starts))) == 4
        assert numpy.all(numpy.isfinite(eng.log_prob()))
        full = eng.state(full=True)
        assert numpy.all(full["s"] >= 0)
        if name != "psdelay":  # PS queues report their initial links, as in the reference (deviation 4)
            assert any(not numpy.array_equal(full["next_byq"][0], full["next_byq"][c]) for c in (1, 2, 3))
        eng.sweep(5, "SLICE", "BAYES")
        d2 = eng.state()
        assert numpy.all(numpy.isfinite(eng.log_prob()))
        assert not numpy.array_equal(d2[0], d2[4])      # same start, different RNG streams
        assert numpy.array_equal(d2[:, obs], numpy.broadcast_to(soa["d"][obs], (12, int(obs.sum()))))
    net2, smp2, sub2, ini2 = helpers.synthetic(helpers.QNET3, 40, 0.3, seed=5)
    with GibbsEngine(net2, ini2, n_chains=2, seed=1) as eng:
        with pytest.raises(ValueError):
            eng.init_chains_host(2)
