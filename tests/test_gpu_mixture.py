"""Mixture service templates (MixtureTemplate, queues.pyx:1373-1509) on the GPU: the per-event conditional with every
event's own component density against the reference's inner_dfun vectors (tests/golden/cond_mix.npz), whole sweeps
-- component resampling, slice updates, per-component parameter updates, mixing proportions -- against the oracle,
the host surface (parameters layout, YAML round trip, simulation), and what is refused like in the reference."""
import numpy
import pytest

import helpers
from bayes_qnet_b200 import GibbsEngine, capi, estimation, qnetu

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_logcond_and_logprob_match_reference_vectors():
    g = helpers.golden("cond_mix")
    total = 0
    for st, pre, theta in (("s0", "s0", g["theta"]), ("s1", "s1", g["theta"]), ("s2", "s1", g["s2_theta"])):
        net, arrv = helpers.golden_state(g, pre, theta)
        with GibbsEngine(net, arrv, n_chains=2, seed=1) as eng:
            assert eng.has_mix and eng.NS == 5
            assert numpy.array_equal(eng.aux(1, 2)[0], g[pre + "_aux"])
            ref_lp = float(g[st + "_logprob"])
            assert numpy.all(numpy.abs(eng.log_prob() - ref_lp) < TOL * max(1.0, abs(ref_lp)))
            for r, xs, ref in zip(g[st + "_rows"], g[st + "_x"], g[st + "_g"]):
                v = eng.logcond(1, int(r), xs)
                assert numpy.array_equal(numpy.isfinite(v), numpy.isfinite(ref)), "support differs at row %d" % r
                m = numpy.isfinite(ref)
                if m.any():
                    assert helpers.relerr(v[m], ref[m]).max() < TOL
                    total += int(m.sum())
    assert total > 200


@pytest.mark.parametrize("update", ["NONE", "BAYES", "STEM"])
def test_sweeps_follow_the_oracle(update):
    """Same seed -> same path: components (resample_auxiliaries at the start of each sweep), departure times,
    parameters incl. the mixing proportions, log_prob, per-slot sufficient statistics."""
    g = helpers.golden("cond_mix")
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=3, seed=99, chain_offset=4) as eng:
        off, order = eng.schedule()
        eng.sweep(4, "SLICE", update, trace=True)
        d, z, th = eng.state(), eng.aux(), eng.params()
        lp, ss = eng.log_prob(), eng.suffstats()
        tr_th, tr_w, tr_lp = eng.traces()
        for c in (0, 2):
            oc = helpers.oracle_chain(net, arrv)
            for sw in range(4):
                oc.resample_aux(eng.seed, 4 + c, sw)
                oc.slice_sweep(order, eng.seed, 4 + c, sw)
                if update != "NONE":
                    oc.update_params(update, eng.seed, 4 + c, sw)
            assert numpy.array_equal(z[c], oc.aux)
            assert helpers.relerr(d[c], oc.d).max() < TOL
            assert helpers.relerr(th[c], oc.theta).max() < 1e-8
            assert abs(lp[c] - oc.log_prob()) < TOL * max(1.0, abs(oc.log_prob()))
            assert helpers.relerr(ss[c], oc.slot_suffstats()).max() < 1e-8
            assert helpers.relerr(tr_th[-1, c], oc.theta).max() < 1e-8
        if update != "NONE":
            for q, mo in ((1, 1), (2, 6)):  # mixing proportions stay a distribution
                assert numpy.allclose(th[:, mo] + th[:, mo + 1], 1.0)


def test_host_surface_and_reference_api():
    """Parameters = mixing proportions then component parameters (queues.pyx:1470-1474); YAML round trip; forward
    simulation records the components; estimation.bayes runs in place and returns the whole parameter vector."""
    g = helpers.golden("cond_mix")
    net = qnetu.qnet_from_text(str(g["yaml"]))
    assert numpy.allclose(net.parameters, g["theta"])
    net2 = qnetu.qnet_from_text(net.as_yaml())
    assert numpy.allclose(net2.parameters, net.parameters)
    qnetu.set_seed(3)
    smp = net.sample(400)
    soa = smp.soa()
    frac = soa["aux"][soa["qid"] == 1].mean()
    assert 0.3 < frac < 0.5           # component 1 of Q1 has weight 0.4
    assert numpy.all(soa["aux"][soa["qid"] == 0] == 0)
    sub = smp.subset_by_task(0.3)
    ini = net.gibbs_initialize(sub)
    mus, arrvl = estimation.bayes(net, ini, 30, seed=5)
    assert len(mus) == 30 and len(mus[0]) == 11 and numpy.all(numpy.isfinite(mus))
    arrvl[0].validate()
    assert abs(mus[-1][1] + mus[-1][2] - 1.0) < 1e-12


def test_what_the_reference_cannot_do_is_refused():
    g = helpers.golden("cond_mix")
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=1, seed=1) as eng:
        with pytest.raises(capi.BqError) as ei:  # MH proposals take a DistributionTemplate only (queues.pyx:394-442)
            eng.sweep(1, "MH", "NONE")
        assert ei.value.code == capi.BQ_ERR_UNSUPPORTED
        with pytest.raises(capi.BqError):        # parameter_kernel raises NotImplementedError (queues.pyx:1453)
            eng.gibbs_kernel(arrv.soa()["d"])
    mixed = str(g["yaml"]).replace("{ name: Q2, service", "{ name: Q2, type: PS, service")
    net3 = qnetu.qnet_from_text(mixed)
    from bayes_qnet_b200 import qnet as _q
    _q.set_initialization_type("PS")
    try:
        qnetu.set_seed(1)
        ini = net3.gibbs_initialize(net3.sample(30).subset_by_task(0.5))
    finally:
        _q.set_initialization_type("DFN2")
    with pytest.raises(capi.BqError) as ei:
        GibbsEngine(net3, ini, n_chains=1, seed=1)
    assert ei.value.code == capi.BQ_ERR_UNSUPPORTED
