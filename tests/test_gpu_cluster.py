"""The chromatic sweep on thread-block clusters (bq_cluster.cuh): a chain spread over 2 / 4 / 8 CTAs with its state in
distributed shared memory walks exactly the path of the one-CTA kernel (same schedule, same RNG streams) -- bit for
bit while the parameters are fixed, to rounding of the summation order once they are updated -- and the path of the
CPU oracle; the automatic choice (few chains, or a chain too large for one CTA's shared memory) is what it claims."""
import numpy
import pytest

import helpers
from bayes_qnet_b200 import GibbsEngine

pytestmark = pytest.mark.gpu


def _run(net, arrv, monkeypatch, ncta, mode="SLICE", update="NONE", sweeps=3, chains=3, reverse=False, extra_env=None):
    monkeypatch.setenv("BQ_CLUSTER", str(ncta))
    for k, v in (extra_env or {}).items():
        monkeypatch.setenv(k, v)
    with GibbsEngine(net, arrv, n_chains=chains, seed=31, chain_offset=2) as eng:
        info = eng.kernel_info()
        eng.sweep(sweeps, mode, update, trace=True, reverse=reverse)
        out = dict(info=info, d=eng.state(), th=eng.params(), lp=eng.log_prob(), tr=eng.traces(), st=eng.stats(),
                   order=eng.schedule()[1], seed=eng.seed)
    for k in (extra_env or {}):
        monkeypatch.delenv(k)
    return out


@pytest.mark.parametrize("name", ["qnet3", "lng1", "delay", "mm1"])
@pytest.mark.parametrize("ncta", [2, 4, 8])
def test_cluster_kernel_walks_the_single_cta_path(name, ncta, monkeypatch):
    g = helpers.golden("cond_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    one = _run(net, arrv, monkeypatch, 1)
    many = _run(net, arrv, monkeypatch, ncta)
    assert one["info"]["cluster"] == 1 and many["info"]["cluster"] == ncta
    assert numpy.array_equal(one["d"], many["d"])                      # fixed parameters: bit for bit
    assert one["st"]["nevents"] == many["st"]["nevents"] and one["st"]["slice_evals"] == many["st"]["slice_evals"]
    assert numpy.allclose(one["lp"], many["lp"], rtol=1e-12)
    assert numpy.allclose(one["tr"][1], many["tr"][1], rtol=1e-12, atol=1e-15)   # mean waits
    for update in ("BAYES", "STEM"):
        a = _run(net, arrv, monkeypatch, 1, update=update)
        b = _run(net, arrv, monkeypatch, ncta, update=update)
        assert helpers.relerr(b["d"], a["d"]).max() < 1e-9
        assert helpers.relerr(b["th"], a["th"]).max() < 1e-9
        assert helpers.relerr(b["tr"][0], a["tr"][0]).max() < 1e-9
    # and the oracle's path
    oc = helpers.oracle_chain(net, arrv)
    for sw in range(3):
        oc.slice_sweep(many["order"], many["seed"], 2 + 1, sw)
    assert helpers.relerr(many["d"][1], oc.d).max() < 1e-9


@pytest.mark.parametrize("ncta", [2, 8])
def test_cluster_kernel_mh_and_reverse(ncta, monkeypatch):
    g = helpers.golden("cond_qnet3")
    net, arrv = helpers.golden_state(g, "s0")
    for kw in (dict(mode="MH"), dict(reverse=True), dict(mode="MH", update="BAYES")):
        a = _run(net, arrv, monkeypatch, 1, **kw)
        b = _run(net, arrv, monkeypatch, ncta, **kw)
        if kw.get("update", "NONE") == "NONE":
            assert numpy.array_equal(a["d"], b["d"]), kw
            assert a["st"]["nreject"] == b["st"]["nreject"]
        else:
            assert helpers.relerr(b["d"], a["d"]).max() < 1e-9


def test_automatic_choice(monkeypatch):
    """Clusters only while CTAs would otherwise be idle, and only for chains with enough work per class."""
    monkeypatch.delenv("BQ_CLUSTER", raising=False)
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 1500, 0.3, seed=5)
    for chains, want in ((1, 8), (8, 8), (16, 4), (37, 4), (70, 2), (200, 1)):
        with GibbsEngine(net, ini, n_chains=chains, seed=1) as eng:
            assert eng.kernel_info()["cluster"] == want, (chains, eng.kernel_info())
            eng.sweep(2, "SLICE", "BAYES")
            assert numpy.all(eng.state(full=True)["s"] >= 0)
    small = helpers.synthetic(helpers.QNET3, 100, 0.3, seed=5)
    with GibbsEngine(small[0], small[3], n_chains=1, seed=1) as eng:
        assert eng.kernel_info()["cluster"] == 1   # a few hundred events: the cluster barrier would cost more than it saves


def test_chain_too_large_for_one_cta_stays_on_chip(monkeypatch):
    """E = 32 000 events (8 000 tasks): one CTA's shared memory holds about 27 000; with few chains the cluster path keeps
    the state on chip with two CTAs per chain and gives what the global-memory build of the one-CTA kernel gives."""
    monkeypatch.delenv("BQ_CLUSTER", raising=False)
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 8000, 0.2, seed=9)
    with GibbsEngine(net, ini, n_chains=60, seed=4) as eng:
        info = eng.kernel_info()
        assert info["cluster"] == 2 and not info["state_in_smem"]
        eng.sweep(2, "SLICE", "NONE")
        d_cluster = eng.state(0, 4)
    monkeypatch.setenv("BQ_CLUSTER", "1")
    with GibbsEngine(net, ini, n_chains=60, seed=4) as eng:
        assert eng.kernel_info()["cluster"] == 1
        eng.sweep(2, "SLICE", "NONE")
        assert numpy.array_equal(eng.state(0, 4), d_cluster)
