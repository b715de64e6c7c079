"""Host-side logic (no GPU): YAML loading, the SoA Arrivals, forward simulation, subsetting, the initialiser,
diagnostics, and the C-ABI library's exported surface."""
import io
import os
import re

import numpy
import pytest

import helpers
from bayes_qnet_b200 import capi, diagnostics, qnet, qnetu, qstats
from bayes_qnet_b200.arrivals import Arrivals, Event, ps_service

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_yaml_model_loading_matches_reference_conventions():
    net = qnetu.qnet_from_text(helpers.QNET3)
    assert net.num_queues() == 7 and net.num_states() == 4
    assert [q.name for q in net.queues] == ["INITIAL", "WEB1", "WEB2", "WEB3", "APP1", "APP2", "DB"]
    assert all(q.code == 0 and q.k == 1 for q in net.queues)  # default type is GG1 -> QueueGGk(k=1), qnetu.py:103-110
    assert net.parameters == [1.0, 0.2, 0.2, 0.2, 0.05, 0.05, 0.3]
    net.parameters = [0.25, 0.5, 0.5, 0.5, 1.2, 1.2, 0.05]  # test_qnet.py:577-579
    assert net.parameters == [0.25, 0.5, 0.5, 0.5, 1.2, 1.2, 0.05]
    assert net.fsm.possible_observations(1) == {1, 2, 3} and net.fsm.successors(2) == {3}
    assert net.queue_by_name("DB") is net.queues[6] and net.queue_by_name("nope") is None
    net2 = qnetu.qnet_from_text(net.as_yaml())  # round trip through as_yaml
    assert [q.name for q in net2.queues] == [q.name for q in net.queues] and net2.parameters == net.parameters
    mm3 = qnetu.qnet_from_text("""
states:
  - {name: I, queues: [I], successors: [A]}
  - {name: A, queues: [Q1], successors: [B]}
  - {name: B, queues: [Q2], successors: [C]}
  - {name: C, queues: [Q3]}
queues:
  - { name: I, service: [M, 10.0] }
  - { name: Q1, processors: 3, service: [LN, 1.75, 1.0] }
  - { name: Q2, type: DELAY, service: [G, 2.0, 0.5] }
  - { name: Q3, type: PS, service: [M, 0.5] }
""")
    assert [(q.code, q.k, q.service.code) for q in mm3.queues] == [(0, 1, 0), (0, 3, 2), (1, 1, 1), (2, 1, 0)]
    assert mm3.parameters == [10.0, 1.75, 1.0, 2.0, 0.5, 0.5]
    rnd = qnetu.qnet_from_text(helpers.MM1.replace("[M, 0.7]", "[M, 0.7], type: GG1R"))  # qnetu.py:115-116
    assert type(rnd.queues[1]).__name__ == "QueueGG1R" and rnd.queues[1].code == 3
    with pytest.raises(Exception):
        qnetu.qnet_from_text("- 1\n- 2\n")


def test_configuration_keys():
    qnetu.read_configuration(io.StringIO("sampling_type: SLICE\none_d_sampler: SLICE\ninitializer: DFN2\nuse_rtp: 0\n"))
    with pytest.raises(NotImplementedError):
        qnetu.read_configuration(io.StringIO("one_d_sampler: ARS_PWFUN\n"))
    with pytest.raises(Exception):
        qnetu.read_configuration(io.StringIO("bogus: 1\n"))


def test_forward_simulation_is_a_valid_fifo_network():
    net = qnetu.qnet_from_text(helpers.QNET3)
    qnetu.set_seed(3)
    arrv = net.sample(600)
    arrv.validate()
    assert arrv.num_events() == 2400 and arrv.num_tasks() == 600
    ms = qstats.mean_service(arrv)
    for m, p in zip(ms, net.parameters):
        assert abs(m - p) < 5 * p / numpy.sqrt(150)
    # every task walks INITIAL -> TIER1 -> TIER2 -> TIER3, arrivals chained to departures
    evts = arrv.events_of_task(17)
    assert [e.state for e in evts] == [0, 1, 2, 3] and evts[0].a == 0.0
    assert all(x.d == y.a for x, y in zip(evts, evts[1:]))
    # the oracle agrees with the host's service computation
    oc = helpers.oracle_chain(net, arrv)
    assert numpy.abs(oc.s - arrv.soa()["s"]).max() < 1e-12


@pytest.mark.parametrize("yaml_text", [helpers.QNET3, None, "mm3", "psdelay", "delay"])
def test_subset_and_initializer_give_a_feasible_state(yaml_text):
    if yaml_text is None:
        yaml_text = helpers.MM1
    elif yaml_text in ("mm3", "psdelay", "delay"):
        yaml_text = str(helpers.golden("cond_" + yaml_text)["yaml"])
    net, smp, sub, ini = helpers.synthetic(yaml_text, 300, 0.3, seed=11)
    soa, ssoa = ini.soa(), sub.soa()
    n_hidden_tasks = len(numpy.unique(ssoa["tid"][ssoa["obs_d"] == 0]))
    assert 0.55 * 300 < n_hidden_tasks < 0.85 * 300
    ini.validate()
    assert ini.num_tasks() == 300
    # observed events keep their data exactly
    obs = ssoa["obs_d"] == 1
    got = dict((int(e), (a, d)) for e, a, d, o in zip(soa["eid"], soa["a"], soa["d"], soa["obs_d"]) if o)
    for e, a, d in zip(ssoa["eid"][obs], ssoa["a"][obs], ssoa["d"][obs]):
        assert got[int(e)] == (a, d)
    # hidden times were re-drawn (they do not leak the simulated truth)
    hid = dict((int(e), d) for e, d, o in zip(soa["eid"], soa["d"], soa["obs_d"]) if not o)
    truth = dict((int(e), d) for e, d in zip(ssoa["eid"][~obs], ssoa["d"][~obs]))
    same = sum(1 for e in hid if e in truth and hid[e] == truth[e])
    assert same == 0
    oc = helpers.oracle_chain(net, ini)
    assert numpy.all(oc.s >= 0) and numpy.isfinite(oc.log_prob())
    assert numpy.abs(oc.s - soa["s"]).max() < 1e-9 and numpy.array_equal(oc.proc, soa["proc"])


def test_event_facade_and_detached_events():
    net = qnetu.qnet_from_text(helpers.MM1)
    arrv = Arrivals(net)
    q0, q1 = net.queues
    arrv.append(Event(0, 0, 0, q0, 0.0, 0.0, 1.0, 0, obs_a=1, obs_d=1))
    arrv.append(Event(1, 0, 1, q1, 1.0, 0.0, 1.5, 1, obs_a=1, obs_d=1))
    arrv.append(Event(2, 1, 0, q0, 0.0, 0.0, 2.0, 0, obs_a=1, obs_d=1))
    arrv.append(Event(3, 1, 1, q1, 2.0, 0.0, 4.0, 1, obs_a=1, obs_d=1))
    arrv.recompute_all_service()
    arrv.validate()
    e = arrv.event(3)
    assert (e.a, e.d, e.s, e.tid, e.qid) == (2.0, 4.0, 2.0, 1, 1) and e.q is q1 and e.queue().name == "WEB1"
    assert e.previous_by_queue().eid == 1 and e.previous_by_task().eid == 2 and e.next_by_task() is None
    assert arrv.event(2).s == 1.0 and arrv.event(2).wait() == 1.0  # source: waits for task 0
    assert [x.eid for x in arrv.events_of_qid(1)] == [1, 3] and arrv.num_hidden_events() == 0
    assert arrv.as_csv().splitlines()[0].startswith("EID TID QID A D S W")
    dup = arrv.duplicate()
    dup.event(3).d = 5.0
    assert arrv.event(3).d == 4.0
    rt = qnetu.read_table_from_string(net, arrv.as_csv())
    assert numpy.allclose(sorted(rt.soa()["d"]), sorted(arrv.soa()["d"])) and rt.num_tasks() == 2


def test_arrivals_yaml_and_multifile_formats():
    net = qnetu.qnet_from_text(helpers.MM1)
    text = """!Arrivals
events:
- !Event {arrival: 0.0, departure: 1.0, obs: 1, queue: INITIAL, state: 0, task: 0}
- !Event {arrival: 0.0, departure: 2.5, obs: 1, queue: INITIAL, state: 0, task: 1}
- !Event {arrival: 1.0, departure: 1.75, obs: 1, queue: WEB1, state: 1, task: 0}
- !Event {arrival: 2.5, departure: 3.0, obs: 1, queue: WEB1, state: 1, task: 1}
"""
    arrv = qnetu.load_arrivals(net, io.StringIO(text))
    assert arrv.num_events() == 4 and arrv.num_tasks() == 2
    assert numpy.allclose(sorted(arrv.soa()["s"]), [0.5, 0.75, 1.0, 1.5])
    sf, af, df = io.StringIO(), io.StringIO(), io.StringIO()
    qnetu.write_multifile_arrv(arrv, sf, af, df, obs_only=False)
    back = qnetu.read_multifile_arrv(net, io.StringIO(sf.getvalue()), io.StringIO(af.getvalue()), io.StringIO(df.getvalue()))
    assert numpy.allclose(sorted(back.soa()["d"]), sorted(arrv.soa()["d"])) and back.num_tasks() == 2
    back.validate()


def test_ps_service_integral():
    """Brute-force check of s = integral dt/N(t) and the test_ps.py:189-212 style zero-length case."""
    rng = numpy.random.RandomState(0)
    a = numpy.sort(rng.uniform(0, 20, 40))
    d = a + rng.exponential(1.0, 40)
    s = ps_service(a, d)
    grid = numpy.linspace(0, d.max(), 400001)
    N = (a[:, None] <= grid[None, ::]).sum(0) - (d[:, None] <= grid[None, :]).sum(0)
    for i in (0, 7, 23, 39):
        m = (grid >= a[i]) & (grid < d[i])
        assert abs(s[i] - numpy.sum(1.0 / N[m]) * (grid[1] - grid[0])) < 2e-3
    assert numpy.all(s <= d - a + 1e-12) and numpy.all(s > 0)


def test_diagnostics():
    rng = numpy.random.RandomState(1)
    x = rng.normal(size=(2000, 8, 2))
    assert numpy.all(numpy.abs(diagnostics.rhat(x) - 1.0) < 0.01)
    e = diagnostics.ess(x)
    assert numpy.all(e > 0.8 * 16000) and numpy.all(e < 1.25 * 16000)
    y = numpy.zeros((4000, 4))  # AR(1) with rho = 0.9 -> ESS ~ n (1-rho)/(1+rho)
    eps = rng.normal(size=y.shape)
    for t in range(1, len(y)):
        y[t] = 0.9 * y[t - 1] + eps[t]
    assert 0.5 * 16000 / 19 < diagnostics.ess(y[:, :, None])[0] < 2.0 * 16000 / 19
    z = x.copy()
    z[:, 0] += 3.0  # one chain somewhere else
    assert numpy.all(diagnostics.rhat(z) > 1.2)


def test_capi_exports_every_declared_symbol():
    """include/bq_capi.h is the boundary: each function it declares is exported by the library and bound by
    capi.SYMBOLS (and vice versa).  No compute call is made -- there is no GPU here."""
    header = open(os.path.join(ROOT, "include", "bq_capi.h")).read()
    declared = set(re.findall(r"\b(bq_[a-z_0-9]+)\s*\(", header))
    assert declared == set(capi.SYMBOLS), declared.symmetric_difference(capi.SYMBOLS)
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.bq_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of silently computing on the host."""
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    net, smp, sub, ini = helpers.synthetic(helpers.MM1, 20, 0.5, seed=1)
    with pytest.raises(capi.BqError) as ei:
        net.engine(ini, n_chains=2)
    assert ei.value.code == capi.BQ_ERR_NO_DEVICE
    with pytest.raises(capi.BqError):
        net.slice_resample(ini, 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bayes_qnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("Py3 oracle", ""), "%s mentions the oracle" % f


@pytest.mark.parametrize("name", ["qnet3", "mm3", "psdelay"])
def test_qstats_match_reference(name):
    """Every exported statistic of qstats.py (qstats.py:174-184) on a state the reference produced, against the values
    the reference's own qstats computed for it (tests/golden/qstats_*.npz)."""
    import helpers
    from bayes_qnet_b200 import qstats
    g = helpers.golden("qstats_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    for n in g["stat_names"]:
        ours = numpy.atleast_1d(numpy.array(getattr(qstats, str(n))(arrv), dtype=numpy.float64))
        ref = g["stat_" + str(n)]
        assert ours.shape == ref.shape, n
        m = numpy.isfinite(ref)
        assert numpy.array_equal(numpy.isfinite(ours), m), n
        assert numpy.allclose(ours[m], ref[m], rtol=1e-9, atol=1e-12), (n, ours, ref)
    assert qstats.all_stats_string().split() == [str(n) for n in g["stat_names"]][:len(qstats.STATS)]


def test_mixture_template_host_surface():
    """MixtureTemplate on the host (queues.pyx:1373-1509): the reference's own pins for mean / std
    (test_mixture_template.py:20-26), the parameter layout (mixing proportions, then each component,
    queues.pyx:1470-1486), YAML round trip, components recorded by the forward simulation."""
    text = """
states:
  - name: INITIAL
    queues: [ INITIAL ]
    successors: [ TIER1 ]
  - name: TIER1
    queues: [ Q ]
queues:
  - { name: INITIAL, service: [M, 10.0]  }
  - { name: Q, service: [ MIX, 0.75, [M, 3.0], 0.25, [LN, 0.0, 1.0] ] }
"""
    net = qnetu.qnet_from_text(text)
    q = net.queue_by_name("Q")
    assert abs(q.service.mean() - 2.662180) < 1e-5 and abs(q.service.std() - 2.813840) < 1e-5
    assert net.parameters == [10.0, 0.75, 0.25, 3.0, 0.0, 1.0]
    net.parameters = [9.0, 0.6, 0.4, 2.0, 0.1, 0.9]
    assert q.service.mixing == [0.6, 0.4] and q.service.dists[0].parameters == [2.0] and q.service.dists[1].parameters == [0.1, 0.9]
    again = qnetu.qnet_from_text(net.as_yaml())
    assert numpy.allclose(again.parameters, net.parameters, atol=1e-5)
    assert q.service.component_codes() == [0, 2] and q.is_mixture() and not net.queue_by_name("INITIAL").is_mixture()
    qnetu.set_seed(7)
    smp = net.sample(600)
    soa = smp.soa()
    frac = soa["aux"][soa["qid"] == 1].mean()
    assert abs(frac - 0.4) < 0.08 and numpy.all(soa["aux"][soa["qid"] == 0] == 0)
    ev = [e for e in smp if e.qid == 1][0]
    assert ev.variable("COMPONENT_Q") in (0, 1)
    dup = smp.duplicate()
    assert numpy.array_equal(dup.soa()["aux"], soa["aux"])
    ini = net.gibbs_initialize(smp.subset_by_task(0.5))
    isoa = ini.soa()
    obs = isoa["obs_d"] == 1
    assert set(numpy.unique(isoa["aux"][isoa["qid"] == 1])) <= {0, 1} and obs.sum() > 0


def test_pooled_sufficient_statistics_and_engine_fingerprint():
    """Qnet.estimate / sample_parameters given several Arrivals pool the data per queue (qnet.pyx:844-896): the host
    adds the per-Arrivals statistics and re-centres the squared log deviations on the pooled mean; the engine cache
    key changes with the structure an engine freezes."""
    from bayes_qnet_b200 import qnet as _q
    rs = numpy.random.RandomState(0)
    data = [numpy.exp(rs.normal(0.3, 0.8, n)) for n in (40, 75, 13)]

    def stats(x):
        ls = numpy.log(x)
        return numpy.array([[len(x), x.sum(), ls.sum(), ((ls - ls.mean()) ** 2).sum(), len(x), 0.0, len(x), 0.0]])

    pooled = _q.pool_suffstats([stats(x) for x in data])
    assert numpy.allclose(pooled, stats(numpy.concatenate(data)), rtol=1e-12)
    net, smp, sub, ini = __import__("helpers").synthetic(__import__("helpers").MM1, 30, 0.5, seed=1)
    f0 = _q._fingerprint(ini.soa())
    assert f0 == _q._fingerprint(ini.duplicate().soa())
    ev = [e for e in ini if not e.obs_d][0]
    ev.obs_d = 1
    assert _q._fingerprint(ini.soa()) != f0


def test_gg1r_loader_simulator_and_state():
    """QueueGG1R on the host (queues.pyx:788-1027): YAML `type: GG1R`, the simulator's random choice among the waiting
    jobs (select_job_for_service, :1026), by-queue links in departure order, s = d - max(a, previous departure); the
    simulated state satisfies the start-on-arrival rule, so the oracle's log_prob (which carries it) is finite."""
    import helpers
    from bayes_qnet_b200 import queues
    net = qnetu.qnet_from_text(helpers.MIXED["gg1r_ps_gg2"])
    assert [q.code for q in net.queues] == [0, 3, 1, 2, 3, 0]
    assert net.queues[1].queue_order() == queues.ARRV_BY_D and "type: GG1R" in net.as_yaml()
    net2 = qnetu.qnet_from_text(net.as_yaml())
    assert [type(q).__name__ for q in net2.queues] == [type(q).__name__ for q in net.queues]
    qnetu.set_seed(3)
    smp = net.sample(400)
    smp.validate()
    soa = smp.soa()
    inversions = 0
    for q in (1, 4):
        rows = smp._rows_of_queue(q)
        d, a, s = soa["d"][rows], soa["a"][rows], soa["s"][rows]
        assert numpy.all(numpy.diff(d) >= 0)                       # the list is the departure order
        assert numpy.all(s > 0)
        want = d - numpy.maximum(a, numpy.concatenate(([0.0], d[:-1])))
        assert numpy.abs(want - s).max() < 1e-12
        inversions += int(numpy.sum(numpy.diff(a) < 0))
    assert inversions > 0                                          # jobs do overtake each other
    oc = helpers.oracle_chain(net, smp)
    assert numpy.isfinite(oc.log_prob())
    sub = smp.subset_by_task(0.3)
    assert numpy.isfinite(helpers.oracle_chain(net, sub).log_prob())


def test_dirac_service_simulates_and_is_refused_by_the_samplers():
    """[D, loc] (qnetu.py:183-184, distributions.pyx:347-375): constant service times in Qnet.sample; no density, so the
    engine refuses the network before touching the device."""
    from bayes_qnet_b200 import GibbsEngine, distributions
    text = """
states:
  - {name: I0, queues: [I0], successors: [S1], initial: TRUE}
  - {name: S1, queues: [Q1]}
queues:
  - { name: I0, service: [M, 1.0] }
  - { name: Q1, service: [D, 0.25] }
"""
    net = qnetu.qnet_from_text(text)
    assert isinstance(net.queues[1].service.f, distributions.Dirac) and net.parameters == [1.0, 0.25]
    assert "[D, 0.25]" in net.as_yaml()
    qnetu.set_seed(2)
    smp = net.sample(50)
    smp.validate()
    soa = smp.soa()
    assert numpy.allclose(soa["s"][soa["qid"] == 1], 0.25)
    with pytest.raises(NotImplementedError):
        GibbsEngine(net, smp.subset_by_task(0.5), n_chains=1)


def test_initialiser_on_random_order_networks_adds_no_violation():
    """gibbs_initialize on a GG1R network: hidden jobs are placed so that none of them starts on arrival with other
    jobs present and none makes a job already placed do so; what remains are observed jobs whose waiting was caused by
    a hidden departure (the reference's initialisers stop there with an assertion, queues.pyx:843-853)."""
    import helpers
    from bayes_qnet_b200 import initialize
    g = helpers.golden("cond_gg1r")
    net = qnetu.qnet_from_text(str(g["yaml"]))
    qnetu.set_seed(5)
    smp = net.sample(300)
    assert initialize.random_order_violations(net, smp) == []
    sub = smp.subset_by_task(0.3)
    ini = net.gibbs_initialize(sub)
    ini.validate()
    soa = ini.soa()
    assert numpy.all(soa["s"] >= 0)
    bad = initialize.random_order_violations(net, ini)
    assert all(soa["obs_d"][r] == 1 for r in bad)
    assert len(bad) < 0.05 * len(soa["d"])
