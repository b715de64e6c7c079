"""Parity of the CUDA path, called through the C-ABI, against (i) vectors produced by the reference's own code
(tests/golden), (ii) the CPU oracle on the same seeded inputs, (iii) size-independent properties at the
benchmark's full size, (iv) long reference runs (posterior agreement within Monte-Carlo error).

Tolerances: conditional log-densities, log_prob, sufficient statistics and whole-sweep trajectories 1e-9
relative in fp64 (BASELINE.json north_star); posterior means within 4 combined Monte-Carlo standard errors."""
import numpy
import pytest

import helpers
from bayes_qnet_b200 import GibbsEngine, capi, diagnostics, estimation, qnet, qnetu

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("name", helpers.STATIC_NETS + helpers.DYNAMIC_NETS)
def test_logcond_matches_reference_vectors(name):
    """bq_logcond == GGkGibbs.inner_dfun - lik0 (qnet.pyx:1340-1357) on states the reference produced."""
    g = helpers.golden("cond_" + name)
    total = 0
    for st, pre, theta in (("s0", "s0", g["theta"]), ("s1", "s1", g["theta"]), ("s2", "s1", g["s2_theta"])):
        net, arrv = helpers.golden_state(g, pre, theta)
        with GibbsEngine(net, arrv, n_chains=2, seed=1) as eng:
            ref_lp = float(g[st + "_logprob"])
            lp = eng.log_prob()
            assert numpy.all(numpy.abs(lp - ref_lp) < TOL * max(1.0, abs(ref_lp)))
            for r, xs, ref in zip(g[st + "_rows"], g[st + "_x"], g[st + "_g"]):
                v = eng.logcond(1, int(r), xs)
                assert numpy.array_equal(numpy.isfinite(v), numpy.isfinite(ref)), "support differs at row %d" % r
                m = numpy.isfinite(ref)
                if m.any():
                    assert helpers.relerr(v[m], ref[m]).max() < TOL
                    total += int(m.sum())
    assert total > 200


@pytest.mark.parametrize("name", helpers.STATIC_NETS + helpers.DYNAMIC_NETS)
def test_sweep_trajectory_matches_oracle(name):
    """Same seed, same schedule -> the kernel and the CPU oracle walk the same path (5 sweeps, 3 chains)."""
    g = helpers.golden("cond_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=3, seed=4242, chain_offset=5) as eng:
        off, order = eng.schedule()
        eng.sweep(5, "SLICE", "NONE")
        d = eng.state()
        st = eng.stats()
        lp = eng.log_prob()
        ss = eng.suffstats()
        for c in range(3):
            oc = helpers.oracle_chain(net, arrv)
            evals = 0
            for s in range(5):
                ne, ev = oc.slice_sweep(order, eng.seed, 5 + c, s)
                evals += ev
            assert helpers.relerr(d[c], oc.d).max() < TOL
            assert abs(lp[c] - oc.log_prob()) < TOL * max(1.0, abs(oc.log_prob()))
            oss = oc.suffstats()
            assert helpers.relerr(ss[c], oss).max() < 1e-8
            full = eng.state(c, c + 1, full=True)
            assert helpers.relerr(full["s"][0], oc.s).max() < TOL and helpers.relerr(full["a"][0], oc.a).max() < TOL
            assert numpy.array_equal(full["proc"][0], oc.proc)
            assert numpy.array_equal(full["next_byq"][0], oc.next_byq)
            assert numpy.array_equal(full["prev_byq"][0], oc.prev_byq)
        assert st["nevents"] == 3 * 5 * len(order)
        assert not numpy.array_equal(d[0], d[1])  # chains differ only by their RNG counters


@pytest.mark.parametrize("batch", ["1", "0"])
@pytest.mark.parametrize("name", ["qnet3", "lng1", "mm3", "psdelay"])
def test_reverse_scan_matches_oracle(name, batch, monkeypatch):
    """slice_resample(reverse=True) (qnet.pyx:683-684): the scan order backwards, alternating with forward sweeps as
    estimation.chib_evidence does; chromatic, batch and group kernels against the oracle run over the reversed list."""
    monkeypatch.setenv("BQ_DYN_BATCH", batch)
    g = helpers.golden("cond_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=2, seed=99) as eng:
        off, order = eng.schedule()
        for s in range(4):
            eng.sweep(1, "SLICE", "NONE", reverse=(s % 2 == 0))
        d = eng.state()
        for c in range(2):
            oc = helpers.oracle_chain(net, arrv)
            for s in range(4):
                oc.slice_sweep(order[::-1] if s % 2 == 0 else order, eng.seed, c, s)
            assert helpers.relerr(d[c], oc.d).max() < TOL
        fwd = GibbsEngine(net, arrv, n_chains=2, seed=99)
        fwd.sweep(1, "SLICE", "NONE")
        assert not numpy.array_equal(fwd.state(), d)
        fwd.close()
    mus, arrvl = estimation.bayes(net, arrv.duplicate(), 3, reverse=True)
    assert len(mus) == 3 and numpy.all(numpy.isfinite(mus))


def test_bayes_and_stem_updates_match_oracle():
    """Fused sufficient statistics + parameter update after every sweep (qnet.pyx:882 / :844)."""
    for name, upd in (("qnet3", "BAYES"), ("lng1", "BAYES"), ("lng1", "STEM"), ("delay", "STEM")):
        g = helpers.golden("cond_" + name)
        net, arrv = helpers.golden_state(g, "s0")
        with GibbsEngine(net, arrv, n_chains=2, seed=99) as eng:
            off, order = eng.schedule()
            eng.sweep(4, "SLICE", upd, trace=True)
            th, mw, lp = eng.traces()
            assert th.shape == (4, 2, len(net.parameters))
            oc = helpers.oracle_chain(net, arrv)
            for s in range(4):
                oc.slice_sweep(order, eng.seed, 1, s)
                ss = oc.suffstats()
                assert helpers.relerr(mw[s, 1], ss[:, 5] / ss[:, 6]).max() < 1e-8
                assert abs(lp[s, 1] - oc.log_prob()) < 1e-8 * max(1.0, abs(oc.log_prob()))
                oc.update_params(upd, eng.seed, 1, s)
                assert helpers.relerr(th[s, 1], oc.theta).max() < 1e-7, (name, upd, s, th[s, 1], oc.theta)
            assert helpers.relerr(eng.params()[1], oc.theta).max() < 1e-7
            assert helpers.relerr(eng.state()[1], oc.d).max() < 1e-7


def test_schedule_is_conflict_free_and_complete():
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 400, 0.2, seed=2)
    with GibbsEngine(net, ini, n_chains=1, seed=1) as eng:
        off, order = eng.schedule()
    oc = helpers.oracle_chain(net, ini)
    assert sorted(order) == sorted(oc.eligible())  # every eligible event exactly once (qnet.pyx:688)
    soa = ini.soa()
    pb, nb, pq, nq = soa["prev_byt"], soa["next_byt"], soa["prev_byq"], soa["next_byq"]

    def blanket(e):
        out = {pb[e], pq[e], nq[e], nb[e]}
        if nq[e] >= 0:
            out.add(pb[nq[e]])
        e1 = nb[e]
        if e1 >= 0:
            out.add(pq[e1])
            if pq[e1] >= 0:
                out.add(pb[pq[e1]])
            if nq[e1] >= 0:
                out.add(pb[nq[e1]])
        out.discard(-1)
        return out

    for c in range(len(off) - 1):
        members = set(int(x) for x in order[off[c]:off[c + 1]])
        for e in members:
            assert not (blanket(e) & members), "colour class %d is not conflict free" % c


def test_sharding_is_invisible():
    """Chains are addressed by GLOBAL id: 8 chains in one handle == 2 handles of 4 with chain_offset."""
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 120, 0.3, seed=8)
    with GibbsEngine(net, ini, n_chains=8, seed=31337) as whole:
        whole.sweep(3, "SLICE", "BAYES")
        d, th = whole.state(), whole.params()
    for k in range(2):
        with GibbsEngine(net, ini, n_chains=4, seed=31337, chain_offset=4 * k) as part:
            part.sweep(3, "SLICE", "BAYES")
            assert numpy.array_equal(part.state(), d[4 * k:4 * k + 4])
            assert numpy.array_equal(part.params(), th[4 * k:4 * k + 4])


def test_global_memory_path_equals_shared_memory_path(monkeypatch):
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 150, 0.3, seed=9)
    with GibbsEngine(net, ini, n_chains=4, seed=5) as a:
        assert a.kernel_info()["state_in_smem"]
        a.sweep(3, "SLICE", "BAYES")
        da, ta = a.state(), a.params()
    monkeypatch.setenv("BQ_FORCE_GLOBAL", "1")
    with GibbsEngine(net, ini, n_chains=4, seed=5) as b:
        assert not b.kernel_info()["state_in_smem"]
        b.sweep(3, "SLICE", "BAYES")
        assert numpy.array_equal(b.state(), da) and numpy.array_equal(b.params(), ta)


def test_edge_cases():
    # nothing hidden: a sweep is a no-op and counts nothing
    net, smp, sub, ini = helpers.synthetic(helpers.MM1, 30, 1.1, seed=3)
    with GibbsEngine(net, ini, n_chains=2, seed=1) as eng:
        d0 = eng.state()
        eng.sweep(2, "SLICE", "BAYES")
        assert numpy.array_equal(eng.state(), d0) and eng.stats()["nevents"] == 0
        assert eng.schedule()[1].size == 0
    # everything hidden, odd event count, single task
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 1, -1.0, seed=3)
    with GibbsEngine(net, ini, n_chains=3, seed=1) as eng:
        eng.sweep(10, "SLICE", "NONE")
        full = eng.state(full=True)
        assert numpy.all(full["s"] >= 0) and eng.stats()["nevents"] == 3 * 10 * 4
    # bad arguments come back as error codes, not crashes
    with GibbsEngine(net, ini, n_chains=1, seed=1) as eng:
        with pytest.raises(capi.BqError):
            eng.logcond(5, 0, [1.0])
        with pytest.raises(capi.BqError):
            eng.logcond(0, 10 ** 6, [1.0])
        with pytest.raises(capi.BqError):
            eng.params(0, 7)


def test_reference_api_in_place_semantics():
    """Qnet.slice_resample mutates arrv in place and returns [arrv] (qnet.pyx:659-730); qnet.stats() counts."""
    from bayes_qnet_b200 import qnet
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 60, 0.3, seed=12)
    d_before = ini.soa()["d"].copy()
    qnet.reset_stats()
    out, lp = net.slice_resample(ini, 4, return_lp=True)
    assert out[0] is ini and numpy.isfinite(lp)
    ini.validate()
    hidden = ini.soa()["obs_d"] == 0
    assert numpy.array_equal(ini.soa()["d"][~hidden], d_before[~hidden])
    assert numpy.mean(ini.soa()["d"][hidden] != d_before[hidden]) > 0.9
    assert qnet.stats()["nevents"] == 4 * ini.num_hidden_events()
    assert abs(net.log_prob(ini) - lp) < 1e-9 * abs(lp)
    theta = net.estimate(ini)
    ms = [float(numpy.mean(ini.soa()["s"][(ini.soa()["qid"] == q) & (ini.soa()["s"] > 0)])) for q in range(7)]
    assert numpy.allclose(theta, ms, rtol=1e-12)
    mus, arrvl = estimation.bayes(net, ini, 20)
    assert len(mus) == 20 and len(mus[0]) == 7 and arrvl[0] is ini
    ini.validate()


def test_full_size_invariants_config2():
    """BASELINE.json configs[1] at full size (5k tasks, 20% observed, 1024 chains): properties that do not need
    the oracle -- observed times untouched, every service >= 0, FIFO order kept, arrivals chained to departures,
    counters consistent, log_prob finite and equal to a host recomputation, chains distinct."""
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 5000, 0.2, seed=13)
    soa = ini.soa()
    C = 1024
    with GibbsEngine(net, ini, n_chains=C, seed=2024) as eng:
        off, order = eng.schedule()
        eng.sweep(5, "SLICE", "BAYES", trace=True)
        st = eng.stats()
        assert st["nevents"] == C * 5 * len(order) and st["n_stuck"] < 1e-4 * st["nevents"]
        assert 3.0 < st["slice_evals"] / st["nevents"] < 12.0
        lp = eng.log_prob()
        assert numpy.all(numpy.isfinite(lp))
        th = eng.params()
        assert numpy.all(th > 0) and numpy.all(numpy.std(th, axis=0) > 0)
        for lo in (0, 500, C - 4):
            full = eng.state(lo, lo + 4, full=True)
            d, a, s = full["d"], full["a"], full["s"]
            obs = soa["obs_d"] == 1
            assert numpy.array_equal(d[:, obs], numpy.broadcast_to(soa["d"][obs], (4, obs.sum())))
            assert numpy.all(s >= 0)
            has = soa["prev_byt"] >= 0
            assert numpy.array_equal(a[:, has], d[:, soa["prev_byt"][has]])
            nq = soa["next_byq"]
            hasn = nq >= 0
            assert numpy.all(d[:, nq[hasn]] >= d[:, hasn]) and numpy.all(a[:, nq[hasn]] >= a[:, hasn])
            for c in range(4):
                scale = th[lo + c][soa["qid"]]
                host_lp = float(numpy.sum(-numpy.log(scale) - s[c] / scale))
                assert abs(host_lp - lp[lo + c]) < 1e-9 * abs(host_lp)
        d_all = eng.state(0, 8)
        assert len(set(d_all[:, order[0]].tolist())) == 8


@pytest.mark.parametrize("workload", ["config3", "config4"])
def test_full_size_invariants_dynamic(workload):
    """BASELINE.json configs[2] and [3] at their full event counts (fewer chains): after sweeps of the batch kernel
    the state the device maintained incrementally -- queue order, free-time vectors, cached log services, sorted PS
    times -- must agree with a from-scratch host recomputation: observed times untouched, services >= 0, arrivals
    chained to departures, queues in arrival order, counters consistent, log_prob equal to the sum of the service
    log-densities recomputed on the host."""
    import bench
    from bayes_qnet_b200 import qnet
    wl = bench.WORKLOADS[workload]
    tasks = wl["tasks"]  # 20 000 (config 3) / 50 000 (config 4)
    assert tasks == (20000 if workload == "config3" else 50000)
    net, ini, _ = bench.make_input(tasks, yaml_text=wl["yaml"], init=wl["init"], fixture=wl["fixture"], pct=wl["pct"])
    soa = ini.soa()
    C = 64
    with GibbsEngine(net, ini, n_chains=C, seed=77) as eng:
        order = eng.sequential_order()
        assert eng.kernel_info()["threads"] > 0
        eng.sweep(3, "SLICE", "BAYES")
        st = eng.stats()
        assert st["nevents"] == C * 3 * len(order)
        assert st["n_redone"] < 0.05 * st["nevents"] and st["n_stuck"] < 1e-3 * st["nevents"]
        lp = eng.log_prob()
        th = eng.params()
        assert numpy.all(numpy.isfinite(lp)) and numpy.all(numpy.isfinite(th))  # a LogNormal mean-log may be < 0
        full = eng.state(0, 4, full=True)
        d, a, s, nq = full["d"], full["a"], full["s"], full["next_byq"]
        obs = soa["obs_d"] == 1
        assert numpy.array_equal(d[:, obs], numpy.broadcast_to(soa["d"][obs], (4, obs.sum())))
        assert numpy.all(s >= 0)
        has = soa["prev_byt"] >= 0
        assert numpy.array_equal(a[:, has], d[:, soa["prev_byt"][has]])
        qtype = numpy.array([q.code for q in net.queues])[soa["qid"]]
        qdist = numpy.array([q.service.code for q in net.queues])
        npar = numpy.where(qdist == 0, 1, 2)
        poff = numpy.concatenate(([0], numpy.cumsum(npar)[:-1]))
        for c in range(4):
            hasn = (nq[c] >= 0) & (qtype == 0)  # FIFO queues stay in arrival order
            assert numpy.all(a[c, nq[c, hasn]] >= a[c, hasn])
            host_lp = 0.0
            for q in range(len(qdist)):
                sel = soa["qid"] == q
                sq = s[c, sel]
                p0 = th[c, poff[q]]
                if qdist[q] == 0:
                    host_lp += float(numpy.sum(-numpy.log(p0) - sq / p0))
                else:  # LogNormal(mu, sd)   (distributions.pyx:266-270)
                    assert qdist[q] == 2
                    sd = th[c, poff[q] + 1]
                    ls = numpy.log(sq)
                    host_lp += float(numpy.sum(-0.5 * numpy.log(2 * numpy.pi) - numpy.log(sd) - ls
                                               - (ls - p0) ** 2 / (2 * sd * sd)))
            assert abs(host_lp - lp[c]) < 1e-9 * abs(host_lp), (host_lp, lp[c])
        assert len(set(eng.state(0, 8)[:, order[0]].tolist())) == 8
    qnet.set_initialization_type("DFN2")


@pytest.mark.parametrize("name", ["mm1", "qnet3", "lng1", "mm3", "psdelay", "gg1rm"])
def test_posterior_agrees_with_long_reference_runs(name):
    """estimation.bayes from the same initial state: posterior means of the service parameters and of the
    per-queue mean waiting time vs the reference's own long runs (tests/golden/post_*.npz), within 4 combined
    Monte-Carlo standard errors (ESS-based)."""
    g = helpers.golden("post_" + name)
    net, arrv = helpers.golden_state(dict((k, g[k]) for k in g.files if k != "theta") | {"theta": g["theta0"]}, "s0")
    ref_th, ref_mw = g["theta"], g["mean_wait"]  # [R, n, P], [R, n, Q]
    n = ref_th.shape[1]
    burn = n // 5
    C = 32
    with GibbsEngine(net, arrv, n_chains=C, seed=777) as eng:
        eng.sweep(n, "SLICE", "BAYES", trace=True)
        th, mw, lp = eng.traces()
    ref_th = numpy.swapaxes(ref_th[:, burn:], 0, 1)
    ref_mw = numpy.swapaxes(ref_mw[:, burn:], 0, 1)
    th, mw = th[burn:], mw[burn:]
    for ours, ref, what in ((th, ref_th, "theta"), (mw[:, :, 1:], ref_mw[:, :, 1:], "mean wait")):
        m1, m2 = ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1))
        se = numpy.sqrt(diagnostics.mcse(ours) ** 2 + diagnostics.mcse(ref) ** 2)
        z = numpy.abs(m1 - m2) / numpy.maximum(se, 1e-12)
        assert numpy.all(z < 4.0), "%s %s: ours %s ref %s z %s" % (name, what, m1, m2, z)
    # the sampler mixes slowly on heavily loaded, mostly hidden queues (same algorithm as the reference), so
    # R-hat is a sanity bound here, not a convergence claim; the MCSEs above already account for autocorrelation
    assert numpy.all(diagnostics.rhat(th) < 2.0)


@pytest.mark.parametrize("width", ["4", "8", "16", "32"])
@pytest.mark.parametrize("name", ["qnet3", "lng1", "mm3", "lnk", "psdelay"])
def test_dynamic_kernel_group_widths(name, width, monkeypatch):
    """The dynamic-order kernel (per-chain queue order, free-time vectors) with W = 8 / 16 / 32 lanes per chain:
    the speculative candidates of a group reproduce the sequential slice sampler, so every width walks the
    oracle's path; static-order networks are forced through the same kernel as a cross-check."""
    monkeypatch.setenv("BQ_DYN_BATCH", "0")
    monkeypatch.setenv("BQ_DYN_W", width)
    monkeypatch.setenv("BQ_FORCE_DYNAMIC", "1")
    g = helpers.golden("cond_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=5, seed=911, chain_offset=2) as eng:
        off, order = eng.schedule()
        assert len(off) == len(order) + 1  # sequential scan: one event per class
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
        assert eng.stats()["nevents"] == 5 * 4 * len(order)


def test_dynamic_set_state_and_sharding():
    """set_state / broadcast_state re-rank the queues and rebuild the free-time vectors; chains are addressed by
    global id, so shards reproduce the unsharded run."""
    g = helpers.golden("cond_mm3")
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=6, seed=5150) as whole:
        whole.sweep(3, "SLICE", "BAYES")
        d, th = whole.state(), whole.params()
        full = whole.state(full=True)
    for k in range(2):
        with GibbsEngine(net, arrv, n_chains=3, seed=5150, chain_offset=3 * k) as part:
            part.sweep(3, "SLICE", "BAYES")
            assert numpy.array_equal(part.state(), d[3 * k:3 * k + 3])
            assert numpy.array_equal(part.params(), th[3 * k:3 * k + 3])
    # feed chain 4's state to a fresh engine: same derived quantities, and it keeps sampling validly
    with GibbsEngine(net, arrv, n_chains=2, seed=1) as eng:
        eng.set_state(d[4])
        f2 = eng.state(full=True)
        for k in ("s", "a", "proc", "next_byq"):
            assert numpy.array_equal(f2[k][1], full[k][4]), k
        eng.broadcast_state(d[2])
        f3 = eng.state(full=True)
        assert numpy.array_equal(f3["s"][0], full["s"][2]) and numpy.array_equal(f3["s"][1], full["s"][2])
        eng.sweep(2, "SLICE", "NONE")
        assert numpy.all(eng.state(full=True)["s"] >= 0) and numpy.all(numpy.isfinite(eng.log_prob()))


def test_per_chain_states_are_staged_in_slices(monkeypatch):
    """set_state with one state per chain stages the derived arrays slice by slice on the host; the result does not
    depend on the slice size (BQ_STAGE_MB=0: one chain per slice)."""
    g = helpers.golden("cond_mm3")
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=5, seed=9) as eng:
        eng.sweep(2, "SLICE", "BAYES")
        d, th = eng.state(), eng.params()
    outs = []
    for mb in (None, "0"):
        if mb is None:
            monkeypatch.delenv("BQ_STAGE_MB", raising=False)
        else:
            monkeypatch.setenv("BQ_STAGE_MB", mb)
        with GibbsEngine(net, arrv, n_chains=5, seed=9) as eng:
            eng.set_state(d)
            eng.set_params(th)
            full = eng.state(full=True)
            eng.sweep(2, "SLICE", "BAYES")
            outs.append((full["s"], full["proc"], full["next_byq"], eng.state(), eng.params()))
    for a, b in zip(outs[0], outs[1]):
        assert numpy.array_equal(a, b)


@pytest.mark.parametrize("name", sorted(helpers.MIXED))
def test_mixed_disciplines_match_oracle(name):
    """PS with LogNormal / Gamma service, PS -> PS, multi-server FIFO next to PS, delay stations in parallel with
    FIFO queues: conditional, sweeps with STEM updates, and the derived state against the oracle."""
    net, smp, sub, ini = helpers.mixed(name)
    with GibbsEngine(net, ini, n_chains=3, seed=2718) as eng:
        off, order = eng.schedule()
        oc = helpers.oracle_chain(net, ini)
        rs = numpy.random.RandomState(1)
        for e in rs.choice(order, size=40, replace=False):
            e1 = oc.next_byt[e]
            hi = oc.d[e1] if e1 >= 0 else oc.d[e] + 3.0
            xs = rs.uniform(oc.a[e] - 0.2, hi + 0.2, size=8)
            v, r = eng.logcond(2, int(e), xs), oc.logcond(int(e), xs)
            assert numpy.array_equal(numpy.isfinite(v), numpy.isfinite(r)), (name, e)
            m = numpy.isfinite(r)
            if m.any():
                assert helpers.relerr(v[m], r[m]).max() < TOL, (name, e, v, r)
        eng.sweep(4, "SLICE", "STEM", trace=True)
        th = eng.traces()[0]
        for s in range(4):
            oc.slice_sweep(order, eng.seed, 2, s)
            oc.update_params("STEM", eng.seed, 2, s)
            assert helpers.relerr(th[s, 2], oc.theta).max() < 1e-7, (name, s, th[s, 2], oc.theta)
        full = eng.state(2, 3, full=True)
        assert helpers.relerr(full["d"][0], oc.d).max() < 1e-7
        assert helpers.relerr(full["s"][0], oc.s).max() < 1e-7
        assert numpy.array_equal(full["proc"][0], oc.proc) and numpy.array_equal(full["next_byq"][0], oc.next_byq)


@pytest.mark.parametrize("impl", ["native", "dynamic"])
@pytest.mark.parametrize("name", helpers.MH_NETS)
def test_mh_proposal_and_sweeps(name, impl, monkeypatch):
    """MH-within-Gibbs (Qnet.gibbs_resample, qnet.pyx:337): bq_mh_proposal against the reference's Pwfun values
    (tests/golden/mh_*.npz); MH sweeps -- two slice steps on the proposal, exact-conditional accept ratio, commit --
    against the oracle, interleaved with slice sweeps (static-order handles switch kernels in between)."""
    if impl == "dynamic":  # static-order networks: the MH update of the sequential dynamic-order kernel instead of
        if name in helpers.DYNAMIC_NETS:  # the chromatic kernel's (two implementations of one algorithm)
            pytest.skip("dynamic-order networks have one implementation")
        monkeypatch.setenv("BQ_MH_DYNAMIC", "1")
    g = helpers.golden("mh_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=4, seed=606, chain_offset=1) as eng:
        for r, xs, ref in zip(g["rows"], g["x"], g["h"]):
            v = eng.mh_proposal(3, int(r), xs)
            assert numpy.array_equal(numpy.isfinite(v), numpy.isfinite(ref)), "support differs at row %d" % r
            m = numpy.isfinite(ref)
            assert helpers.relerr(v[m], ref[m]).max() < TOL
        seq = eng.mh_order()
        off, order = eng.schedule()
        assert sorted(seq) == sorted(order) == sorted(eng.sequential_order())
        ocs = [helpers.oracle_chain(net, arrv) for _ in range(2)]
        chains = (0, 3)
        props = rejects = 0
        sweep = 0
        for mode, n in (("MH", 3), ("SLICE", 2), ("MH", 2)):
            eng.sweep(n, mode, "NONE")
            for s in range(n):
                for oc, c in zip(ocs, chains):
                    if mode == "MH":
                        a = oc.mh_sweep(seq, eng.seed, 1 + c, sweep)
                        props, rejects = props + a[0], rejects + a[1]
                    else:
                        oc.slice_sweep(order, eng.seed, 1 + c, sweep)
                sweep += 1
        d = eng.state()
        for oc, c in zip(ocs, chains):
            assert helpers.relerr(d[c], oc.d).max() < 1e-7
        st = eng.stats()
        assert st["nreject"] > 0 and st["nevents"] > st["nreject"]
        lp = eng.log_prob()
        assert abs(lp[3] - ocs[1].log_prob()) < 1e-7 * max(1.0, abs(ocs[1].log_prob()))


def test_mh_counts_and_group_widths(monkeypatch):
    """nevents / nreject of one chain equal the oracle's; every group width walks the same path."""
    g = helpers.golden("mh_mm3")
    net, arrv = helpers.golden_state(g, "s0")
    ref = None
    monkeypatch.setenv("BQ_DYN_BATCH", "0")
    for width in ("4", "8", "16", "32"):
        monkeypatch.setenv("BQ_DYN_W", width)
        with GibbsEngine(net, arrv, n_chains=1, seed=31) as eng:
            seq = eng.mh_order()
            eng.sweep(5, "MH", "BAYES")
            st, d, th = eng.stats(), eng.state(), eng.params()
        if ref is None:
            oc = helpers.oracle_chain(net, arrv)
            props = rejects = 0
            for s in range(5):
                a = oc.mh_sweep(seq, 31, 0, s)
                props, rejects = props + a[0], rejects + a[1]
                oc.update_params("BAYES", 31, 0, s)
            assert (st["nevents"], st["nreject"]) == (props, rejects)
            assert helpers.relerr(d[0], oc.d).max() < 1e-7 and helpers.relerr(th[0], oc.theta).max() < 1e-7
            ref = (d, th)
        else:  # the sufficient statistics are summed in a width-dependent order: equal up to rounding
            assert helpers.relerr(d, ref[0]).max() < 1e-9 and helpers.relerr(th, ref[1]).max() < 1e-9


def test_mh_refused_for_ps():
    g = helpers.golden("cond_psdelay")
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=1, seed=1) as eng:
        with pytest.raises(capi.BqError):
            eng.sweep(1, "MH", "NONE")
        with pytest.raises(capi.BqError):
            eng.mh_proposal(0, 0, [1.0])


def test_reference_api_gibbs_resample():
    """Qnet.gibbs_resample(arrv, burn, num_iter) mutates arrv, returns the per-iteration copies, counts rejects;
    estimation.bayes honours set_sampler("MH") (estimation.py:20-25, 131-132)."""
    from bayes_qnet_b200 import qnet
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 60, 0.3, seed=12)
    qnet.reset_stats()
    out = net.gibbs_resample(ini, 2, 3)
    assert len(out) == 3
    ini.validate()
    st = qnet.stats()
    assert 0 < st["nreject"] < st["nevents"] <= 5 * ini.num_hidden_events()
    estimation.set_sampler("MH")
    try:
        mus, arrvl = estimation.bayes(net, ini, 15)
    finally:
        estimation.set_sampler("SLICE")
    assert len(mus) == 15 and estimation.last_run().stats["nreject"] > 0
    ini.validate()


@pytest.mark.parametrize("name", ["mm1", "qnet3", "mm3"])
def test_mh_posterior_agrees_with_long_reference_mh_runs(name):
    """estimation.bayes with sampling_type MH from the same initial state vs the reference's own long MH runs
    (tests/golden/postmh_*.npz), within 4 combined Monte-Carlo standard errors."""
    g = helpers.golden("postmh_" + name)
    net, arrv = helpers.golden_state(dict((k, g[k]) for k in g.files if k != "theta") | {"theta": g["theta0"]}, "s0")
    ref_th, ref_mw = g["theta"], g["mean_wait"]
    n = ref_th.shape[1]
    burn = n // 5
    with GibbsEngine(net, arrv, n_chains=32, seed=4711) as eng:
        eng.sweep(n, "MH", "BAYES", trace=True)
        th, mw, lp = eng.traces()
    ref_th = numpy.swapaxes(ref_th[:, burn:], 0, 1)
    ref_mw = numpy.swapaxes(ref_mw[:, burn:], 0, 1)
    th, mw = th[burn:], mw[burn:]
    for ours, ref, what in ((th, ref_th, "theta"), (mw[:, :, 1:], ref_mw[:, :, 1:], "mean wait")):
        m1, m2 = ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1))
        se = numpy.sqrt(diagnostics.mcse(ours) ** 2 + diagnostics.mcse(ref) ** 2)
        z = numpy.abs(m1 - m2) / numpy.maximum(se, 1e-12)
        assert numpy.all(z < 4.0), "%s %s: ours %s ref %s z %s" % (name, what, m1, m2, z)


@pytest.mark.parametrize("mode", ["SLICE", "MH"])
@pytest.mark.parametrize("name", ["qnet3", "lng1", "mm3", "lnk", "psdelay"])
def test_batch_kernel_is_the_sequential_sweep(name, mode, monkeypatch):
    """The batch kernel (a warp per chain, 32 events of the chain updated at once, conflicting updates redone) must
    give exactly the sequential sweep in its fixed order.  On these small networks the events of a batch are close
    together, so a large share of the updates conflicts and is redone -- the path under test."""
    if mode == "MH" and name == "psdelay":
        pytest.skip("PS has no MH proposal")
    monkeypatch.setenv("BQ_FORCE_DYNAMIC", "1")
    monkeypatch.setenv("BQ_MH_DYNAMIC", "1")
    g = helpers.golden("cond_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=3, seed=8086, chain_offset=4) as eng:
        order = eng.mh_order() if mode == "MH" else eng.schedule()[1]
        assert numpy.array_equal(order, eng.sequential_order())
        eng.sweep(4, mode, "BAYES", trace=True)
        d, th, st = eng.state(), eng.params(), eng.stats()
        assert st["n_redone"] > 0
        for c in (0, 2):
            oc = helpers.oracle_chain(net, arrv)
            props = 0
            for s in range(4):
                if mode == "MH":
                    props += oc.mh_sweep(order, eng.seed, 4 + c, s)[0]
                else:
                    props += oc.slice_sweep(order, eng.seed, 4 + c, s)[0]
                oc.update_params("BAYES", eng.seed, 4 + c, s)
            assert helpers.relerr(d[c], oc.d).max() < 1e-7
            assert helpers.relerr(th[c], oc.theta).max() < 1e-7
        full = eng.state(2, 3, full=True)
        assert helpers.relerr(full["s"][0], oc.s).max() < 1e-7 and numpy.array_equal(full["proc"][0], oc.proc)


def test_gibbs_front_end(tmp_path, monkeypatch, capsys):
    """The gibbs.py front end (gibbs.py:41-248): same options, same output files -- parameter series, per-iteration
    statistics, arrivals snapshots, LOG_PROB / AIC / BIC -- here with 8 chains on the GPU."""
    from bayes_qnet_b200 import gibbs
    monkeypatch.chdir(tmp_path)
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 60, 0.3, seed=12)
    with open("net.yml", "w") as f:
        f.write(helpers.QNET3)
    qnetu.write_multif_to_prefix("obs_", sub, obs_only=False)
    rc = gibbs.main(["--network", "net.yml", "--multiarr", "obs_", "--bayes", "1", "--gibbs-iter", "12", "--seed", "5",
                     "--chains", "8", "--output-mu", "mu.txt", "--statistics", "mean_wait, mean_service",
                     "--output-arrv", "arrv_", "--arrv-iter", "5", "--stats-iter", "3"])
    assert rc == 0
    out = capsys.readouterr().out
    assert "LOG_PROB" in out and "AIC" in out and "BIC" in out and "RHAT" in out
    mu = numpy.loadtxt("mu.txt")
    assert mu.shape == (12, 7) and numpy.all(mu > 0)
    tw = numpy.loadtxt("train_mean_wait.txt")
    assert tw.shape == (12, 8) and numpy.array_equal(tw[:, 0], numpy.arange(12))
    assert numpy.loadtxt("mean_service.txt").shape == (3, 7)
    for name in ("initial.txt", "arrv_0.txt", "arrv_5.txt", "arrv_10.txt"):
        with open(name) as f:
            back = qnetu.read_csv_from_lines(net, f.readlines())
        assert back.num_events() == ini.num_events()
    # StEM through the same front end, MH sampler, single chain
    rc = gibbs.main(["-n", "net.yml", "-A", "obs_", "--burn", "3", "--gibbs-iter", "5", "--thin", "2", "--sampler", "MH",
                     "--output-mu", "mu2.txt"])
    estimation.set_sampler("SLICE")
    assert rc == 0 and numpy.loadtxt("mu2.txt").shape == (5, 7)


@pytest.mark.parametrize("name", ["qnet3", "lng1", "mm3", "lnk", "delay_gg1_gg4", "gg1r"])
def test_incremental_conditional_equals_log_prob_difference(name):
    """The identity the reference checks in test_likdelta.py:43-75: for hidden events and several offsets, the
    incrementally evaluated conditional inner_dfun(d1) - lik0 equals log_prob after physically moving the departure
    (queues re-ranked by arrival, services recomputed from scratch) minus log_prob before.  Here at a few hundred
    tasks, on static- and dynamic-order networks, without the oracle.  (Not on PS networks: there the reference's
    sampling target carries log N terms that its log_prob lacks)"""
    if name in helpers.MIXED:
        net, smp, sub, ini = helpers.mixed(name, ntasks=300, pct=0.2, seed=33)
    else:
        g = helpers.golden("cond_" + name)
        net = qnetu.qnet_from_text(str(g["yaml"]))
        net.parameters = list(g["theta"])
        qnetu.set_seed(44)
        ini = net.sample(300).subset_by_task(0.2)
        if name != "gg1r":  # random-order networks start from the sampled state, as in test_likdelta.py:43-48
            ini = net.gibbs_initialize(ini)
    rs = numpy.random.RandomState(9)
    # only feasible moves are compared: where the conditional is -inf (a service would go negative in the current
    # order) "the state after the move" is not defined -- re-ranking the queue from scratch can make it feasible again
    with GibbsEngine(net, ini, n_chains=2, seed=3) as eng:
        eng.sweep(3, "SLICE", "NONE")       # move away from the initialiser's state
        base = eng.state(0, 1, full=True)
        d0 = base["d"][0].copy()
        eng.set_state(d0, 1, 2)              # chain 1 = scratch copy of chain 0
        lp0 = eng.log_prob()[1]
        order = eng.schedule()[1]
        soa = ini.soa()
        checked = 0
        for e in rs.choice(order, size=40, replace=False):
            e1 = soa["next_byt"][e]
            lo, hi = base["a"][0][e], (d0[e1] if e1 >= 0 else d0[e] + 2.0)
            xs = numpy.concatenate((d0[e] + numpy.array([-0.3, 0.3]) * min(d0[e] - lo, hi - d0[e]), rs.uniform(lo, hi, 3)))
            inc = eng.logcond(0, int(e), xs)
            for x, v in zip(xs, inc):
                d1 = d0.copy()
                d1[e] = x
                eng.set_state(d1, 1, 2)
                lp1 = eng.log_prob()[1]
                if numpy.isfinite(v):
                    assert abs((lp1 - lp0) - v) < 1e-8 * max(1.0, abs(lp0)), (name, e, x, v, lp1 - lp0)
                    checked += 1
        assert checked > 60


def test_disperse_spreads_the_chains():
    """GibbsEngine.disperse: burn-in under per-chain perturbed parameters, parameters restored afterwards; the
    chains' states end up more spread than after the same number of ordinary sweeps, and stay valid."""
    net, smp, sub, ini = helpers.synthetic(helpers.QNET3, 200, 0.2, seed=4)
    hidden = ini.soa()["obs_d"] == 0
    with GibbsEngine(net, ini, n_chains=64, seed=9) as a, GibbsEngine(net, ini, n_chains=64, seed=9) as b:
        th0 = a.params().copy()
        a.disperse(sweeps=15, scale=0.7, seed=1)
        b.sweep(15, "SLICE", "NONE")
        assert numpy.array_equal(a.params(), th0)
        sa, sb = a.state(full=True), b.state(full=True)
        assert numpy.all(sa["s"] >= 0) and numpy.all(numpy.isfinite(a.log_prob()))
        spread_a = numpy.std(sa["d"][:, hidden], axis=0).mean()
        spread_b = numpy.std(sb["d"][:, hidden], axis=0).mean()
        assert spread_a > 1.2 * spread_b


@pytest.mark.parametrize("batch", ["1", "0"])
def test_dynamic_edge_cases(batch, monkeypatch):
    """Dynamic-order kernels on degenerate inputs: nothing hidden, everything hidden, a single task, odd chain counts,
    too many processors -- for a multi-server network and a PS network."""
    monkeypatch.setenv("BQ_DYN_BATCH", batch)
    for name, init in (("mm3", "DFN2"), ("psdelay", "PS")):
        g = helpers.golden("cond_" + name)
        yaml_text = str(g["yaml"])
        old = qnet.get_initializer()
        qnet.set_initialization_type(init)
        try:
            # nothing hidden: no eligible event, sweeps are no-ops, reporting still works
            net, smp, sub, ini = helpers.synthetic(yaml_text, 30, 2.0, seed=3, theta=g["theta"])
            with GibbsEngine(net, ini, n_chains=3, seed=1) as eng:
                assert len(eng.schedule()[1]) == 0
                d0 = eng.state()
                eng.sweep(2, "SLICE", "BAYES", trace=True)
                assert numpy.array_equal(eng.state(), d0) and eng.stats()["nevents"] == 0
                assert numpy.all(numpy.isfinite(eng.log_prob()))
            # everything hidden, odd chain count
            net, smp, sub, ini = helpers.synthetic(yaml_text, 30, -1.0, seed=3, theta=g["theta"])
            with GibbsEngine(net, ini, n_chains=5, seed=1) as eng:
                n = len(eng.schedule()[1])
                assert n == ini.num_events()
                eng.sweep(3, "SLICE", "NONE")
                full = eng.state(full=True)
                assert numpy.all(full["s"] >= 0) and eng.stats()["nevents"] == 5 * 3 * n
                oc = helpers.oracle_chain(net, ini)
                for s in range(3):
                    oc.slice_sweep(eng.schedule()[1], eng.seed, 4, s)
                assert helpers.relerr(full["d"][4], oc.d).max() < 1e-7
            # a single hidden task
            net, smp, sub, ini = helpers.synthetic(yaml_text, 1, -1.0, seed=3, theta=g["theta"])
            with GibbsEngine(net, ini, n_chains=2, seed=1) as eng:
                eng.sweep(5, "SLICE", "NONE")
                assert numpy.all(eng.state(full=True)["s"] >= 0)
        finally:
            qnet.set_initialization_type(old)
    wide = str(helpers.golden("cond_mm3")["yaml"]).replace("processors: 3", "processors: 9")
    net, smp, sub, ini = helpers.synthetic(wide, 200, 0.5, seed=1)  # 9 servers, ~100 events per queue
    with pytest.raises(capi.BqError):
        GibbsEngine(net, ini, n_chains=1, seed=1)


@pytest.mark.parametrize("name", ["qnet3", "lng1", "mm3", "lnk"])
def test_checkpoint_resume_is_exact(name, tmp_path):
    """GibbsEngine.save / load: 4 sweeps, checkpoint, 4 more sweeps in a NEW engine == 8 sweeps in one go, bit for bit
    (departure times, parameters, traces) -- for static- and dynamic-order networks, slice and MH."""
    g = helpers.golden("cond_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    for mode in ("SLICE", "MH"):
        with GibbsEngine(net, arrv, n_chains=3, seed=555, chain_offset=2) as whole:
            whole.sweep(8, mode, "BAYES", trace=True)
            d8, th8 = whole.state(), whole.params()
            tr8 = whole.traces()[0]
        ck = str(tmp_path / ("ck_%s_%s.npz" % (name, mode)))
        with GibbsEngine(net, arrv, n_chains=3, seed=555, chain_offset=2) as first:
            first.sweep(4, mode, "BAYES")
            first.save(ck)
        with GibbsEngine(net, arrv, n_chains=3, seed=555, chain_offset=2) as second:
            second.load(ck)
            second.sweep(4, mode, "BAYES", trace=True)
            assert numpy.array_equal(second.state(), d8) and numpy.array_equal(second.params(), th8)
            assert numpy.array_equal(second.traces()[0], tr8[4:])
        with GibbsEngine(net, arrv, n_chains=3, seed=556, chain_offset=2) as other:
            with pytest.raises(ValueError):
                other.load(ck)


@pytest.mark.parametrize("name", ["mm3", "lnk", "psdelay", "gg2_to_ps_to_ps", "delay_gg1_gg4"])
def test_two_warps_per_chain_is_the_sequential_sweep(name, monkeypatch):
    """sweep_batch_kernel_nw (rounds of 64 events shared by two warps, named barriers, first conflict found through
    shared memory), forced onto small networks where most of a round conflicts and is redone: the result must still
    be the oracle's sequential sweep in the handle's order, with parameter updates, for every chain -- also when the
    number of chains is odd (the second chain slot of the last CTA is empty)."""
    monkeypatch.setenv("BQ_BATCH_NW", "2")
    if name in helpers.MIXED:
        net, smp, sub, arrv = helpers.mixed(name)
    else:
        net, arrv = helpers.golden_state(helpers.golden("cond_" + name), "s0")
    with GibbsEngine(net, arrv, n_chains=5, seed=321, chain_offset=1) as eng:
        off, order = eng.schedule()
        eng.sweep(4, "SLICE", "BAYES", trace=True)
        d = eng.state()
        tr_th = eng.traces()[0]
        st = eng.stats()
        assert st["nevents"] == 5 * 4 * len(order) and st["n_redone"] > 0
        for c in (0, 3, 4):
            oc = helpers.oracle_chain(net, arrv)
            for s in range(4):
                oc.slice_sweep(order, eng.seed, 1 + c, s)
                oc.update_params("BAYES", eng.seed, 1 + c, s)
                assert helpers.relerr(tr_th[s, c], oc.theta).max() < 1e-7
            assert helpers.relerr(d[c], oc.d).max() < 1e-7


@pytest.mark.parametrize("name", ["lnk", "psdelay"])
def test_long_calls_split_into_several_launches(name, monkeypatch):
    """A call with more sweeps than one launch of the batch kernels may hold (round stamps are ints) runs as several
    launches: with Tmax taken every sweep (what estimation.bayes does) the result -- states, parameters, traces,
    counters -- is that of the single launch."""
    g = helpers.golden("cond_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    outs = []
    for hook in (None, "3"):
        if hook:
            monkeypatch.setenv("BQ_MAX_SWEEPS_PER_LAUNCH", hook)
        with GibbsEngine(net, arrv, n_chains=3, seed=17) as eng:
            eng.sweep(8, "SLICE", "BAYES", trace=True)      # 3 + 3 + 2 with the hook
            outs.append((eng.state(), eng.params(), eng.traces(), eng.stats()["nevents"], eng.stats()["n_launches"]))
    assert numpy.array_equal(outs[0][0], outs[1][0]) and numpy.array_equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][2], outs[1][2]):
        assert numpy.array_equal(a, b)
    assert outs[0][3] == outs[1][3] and outs[1][4] == outs[0][4] + 2
