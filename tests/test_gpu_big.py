"""GPU leg of the config-size parity (see tests/test_big_golden.py): bq_logcond through the C-ABI against the
reference's inner_dfun vectors on states of BASELINE size -- 5 000-task 3-tier network (chromatic kernel's closed
form), G/G/3 LogNormal tandem at 5 000 / 20 000 tasks, PS + delay network at 10 000 / 50 000 tasks (position-indexed
walks: galloping, flag scans, re-ranking), 1e-9 relative with identical support; then one full sweep of the batch
kernel and the group kernel, bit for bit, and of the host emulation of the same walk code (to rounding), at full size; and BASELINE
configs[0] as the reference states it (one chain, 1000 sweeps)."""
import os

import numpy
import pytest

import emul_helper
import helpers
from bayes_qnet_b200 import GibbsEngine, estimation, diagnostics

pytestmark = pytest.mark.gpu


def _load(name):
    path = os.path.join(helpers.GOLDEN, "big_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("fixture big_%s.npz not generated" % name)
    return helpers.golden("big_" + name)


@pytest.mark.parametrize("name", helpers.BIG)
def test_logcond_matches_reference_at_config_size(name):
    g = _load(name)
    total = 0
    for tag in ("s0", "s1"):
        net, arrv = helpers.big_state(g, tag)
        with GibbsEngine(net, arrv, n_chains=2, seed=1) as eng:
            lp, ref_lp = eng.log_prob(), float(g[tag + "_logprob"])
            assert numpy.all(numpy.abs(lp - ref_lp) < 1e-9 * abs(ref_lp))
            total += helpers.check_grid(lambda r, xs: eng.logcond(1, r, xs), g, tag)
    assert total > 1500


@pytest.mark.parametrize("name", ["qnet3_5k"])
def test_logcond_static_global_memory_path_at_config_size(name, monkeypatch):
    """The same vectors through the chain-state-in-global-memory build of the chromatic kernel's conditional."""
    monkeypatch.setenv("BQ_FORCE_GLOBAL", "1")
    g = _load(name)
    net, arrv = helpers.big_state(g, "s1")
    with GibbsEngine(net, arrv, n_chains=2, seed=1) as eng:
        assert not eng.kernel_info()["state_in_smem"]
        assert helpers.check_grid(lambda r, xs: eng.logcond(0, r, xs), g, "s1") > 500
        eng.sweep(1, "SLICE", "NONE")
        assert numpy.all(eng.state(0, 2, full=True)["s"] >= 0)


@pytest.mark.parametrize("name", ["lnk_20k", "psdelay_50k", "lnk_5k"])
def test_full_sweep_batch_group_emulation_identical(name, monkeypatch):
    """One full sweep of 4 chains from the reference's config-size state: the batch kernel (32 events of a chain per
    round, conflicts redone), the group kernel (W = 32 lanes on one event) walking the same order, and the host
    emulation of the same walk / commit code run sequentially give the SAME departure times: the two kernels bit for bit,
    the host run to the last-bit differences of libm vs libdevice."""
    g = _load(name)
    net, arrv = helpers.big_state(g, "s0")
    with GibbsEngine(net, arrv, n_chains=4, seed=99, chain_offset=3) as eng:
        order = eng.sequential_order()
        eng.sweep(1, "SLICE", "NONE")
        d_batch = eng.state()
        st = eng.stats()
        assert st["nevents"] == 4 * len(order)
    monkeypatch.setenv("BQ_DYN_BATCH", "0")
    monkeypatch.setenv("BQ_DYN_ORDER", "batch")
    monkeypatch.setenv("BQ_DYN_W", "32")
    with GibbsEngine(net, arrv, n_chains=4, seed=99, chain_offset=3) as eng2:
        assert numpy.array_equal(eng2.sequential_order(), order)
        eng2.sweep(1, "SLICE", "NONE")
        d_group = eng2.state()
    assert numpy.array_equal(d_batch, d_group)
    tmax = 2.0 * float(arrv.soa()["d"].max())
    for c in (0, 3):
        em = emul_helper.EmulChain(net, arrv)
        em.slice_sweep(order, 99, 3 + c, 0, tmax=tmax)
        assert em.check() == 0
        # host libm vs device libdevice and FMA contraction differ in the last bit: 1e-12 instead of bit equality (a
        # flipped accept decision anywhere in the sweep would show as a difference of the order of a service time)
        assert helpers.relerr(em.d, d_batch[c]).max() < 1e-12, "chain %d differs from the host emulation" % c


def test_config1_single_chain_1000_sweeps():
    """BASELINE.json configs[0]: single M/M/1 queue, 200 tasks, 50 % observed, ONE chain x 1000 Gibbs sweeps through
    estimation.bayes -- posterior means of the two service parameters and of the mean wait against the reference's
    own long runs from the same initial state (tests/golden/post_mm1.npz), within 4 combined Monte-Carlo standard
    errors."""
    g = helpers.golden("post_mm1")
    net, arrv = helpers.golden_state(dict((k, g[k]) for k in g.files if k != "theta") | {"theta": g["theta0"]}, "s0")
    assert arrv.num_events() == 400
    mus, _ = estimation.bayes(net, arrv, 1000, seed=13, report_fn=None)
    run = estimation.last_run()
    assert run.theta.shape == (1000, 1, 2) and run.stats["nevents"] == 1000 * len(run.engine.schedule()[1])
    burn = 200
    ours = numpy.concatenate((run.theta[burn:, :, :], run.mean_wait[burn:, :, 1:]), axis=2)  # [n, 1, 3]
    ref = numpy.concatenate((g["theta"][:, burn:, :], g["mean_wait"][:, burn:, 1:]), axis=2).transpose(1, 0, 2)  # [n, R, 3]
    m_o, m_r = ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1))
    se = numpy.sqrt(diagnostics.mcse(ours) ** 2 + diagnostics.mcse(ref) ** 2)
    assert numpy.all(numpy.abs(m_o - m_r) < 4.0 * se), (m_o, m_r, se)
