"""The dynamic-order walk / commit routines of the CUDA kernels (bq_dynamic.cuh), compiled for the host by
tests/emul and run against (i) the reference's own inner_dfun vectors and (ii) the CPU oracle's trajectories.
This is the no-GPU check of the kernel logic; the kernels themselves are checked by test_gpu_parity.py."""
import numpy
import pytest

import emul_helper
import helpers

TOL = 1e-9


@pytest.mark.parametrize("name", helpers.COND_NETS)
def test_emulated_logcond_matches_reference_vectors(name):
    g = helpers.golden("cond_" + name)
    total = 0
    for st, pre, theta in (("s0", "s0", g["theta"]), ("s1", "s1", g["theta"]), ("s2", "s1", g["s2_theta"])):
        net, arrv = helpers.golden_state(g, pre, theta)
        ec = emul_helper.EmulChain(net, arrv, theta)
        for r, xs, ref in zip(g[st + "_rows"], g[st + "_x"], g[st + "_g"]):
            v = ec.logcond(int(r), xs)
            assert numpy.array_equal(numpy.isfinite(v), numpy.isfinite(ref)), "support differs at row %d" % r
            m = numpy.isfinite(ref)
            if m.any():
                assert helpers.relerr(v[m], ref[m]).max() < TOL
                total += int(m.sum())
    assert total > 200


@pytest.mark.parametrize("name", helpers.COND_NETS)
def test_emulated_sweeps_match_oracle(name):
    """Same stream, same scan order: identical trajectories, and the incrementally maintained queue order and
    free-time vectors stay consistent with a from-scratch rebuild after every sweep."""
    g = helpers.golden("cond_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    oc = helpers.oracle_chain(net, arrv)
    ec = emul_helper.EmulChain(net, arrv)
    order = oc.eligible()
    for s in range(6):
        oc.slice_sweep(order, 4242, 3, s)
        ec.slice_sweep(order, 4242, 3, s)
        assert ec.check() == 0
    dv = ec.derive()
    assert helpers.relerr(ec.d, oc.d).max() < TOL
    assert helpers.relerr(dv["s"], oc.s).max() < TOL
    assert numpy.array_equal(dv["proc"], oc.proc)
    assert numpy.array_equal(dv["next_byq"], oc.next_byq) and numpy.array_equal(dv["prev_byq"], oc.prev_byq)


@pytest.mark.parametrize("name", sorted(helpers.MIXED))
def test_emulated_mixed_networks_match_oracle(name):
    """Every pairing of disciplines (PS with LogNormal / Gamma service next to multi-server FIFO, PS -> PS,
    delay stations, parallel queues): conditional on random grids and whole sweeps against the oracle."""
    net, smp, sub, ini = helpers.mixed(name)
    oc = helpers.oracle_chain(net, ini)
    ec = emul_helper.EmulChain(net, ini)
    assert helpers.relerr(ec.derive()["s"], oc.s).max() < 1e-12
    order = oc.eligible()
    rs = numpy.random.RandomState(5)
    for sweep in range(4):
        for e in rs.choice(order, size=25, replace=False):
            e1 = oc.next_byt[e]
            hi = oc.d[e1] if e1 >= 0 else oc.d[e] + 3.0
            xs = numpy.concatenate((rs.uniform(oc.a[e] - 0.2, hi + 0.2, size=6), [oc.d[e]]))
            v, r = ec.logcond(int(e), xs), oc.logcond(int(e), xs)
            assert numpy.array_equal(numpy.isfinite(v), numpy.isfinite(r)), (name, e, xs, v, r)
            m = numpy.isfinite(r)
            assert helpers.relerr(v[m], r[m]).max() < TOL, (name, e, xs, v, r)
        oc.slice_sweep(order, 99, 1, sweep)
        ec.slice_sweep(order, 99, 1, sweep)
        assert ec.check() == 0
        assert helpers.relerr(ec.d, oc.d).max() < TOL
    dv = ec.derive()
    assert helpers.relerr(dv["s"], oc.s).max() < TOL
    assert numpy.array_equal(dv["next_byq"], oc.next_byq)


@pytest.mark.parametrize("name", helpers.MH_NETS)
def test_emulated_mh_matches_reference_and_oracle(name):
    """MH-within-Gibbs (Qnet.gibbs_resample): the product's proposal log-density against the reference's Pwfun
    values, and whole MH sweeps (proposal draws, accept / reject, commits) against the oracle."""
    g = helpers.golden("mh_" + name)
    net, arrv = helpers.golden_state(g, "s0")
    ec = emul_helper.EmulChain(net, arrv)
    for r, xs, ref in zip(g["rows"], g["x"], g["h"]):
        v = ec.mh_proposal(int(r), xs)
        assert numpy.array_equal(numpy.isfinite(v), numpy.isfinite(ref)), "support differs at row %d" % r
        m = numpy.isfinite(ref)
        assert helpers.relerr(v[m], ref[m]).max() < 1e-12
    oc = helpers.oracle_chain(net, arrv)
    order = oc.eligible()
    props = rejects = 0
    for s in range(6):
        a, b = oc.mh_sweep(order, 77, 2, s), ec.mh_sweep(order, 77, 2, s)
        assert a == b
        props, rejects = props + a[0], rejects + a[1]
        assert ec.check() == 0
    assert helpers.relerr(ec.d, oc.d).max() < TOL
    assert 0 < rejects < props
    # sample_final = False skips the tasks' last events (qnet.pyx:361)
    n_final = int(sum(1 for e in order if oc.next_byt[e] < 0))
    a = oc.mh_sweep(order, 77, 2, 9, sample_final=False)
    assert a == ec.mh_sweep(order, 77, 2, 9, sample_final=False) and a[0] <= len(order) - n_final


@pytest.mark.parametrize("name", helpers.COND_NETS + sorted(helpers.MIXED))
def test_emulated_batches_are_the_sequential_sweep(name):
    """lane_update + range claims (what sweep_batch_kernel does with 32 lanes), emulated for several batch sizes: the
    committed result must be the oracle's sequential sweep whatever the batch boundaries, for slice and MH updates."""
    if name in helpers.MIXED:
        net, smp, sub, arrv = helpers.mixed(name)
    else:
        net, arrv = helpers.golden_state(helpers.golden("cond_" + name), "s0")
    has_ps = any(q.code >= 2 for q in net.queues)  # PS and random-order queues: slice only
    for B in (2, 7, 32):
        for mh in ((False,) if has_ps else (False, True)):
            oc = helpers.oracle_chain(net, arrv)
            ec = emul_helper.EmulChain(net, arrv)
            order = oc.eligible()
            redone = 0
            for s in range(3):
                if mh:
                    a, b = oc.mh_sweep(order, 5, 1, s), ec.sweep_batch(order, 5, 1, s, B, mh=True)
                else:
                    tmax = 2.0 * oc.d.max()
                    a, b = (oc.slice_sweep(order, 5, 1, s, tmax=tmax)[0], 0), ec.sweep_batch(order, 5, 1, s, B, tmax=tmax)
                assert a == b[:2]
                redone += b[2]
                assert ec.check() == 0
            assert redone > 0  # task-major order: neighbours in a batch do conflict
            assert helpers.relerr(ec.d, oc.d).max() < TOL


def test_emulated_code_is_clean_under_address_sanitizer(tmp_path):
    """The same routines, compiled with -fsanitize=address, driven over every fixture network with probes outside
    the support, slice / MH / batch sweeps (tests/emul/asan_check.py)."""
    import os
    import subprocess
    import sys
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    so = str(tmp_path / "libbq_emul_asan.so")
    subprocess.check_call([emul_helper.NVCC, "-O1", "-g", "-std=c++17", "-Wno-deprecated-gpu-targets",
                           "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fsanitize=address,-fno-omit-frame-pointer",
                           "-shared", "-o", so, emul_helper.SRC])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", BQ_EMUL_SO=so)
    r = subprocess.run([sys.executable, os.path.join(emul_helper.HERE, "emul", "asan_check.py")], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ERROR: AddressSanitizer" not in r.stderr, r.stderr[-3000:]
    assert r.stdout.count(" ok") == 12
