"""Run the dynamic-order walk / commit / proposal / batch routines (host build of bq_dynamic.cuh) under
AddressSanitizer: conditionals on grids that reach outside the support, slice / MH / batch sweeps on every fixture
network.  Started by tests/test_emul.py with LD_PRELOAD=libasan and BQ_EMUL_SO pointing at the ASan build
(compute-sanitizer is not available on the GPU pool; this covers the same index arithmetic on the CPU)."""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import emul_helper  # noqa: E402

emul_helper.SO = os.environ["BQ_EMUL_SO"]
emul_helper.build = lambda: emul_helper.SO
import helpers  # noqa: E402

nets = [("cond_"+n, None) for n in helpers.COND_NETS] + [(None, n) for n in sorted(helpers.MIXED)]
for gname, mname in nets:
    if gname:
        g = helpers.golden(gname); net, arrv = helpers.golden_state(g, "s0"); name=gname
    else:
        net, smp, sub, arrv = helpers.mixed(mname); name=mname
    has_ps = any(q.code == 2 for q in net.queues)
    oc = helpers.oracle_chain(net, arrv)
    order = oc.eligible()
    ec = emul_helper.EmulChain(net, arrv)
    rs = numpy.random.RandomState(1)
    for e in order:
        e1 = oc.next_byt[e]; hi = oc.d[e1] if e1 >= 0 else oc.d[e] + 50.0
        xs = numpy.concatenate((rs.uniform(oc.a[e]-1, hi+1, 6), [oc.a[e], hi, 1e9, 0.0]))
        ec.logcond(int(e), xs)
        if not has_ps: ec.mh_proposal(int(e), xs)
    for s in range(3):
        ec.slice_sweep(order, 5, 1, s)
        ec.sweep_batch(order, 5, 1, s+10, 32)
        if not has_ps: ec.mh_sweep(order, 5, 1, s+20); ec.sweep_batch(order, 5, 1, s+30, 7, mh=True)
        assert ec.check() == 0
    print(name, "ok")
