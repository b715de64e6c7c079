import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy, helpers
from bayes_qnet_b200 import GibbsEngine, capi, qnetu
T = """
states:
  - {name: INITIAL, queues: [INITIAL], successors: [A], initial: TRUE}
  - {name: A, queues: [QA]}
queues:
  - { name: INITIAL, service: [M, 1.0] }
  - { name: QA, service: [%s] }
"""
which = sys.argv[1]
yaml_text = T % {"ln": "LN, 0.0, 0.5", "g": "G, 3.0, 0.1"}[which]
net, smp, sub, ini = helpers.synthetic(yaml_text, 60, 0.3, seed=3)
with GibbsEngine(net, ini, n_chains=3, seed=5) as eng:
    eng.sweep(2, "SLICE", "BAYES")
    print(which, "fused BAYES sweep ok", eng.params()[0], flush=True)
    ss = eng.suffstats()
    print(ss[0])
    arr = numpy.ascontiguousarray(ss[0])
    capi.check(eng._L.bq_update_params_from(eng._h, capi.BQ_UPD_BAYES, capi.ptr(arr, capi.c_f64p), 1))
    print(which, "pooled BAYES ok", eng.params()[0], flush=True)
