import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, bench
from bayes_qnet_b200 import GibbsEngine
wl = bench.WORKLOADS["config2"]
net, ini, how = bench.make_input(5000, yaml_text=wl["yaml"], init=wl["init"], fixture=wl["fixture"])
soa = ini.soa()
eng = GibbsEngine(net, ini, n_chains=8, seed=777)
print("theta0", eng.params()[0])
eng.init_chains(0.7)
full = eng.state(full=True)
for c in range(8):
    s = full["s"][c]; d = full["d"][c]
    src = soa["qid"] == 0
    print(c, "src zero-s", int((s[src] == 0).sum()), "tiny", int((s[src] < 1e-9).sum()), "d_last_src", d[src].max(), "max d", d.max(),
          "sum s src", s[src].sum(), "zeros by queue", [int((s[soa["qid"] == q] == 0).sum()) for q in range(7)])
eng.sweep(200, "SLICE", "BAYES")
print("after 200 sweeps theta:", eng.params()[:, 0], "stuck", eng.stats()["n_stuck"], "nev", eng.stats()["nevents"])
full = eng.state(full=True)
for c in range(3):
    s = full["s"][c]; d = full["d"][c]; src = soa["qid"] == 0
    print(c, "src zero-s", int((s[src] == 0).sum()), "d_last_src", d[src].max(), "max d", d.max())
