"""Multi-rank host logic on CPU: chain sharding and the summary gather, world_size 2 over gloo."""
import os
import socket
import subprocess
import sys

import numpy

from bayes_qnet_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, numpy
sys.path.insert(0, %r)
import torch.distributed as dist
from bayes_qnet_b200 import parallel, diagnostics
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = parallel.shard(7, r, w)
rng = numpy.random.RandomState(0)
full = rng.normal(size=(50, 7, 3))
th, mw, lp = parallel.gather_summaries(full[:, lo:hi], full[:, lo:hi, :2], full[:, lo:hi, 0])
assert numpy.array_equal(th, full) and numpy.array_equal(mw, full[:, :, :2]) and numpy.array_equal(lp, full[:, :, 0])
assert numpy.allclose(diagnostics.rhat(th), diagnostics.rhat(full))
print("rank %%d ok %%d..%%d" %% (r, lo, hi))
dist.destroy_process_group()
"""


def test_shard_partitions_all_chains():
    for total, world in ((1024, 8), (7, 2), (5, 8), (65536, 4)):
        ranges = [parallel.shard(total, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_single_process_gather_is_identity():
    x = numpy.arange(24.0).reshape(2, 3, 4)
    assert numpy.array_equal(parallel.gather_traces(x), x)


def test_gather_world_size_2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 ok 0..4" in out.stdout and "rank 1 ok 4..7" in out.stdout


def test_device_diagnostics_match_the_host_versions():
    """parallel.device_diagnostics (torch ops, meant to run where the gathered summaries live) on AR(1) chains: split
    R-hat equals diagnostics.rhat; the summed per-chain ESS is near n m (1 - phi) / (1 + phi)."""
    import torch
    from bayes_qnet_b200 import diagnostics
    rs = numpy.random.RandomState(0)
    n, m = 2000, 24
    phi = numpy.array([0.0, 0.5, 0.9])
    x = numpy.zeros((n, m, 3))
    e = rs.standard_normal((n, m, 3))
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    d = parallel.device_diagnostics(torch.from_numpy(x))
    assert numpy.allclose(d["rhat"], diagnostics.rhat(x), rtol=1e-10)
    theory = n * m * (1 - phi) / (1 + phi)
    assert numpy.all(numpy.abs(d["ess_per_chain_sum"] / theory - 1.0) < 0.2), (d["ess_per_chain_sum"], theory)
    assert numpy.all(d["ess_per_chain_min"] > 0)


def test_allgather_into_tensor_layout_world_size_2_gloo(tmp_path):
    """The rank-major -> chain-major reshape of gather_summaries_device, exercised with the same torch calls on gloo."""
    worker = r"""
import os, sys, numpy, torch
sys.path.insert(0, %r)
import torch.distributed as dist
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
full = torch.arange(5 * 6 * 2, dtype=torch.float64).reshape(5, 6, 2)
local = full[:, 3 * r:3 * r + 3].contiguous()
n = local.shape[0]
out = torch.empty((w * n,) + tuple(local.shape[1:]), dtype=local.dtype)
dist.all_gather_into_tensor(out, local)
got = out.view(w, n, local.shape[1], local.shape[2]).permute(1, 0, 2, 3).reshape(n, w * local.shape[1], local.shape[2])
assert torch.equal(got, full)
print("rank %%d layout ok" %% r)
dist.destroy_process_group()
"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker2.py"
    script.write_text(worker % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 layout ok" in out.stdout and "rank 1 layout ok" in out.stdout
