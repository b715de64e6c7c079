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
