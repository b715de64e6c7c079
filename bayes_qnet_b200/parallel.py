"""Chain sharding over the GPUs of one box, one process per GPU.

Chains are independent (same data, different RNG counters), so the sweep needs no collective at all: rank r
owns the global chain ids [r*C, (r+1)*C) (``GibbsEngine(chain_offset=r*C)``; the RNG is keyed by the GLOBAL id,
so a sharded run reproduces the unsharded one chain for chain).  The only exchange is the gather of per-chain
summaries -- parameter / mean-wait / log-prob traces, a few KB to MB -- needed to compute R-hat and ESS over
the whole ensemble; that is a plain all_gather over NCCL (NVLink5 / NVSwitch), or gloo on CPU in the tests.
"""
import numpy


def shard(n_chains_total, rank, world):
    """Contiguous chain range [lo, hi) of `rank` when n_chains_total chains are split over `world` ranks."""
    base, rem = divmod(int(n_chains_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_traces(x, world=None):
    """x: this rank's [n_rec, C_local, K] trace (numpy).  Returns [n_rec, sum C_local, K] on every rank, chains
    in global order.  Uses torch.distributed when a process group is up; identity for a single process."""
    x = numpy.ascontiguousarray(x)
    try:
        import torch
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return x
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    w = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    sizes = torch.zeros(w, dtype=torch.int64, device=dev)
    sizes[dist.get_rank()] = x.shape[1]
    dist.all_reduce(sizes)
    cmax = int(sizes.max())
    pad = numpy.zeros((x.shape[0], cmax) + x.shape[2:], dtype=x.dtype)
    pad[:, :x.shape[1]] = x
    mine = torch.from_numpy(pad).to(dev)
    parts = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(parts, mine)
    out = [p.cpu().numpy()[:, :int(sizes[r])] for r, p in enumerate(parts)]
    return numpy.concatenate(out, axis=1)


def gather_summaries(theta, mean_wait, logp):
    """Gather the three trace arrays of GibbsEngine.traces() over all ranks."""
    return gather_traces(theta), gather_traces(mean_wait), gather_traces(logp[:, :, None])[:, :, 0]
