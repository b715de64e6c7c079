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


# ---- summaries on the device: no host bounce ---------------------------------------------------------------------
class _DeviceArray(object):
    """A raw device pointer dressed for torch.as_tensor (CUDA array interface, fp64, C order)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_traces(eng):
    """Zero-copy torch views of the engine's device-resident traces (bq_trace_dev): theta [n_rec, C, P],
    mean_wait [n_rec, C, Q], logp [n_rec, C].  The engine must be synchronised."""
    import torch
    n = eng.trace_len()
    p = eng.device_pointers()
    dev = torch.device("cuda", torch.cuda.current_device())
    mk = lambda ptr, shape: torch.as_tensor(_DeviceArray(ptr, shape), device=dev)
    return (mk(p["tr_theta"], (n, eng.n_chains, eng.P)), mk(p["tr_wait"], (n, eng.n_chains, eng.Q)),
            mk(p["tr_logp"], (n, eng.n_chains)))


def gather_summaries_device(eng, world=None):
    """Per-chain summaries of every rank -- the traces of the service parameters and of the per-queue mean waits
    (source queue dropped) -- as ONE device tensor [n_rec, total chains, P + Q - 1], chains in global order:
    ncclAllGather straight from the engines' trace buffers, no host copy."""
    import torch
    import torch.distributed as dist
    th, mw, _ = device_traces(eng)
    local = torch.cat((th, mw[:, :, 1:]), dim=2).contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    w = dist.get_world_size()
    n = local.shape[0]
    out = torch.empty((w * n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)  # rank-major concatenation; equal chain counts per rank
    return out.view(w, n, local.shape[1], local.shape[2]).permute(1, 0, 2, 3).reshape(n, w * local.shape[1], local.shape[2])


def device_diagnostics(x, budget_bytes=1 << 29):
    """Split R-hat over the ensemble and per-chain effective sample sizes, computed where the summaries live (any torch
    device).  x: [n_iter, n_chains, K].  Returns small host arrays: rhat [K]; ess_per_chain_sum [K] = sum over chains of
    the chain's own ESS (Geyer's initial positive sequence on its autocorrelation); ess_per_chain_min [K].
    The summaries are processed in groups as large as `budget_bytes` of padded-FFT workspace allows (all K at once for
    the short windows of the timed path: a dozen kernel launches instead of a dozen per summary) and the three results
    cross to the host in one copy."""
    import math

    import torch
    n, m, K = x.shape
    h = n // 2
    if h >= 2:
        halves = torch.cat((x[:h], x[h:2 * h]), dim=1)
        cm = halves.mean(dim=0)
        W = halves.var(dim=0, unbiased=True).mean(dim=0)
        B = h * cm.var(dim=0, unbiased=True)
        rhat = torch.sqrt(((h - 1.0) / h * W + B / h) / W)
    else:
        rhat = torch.full((K,), float("nan"), dtype=x.dtype, device=x.device)
    if n >= 4:
        floor = 1.0 / max(math.log10(float(max(n, 10))), 1.0)
        group = max(1, min(K, int(budget_bytes // max(1, 2 * n * m * 16 * 4))))  # complex spectrum + products + real copies
        sums, mins = [], []
        t2 = (n // 2) * 2
        for k0 in range(0, K, group):
            xs = x[:, :, k0:k0 + group]
            xc = xs - xs.mean(dim=0, keepdim=True)
            f = torch.fft.rfft(xc, n=2 * n, dim=0)
            ac = torch.fft.irfft(f * torch.conj(f), n=2 * n, dim=0)[:n] / n     # [n, m, g]
            rho = ac / ac[0].clamp_min(1e-300)
            pairs = rho[0:t2:2] + rho[1:t2:2]                                   # [n/2, m, g]
            alive = torch.cumprod((pairs >= 0).to(x.dtype), dim=0)
            tau = -1.0 + 2.0 * (pairs * alive).sum(dim=0)                       # [m, g]
            ess = n / tau.clamp_min(floor)
            ess = torch.where(ac[0] > 0, ess, torch.zeros_like(ess))
            sums.append(ess.sum(dim=0))
            mins.append(ess.min(dim=0).values)
        ess_sum, ess_min = torch.cat(sums), torch.cat(mins)
    else:
        ess_sum = torch.full((K,), float("nan"), dtype=x.dtype, device=x.device)
        ess_min = ess_sum
    host = torch.stack((rhat, ess_sum, ess_min)).cpu().numpy()
    return {"rhat": host[0], "ess_per_chain_sum": host[1], "ess_per_chain_min": host[2]}
