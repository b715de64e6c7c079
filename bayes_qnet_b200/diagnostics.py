"""Convergence diagnostics over the chain ensemble: split R-hat and effective sample size.

New relative to the reference (which runs one chain and computes neither); these are
what the per-chain summaries gathered over NCCL are for (``parallel.gather_traces``).
"""
import numpy


def rhat(x):
    """Split R-hat.  x: [n_iter, n_chains, ...] -> [...]."""
    x = numpy.asarray(x, dtype=numpy.float64)
    n = x.shape[0] // 2
    if n < 2:
        return numpy.full(x.shape[2:], numpy.nan)
    halves = numpy.concatenate((x[:n], x[n:2 * n]), axis=1)  # [n, 2C, ...]
    m = halves.shape[1]
    chain_mean = halves.mean(axis=0)
    chain_var = halves.var(axis=0, ddof=1)
    W = chain_var.mean(axis=0)
    B = n * chain_mean.var(axis=0, ddof=1) if m > 1 else numpy.zeros_like(W)
    var_plus = (n - 1.0) / n * W + B / n
    with numpy.errstate(divide="ignore", invalid="ignore"):
        return numpy.sqrt(var_plus / W)


def _autocov(x):
    """Autocovariance along axis 0 by FFT.  x: [n, ...]."""
    n = x.shape[0]
    xc = x - x.mean(axis=0, keepdims=True)
    f = numpy.fft.rfft(xc, n=2 * n, axis=0)
    ac = numpy.fft.irfft(f * numpy.conj(f), n=2 * n, axis=0)[:n]
    return ac / n


def ess(x):
    """Effective sample size pooled over chains (Geyer initial-positive-sequence on the chain-averaged
    autocorrelation, with the between-chain variance folded in).  x: [n_iter, n_chains, ...] -> [...]."""
    x = numpy.asarray(x, dtype=numpy.float64)
    n, m = x.shape[0], x.shape[1]
    if n < 4:
        return numpy.full(x.shape[2:], numpy.nan)
    acov = _autocov(x)  # [n, m, ...]
    chain_mean = x.mean(axis=0)
    W = acov[0].mean(axis=0) * n / (n - 1.0)
    B_over_n = chain_mean.var(axis=0, ddof=1) if m > 1 else numpy.zeros_like(W)
    var_plus = (n - 1.0) / n * W + B_over_n
    rho = 1.0 - (W - acov.mean(axis=1)) / numpy.where(var_plus > 0, var_plus, 1.0)  # [n, ...]
    flat = rho.reshape(n, -1)
    out = numpy.empty(flat.shape[1])
    for j in range(flat.shape[1]):
        r = flat[:, j]
        tau = -1.0
        t = 0
        while t + 1 < n:
            pair = r[t] + r[t + 1]
            if pair < 0:
                break
            tau += 2.0 * pair
            t += 2
        out[j] = n * m / max(tau, 1.0 / numpy.log10(max(n * m, 10)))
    return out.reshape(x.shape[2:])


def mcse(x):
    """Monte Carlo standard error of the pooled mean."""
    x = numpy.asarray(x, dtype=numpy.float64)
    sd = x.reshape(-1, *x.shape[2:]).std(axis=0, ddof=1)
    return sd / numpy.sqrt(ess(x))
