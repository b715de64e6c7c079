"""Per-queue statistics of an Arrivals (qstats.py:41-76), vectorised over the flat arrays.  The per-chain
mean waiting time and mean service used by the sampler's summaries are reduced on the GPU instead
(``GibbsEngine.mean_wait`` / ``suffstats``)."""
import numpy


def _by_q(arrv, values, mask=None, agg=numpy.mean):
    soa = arrv.soa()
    out = []
    for q in range(arrv.num_queues()):
        m = soa["qid"] == q
        if mask is not None:
            m = m & mask
        v = values[m]
        out.append(float(agg(v)) if len(v) else float("nan"))
    return out


def _obs(arrv):
    soa = arrv.soa()
    return (soa["obs_a"] != 0) & (soa["obs_d"] != 0)


def utilization(arrv):
    soa = arrv.soa()
    return [x / soa["d"].max() for x in _by_q(arrv, soa["s"], agg=numpy.sum)]


def mean_wait(arrv):
    soa = arrv.soa()
    return _by_q(arrv, numpy.maximum(0.0, soa["d"] - soa["a"] - soa["s"]))


def mean_response_time(arrv):
    soa = arrv.soa()
    return _by_q(arrv, soa["d"] - soa["a"])


def std_response_time(arrv):
    soa = arrv.soa()
    return _by_q(arrv, soa["d"] - soa["a"], agg=numpy.std)


def mean_service(arrv):
    return _by_q(arrv, arrv.soa()["s"])


def std_service(arrv):
    return _by_q(arrv, arrv.soa()["s"], agg=numpy.std)


def max_service(arrv):
    return _by_q(arrv, arrv.soa()["s"], agg=numpy.max)


def mean_obs_service(arrv):
    return _by_q(arrv, arrv.soa()["s"], mask=_obs(arrv))


def mean_latent_service(arrv):
    return _by_q(arrv, arrv.soa()["s"], mask=~_obs(arrv))


def mean_obs_response(arrv):
    soa = arrv.soa()
    return _by_q(arrv, soa["d"] - soa["a"], mask=_obs(arrv))
