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


def std_obs_response(arrv):
    soa = arrv.soa()
    return _by_q(arrv, soa["d"] - soa["a"], mask=_obs(arrv), agg=numpy.std)


def std_obs_service(arrv):
    return _by_q(arrv, arrv.soa()["s"], mask=_obs(arrv), agg=numpy.std)


def nobs_by_queue(arrv):
    """Number of fully observed events per queue (qstats.py:81)."""
    soa = arrv.soa()
    obs = _obs(arrv)
    return [int(numpy.sum(obs & (soa["qid"] == q))) for q in range(arrv.num_queues())]


def mean_wait_percentage(arrv):
    """Per queue, the mean over events of wait time / response time (qstats.py:87-95)."""
    soa = arrv.soa()
    r = soa["d"] - soa["a"]
    pct = numpy.where(r < 1e-50, 0.0, 1.0 - soa["s"] / numpy.where(r < 1e-50, 1.0, r))
    return _by_q(arrv, pct)


def mean_qsize_for_qid(arrv, qi):
    """Mean number in queue qi as seen by an arriving customer (qstats.py:98-116): arrivals and departures merged in
    time order (ties: departures first, as the reference's tuple sort gives), the running count before each arrival."""
    soa = arrv.soa()
    m = soa["qid"] == qi
    a, d = soa["a"][m], soa["d"][m]
    if len(a) == 0:
        return float("nan")
    t = numpy.concatenate((a, d))
    k = numpy.concatenate((numpy.ones(len(a)), -numpy.ones(len(d))))
    o = numpy.lexsort((k, t))
    cum = numpy.concatenate(([0.0], numpy.cumsum(k[o])[:-1]))
    return float(numpy.mean(cum[k[o] == 1]))


def mean_qsize_arriving(arrv):
    return [mean_qsize_for_qid(arrv, qi) for qi in range(arrv.num_queues())]


def _by_task(arrv, fn, agg=numpy.mean, keep=None):
    soa = arrv.soa()
    order = numpy.argsort(soa["tid"], kind="stable")
    tids, starts = numpy.unique(soa["tid"][order], return_index=True)
    xs = []
    for rows in numpy.split(order, starts[1:]):
        if keep is None or keep(soa, rows):
            xs.append(fn(soa, rows))
    return agg(xs) if xs else 0


def mean_task_time(arrv):
    """Mean over tasks of last departure - first departure (qstats.py:122-127)."""
    return float(_by_task(arrv, lambda s, r: s["d"][r].max() - s["d"][r].min()))


def mean_task_time_in_range(arrv, l, r):
    return float(_by_task(arrv, lambda s, rows: s["d"][rows].max() - s["d"][rows].min(),
                          keep=lambda s, rows: l <= s["d"][rows].min() < r))


def num_tasks_in_arrival_range(arrv, l, r):
    return int(_by_task(arrv, lambda s, rows: 1, agg=len, keep=lambda s, rows: l <= s["d"][rows].min() < r))


def mean_task_service(arrv):
    return float(_by_task(arrv, lambda s, r: s["s"][r].sum()))


def mean_task_length(arrv):
    return float(_by_task(arrv, lambda s, r: len(r)))


def total_service(arrv):
    return float(arrv.soa()["s"].sum())


# "exported" statistics (qstats.py:174-184)
STATS = [utilization, mean_wait, mean_response_time, std_response_time, mean_service, std_service,
         mean_obs_response, std_obs_response, mean_obs_service, std_obs_service, mean_latent_service, nobs_by_queue,
         mean_wait_percentage, mean_qsize_arriving, mean_task_time]


def all_stats_string():
    return " ".join(fn.__name__ for fn in STATS)


def stringify(val):
    if isinstance(val, (list, tuple, numpy.ndarray)):
        return " ".join(map(str, val))
    return str(val)
