"""bayes_qnet_b200 -- B200-native Gibbs sampling for queueing networks with missing data.

Drop-in for the hot path of casutton/bayes-qnet: same YAML models (``qnetu``), same ``Qnet`` / ``Arrivals`` /
``Event`` objects and ``estimation.bayes|sem|sample_departures`` entry points, with the resampling of hidden
arrival/departure times and the service-parameter updates executed by hand-written sm_100a CUDA kernels for
thousands of independent chains (``GibbsEngine``).  There is no CPU fallback.
"""
from . import distributions, queues, hmm, sampling, arrivals, capi, engine, qnet, qstats, diagnostics, estimation, qnetu  # noqa: F401
from .engine import GibbsEngine  # noqa: F401

__all__ = ["distributions", "queues", "hmm", "sampling", "arrivals", "capi", "engine", "qnet", "qstats",
           "diagnostics", "estimation", "qnetu", "GibbsEngine"]
