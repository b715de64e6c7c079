"""Service-time families (host side): parameter holders, log-density, forward sampling.

Mirrors the reference's ``distributions.pyx`` classes (Exponential :52, Gamma :107, LogNormal :245) for
everything the host needs -- model description, forward simulation, reporting.  The hot-path evaluation
of these densities, the posterior draws and the ML updates run on the GPU (``csrc/bq_dist.cuh``).
"""
import math

import numpy

D_EXP, D_GAMMA, D_LOGNORMAL, D_DIRAC = 0, 1, 2, 3
HALF_LOG_2PI = -0.918938533204673  # distributions.pyx:243


class Distribution(object):
    code = -1

    def lpdf(self, x):
        raise NotImplementedError()

    def mean(self):
        raise NotImplementedError()

    def sample(self, N):
        raise NotImplementedError()

    def numParameters(self):
        return len(self.parameters)

    def parameter_kernel(self, theta_from, theta_to, data):
        """Log-density of the posterior draw theta_to given the services `data`: the transition kernel of
        sample_parameters (used by Qnet.parameter_kernel / estimation.chib_evidence)."""
        raise NotImplementedError()

    def _rng(self):
        return numpy.random


class Exponential(Distribution):
    """distributions.pyx:52-105.  parameters = [scale]."""
    code = D_EXP

    def __init__(self, scale=1.0):
        self.scale = float(scale)

    def __repr__(self):
        return "<EXP SCALE=%s/>" % self.scale

    def lpdf(self, x):
        return -math.log(self.scale) - x / self.scale

    def mean(self):
        return self.scale

    def std(self):
        return self.scale

    def sample(self, N):
        return numpy.random.exponential(scale=self.scale, size=N)

    @property
    def parameters(self):
        return [self.scale]

    @parameters.setter
    def parameters(self, v):
        self.scale = float(v[0])

    def as_yaml(self):
        return "[M, %s]" % self.scale

    def parameter_kernel(self, theta_from, theta_to, data):
        """distributions.pyx:98-104: the new scale ~ InverseGamma(N, sum of the services)."""
        from math import lgamma
        mu, N = float(numpy.sum(data)), len(data)
        x = theta_to[0]
        return -(N + 1) * math.log(x) - mu / x - lgamma(N) + N * math.log(mu)  # InverseGamma.lpdf, distributions.pyx:224


class Gamma(Distribution):
    """distributions.pyx:107-207.  parameters = [shape, scale]."""
    code = D_GAMMA

    def __init__(self, shape=1.0, scale=1.0):
        self.shape = float(shape)
        self.scale = float(scale)

    def __repr__(self):
        return "<GAMMA SHAPE=%s SCALE=%s/>" % (self.shape, self.scale)

    def lpdf(self, x):
        if self.shape == 1:  # distributions.pyx:130
            return -x / self.scale - self.shape * math.log(self.scale)
        lx = math.log(x) if x > 0 else -numpy.inf
        return (self.shape - 1) * lx - x / self.scale - math.lgamma(self.shape) - self.shape * math.log(self.scale)

    def mean(self):
        return self.shape * self.scale

    def std(self):
        return math.sqrt(self.shape) * self.scale

    def sample(self, N):
        return numpy.random.gamma(self.shape, size=N, scale=self.scale)

    @property
    def parameters(self):
        return [self.shape, self.scale]

    @parameters.setter
    def parameters(self, v):
        self.shape, self.scale = float(v[0]), float(v[1])

    def as_yaml(self):
        return "[G, %s, %s]" % (self.shape, self.scale)

    def parameter_kernel(self, theta_from, theta_to, data):
        """netutils.gamma_transition_kernel (netutils.py:44-60): the two conditionals of the slice-within-Gibbs draw."""
        from math import lgamma
        shape0, scale0 = theta_from
        shape1, scale1 = theta_to
        data = numpy.asarray(data, dtype=numpy.float64)
        S, Slog, N = float(data.sum()), float(numpy.log(data).sum()), len(data)
        logP, q, r, s = 1.0 + Slog, 1.0 + S, 1.0 + N, 1.0 + N
        p_shape = (shape1 - 1) * logP - q / scale0 - r * lgamma(shape1) - (shape1 * s) * math.log(scale0)
        p_scale = (shape1 - 1) * logP - q / scale1 - r * lgamma(shape1) - (shape1 * s) * math.log(scale1)
        return p_shape + p_scale


class LogNormal(Distribution):
    """distributions.pyx:245-345.  parameters = [meanlog, sdlog]."""
    code = D_LOGNORMAL

    def __init__(self, meanlog=0.0, sdlog=1.0):
        self.meanlog = float(meanlog)
        self.sdlog = float(sdlog)

    def __repr__(self):
        return "<LOGNORMAL MU=%s SD=%s/>" % (self.meanlog, self.sdlog)

    def lpdf(self, x):
        if x == 0:
            return -numpy.inf
        lx = math.log(x)
        return HALF_LOG_2PI - math.log(self.sdlog) - lx - ((lx - self.meanlog) ** 2) / (2 * self.sdlog * self.sdlog)

    def mean(self):
        return math.exp(self.meanlog + self.sdlog * self.sdlog / 2)

    def std(self):
        s2 = self.sdlog * self.sdlog
        return math.sqrt((math.exp(s2) - 1) * math.exp(2 * self.meanlog + s2))

    def sample(self, N):
        return numpy.exp(numpy.random.normal(self.meanlog, self.sdlog, size=N))

    @property
    def parameters(self):
        return [self.meanlog, self.sdlog]

    @parameters.setter
    def parameters(self, v):
        self.meanlog, self.sdlog = float(v[0]), float(v[1])

    def as_yaml(self):
        return "[LN, %s, %s]" % (self.meanlog, self.sdlog)

    def parameter_kernel(self, theta_from, theta_to, data):
        """distributions.pyx:323-338: tau ~ Gamma((n+1)/2, 2/(n sd^2)), mu ~ N(mean, sd_new) of the logs."""
        from math import lgamma
        logs = numpy.log(numpy.asarray([x for x in data if x >= 1e-50], dtype=numpy.float64))
        ns = len(data)
        if ns == 0 or len(logs) == 0:
            return 0.0
        mean = float(logs.mean())
        sd = max(float(numpy.sqrt(numpy.mean((logs - mean) ** 2))), 1e-10)  # distributions.pyx:276-299
        if sd <= 1e-5:
            return 0.0
        alpha, beta = (ns + 1) / 2.0, 2.0 / (ns * sd * sd)
        mu_new, sd_new = theta_to
        tau = 1.0 / (sd_new * sd_new)
        lpdf_tau = (alpha - 1) * math.log(tau) - tau / beta - lgamma(alpha) - alpha * math.log(beta)
        lpdf_mu = -0.5 * math.log(2 * math.pi) - math.log(sd_new) - (mu_new - mean) ** 2 / (2 * sd_new * sd_new)
        return lpdf_mu + lpdf_tau


class Dirac(Distribution):
    """distributions.pyx:347-375: a constant service time.  parameters = [loc].  Like the reference's class it has no
    log-density -- it serves forward simulation (Qnet.sample) and reporting; the samplers refuse a network that
    carries one (bq_create: unknown service family), where the reference fails at its first lpdf call."""
    code = D_DIRAC

    def __init__(self, loc):
        self.loc = float(loc)

    def __repr__(self):
        return "<DIRAC LOC=%s/>" % self.loc

    def mean(self):
        return self.loc

    def std(self):
        return 0.0

    def sample(self, N):
        return numpy.full(N, self.loc)

    def quantile(self, p):
        return self.loc

    def estimate(self, data):
        self.loc = float(numpy.mean(data))

    @property
    def parameters(self):
        return [self.loc]

    @parameters.setter
    def parameters(self, v):
        self.loc = float(v[0])

    def as_yaml(self):
        return "[D, %s]" % self.loc
