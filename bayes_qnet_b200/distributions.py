"""Service-time families (host side): parameter holders, log-density, forward sampling.

Mirrors the reference's ``distributions.pyx`` classes (Exponential :52, Gamma :107, LogNormal :245) for
everything the host needs -- model description, forward simulation, reporting.  The hot-path evaluation
of these densities, the posterior draws and the ML updates run on the GPU (``csrc/bq_dist.cuh``).
"""
import math

import numpy

D_EXP, D_GAMMA, D_LOGNORMAL = 0, 1, 2
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
