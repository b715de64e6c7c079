"""Queue descriptors (host side).

The reference's ``queues.pyx`` Queue vtable (queues.pxd:19-37) -- diff lists, likelihood deltas, service
recomputation -- is what the CUDA kernels subsume; what stays on the host is the *description* of each
station: its discipline, processor count and service template.  Class names follow the reference so that
``qnetu`` can build the same objects: QueueGGk (queues.pyx:444), DelayStation (:745), QueuePS (:1032),
DistributionTemplate (:1305).
"""
from . import distributions

Q_GGK, Q_DELAY, Q_PS = 0, 1, 2
ARRV_BY_A, ARRV_BY_D = 0, 1


class DistributionTemplate(object):
    """A univariate service distribution attached to a queue (queues.pyx:1305-1371)."""

    def __init__(self, f):
        self.f = f
        self.zero_prob = 0.01

    def __repr__(self):
        return "<DISTRIBUTION ZERO_PROB=\"%s\">%s</DISTRIBUTION>" % (self.zero_prob, self.f)

    def numParameters(self):
        return self.f.numParameters()

    def getParameters(self):
        return list(self.f.parameters)

    def setParameters(self, v):
        self.f.parameters = v

    parameters = property(getParameters, setParameters)

    def sample_service(self):
        return float(self.f.sample(1)[0])

    def mean(self):
        return self.f.mean()

    def std(self):
        return self.f.std()

    def as_yaml(self):
        return self.f.as_yaml()

    @property
    def code(self):
        return self.f.code


class Queue(object):
    code = -1
    k = 1

    def __init__(self, name, service):
        self.name = name
        self.service = service

    def __repr__(self):
        return "[QUEUE: %s S: %s ]" % (self.name, self.service)

    def sample_service(self):
        return self.service.sample_service()

    def queue_order(self):
        return ARRV_BY_A

    def is_markov(self):
        return isinstance(self.service.f, distributions.Exponential)

    def service_lpdf(self, s):
        return self.service.f.lpdf(s)


class QueueGGk(Queue):
    """FIFO queue with k identical servers; YAML ``GG1`` maps to k = 1 (qnetu.py:109-110)."""
    code = Q_GGK

    def __init__(self, name, service, k=1):
        Queue.__init__(self, name, service)
        self.k = int(k)

    def as_yaml(self):
        return "  - { name: %s, type: GGk, processors: %d, service: %s }" % (self.name, self.k, self.service.as_yaml())


class DelayStation(Queue):
    """Infinite-server station: s = d - a (queues.pyx:745-784)."""
    code = Q_DELAY

    def as_yaml(self):
        return "  - { name: %s, type: DELAY, service: %s }" % (self.name, self.service.as_yaml())


class QueuePS(Queue):
    """Processor sharing: s = integral over [a, d] of dt / N(t) (queues.pyx:1032-1232)."""
    code = Q_PS

    def as_yaml(self):
        return "  - { name: %s, type: PS, service: %s }" % (self.name, self.service.as_yaml())
