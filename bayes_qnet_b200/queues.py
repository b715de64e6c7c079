"""Queue descriptors (host side).

The reference's ``queues.pyx`` Queue vtable (queues.pxd:19-37) -- diff lists, likelihood deltas, service
recomputation -- is what the CUDA kernels subsume; what stays on the host is the *description* of each
station: its discipline, processor count and service template.  Class names follow the reference so that
``qnetu`` can build the same objects: QueueGGk (queues.pyx:444), DelayStation (:745), QueueGG1R (:788), QueuePS (:1032),
DistributionTemplate (:1305).
"""
from . import distributions

Q_GGK, Q_DELAY, Q_PS, Q_GG1R = 0, 1, 2, 3
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

    def component_codes(self):
        return [self.f.code]

    def component_dists(self):
        return [self.f]

    def initialize_auxiliaries(self, evt):
        pass


class MixtureTemplate(object):
    """A finite mixture of service distributions with a per-event component indicator (queues.pyx:1373-1509; YAML
    ``[MIX, w0, [M, 3.0], w1, [LN, 0.0, 1.0], ...]``).  An event's service density is the one of ITS component
    (serviceLikelihood, :1455-1457); the components are resampled at the start of every sweep and the mixing
    proportions follow their counts -- both on the GPU (bq_static.cuh).  Parameters: the mixing proportions, then each
    component's parameters (:1470-1474)."""

    def __init__(self, var_name, mixing, dists):
        self.var_name = var_name
        self.mixing = [float(w) for w in mixing]
        self.dists = list(dists)
        self.zero_prob = 0

    def __repr__(self):
        result = "<MIXTURE>\n  probs: %s\n  zero_prob: %s\n" % (self.mixing, self.zero_prob)
        for d in self.dists:
            result += "  %s\n" % d
        return result + "</MIXTURE>"

    f = property(lambda self: self.dists[0])   # family of the first component, where a single family is asked for
    code = property(lambda self: self.dists[0].code)

    def component_codes(self):
        return [d.code for d in self.dists]

    def component_dists(self):
        return list(self.dists)

    def numParameters(self):
        return len(self.mixing) + sum(len(d.parameters) for d in self.dists)

    def getParameters(self):
        l = list(self.mixing)
        for d in self.dists:
            l.extend(d.parameters)
        return l

    def setParameters(self, v):
        j = len(self.mixing)
        self.mixing = [float(x) for x in v[0:j]]
        i = j
        for d in self.dists:
            j = i + len(d.parameters)
            d.parameters = v[i:j]
            i = j

    parameters = property(getParameters, setParameters)

    def sample_component(self):
        import numpy
        p = numpy.asarray(self.mixing, dtype=float)
        u = p.sum() * numpy.random.rand()
        return int(min(numpy.searchsorted(numpy.cumsum(p), u, side="right"), len(p) - 1))

    def sample_service(self):
        """(service time, component): the forward simulator records the component as the event's auxiliary."""
        i = self.sample_component()
        return float(self.dists[i].sample(1)[0]), i

    def initialize_auxiliaries(self, evt):  # a uniformly random component (queues.pyx:1394-1400)
        import numpy
        evt.set_variable(self.var_name, int(numpy.random.randint(len(self.mixing))))

    def mean(self):
        return sum(w * f.mean() for w, f in zip(self.mixing, self.dists))

    def std(self):  # sic (queues.pyx:1500-1503): the weighted component variances only
        import math
        return math.sqrt(sum(w * f.std() ** 2 for w, f in zip(self.mixing, self.dists)))

    def as_yaml(self):
        ret = "[ MIX, "
        for w, d in zip(self.mixing, self.dists):
            ret += "%.5f, %s, " % (w, d.as_yaml())
        return ret + "]"


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

    def is_mixture(self):
        return isinstance(self.service, MixtureTemplate)

    def queue_order(self):
        return ARRV_BY_A

    def is_markov(self):
        return isinstance(self.service.f, distributions.Exponential)

    def service_lpdf(self, s, component=0):
        return self.service.component_dists()[component].lpdf(s)


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


class QueueGG1R(Queue):
    """One server that takes the next job uniformly at random among those waiting (queues.pyx:788-1027): the
    queue's list is ordered by departure, s = d - max(a, previous departure); the likelihood carries
    - log N(commencement) per event and the rule "started on arrival => arrived to an empty queue"."""
    code = Q_GG1R

    def queue_order(self):
        return ARRV_BY_D

    def as_yaml(self):
        return "  - { name: %s, type: GG1R, service: %s }" % (self.name, self.service.as_yaml())
