"""Routing FSM of a queueing network (host side only): which states follow which, which queues a state
may emit.  Same surface as the reference's ``hmm.HMM`` (hmm.py:3-66); uniform probabilities are filled in by
``qnetu.qnet_from_text`` (qnetu.py:63-74).  Used by the forward simulator and the initializer -- never by the
kernels: queue membership of every event is fixed before sampling starts.
"""
import numpy


def roll_die(p):
    """Index drawn from the probability vector p (utils.py:14-24)."""
    p = numpy.asarray(p, dtype=numpy.float64)
    c = numpy.cumsum(p)
    return int(min(numpy.searchsorted(c, numpy.random.rand() * c[-1], side="right"), len(p) - 1))


class HMM(object):
    def __init__(self, a, o):
        self.a = numpy.asarray(a, dtype=numpy.float64)
        self.o = numpy.asarray(o, dtype=numpy.float64)
        self.ns = self.a.shape[0]
        self.n_obs = self.o.shape[1]
        assert self.a.ndim == 2 and self.o.ndim == 2
        assert self.a.shape[0] == self.a.shape[1] == self.o.shape[0]

    def initial_state(self):
        return 0

    def is_final(self, state):
        return self.a[state, :].sum() < 1e-10

    def num_states(self):
        return self.ns

    def successors(self, sid):
        return set(int(i) for i in numpy.nonzero(self.a[sid, :] > 0)[0])

    def possible_observations(self, sid):
        return set(int(i) for i in numpy.nonzero(self.o[sid, :] > 0)[0])

    def sample_one_obs(self, state):
        return roll_die(self.o[state, :])

    def sample_state(self, state):
        if self.is_final(state):
            return None
        return roll_die(self.a[state, :])
