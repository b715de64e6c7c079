"""ctypes wrapper of oracle/libbq_oracle.so (bq_oracle.c).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg; never by bayes_qnet_b200/."""
import ctypes
import os
import subprocess

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libbq_oracle.so")
SRC = os.path.join(HERE, "bq_oracle.c")

i32p = ctypes.POINTER(ctypes.c_int32)
u8p = ctypes.POINTER(ctypes.c_uint8)
f64p = ctypes.POINTER(ctypes.c_double)


class _State(ctypes.Structure):
    _fields_ = [("E", ctypes.c_int), ("Q", ctypes.c_int), ("q_type", i32p), ("q_k", i32p), ("q_dist", i32p),
                ("q_poff", i32p), ("theta", f64p), ("qid", i32p), ("prev_byt", i32p), ("next_byt", i32p),
                ("obs_a", u8p), ("obs_d", u8p), ("a", f64p), ("d", f64p), ("s", f64p), ("proc", i32p),
                ("prev_byq", i32p), ("next_byq", i32p), ("q_slot", i32p), ("slot_dist", i32p), ("slot_poff", i32p),
                ("aux", u8p)]


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, SRC, "-lm"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.bqo_lpdf.restype = ctypes.c_double
        L.bqo_lpdf.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.bqo_recompute.argtypes = [ctypes.POINTER(_State)]
        L.bqo_log_prob.restype = ctypes.c_double
        L.bqo_log_prob.argtypes = [ctypes.POINTER(_State)]
        L.bqo_logcond.argtypes = [ctypes.POINTER(_State), ctypes.c_int, f64p, ctypes.c_int, f64p]
        L.bqo_apply.argtypes = [ctypes.POINTER(_State), ctypes.c_int, ctypes.c_double]
        L.bqo_eligible.restype = ctypes.c_int
        L.bqo_eligible.argtypes = [ctypes.POINTER(_State), ctypes.c_int]
        L.bqo_slice_sweep.restype = ctypes.c_long
        L.bqo_slice_sweep.argtypes = [ctypes.POINTER(_State), i32p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64,
                                      ctypes.c_uint32, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_long)]
        L.bqo_mh_sweep.restype = ctypes.c_long
        L.bqo_mh_sweep.argtypes = [ctypes.POINTER(_State), i32p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64,
                                   ctypes.c_uint32, ctypes.c_int, ctypes.POINTER(ctypes.c_long),
                                   ctypes.POINTER(ctypes.c_long)]
        L.bqo_mh_proposal.argtypes = [ctypes.POINTER(_State), ctypes.c_int, f64p, ctypes.c_int, f64p]
        L.bqo_suffstats.argtypes = [ctypes.POINTER(_State), f64p]
        L.bqo_estimate.argtypes = [ctypes.c_int, f64p, f64p, f64p]
        L.bqo_sample_params.argtypes = [ctypes.c_int, f64p, ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32,
                                        ctypes.c_uint32, f64p, f64p]
        L.bqo_rng_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_uint32, ctypes.c_int, f64p]
        L.bqo_philox_block.argtypes = [ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_uint32)]
        L.bqo_rng_slice_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_uint32, ctypes.c_int, f64p]
        _lib = L
    return _lib


def lpdf(dist, p0, p1, x):
    return lib().bqo_lpdf(int(dist), float(p0), float(p1), float(x))


def rng_uniforms(seed, chain, sweep, ident, tag, n):
    out = numpy.empty(n)
    lib().bqo_rng_uniforms(seed, chain, sweep, ident, tag, n, out.ctypes.data_as(f64p))
    return out


def rng_slice_uniforms(seed, chain, sweep, ident, tag, n):
    out = numpy.empty(n)
    lib().bqo_rng_slice_uniforms(seed, chain, sweep, ident, tag, n, out.ctypes.data_as(f64p))
    return out


def philox_block(ctr, key, rounds):
    c = (ctypes.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (ctypes.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (ctypes.c_uint32 * 4)()
    lib().bqo_philox_block(c, k, int(rounds), o)
    return [int(x) for x in o]


class OracleChain(object):
    """One chain of the CPU oracle.  Built from flat arrays (row indices as event ids)."""

    def __init__(self, q_type, q_k, q_dist, theta, qid, prev_byt, next_byt, obs_a, obs_d, d, prev_byq, next_byq,
                 comp_dists=None, aux=None):
        """comp_dists: per queue the list of its components' families (a one-element list for a plain queue) when the
        network has mixture templates; aux: component of every event."""
        I = lambda x: numpy.ascontiguousarray(x, dtype=numpy.int32).copy()
        self.q_type, self.q_k, self.q_dist = I(q_type), I(q_k), I(q_dist)
        self.mix = comp_dists is not None and any(len(c) > 1 for c in comp_dists)
        if self.mix:
            ncomp = [len(c) for c in comp_dists]
            self.q_slot = I(numpy.concatenate(([0], numpy.cumsum(ncomp))))
            self.slot_dist = I([f for c in comp_dists for f in c])
            self.q_mixoff, self.slot_poff, off, poff = [], [], 0, []
            for c in comp_dists:
                poff.append(off)
                if len(c) > 1:
                    self.q_mixoff.append(off)
                    off += len(c)
                else:
                    self.q_mixoff.append(-1)
                for f in c:
                    self.slot_poff.append(off)
                    off += 1 if f == 0 else 2
            self.slot_poff, self.q_poff = I(self.slot_poff), I(poff)
            self.q_dist = I([c[0] for c in comp_dists])
            self.aux = numpy.ascontiguousarray(numpy.zeros(len(d)) if aux is None else aux, dtype=numpy.uint8).copy()
        else:
            npar = numpy.where(self.q_dist == 0, 1, 2)
            self.q_poff = I(numpy.concatenate(([0], numpy.cumsum(npar)[:-1])))
            self.q_slot = I(numpy.arange(len(self.q_type) + 1))
            self.slot_dist, self.slot_poff = self.q_dist, self.q_poff
            self.q_mixoff = [-1] * len(self.q_type)
            self.aux = None
        self.theta = numpy.ascontiguousarray(theta, dtype=numpy.float64).copy()
        self.qid, self.prev_byt, self.next_byt = I(qid), I(prev_byt), I(next_byt)
        self.obs_a = numpy.ascontiguousarray(obs_a, dtype=numpy.uint8).copy()
        self.obs_d = numpy.ascontiguousarray(obs_d, dtype=numpy.uint8).copy()
        self.d = numpy.ascontiguousarray(d, dtype=numpy.float64).copy()
        self.E, self.Q = len(self.d), len(self.q_type)
        self.a = numpy.zeros(self.E)
        self.s = numpy.zeros(self.E)
        self.proc = numpy.zeros(self.E, dtype=numpy.int32)
        self.prev_byq, self.next_byq = I(prev_byq), I(next_byq)
        self._st = _State(self.E, self.Q, self.q_type.ctypes.data_as(i32p), self.q_k.ctypes.data_as(i32p),
                          self.q_dist.ctypes.data_as(i32p), self.q_poff.ctypes.data_as(i32p),
                          self.theta.ctypes.data_as(f64p), self.qid.ctypes.data_as(i32p),
                          self.prev_byt.ctypes.data_as(i32p), self.next_byt.ctypes.data_as(i32p),
                          self.obs_a.ctypes.data_as(u8p), self.obs_d.ctypes.data_as(u8p),
                          self.a.ctypes.data_as(f64p), self.d.ctypes.data_as(f64p), self.s.ctypes.data_as(f64p),
                          self.proc.ctypes.data_as(i32p), self.prev_byq.ctypes.data_as(i32p),
                          self.next_byq.ctypes.data_as(i32p), self.q_slot.ctypes.data_as(i32p),
                          self.slot_dist.ctypes.data_as(i32p), self.slot_poff.ctypes.data_as(i32p),
                          self.aux.ctypes.data_as(u8p) if self.mix else None)
        self.recompute()

    @classmethod
    def from_arrivals(cls, net, arrv, theta=None):
        soa = arrv.soa()
        comps = [q.service.component_codes() for q in net.queues]
        mix = any(len(c) > 1 for c in comps)
        return cls([q.code for q in net.queues], [q.k for q in net.queues], [q.service.code for q in net.queues],
                   net.parameters if theta is None else theta, soa["qid"], soa["prev_byt"], soa["next_byt"],
                   soa["obs_a"], soa["obs_d"], soa["d"], soa["prev_byq"], soa["next_byq"],
                   comp_dists=comps if mix else None, aux=soa["aux"] if mix else None)

    # ---- mixture templates (MixtureTemplate, queues.pyx:1373-1509), restated in Python on top of the C density ----
    def resample_aux(self, seed, chain, sweep):
        """z_e ~ Cat(pi_i exp(lpdf_i(s_e))) for every event of every mixture queue, from uniform 0 of the stream
        (chain, sweep, event, tag 6); all weights below 1e-50 in total: a uniformly random component
        (_roll_die_unnorm, sampling.pyx:215-229)."""
        if not self.mix:
            return
        for e in range(self.E):
            q = int(self.qid[e])
            s0, nc = int(self.q_slot[q]), int(self.q_slot[q + 1] - self.q_slot[q])
            if nc < 2:
                continue
            w = []
            for c in range(nc):
                o, f = int(self.slot_poff[s0 + c]), int(self.slot_dist[s0 + c])
                lp = lpdf(f, self.theta[o], self.theta[o + 1] if f != 0 else 0.0, self.s[e])
                w.append(self.theta[self.q_mixoff[q] + c] * numpy.exp(lp))
            u = rng_uniforms(seed, chain, sweep, e, 6, 1)[0]
            Z = sum(w)
            if Z < 1e-50:
                pick = min(int(u * nc), nc - 1)
            else:
                t, S, pick = Z * u, 0.0, nc - 1
                for c in range(nc):
                    S += w[c]
                    if t < S:
                        pick = c
                        break
            self.aux[e] = pick

    def slot_suffstats(self):
        """[NS, 8] in engine.SUFFSTAT_FIELDS order, each slot over the events currently assigned to it."""
        out = numpy.zeros((len(self.slot_dist), 8))
        for e in range(self.E):
            q = int(self.qid[e])
            sl = int(self.q_slot[q]) + (int(self.aux[e]) if self.mix else 0)
            f, s = int(self.slot_dist[sl]), self.s[e]
            w = self.d[e] - self.a[e] - s
            o = out[sl]
            o[5] += w if w > 0 else 0.0
            o[6] += 1.0
            if s > 0:
                o[0] += 1.0; o[1] += s
                if f == 1 or (f == 2 and s >= 1e-50):
                    o[2] += numpy.log(s); o[4] += 1.0
            elif s == 0:
                o[7] += 1.0
        for sl in range(len(self.slot_dist)):
            if self.slot_dist[sl] == 2 and out[sl, 4] > 0:
                mean = out[sl, 2] / out[sl, 4]
                q = int(numpy.searchsorted(self.q_slot, sl, side="right") - 1)
                for e in range(self.E):
                    if self.qid[e] == q and (not self.mix or self.q_slot[q] + self.aux[e] == sl) and self.s[e] >= 1e-50:
                        out[sl, 3] += (numpy.log(self.s[e]) - mean) ** 2
        return out

    def set_theta(self, theta):
        self.theta[:] = theta

    def recompute(self):
        lib().bqo_recompute(ctypes.byref(self._st))

    def log_prob(self):
        return lib().bqo_log_prob(ctypes.byref(self._st))

    def logcond(self, e, xs):
        xs = numpy.ascontiguousarray(xs, dtype=numpy.float64)
        out = numpy.empty(len(xs))
        lib().bqo_logcond(ctypes.byref(self._st), int(e), xs.ctypes.data_as(f64p), len(xs), out.ctypes.data_as(f64p))
        return out

    def apply(self, e, x):
        lib().bqo_apply(ctypes.byref(self._st), int(e), float(x))

    def eligible(self):
        return numpy.array([e for e in range(self.E) if lib().bqo_eligible(ctypes.byref(self._st), e)], dtype=numpy.int32)

    def slice_sweep(self, order, seed, chain, sweep, tmax=None):
        order = numpy.ascontiguousarray(order, dtype=numpy.int32)
        if tmax is None:
            tmax = 2.0 * self.d.max()  # qnet.pyx:672
        ev = ctypes.c_long(0)
        ne = lib().bqo_slice_sweep(ctypes.byref(self._st), order.ctypes.data_as(i32p), len(order), seed, chain, sweep,
                                   float(tmax), 0, ctypes.byref(ev))
        return int(ne), int(ev.value)

    def mh_sweep(self, order, seed, chain, sweep, sample_final=True):
        """One sweep of Qnet.gibbs_resample; returns (proposals, rejections)."""
        order = numpy.ascontiguousarray(order, dtype=numpy.int32)
        rej, ev = ctypes.c_long(0), ctypes.c_long(0)
        ne = lib().bqo_mh_sweep(ctypes.byref(self._st), order.ctypes.data_as(i32p), len(order), seed, chain, sweep,
                                1 if sample_final else 0, ctypes.byref(rej), ctypes.byref(ev))
        return int(ne), int(rej.value)

    def mh_proposal(self, e, xs):
        xs = numpy.ascontiguousarray(xs, dtype=numpy.float64)
        out = numpy.empty(len(xs))
        lib().bqo_mh_proposal(ctypes.byref(self._st), int(e), xs.ctypes.data_as(f64p), len(xs), out.ctypes.data_as(f64p))
        return out

    def gibbs_kernel(self, order, d_to, strict=False):
        """log T(d_to <- this state) of one slice sweep over `order` (Qnet.gibbs_kernel); moves the chain to d_to."""
        order = numpy.ascontiguousarray(order, dtype=numpy.int32)
        d_to = numpy.ascontiguousarray(d_to, dtype=numpy.float64)
        f = lib().bqo_gibbs_kernel
        f.restype = ctypes.c_double
        return float(f(ctypes.byref(self._st), order.ctypes.data_as(i32p), len(order), d_to.ctypes.data_as(f64p),
                       1 if strict else 0))

    def suffstats(self):
        out = numpy.zeros((self.Q, 8))
        lib().bqo_suffstats(ctypes.byref(self._st), out.ctypes.data_as(f64p))
        return out

    def update_params(self, update, seed=0, chain=0, sweep=0):
        """update in {'BAYES','STEM'}: applies Qnet.sample_parameters / Qnet.estimate to self.theta (per distribution
        slot; mixing proportions (1 + count_i) / (n_components + N), queues.pyx:1416-1451)."""
        ss = self.slot_suffstats() if self.mix else self.suffstats()
        for sl in range(len(self.slot_dist)):
            o, f = int(self.slot_poff[sl]), int(self.slot_dist[sl])
            p0 = ctypes.c_double(self.theta[o])
            p1 = ctypes.c_double(self.theta[o + 1] if f != 0 else 0.0)
            row = numpy.ascontiguousarray(ss[sl])
            if update == "STEM":
                lib().bqo_estimate(f, row.ctypes.data_as(f64p), ctypes.byref(p0), ctypes.byref(p1))
            else:
                lib().bqo_sample_params(f, row.ctypes.data_as(f64p), seed, chain, sweep, sl,
                                        ctypes.byref(p0), ctypes.byref(p1))
            self.theta[o] = p0.value
            if f != 0:
                self.theta[o + 1] = p1.value
        if self.mix:
            for q in range(self.Q):
                mo, s0, nc = self.q_mixoff[q], int(self.q_slot[q]), int(self.q_slot[q + 1] - self.q_slot[q])
                if mo < 0:
                    continue
                N = nc + ss[s0:s0 + nc, 6].sum()
                for c in range(nc):
                    self.theta[mo + c] = (1.0 + ss[s0 + c, 6]) / N
        return self.theta.copy()
