"""ctypes wrapper of tests/emul/libbq_emul.so: the product's dynamic-order walk / commit routines
(bayes_qnet_b200/csrc/bq_dynamic.cuh, __host__ __device__) compiled for the host.  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "bq_emul.cu")
SO = os.path.join(HERE, "emul", "libbq_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "bayes_qnet_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

i32p = ctypes.POINTER(ctypes.c_int32)
u8p = ctypes.POINTER(ctypes.c_uint8)
f64p = ctypes.POINTER(ctypes.c_double)


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not os.path.exists(SO) or any(os.path.getmtime(f) > os.path.getmtime(SO) for f in deps):
        subprocess.check_call([NVCC, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
                               "-shared", "-o", SO, SRC])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.bqe_create.restype = ctypes.c_void_p
        L.bqe_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p, i32p, i32p, f64p, i32p, i32p,
                                 i32p, i32p, i32p, u8p, u8p, f64p]
        L.bqe_destroy.argtypes = [ctypes.c_void_p]
        L.bqe_set_theta.argtypes = [ctypes.c_void_p, f64p]
        L.bqe_logcond.argtypes = [ctypes.c_void_p, ctypes.c_int, f64p, ctypes.c_int, f64p]
        L.bqe_apply.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
        L.bqe_slice_sweep.restype = ctypes.c_long
        L.bqe_slice_sweep.argtypes = [ctypes.c_void_p, i32p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64,
                                      ctypes.c_uint32, ctypes.c_double]
        L.bqe_mh_sweep.restype = ctypes.c_long
        L.bqe_mh_sweep.argtypes = [ctypes.c_void_p, i32p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32,
                                   ctypes.c_int, ctypes.POINTER(ctypes.c_long)]
        L.bqe_mh_proposal.argtypes = [ctypes.c_void_p, ctypes.c_int, f64p, ctypes.c_int, f64p]
        L.bqe_sweep_batch.restype = ctypes.c_long
        L.bqe_sweep_batch.argtypes = [ctypes.c_void_p, i32p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32,
                                      ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long)]
        L.bqe_get.argtypes = [ctypes.c_void_p, f64p, i32p, f64p, f64p]
        L.bqe_derive.argtypes = [ctypes.c_void_p, f64p, f64p, i32p, i32p, i32p]
        L.bqe_check.restype = ctypes.c_int
        L.bqe_check.argtypes = [ctypes.c_void_p]
        L.bqe_evals.restype = ctypes.c_long
        L.bqe_evals.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


class EmulChain(object):
    def __init__(self, net, arrv, theta=None):
        soa = arrv.soa()
        I = lambda x: numpy.ascontiguousarray(x, dtype=numpy.int32)
        self.q_type = I([q.code for q in net.queues])
        self.q_k = I([q.k for q in net.queues])
        self.q_dist = I([q.service.code for q in net.queues])
        npar = numpy.where(self.q_dist == 0, 1, 2)
        self.q_poff = I(numpy.concatenate(([0], numpy.cumsum(npar)[:-1])))
        self.theta = numpy.ascontiguousarray(net.parameters if theta is None else theta, dtype=numpy.float64)
        self.E, self.Q, self.P = len(soa["d"]), len(self.q_type), len(self.theta)
        arrs = [I(soa[k]) for k in ("qid", "prev_byt", "next_byt", "prev_byq", "next_byq")]
        oa = numpy.ascontiguousarray(soa["obs_a"], dtype=numpy.uint8)
        od = numpy.ascontiguousarray(soa["obs_d"], dtype=numpy.uint8)
        d = numpy.ascontiguousarray(soa["d"], dtype=numpy.float64)
        P = lambda a, t: a.ctypes.data_as(t)
        self._h = lib().bqe_create(self.E, self.Q, self.P, P(self.q_type, i32p), P(self.q_k, i32p),
                                   P(self.q_dist, i32p), P(self.q_poff, i32p), P(self.theta, f64p),
                                   *([P(a, i32p) for a in arrs] + [P(oa, u8p), P(od, u8p), P(d, f64p)]))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bqe_destroy(self._h)
            self._h = None

    def set_theta(self, theta):
        th = numpy.ascontiguousarray(theta, dtype=numpy.float64)
        lib().bqe_set_theta(self._h, th.ctypes.data_as(f64p))

    def logcond(self, e, xs):
        xs = numpy.ascontiguousarray(xs, dtype=numpy.float64)
        out = numpy.empty(len(xs))
        lib().bqe_logcond(self._h, int(e), xs.ctypes.data_as(f64p), len(xs), out.ctypes.data_as(f64p))
        return out

    def apply(self, e, x):
        lib().bqe_apply(self._h, int(e), float(x))

    def slice_sweep(self, order, seed, chain, sweep, tmax=None):
        order = numpy.ascontiguousarray(order, dtype=numpy.int32)
        if tmax is None:
            tmax = 2.0 * self.d.max()
        return int(lib().bqe_slice_sweep(self._h, order.ctypes.data_as(i32p), len(order), seed, chain, sweep,
                                         float(tmax)))

    def mh_sweep(self, order, seed, chain, sweep, sample_final=True):
        order = numpy.ascontiguousarray(order, dtype=numpy.int32)
        rej = ctypes.c_long(0)
        ne = lib().bqe_mh_sweep(self._h, order.ctypes.data_as(i32p), len(order), seed, chain, sweep,
                                1 if sample_final else 0, ctypes.byref(rej))
        return int(ne), int(rej.value)

    def sweep_batch(self, order, seed, chain, sweep, B, mh=False, sample_final=True, tmax=None):
        """One sweep with the batch semantics of sweep_batch_kernel; returns (resamples, rejections, redone)."""
        order = numpy.ascontiguousarray(order, dtype=numpy.int32)
        if tmax is None:
            tmax = 2.0 * self.d.max()
        rej, redo = ctypes.c_long(0), ctypes.c_long(0)
        ne = lib().bqe_sweep_batch(self._h, order.ctypes.data_as(i32p), len(order), seed, chain, sweep, float(tmax),
                                   int(B), 1 if mh else 0, 1 if sample_final else 0, ctypes.byref(rej), ctypes.byref(redo))
        return int(ne), int(rej.value), int(redo.value)

    def mh_proposal(self, e, xs):
        xs = numpy.ascontiguousarray(xs, dtype=numpy.float64)
        out = numpy.empty(len(xs))
        lib().bqe_mh_proposal(self._h, int(e), xs.ctypes.data_as(f64p), len(xs), out.ctypes.data_as(f64p))
        return out

    @property
    def d(self):
        out = numpy.empty(self.E)
        lib().bqe_get(self._h, out.ctypes.data_as(f64p), None, None, None)
        return out

    @property
    def ord(self):
        out = numpy.empty(self.E, dtype=numpy.int32)
        lib().bqe_get(self._h, None, out.ctypes.data_as(i32p), None, None)
        return out

    @property
    def svc(self):
        out = numpy.empty(self.E)
        lib().bqe_get(self._h, None, None, None, out.ctypes.data_as(f64p))
        return out

    def derive(self):
        a, s = numpy.empty(self.E), numpy.empty(self.E)
        proc, pq, nq = (numpy.empty(self.E, dtype=numpy.int32) for _ in range(3))
        lib().bqe_derive(self._h, a.ctypes.data_as(f64p), s.ctypes.data_as(f64p), proc.ctypes.data_as(i32p),
                         pq.ctypes.data_as(i32p), nq.ctypes.data_as(i32p))
        return dict(a=a, s=s, proc=proc, prev_byq=pq, next_byq=nq)

    def check(self):
        return int(lib().bqe_check(self._h))

    def evals(self):
        return int(lib().bqe_evals(self._h))
