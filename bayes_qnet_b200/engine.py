"""GibbsEngine: many independent chains of the reference's sampler on one GPU.

One engine = one ``bq_handle`` = one CUDA device + stream.  It is created from a ``Qnet`` (the model) and an
initialised ``Arrivals`` (the data and the starting state, replicated to every chain); chains differ only in
their RNG counters.  ``sweep`` runs the reference's per-iteration work for all chains in one launch:
``Qnet.slice_resample`` (qnet.pyx:659) followed by ``Qnet.sample_parameters`` (qnet.pyx:882, "BAYES"),
``Qnet.estimate`` (qnet.pyx:844, "STEM") or nothing ("NONE", ``estimation.sample_departures``).
"""
import ctypes

import numpy

from . import capi
from . import sampling

_UPD = {"NONE": capi.BQ_UPD_NONE, "BAYES": capi.BQ_UPD_BAYES, "STEM": capi.BQ_UPD_STEM}
_MODE = {"SLICE": capi.BQ_MODE_SLICE, "MH": capi.BQ_MODE_MH}
SUFFSTAT_FIELDS = ("n", "sum", "sumlog", "sumsq", "nlog", "wait", "nall", "nzero")


def flatten_model(net):
    """The bq_model arrays: per queue discipline / servers / family / parameter offset, and -- when some queue has a
    mixture template -- the distribution slots (one per plain queue, one per mixture component)."""
    q_type = capi.as_i32([q.code for q in net.queues])
    q_k = capi.as_i32([q.k for q in net.queues])
    q_dist = capi.as_i32([q.service.code for q in net.queues])
    npar = [q.service.numParameters() for q in net.queues]
    q_poff = capi.as_i32(numpy.concatenate(([0], numpy.cumsum(npar)[:-1])))
    comps = [q.service.component_codes() for q in net.queues]
    q_slot = capi.as_i32(numpy.concatenate(([0], numpy.cumsum([len(c) for c in comps]))))
    slot_dist = capi.as_i32([f for c in comps for f in c])
    return q_type, q_k, q_dist, q_poff, int(sum(npar)), q_slot, slot_dist


class GibbsEngine(object):
    def __init__(self, net, arrv, n_chains=1, seed=None, device=0, chain_offset=0):
        self.net = net
        self.arrv = arrv
        self.n_chains = int(n_chains)
        self.chain_offset = int(chain_offset)
        self.seed = sampling.next_engine_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        L = capi.lib()
        soa = arrv.soa()
        self.E = len(soa["d"])
        self.Q = len(net.queues)
        q_type, q_k, q_dist, q_poff, self.P, q_slot, slot_dist = flatten_model(net)
        if numpy.any(slot_dist == 3):
            raise NotImplementedError("a Dirac service time has no density (distributions.pyx:347-375): such a network can "
                                      "be simulated, not sampled")
        self._keep = [q_type, q_k, q_dist, q_poff, q_slot, slot_dist]
        self.q_slot = q_slot
        self.NS = int(q_slot[-1])
        self.has_mix = self.NS > self.Q
        m = capi.BqModel(self.Q, self.P, capi.ptr(q_type, capi.c_i32p), capi.ptr(q_k, capi.c_i32p),
                         capi.ptr(q_dist, capi.c_i32p), capi.ptr(q_poff, capi.c_i32p),
                         self.NS if self.has_mix else 0, capi.ptr(q_slot, capi.c_i32p) if self.has_mix else None,
                         capi.ptr(slot_dist, capi.c_i32p) if self.has_mix else None)
        arrs = dict((k, capi.as_i32(soa[k])) for k in ("qid", "tid", "prev_byt", "next_byt", "prev_byq", "next_byq"))
        oa = numpy.ascontiguousarray(soa["obs_a"], dtype=numpy.uint8)
        od = numpy.ascontiguousarray(soa["obs_d"], dtype=numpy.uint8)
        d0 = capi.as_f64(soa["d"])
        a = capi.BqArrivals(self.E, capi.ptr(arrs["qid"], capi.c_i32p), capi.ptr(arrs["tid"], capi.c_i32p),
                            capi.ptr(oa, capi.c_u8p), capi.ptr(od, capi.c_u8p),
                            capi.ptr(arrs["prev_byt"], capi.c_i32p), capi.ptr(arrs["next_byt"], capi.c_i32p),
                            capi.ptr(arrs["prev_byq"], capi.c_i32p), capi.ptr(arrs["next_byq"], capi.c_i32p),
                            capi.ptr(d0, capi.c_f64p))
        theta0 = capi.as_f64(net.parameters)
        h = ctypes.c_void_p()
        capi.check(L.bq_create(ctypes.byref(m), ctypes.byref(a), capi.ptr(theta0, capi.c_f64p), self.n_chains,
                               self.chain_offset, self.seed, int(device), ctypes.byref(h)))
        self._h = h
        self._L = L
        if self.has_mix:  # the events' components travel with the state (Event.auxiliaries)
            self.set_aux(soa["aux"])

    # ---- lifetime -----------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.bq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- sampling -----------------------------------------------------------------------------------
    def sweep(self, n_sweeps=1, mode="SLICE", update="NONE", trace=False, tmax_per_sweep=True, tight=False,
              sample_final=True, sync=True, reverse=False):
        """n_sweeps Gibbs sweeps of every chain, each followed by the parameter update.  reverse: visit the events in
        the reverse of the scan order (slice_resample(reverse=True), qnet.pyx:683-684)."""
        flags = 0
        if tmax_per_sweep:
            flags |= capi.BQ_FLAG_TMAX_PER_SWEEP
        if trace:
            flags |= capi.BQ_FLAG_TRACE
        if tight:
            flags |= capi.BQ_FLAG_TIGHT
        if not sample_final:
            flags |= capi.BQ_FLAG_NO_FINAL
        if reverse:
            flags |= capi.BQ_FLAG_REVERSE
        capi.check(self._L.bq_sweep(self._h, int(n_sweeps), _MODE[mode], _UPD[update], flags))
        if sync:
            self.synchronize()

    def init_chains(self, sigma=0.7):
        """Every chain gets a starting state of its own (new; the reference initialises one state per call of
        Qnet.gibbs_initialize): all hidden departures redrawn on the device, chain by chain on separate RNG streams,
        from exponentials scaled by a per-chain factor exp(sigma N(0,1)), keeping the observed times, the queue
        order and s >= 0.  Static-order networks only (raises BqError otherwise; use `disperse` there)."""
        capi.check(self._L.bq_init_chains(self._h, float(sigma)))

    def init_chains_host(self, n_starts=8, seed=0, kind=None, slice_chains=256):
        """Starting states of their own for the chains of a dynamic-order network (multi-server, processor-sharing or
        random-order queues), where the device initialiser does not apply: `n_starts` independent runs of the host
        initialiser on the handle's data -- same observed times, same task paths, different random draws, hence different
        hidden times AND different queue orders -- chain c starting from run c % n_starts.  The reference initialises one
        state per call of Qnet.gibbs_initialize (qnet.pyx:531-540); this is that call repeated.  Returns the starts
        [n_starts, E]."""
        from . import initialize as _init
        from . import queues as _queues
        codes = [q.code for q in self.net.queues]
        if not any(c in (_queues.Q_PS, _queues.Q_GG1R) or (c == _queues.Q_GGK and q.k > 1)
                   for c, q in zip(codes, self.net.queues)):
            raise ValueError("static-order network: the queue order is shared by all chains of a handle, use init_chains()")
        if kind is None:
            kind = "PS" if _queues.Q_PS in codes else "DFN2"
        soa = self.arrv.soa()
        row_of = dict((int(e), i) for i, e in enumerate(soa["eid"]))
        keep = numpy.random.get_state()
        starts = numpy.empty((int(n_starts), self.E), dtype=numpy.float64)
        try:
            for i in range(int(n_starts)):
                numpy.random.seed((int(seed) * 1000003 + i) % (2 ** 32))
                s2 = _init.initialize(self.net, self.arrv, kind=kind, sample_paths=False).soa()
                rows = numpy.fromiter((row_of[int(e)] for e in s2["eid"]), dtype=numpy.int64, count=self.E)
                starts[i, rows] = s2["d"]
        finally:
            numpy.random.set_state(keep)
        obs = soa["obs_d"] == 1
        assert numpy.all(starts[:, obs] == soa["d"][obs]), "the initialiser moved an observed departure"
        for lo in range(0, self.n_chains, int(slice_chains)):
            hi = min(self.n_chains, lo + int(slice_chains))
            self.set_state(starts[numpy.arange(lo, hi) % int(n_starts)], lo, hi)
        return starts

    def repair_random_order(self, max_sweeps=400, check_every=10):
        """Walk every chain into the support of a random-order network's likelihood (QueueGG1R: an event that starts
        on arrival must have found the queue empty).  gibbs_initialize cannot always satisfy that rule for observed
        jobs whose wait was caused by a hidden departure; here slice sweeps run with the rule softened to a penalty
        (bq_set_relaxed) until every chain's log_prob is finite, then the reference's rule is restored.  Returns the
        number of sweeps used; raises if violations remain after max_sweeps."""
        if numpy.all(numpy.isfinite(self.log_prob())):
            return 0
        capi.check(self._L.bq_set_relaxed(self._h, 1))
        used = 0
        try:
            while used < max_sweeps:
                self.sweep(check_every, "SLICE", "NONE")
                used += check_every
                if numpy.all(numpy.isfinite(self.log_prob())):
                    break
        finally:
            capi.check(self._L.bq_set_relaxed(self._h, 0))
        if not numpy.all(numpy.isfinite(self.log_prob())):
            raise RuntimeError("random-order queues: %d of %d chains still break the start-on-arrival rule after %d repair "
                               "sweeps" % (int(numpy.sum(~numpy.isfinite(self.log_prob()))), self.n_chains, used))
        return used

    def disperse(self, sweeps=20, scale=0.7, seed=0, mode="SLICE"):
        """Over-dispersed starting points for convergence diagnostics (new; the reference is single-chain): every
        chain runs `sweeps` sweeps under its own perturbed parameters theta * exp(scale * N(0, 1)) -- held fixed --
        then gets its parameters back.  Chains that started from one state are then spread wider than the posterior,
        which is what R-hat assumes.  Queue membership and, for static-order networks, the queue order stay those
        of the initial state (they are shared by all chains of a handle)."""
        th = self.params()
        rs = numpy.random.RandomState(seed)
        pert = th * numpy.exp(scale * rs.standard_normal(th.shape))
        # location parameters (LogNormal meanlog) are shifted, not scaled
        col = 0
        for q in self.net.queues:
            if q.service.code == 2:
                pert[:, col] = th[:, col] + scale * rs.standard_normal(self.n_chains)
            col += q.service.numParameters()
        self.set_params(pert)
        self.sweep(sweeps, mode, "NONE")
        self.set_params(th)

    # ---- checkpoint / resume ----------------------------------------------------------------------------
    def save(self, path):
        """Lossless checkpoint (fp64 .npz): departure times and parameters of every chain, seed, chain offset and the
        sweep counter.  `load` into an engine built from the same model and data continues the run exactly (the
        reference restarts from 5-decimal CSV snapshots, gibbs.py:137-160)."""
        ctr = ctypes.c_uint32()
        capi.check(self._L.bq_get_sweep_counter(self._h, ctypes.byref(ctr)))
        numpy.savez_compressed(path, d=self.state(), theta=self.params(), seed=numpy.uint64(self.seed),
                               chain_offset=numpy.int64(self.chain_offset), sweep_counter=numpy.uint32(ctr.value),
                               n_events=numpy.int64(self.E))

    def load(self, path):
        z = numpy.load(path)
        if int(z["n_events"]) != self.E or z["d"].shape != (self.n_chains, self.E):
            raise ValueError("checkpoint %s does not fit this engine (%s vs %d chains x %d events)" % (
                path, z["d"].shape, self.n_chains, self.E))
        if int(z["seed"]) != self.seed or int(z["chain_offset"]) != self.chain_offset:
            raise ValueError("checkpoint was written with seed %d / chain offset %d" % (int(z["seed"]), int(z["chain_offset"])))
        self.set_state(z["d"])
        self.set_params(z["theta"])
        capi.check(self._L.bq_set_sweep_counter(self._h, int(z["sweep_counter"])))

    def synchronize(self):
        capi.check(self._L.bq_synchronize(self._h))

    def last_sweep_ms(self):
        ms = ctypes.c_float()
        capi.check(self._L.bq_last_sweep_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    # ---- parameters and state -------------------------------------------------------------------------
    def params(self, lo=0, hi=None):
        hi = self.n_chains if hi is None else hi
        out = numpy.empty((hi - lo, self.P), dtype=numpy.float64)
        capi.check(self._L.bq_get_params(self._h, lo, hi, capi.ptr(out, capi.c_f64p)))
        return out

    def set_params(self, theta, lo=0, hi=None):
        hi = self.n_chains if hi is None else hi
        th = numpy.ascontiguousarray(numpy.broadcast_to(numpy.asarray(theta, dtype=numpy.float64),
                                                        (hi - lo, self.P)))
        capi.check(self._L.bq_set_params(self._h, lo, hi, capi.ptr(th, capi.c_f64p)))

    def state(self, lo=0, hi=None, full=False):
        """Departure times [hi-lo, E]; with full=True a dict with a, s, proc and the by-queue links too."""
        hi = self.n_chains if hi is None else hi
        n = hi - lo
        d = numpy.empty((n, self.E), dtype=numpy.float64)
        if not full:
            capi.check(self._L.bq_get_state(self._h, lo, hi, capi.ptr(d, capi.c_f64p), None, None, None, None, None))
            return d
        a = numpy.empty_like(d)
        s = numpy.empty_like(d)
        proc = numpy.empty((n, self.E), dtype=numpy.int32)
        pq = numpy.empty((n, self.E), dtype=numpy.int32)
        nq = numpy.empty((n, self.E), dtype=numpy.int32)
        capi.check(self._L.bq_get_state(self._h, lo, hi, capi.ptr(d, capi.c_f64p), capi.ptr(a, capi.c_f64p),
                                        capi.ptr(s, capi.c_f64p), capi.ptr(proc, capi.c_i32p),
                                        capi.ptr(pq, capi.c_i32p), capi.ptr(nq, capi.c_i32p)))
        return dict(d=d, a=a, s=s, proc=proc, prev_byq=pq, next_byq=nq)

    def set_state(self, d, lo=0, hi=None):
        hi = self.n_chains if hi is None else hi
        dd = numpy.ascontiguousarray(numpy.broadcast_to(numpy.asarray(d, dtype=numpy.float64), (hi - lo, self.E)))
        capi.check(self._L.bq_set_state(self._h, lo, hi, capi.ptr(dd, capi.c_f64p)))

    def aux(self, lo=0, hi=None):
        """Mixture component of every event [hi-lo, E] (zeros for plain queues); mixture networks only."""
        hi = self.n_chains if hi is None else hi
        out = numpy.empty((hi - lo, self.E), dtype=numpy.uint8)
        capi.check(self._L.bq_get_aux(self._h, lo, hi, capi.ptr(out, capi.c_u8p)))
        return out

    def set_aux(self, aux, lo=0, hi=None):
        hi = self.n_chains if hi is None else hi
        z = numpy.ascontiguousarray(numpy.broadcast_to(numpy.asarray(aux, dtype=numpy.uint8), (hi - lo, self.E)))
        capi.check(self._L.bq_set_aux(self._h, lo, hi, capi.ptr(z, capi.c_u8p)))

    def broadcast_state(self, d):
        """One chain's departure times [E] replicated to every chain (only E*8 bytes cross PCIe)."""
        dd = capi.as_f64(d)
        assert dd.shape == (self.E,)
        capi.check(self._L.bq_broadcast_state(self._h, capi.ptr(dd, capi.c_f64p)))

    def to_arrivals(self, chain=0):
        """A host ``Arrivals`` holding the current state of one chain."""
        st = self.state(chain, chain + 1, full=True)
        out = self.arrv.duplicate()
        out._d[:] = st["d"][0]
        out._a[:] = st["a"][0]
        out._s[:] = st["s"][0]
        out._proc[:] = st["proc"][0]
        out._prev_byq[:] = st["prev_byq"][0]
        out._next_byq[:] = st["next_byq"][0]
        if self.has_mix:
            out._aux[:] = self.aux(chain, chain + 1)[0]
        return out

    def write_back(self, arrv, chain=0):
        """Copy one chain's state into an existing Arrivals (the reference's samplers mutate in place)."""
        st = self.state(chain, chain + 1, full=True)
        arrv._d[:] = st["d"][0]
        arrv._a[:] = st["a"][0]
        arrv._s[:] = st["s"][0]
        arrv._proc[:] = st["proc"][0]
        arrv._prev_byq[:] = st["prev_byq"][0]
        arrv._next_byq[:] = st["next_byq"][0]
        if self.has_mix:
            arrv._aux[:] = self.aux(chain, chain + 1)[0]

    # ---- parity hooks and reporting ---------------------------------------------------------------------
    def logcond(self, chain, row, xs):
        """GGkGibbs.inner_dfun(x) - lik0 (qnet.pyx:1340-1357) for event `row` of `chain` at the points xs."""
        xs = capi.as_f64(xs)
        out = numpy.empty(len(xs), dtype=numpy.float64)
        capi.check(self._L.bq_logcond(self._h, int(chain), int(row), capi.ptr(xs, capi.c_f64p), len(xs),
                                      capi.ptr(out, capi.c_f64p)))
        return out

    def mh_proposal(self, chain, row, xs):
        """The MH proposal log-density of event `row` (queues.pair_proposal / departure_proposal, queues.pyx:1279)."""
        xs = capi.as_f64(xs)
        out = numpy.empty(len(xs), dtype=numpy.float64)
        capi.check(self._L.bq_mh_proposal(self._h, int(chain), int(row), capi.ptr(xs, capi.c_f64p), len(xs),
                                          capi.ptr(out, capi.c_f64p)))
        return out

    def gibbs_kernel(self, d_to, strict=False):
        """log T(d_to <- chain) of one slice sweep in the scan order, per chain (Qnet.gibbs_kernel, qnet.pyx:741).
        d_to: [E] (one target for every chain) or [n_chains, E].  strict: -inf instead of the reference's retry passes
        when an event's target is outside its support at its turn."""
        d_to = numpy.ascontiguousarray(d_to, dtype=numpy.float64)
        n_to = 1 if d_to.ndim == 1 else d_to.shape[0]
        if d_to.size != n_to * self.E:
            raise ValueError("d_to must have n_events entries per target")
        out = numpy.empty(self.n_chains, dtype=numpy.float64)
        capi.check(self._L.bq_gibbs_kernel(self._h, capi.ptr(d_to, capi.c_f64p), n_to, 1 if strict else 0,
                                            capi.ptr(out, capi.c_f64p)))
        return out

    def log_prob(self):
        out = numpy.empty(self.n_chains, dtype=numpy.float64)
        capi.check(self._L.bq_log_prob(self._h, capi.ptr(out, capi.c_f64p)))
        return out

    def suffstats(self):
        """[n_chains, n_queues, 8]: see SUFFSTAT_FIELDS.  With mixture templates one row per distribution slot
        ([n_chains, n_slots, 8]; slots of queue q: q_slot[q] .. q_slot[q+1]-1)."""
        out = numpy.empty((self.n_chains, self.NS, 8), dtype=numpy.float64)
        capi.check(self._L.bq_suffstats(self._h, capi.ptr(out, capi.c_f64p)))
        return out

    def mean_wait(self):
        ss = self.suffstats()
        w = numpy.add.reduceat(ss[:, :, 5], self.q_slot[:-1], axis=1)
        n = numpy.add.reduceat(ss[:, :, 6], self.q_slot[:-1], axis=1)
        return w / numpy.maximum(n, 1.0)

    def stats(self):
        st = capi.BqStats()
        capi.check(self._L.bq_stats(self._h, ctypes.byref(st)))
        return dict((k, int(getattr(st, k))) for k, _ in capi.BqStats._fields_)

    def reset_stats(self):
        capi.check(self._L.bq_reset_stats(self._h))

    def trace_len(self):
        n = ctypes.c_int32()
        capi.check(self._L.bq_trace_len(self._h, ctypes.byref(n)))
        return int(n.value)

    def traces(self):
        """(theta [n_rec, C, P], mean_wait [n_rec, C, Q], logp [n_rec, C]) recorded by sweep(trace=True)."""
        n = self.trace_len()
        th = numpy.empty((n, self.n_chains, self.P))
        mw = numpy.empty((n, self.n_chains, self.Q))
        lp = numpy.empty((n, self.n_chains))
        if n:
            capi.check(self._L.bq_get_trace(self._h, capi.ptr(th, capi.c_f64p), capi.ptr(mw, capi.c_f64p),
                                            capi.ptr(lp, capi.c_f64p)))
        return th, mw, lp

    def clear_traces(self):
        capi.check(self._L.bq_clear_trace(self._h))

    def schedule(self):
        """(color_off [n_colors+1], order [n_eligible]): the conflict-free sweep schedule."""
        nc, ne = ctypes.c_int32(), ctypes.c_int32()
        capi.check(self._L.bq_get_schedule(self._h, ctypes.byref(nc), ctypes.byref(ne), None, None))
        off = numpy.empty(nc.value + 1, dtype=numpy.int32)
        order = numpy.empty(ne.value, dtype=numpy.int32)
        capi.check(self._L.bq_get_schedule(self._h, None, None, capi.ptr(off, capi.c_i32p), capi.ptr(order, capi.c_i32p)))
        return off, order

    def sequential_order(self):
        """Scan order of MH sweeps (and of slice sweeps on networks with multi-server or PS queues)."""
        ne = ctypes.c_int32()
        capi.check(self._L.bq_get_sequential_order(self._h, ctypes.byref(ne), None))
        order = numpy.empty(ne.value, dtype=numpy.int32)
        capi.check(self._L.bq_get_sequential_order(self._h, None, capi.ptr(order, capi.c_i32p)))
        return order

    def mh_order(self):
        """The order in which MH sweeps visit the events (chromatic schedule for static-order networks)."""
        ne = ctypes.c_int32()
        capi.check(self._L.bq_get_mh_order(self._h, ctypes.byref(ne), None))
        order = numpy.empty(ne.value, dtype=numpy.int32)
        capi.check(self._L.bq_get_mh_order(self._h, None, capi.ptr(order, capi.c_i32p)))
        return order

    def kernel_info(self):
        t, s, m = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        capi.check(self._L.bq_kernel_info(self._h, ctypes.byref(t), ctypes.byref(s), ctypes.byref(m)))
        n = ctypes.c_int32()
        capi.check(self._L.bq_cluster_size(self._h, ctypes.byref(n)))
        return dict(threads=int(t.value), smem_bytes=int(s.value), state_in_smem=bool(m.value), cluster=int(n.value))

    def device_pointers(self):
        """Raw device addresses (state, trace arrays) for zero-copy wrapping by torch / NCCL code."""
        p = ctypes.c_void_p()
        stride = ctypes.c_int64()
        capi.check(self._L.bq_state_dev(self._h, ctypes.byref(p), ctypes.byref(stride)))
        t, w, l = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        capi.check(self._L.bq_trace_dev(self._h, ctypes.byref(t), ctypes.byref(w), ctypes.byref(l)))
        return dict(state=p.value, chain_stride=int(stride.value), tr_theta=t.value, tr_wait=w.value, tr_logp=l.value)
