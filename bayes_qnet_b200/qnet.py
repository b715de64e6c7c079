"""``Qnet``: the queueing-network model object, with the reference's method surface (qnet.pyx:143).

What runs where:
* ``slice_resample`` / ``gibbs_resample`` / ``sample_parameters`` / ``estimate`` / ``log_prob`` -- the hot
  path -- go to the GPU through ``GibbsEngine`` (C-ABI -> CUDA kernels).  No CPU implementation exists here.
* ``sample`` (forward simulation, qnet.pyx:184-298) and ``gibbs_initialize`` (qnet.pyx:531) run once per data
  set on the host; they only produce the input of the hot path.
Module-level switches mirror the reference's globals (qnet.pyx:55-109): ``set_sampler``,
``set_initialization_type``, ``set_slice_sorted``, ``set_use_rtp``, ``stats``/``reset_stats``.
"""
import heapq
import weakref

import numpy

from . import arrivals as _arrivals
from . import queues as _queues
from . import initialize as _initialize
from .engine import GibbsEngine

INITIAL_QID = 0
INITIAL_STATE = 0

# ---- configuration (qnet.pyx:55-109) ---------------------------------------------------------------
_sampler = "SLICE"
_initializer = "DFN2"
_use_rtp = 0
_slice_sorted = 0


def set_sampler(name):
    """one_d_sampler of the MH path (qnet.pyx:71).  Only SLICE is built; the others raise."""
    global _sampler
    if name == "SLICE":
        _sampler = name
    elif name in ("SLICE_PWFUN", "ARS_PWFUN", "STUPID_REJECT"):
        raise NotImplementedError("one_d_sampler %s is outside the GPU hot path (SURVEY.md section 2 rows 5-6)" % name)
    else:
        raise Exception("Could not understand one_d_sampler: %s" % name)


def set_use_rtp(doitp):
    global _use_rtp
    if int(doitp):
        raise NotImplementedError("remove_task_proposal (qnet.pyx:922) is not part of the GPU hot path")
    _use_rtp = 0


def set_initialization_type(type_):
    global _initializer
    t = type_.upper() if isinstance(type_, str) else type_
    if callable(t) or t in ("DFN1", "DFN2", "PS", "NONE"):
        _initializer = t
    elif t in ("LP", "MINILP"):
        raise NotImplementedError("LP initializers depend on a solver that the reference itself stubs out (stupidlp.py:6)")
    else:
        raise Exception("Cannot understand type %s" % type_)


def get_initializer():
    return _initializer


def set_slice_sorted(val):
    global _slice_sorted
    if int(val):
        raise NotImplementedError("slice_sorted: the kernels use their own conflict-free scan order")
    _slice_sorted = 0


# ---- statistics (qnet.pyx:115-132) -------------------------------------------------------------------
_stats = dict(nreject=0, nevents=0, nrtp_calls=0, nrtp_reject=0, diff_len=0, num_diff=0)


def reset_stats():
    for k in _stats:
        _stats[k] = 0


def stats():
    return dict(_stats)


def _absorb(before, after):
    for k in ("nevents", "nreject", "num_diff", "diff_len"):
        _stats[k] += after[k] - before[k]


class Qnet(object):
    def __init__(self, queues, templates, universe, qname2id, sname2id, fsm):
        self.queues = queues
        self.templates = templates
        self.universe = universe
        self.fsm = fsm
        self.qname2id = qname2id
        self.sname2id = sname2id
        self.eidx = -1
        self._engines = weakref.WeakKeyDictionary()

    # ---- description ---------------------------------------------------------------------------------
    def as_yaml(self):
        ret = "states:\n"
        for si in range(self.fsm.num_states()):
            ret += "  - name: %s\n" % self.sid_to_name(si)
            ret += "    queues: [ %s ]\n" % ",".join(q.name for q in self.queues_of_state(si))
            succ = [self.sid_to_name(s) for s in sorted(self.fsm.successors(si))]
            if succ:
                ret += "    successors: [ %s ]\n" % ",".join(succ)
        ret += "queues:\n"
        for q in self.queues:
            ret += q.as_yaml() + "\n"
        return ret

    def num_queues(self): return len(self.queues)
    def num_states(self): return self.fsm.num_states()
    def get_fsm(self): return self.fsm
    def sid_by_name(self, sname): return self.sname2id.get(sname, None)

    def sid_to_name(self, sid):
        for k, v in self.sname2id.items():
            if v == sid:
                return k
        raise KeyError("No state with id %s" % sid)

    def queues_of_state(self, sid):
        return [self.queues[qi] for qi in sorted(self.fsm.possible_observations(sid))]

    def qid_of_queue(self, queue):
        for i, q in enumerate(self.queues):
            if q is queue:
                return i
        raise Exception("Can't find %s in net %s" % (queue, self))

    def queue_by_name(self, qname):
        qid = self.qname2id.get(qname, None)
        return self.queues[qid] if qid is not None else None

    def all_queues(self): return self.queues[:]
    def all_templates(self): return self.templates[:]

    @property
    def parameters(self):  # qnet.pyx:898-915
        r = []
        for q in self.queues:
            r.extend(q.service.parameters)
        return r

    @parameters.setter
    def parameters(self, v):
        r = list(v)
        for q in self.queues:
            n = q.service.numParameters()
            q.service.setParameters(r[0:n])
            r = r[n:]

    # ---- forward simulation (host; qnet.pyx:184-298) ---------------------------------------------------
    def sample(self, n):
        """Generate n tasks: all of them queue at the source at t = 0 (qnet.pyx:184-193)."""
        return _simulate(self, int(n))

    def sample_path(self):  # qnet.pyx:1023-1030
        path = []
        sid = self.fsm.sample_state(self.fsm.initial_state())
        while sid:
            path.append((sid, self.fsm.sample_one_obs(sid)))
            sid = self.fsm.sample_state(sid)
        return path

    # ---- initialisation (host; qnet.pyx:531) -------------------------------------------------------------
    def gibbs_initialize(self, a, samplePaths=False):
        init = _initializer
        if callable(init):
            return init(self, a)
        if init == "NONE":
            return a
        return _initialize.initialize(self, a, kind=init)

    # ---- the hot path: delegated to the GPU ----------------------------------------------------------------
    def engine(self, arrv, n_chains=1, seed=None, device=0, chain_offset=0):
        """A fresh multi-chain GPU engine started from arrv (new API; the reference is single-chain)."""
        return GibbsEngine(self, arrv, n_chains=n_chains, seed=seed, device=device, chain_offset=chain_offset)

    def _engine_for(self, arrv):
        """The single-chain engine behind the in-place reference API, cached per Arrivals object and rebuilt when
        the structure the engine froze at creation (observation flags, task and queue links, queue membership)
        has changed since."""
        soa = arrv.soa()
        fp = _fingerprint(soa)
        hit = self._engines.get(arrv)
        if hit is None or hit[1] != fp:
            if hit is not None:
                hit[0].close()
            eng = GibbsEngine(self, arrv, n_chains=1)
            eng.arrv = weakref.proxy(arrv)  # no strong reference: the engine (and its GPU memory) goes with the Arrivals
            self._engines[arrv] = (eng, fp)
        else:
            eng = hit[0]
            eng.set_state(soa["d"])
        if eng.has_mix:
            eng.set_aux(soa["aux"])
        eng.set_params(self.parameters)
        return eng

    def slice_resample(self, arrv, num_iter, report_fn=None, return_arrv=False, reverse=False, return_lp=False):
        """num_iter exact-conditional slice sweeps over the hidden times of arrv, in place (qnet.pyx:659).
        reverse: the kernel's scan order backwards (qnet.pyx:683-684)."""
        eng = self._engine_for(arrv)
        before = eng.stats()
        all_arrv = []
        lik = None
        if report_fn or return_arrv:
            for gi in range(num_iter):
                # Tmax is fixed for the whole call (qnet.pyx:672): only the first launch recomputes it
                eng.sweep(1, "SLICE", "NONE", tmax_per_sweep=(gi == 0), reverse=reverse)
                eng.write_back(arrv)
                lik = float(eng.log_prob()[0])
                if report_fn:
                    report_fn(self, arrv, gi, lik)
                if return_arrv:
                    all_arrv.append(arrv.duplicate())
        else:
            eng.sweep(num_iter, "SLICE", "NONE", tmax_per_sweep=False, reverse=reverse)
            eng.write_back(arrv)
        _absorb(before, eng.stats())
        if not return_arrv:
            all_arrv = [arrv]
        if return_lp:
            if lik is None:
                lik = float(eng.log_prob()[0])
            return all_arrv, lik
        return all_arrv

    def gibbs_resample(self, arrv, burn, num_iter, report_fn=None, return_arrv=True, sample_final=True):
        """MH-within-Gibbs sweeps (qnet.pyx:337)."""
        eng = self._engine_for(arrv)
        before = eng.stats()
        all_arrv = []
        for gi in range(burn + num_iter):
            eng.sweep(1, "MH", "NONE", sample_final=sample_final)
            if burn <= gi and (report_fn or return_arrv):
                eng.write_back(arrv)
                if report_fn:
                    report_fn(self, arrv, gi)
                if return_arrv:
                    all_arrv.append(arrv.duplicate())
        eng.write_back(arrv)
        _absorb(before, eng.stats())
        if not return_arrv:
            all_arrv = [arrv]
        return all_arrv

    def _update_parameters(self, arrvl, update):
        if isinstance(arrvl, _arrivals.Arrivals):
            arrvl = [arrvl]
        if len(arrvl) == 0:
            raise ValueError("no Arrivals given")
        if len(arrvl) == 1:
            eng = self._engine_for(arrvl[0])
            capi_update_only(eng, update)
        else:
            # several Arrivals: the events of all of them are pooled per queue (qnet.pyx:844-858, 882-896); the
            # per-Arrivals sufficient statistics come from the GPU, their sum goes back for the update
            from . import capi
            parts = [self._engine_for(a).suffstats()[0] for a in arrvl]  # each [Q, 8]
            pooled = pool_suffstats(parts)
            eng = self._engine_for(arrvl[0])
            capi.check(eng._L.bq_update_params_from(eng._h, {"BAYES": capi.BQ_UPD_BAYES, "STEM": capi.BQ_UPD_STEM}[update],
                                                    capi.ptr(pooled, capi.c_f64p), 1))
        self.parameters = list(eng.params(0, 1)[0])
        return self.parameters

    def sample_parameters(self, arrvl):
        """Posterior draw of every queue's service parameters given complete data (qnet.pyx:882)."""
        return self._update_parameters(arrvl, "BAYES")

    def estimate(self, arrvl):
        """ML estimate of every queue's service parameters given complete data (qnet.pyx:844)."""
        return self._update_parameters(arrvl, "STEM")

    def log_prob(self, arrv):
        """Sum over events of the service log-density (qnet.pyx:1032)."""
        return float(self._engine_for(arrv).log_prob()[0])

    def gibbs_kernel(self, arrv_from, arrv_to, lik=None, strict=False):
        """log T(arrv_to <- arrv_from) of one slice sweep in the kernel's scan order (qnet.pyx:741-809): minus the sum
        over the resampled events of the log normaliser of their conditional, each event moved to its target before
        the next.  -inf when the target cannot be reached.  Like the reference, events whose target is not reachable at
        their turn are retried after the others; strict=True returns -inf instead (the density of the sweep itself)."""
        return float(self._engine_for(arrv_from).gibbs_kernel(arrv_to.soa()["d"], strict=strict)[0])

    def parameter_kernel(self, theta_from, theta_to, arrvl):
        """Log-density of the transition sample_parameters makes from theta_from to theta_to given arrvl (qnet.pyx:811-837)."""
        if isinstance(arrvl, _arrivals.Arrivals):
            arrvl = [arrvl]
        ret, off = 0.0, 0
        for q in self.queues:
            s_obs = [e.s for a in arrvl for e in a.events_of_queue(q) if e.s > 0]  # queues.pyx:1320-1322
            n = q.service.numParameters()
            ret += q.service.f.parameter_kernel(list(theta_from[off:off + n]), list(theta_to[off:off + n]), s_obs)
            off += n
        return ret


def _fingerprint(soa):
    """Hash of everything a GibbsEngine fixes at creation."""
    import hashlib
    h = hashlib.blake2b(digest_size=16)
    for k in ("qid", "tid", "obs_a", "obs_d", "prev_byt", "next_byt", "prev_byq", "next_byq"):
        h.update(numpy.ascontiguousarray(soa[k]).tobytes())
    return h.digest()


def pool_suffstats(parts):
    """Sum of per-Arrivals sufficient statistics [Q, 8] (engine.SUFFSTAT_FIELDS order: n, sum, sumlog, sumsq, nlog,
    wait, nall, nzero); the squared deviations are re-centred on the pooled mean of the logs (the reference's
    LogNormal estimate is two-pass over the pooled list, distributions.pyx:276-299)."""
    parts = [numpy.asarray(p, dtype=numpy.float64) for p in parts]
    out = numpy.sum(parts, axis=0)
    nlog = out[:, 4]
    mean = numpy.where(nlog > 0, out[:, 2] / numpy.maximum(nlog, 1.0), 0.0)
    sq = numpy.zeros(out.shape[0])
    for p in parts:
        m_i = numpy.where(p[:, 4] > 0, p[:, 2] / numpy.maximum(p[:, 4], 1.0), 0.0)
        sq += p[:, 3] + p[:, 4] * (m_i - mean) ** 2
    out[:, 3] = sq
    return numpy.ascontiguousarray(out)


def capi_update_only(eng, update):
    """Run just the parameter update (no hidden-time sweep) on an engine."""
    from . import capi
    capi.check(eng._L.bq_update_params(eng._h, {"BAYES": capi.BQ_UPD_BAYES, "STEM": capi.BQ_UPD_STEM}[update]))
    eng.synchronize()


# ---------------------------------------------------------------------------------------------------------
# forward simulator
# ---------------------------------------------------------------------------------------------------------
def _simulate(net, n):
    """Discrete-event simulation of n tasks (qnet.pyx:228-298).

    Arrivals are processed in time order, so FIFO stations can be resolved at arrival time from the
    per-server free times; PS stations keep their active set and re-plan their next departure whenever the
    set changes (the reference re-queues stale departures instead, qnet.pyx:280-290).
    """
    nq = len(net.queues)
    fsm = net.fsm
    free = [[0.0] * q.k for q in net.queues]
    # PS state per queue: clock, {row: remaining work}, version
    ps_clock = [0.0] * nq
    ps_active = [dict() for _ in range(nq)]
    ps_version = [0] * nq
    # random-order queues (QueueGG1R, select_job_for_service: queues.pyx:1026-1027): jobs waiting, server busy
    rs_wait = [[] for _ in range(nq)]
    rs_busy = [False] * nq
    rows = dict(tid=[], qid=[], state=[], a=[], d=[], proc=[], aux=[])
    heap = []
    seq = 0
    s0 = fsm.initial_state()
    for i in range(n):
        heap.append((0.0, seq, 0, i, s0))
        seq += 1
    heapq.heapify(heap)

    def ps_advance(q, t):
        act = ps_active[q]
        if act:
            dt = (t - ps_clock[q]) / len(act)
            if dt > 0:
                for r in act:
                    act[r] -= dt
        ps_clock[q] = t

    def ps_plan(q):
        nonlocal seq
        act = ps_active[q]
        ps_version[q] += 1
        if act:
            r = min(act, key=act.get)
            t = ps_clock[q] + max(act[r], 0.0) * len(act)
            heapq.heappush(heap, (t, seq, 1, q, (r, ps_version[q])))
            seq += 1

    while heap:
        t, _, kind, x, y = heapq.heappop(heap)
        if kind == 0:  # arrival of task x in state y
            tid, sid = x, y
            qid = fsm.sample_one_obs(sid)
            q = net.queues[qid]
            s = q.sample_service()
            comp = 0
            if isinstance(s, tuple):  # mixture template: (service time, component)   (queues.pyx:1459-1462)
                s, comp = s
            rows["aux"].append(comp)
            row = len(rows["tid"])
            rows["tid"].append(tid); rows["qid"].append(qid); rows["state"].append(sid); rows["a"].append(t)
            if q.code == _queues.Q_PS:
                rows["d"].append(numpy.nan); rows["proc"].append(0)
                ps_advance(qid, t)
                ps_active[qid][row] = s
                ps_plan(qid)
                continue
            if q.code == _queues.Q_GG1R:
                rows["d"].append(numpy.nan); rows["proc"].append(0)
                if rs_busy[qid]:
                    rs_wait[qid].append((row, s))
                else:
                    rs_busy[qid] = True
                    heapq.heappush(heap, (t + s, seq, 2, qid, row))
                    seq += 1
                continue
            if q.code == _queues.Q_DELAY:
                d, p = t + s, 0
            else:
                f = free[qid]
                p = min(range(q.k), key=f.__getitem__)
                d = max(t, f[p]) + s
                f[p] = d
            rows["d"].append(d); rows["proc"].append(p)
            nxt = fsm.sample_state(sid)
            if nxt:
                heapq.heappush(heap, (d, seq, 0, tid, nxt))
                seq += 1
        elif kind == 2:  # service completion at random-order queue x: the next job is drawn among those waiting
            qid, row = x, y
            rows["d"][row] = t
            if rs_wait[qid]:
                r2, s2 = rs_wait[qid].pop(int(numpy.random.randint(len(rs_wait[qid]))))
                heapq.heappush(heap, (t + s2, seq, 2, qid, r2))
                seq += 1
            else:
                rs_busy[qid] = False
            nxt = fsm.sample_state(rows["state"][row])
            if nxt:
                heapq.heappush(heap, (t, seq, 0, rows["tid"][row], nxt))
                seq += 1
        else:  # planned PS departure from queue x
            qid = x
            row, ver = y
            if ver != ps_version[qid]:
                continue
            ps_advance(qid, t)
            del ps_active[qid][row]
            rows["d"][row] = t
            ps_plan(qid)
            nxt = fsm.sample_state(rows["state"][row])
            if nxt:
                heapq.heappush(heap, (t, seq, 0, rows["tid"][row], nxt))
                seq += 1

    return _arrivals.Arrivals.from_arrays(net, rows["tid"], rows["qid"], rows["state"], rows["a"], rows["d"],
                                          proc=rows["proc"], aux=rows["aux"])
