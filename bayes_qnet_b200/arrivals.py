"""Chain state on the host: ``Arrivals`` / ``Event`` with the reference's surface over structure-of-arrays.

The reference keeps one Python object per event, linked four ways (arrivals.pxd:11-37, arrivals.pyx:240-262).
Here an ``Arrivals`` is a set of flat numpy arrays -- ``a, d, s`` fp64; ``eid, tid, qid, state, proc`` and the
links ``prev_byt, next_byt, prev_byq, next_byq`` int32 row indices (-1 = none); ``obs_a, obs_d`` uint8 -- which
is exactly what crosses the C-ABI (``include/bq_capi.h: bq_arrivals``).  ``Event`` is a light view on one row
so code written against the reference (``for evt in arrv: evt.d``, ``evt.next_by_task()``, ``arrv.validate()``)
keeps working.

Row order inside a task is visit order (the reference links by task in insertion order, arrivals.pyx:430-436);
order inside a queue is by arrival time, except queue 0 -- the source, all a = 0 -- which is ordered by
departure (arrivals.pyx:256).
"""
import sys

import numpy

from . import queues as _queues


class Event(object):
    """One visit of a task to a queue.  Either detached (own values) or a view on row ``i`` of ``arrv``."""
    __slots__ = ("arrv", "i", "_v")

    _FIELDS = ("eid", "tid", "qid", "state", "a", "d", "s", "obs_a", "obs_d", "proc")

    def __init__(self, eid, tid, qid, q, a, s, d, state, obs_a=0, obs_d=0, proc=0, aux=None):
        self.arrv = None
        self.i = -1
        self._v = dict(eid=int(eid), tid=int(tid), qid=int(qid), state=int(state), a=float(a), d=float(d),
                       s=float(s), obs_a=int(obs_a), obs_d=int(obs_d), proc=int(proc) if proc else 0, q=q,
                       aux=int(aux) if isinstance(aux, (int, numpy.integer)) else 0)

    @classmethod
    def _view(cls, arrv, i):
        e = cls.__new__(cls)
        e.arrv = arrv
        e.i = int(i)
        e._v = None
        return e

    def _get(self, k):
        if self.arrv is None:
            return self._v[k]
        return getattr(self.arrv, "_" + k)[self.i].item()

    def _set(self, k, v):
        if self.arrv is None:
            self._v[k] = v
        else:
            getattr(self.arrv, "_" + k)[self.i] = v

    eid = property(lambda self: self._get("eid"), lambda self, v: self._set("eid", v))
    tid = property(lambda self: self._get("tid"), lambda self, v: self._set("tid", v))
    qid = property(lambda self: self._get("qid"), lambda self, v: self._set("qid", v))
    state = property(lambda self: self._get("state"), lambda self, v: self._set("state", v))
    a = property(lambda self: self._get("a"), lambda self, v: self._set("a", v))
    d = property(lambda self: self._get("d"), lambda self, v: self._set("d", v))
    s = property(lambda self: self._get("s"), lambda self, v: self._set("s", v))
    obs_a = property(lambda self: self._get("obs_a"), lambda self, v: self._set("obs_a", v))
    obs_d = property(lambda self: self._get("obs_d"), lambda self, v: self._set("obs_d", v))
    proc = property(lambda self: self._get("proc"), lambda self, v: self._set("proc", v))
    aux = property(lambda self: self._get("aux"), lambda self, v: self._set("aux", v))

    # the one auxiliary variable an event can carry: the component of its queue's mixture template
    # (Event.variable / set_variable, arrivals.pyx; MixtureTemplate.var_name = "COMPONENT_<queue>")
    def variable(self, name=None):
        return self.aux

    def set_variable(self, name, value):
        self.aux = int(value)

    @property
    def q(self):
        if self.arrv is None:
            return self._v["q"]
        return self.arrv.net.queues[self.qid]

    @property
    def c(self):  # commencement, arrivals.pyx:62,190-192
        return max(self.a, self.d - self.s)

    def arrival(self): return self.a
    def service(self): return self.s
    def departure(self): return self.d
    def queue(self): return self.q
    def commencement(self): return self.c
    def wait(self): return max(0, self.d - self.a - self.s)  # arrivals.pyx:187

    def _link(self, name):
        if self.arrv is None:
            return None
        j = int(getattr(self.arrv, "_" + name)[self.i])
        return Event._view(self.arrv, j) if j >= 0 else None

    def previous_by_queue(self): return self._link("prev_byq")
    def next_by_queue(self): return self._link("next_byq")
    def previous_by_task(self): return self._link("prev_byt")
    def next_by_task(self): return self._link("next_byt")
    prev_byq = property(previous_by_queue)
    next_byq = property(next_by_queue)
    prev_byt = property(previous_by_task)
    next_byt = property(next_by_task)

    def duplicate(self):
        return Event(self.eid, self.tid, self.qid, self.q, self.a, self.s, self.d, self.state, self.obs_a,
                     self.obs_d, self.proc)

    def __eq__(self, other):
        if not isinstance(other, Event):
            return NotImplemented
        if self.arrv is not None and other.arrv is self.arrv:
            return self.i == other.i
        return self is other

    def __hash__(self):
        return hash((id(self.arrv), self.i)) if self.arrv is not None else id(self)

    def __lt__(self, other):  # Event.__cmp__, arrivals.pyx:210-214
        return (self.a, self.d, self.tid) < (other.a, other.d, other.tid)

    def __repr__(self):
        return "[EVT %5d: %5.5f .. %5.5f S: %5.5f TID: %5d Q: %s ST: %d PID: %d OBS: %d%d]" % (
            self.eid, self.a, self.d, self.s, self.tid, self.q.name if self.q is not None else self.qid,
            self.state, self.proc, self.obs_a, self.obs_d)

    def dump(self):
        return "[EVT %5d: %23.17g .. %23.17g S: %23.17g TID: %5d Q: %s ST: %d PID: %d  OBS: %d%d ]" % (
            self.eid, self.a, self.d, self.s, self.tid, self.q.name if self.q is not None else self.qid,
            self.state, self.proc, self.obs_a, self.obs_d)

    def as_csv(self):  # arrivals.pyx:96-97
        return "%d %d %d  %.5f %.5f %.5f %.5f %d %d %d %d" % (
            self.eid, self.tid, self.qid, self.a, self.d, self.s, self.wait(), self.state, self.proc,
            self.obs_a, self.obs_d)


_INT = ("eid", "tid", "qid", "state", "proc")
_DBL = ("a", "d", "s")
_FLG = ("obs_a", "obs_d", "aux")  # aux: mixture component of the event (Event.auxiliaries, arrivals.pxd:36)
_LNK = ("prev_byt", "next_byt", "prev_byq", "next_byq")


class Arrivals(object):
    """All events of all tasks, structure-of-arrays (arrivals.pyx:240)."""

    def __init__(self, net, cap=None):
        self.net = net
        self.cap = cap
        self._pending = []
        self._n = 0
        for k in _INT + _LNK:
            setattr(self, "_" + k, numpy.zeros(0, dtype=numpy.int32))
        for k in _DBL:
            setattr(self, "_" + k, numpy.zeros(0, dtype=numpy.float64))
        for k in _FLG:
            setattr(self, "_" + k, numpy.zeros(0, dtype=numpy.uint8))
        self._eid2row = None

    # ---- construction -------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, net, tid, qid, state, a, d, obs_a=None, obs_d=None, eid=None, proc=None, s=None,
                    recompute=True, prev_byq=None, next_byq=None, aux=None):
        """Rows must list each task's events in visit order (any interleaving between tasks).  The by-queue
        order is derived from the times unless explicit prev_byq/next_byq links are given."""
        self = cls(net)
        n = len(tid)
        self._n = n
        self._tid = numpy.asarray(tid, dtype=numpy.int32).copy()
        self._qid = numpy.asarray(qid, dtype=numpy.int32).copy()
        self._state = numpy.asarray(state, dtype=numpy.int32).copy()
        self._a = numpy.asarray(a, dtype=numpy.float64).copy()
        self._d = numpy.asarray(d, dtype=numpy.float64).copy()
        self._s = numpy.zeros(n) if s is None else numpy.asarray(s, dtype=numpy.float64).copy()
        self._obs_a = numpy.ones(n, dtype=numpy.uint8) if obs_a is None else numpy.asarray(obs_a, dtype=numpy.uint8).copy()
        self._obs_d = numpy.ones(n, dtype=numpy.uint8) if obs_d is None else numpy.asarray(obs_d, dtype=numpy.uint8).copy()
        self._eid = numpy.arange(n, dtype=numpy.int32) if eid is None else numpy.asarray(eid, dtype=numpy.int32).copy()
        self._proc = numpy.zeros(n, dtype=numpy.int32) if proc is None else numpy.asarray(proc, dtype=numpy.int32).copy()
        self._aux = numpy.zeros(n, dtype=numpy.uint8) if aux is None else numpy.asarray(aux, dtype=numpy.uint8).copy()
        self._relink()
        if prev_byq is not None:
            self._prev_byq = numpy.asarray(prev_byq, dtype=numpy.int32).copy()
            self._next_byq = numpy.asarray(next_byq, dtype=numpy.int32).copy()
        if recompute:
            self.recompute_all_service()
        return self

    def append(self, evt):
        """Add a detached Event (arrivals.pyx:426).  Links are rebuilt lazily."""
        self._pending.append(evt)

    insert = append

    def _flush(self):
        if not self._pending:
            return
        pend, self._pending = self._pending, []
        n0, n1 = self._n, self._n + len(pend)
        for k in _INT + _DBL + _FLG:
            old = getattr(self, "_" + k)
            new = numpy.zeros(n1, dtype=old.dtype)
            new[:n0] = old[:n0]
            new[n0:] = [e._v[k] for e in pend]
            setattr(self, "_" + k, new)
        self._n = n1
        self._relink()
        for j, e in enumerate(pend):  # detached events become views, as in the reference (evt.arrv = self)
            e.arrv, e.i, e._v = self, n0 + j, None

    def _relink(self):
        n = self._n
        nq = len(self.net.queues)
        idx = numpy.arange(n, dtype=numpy.int64)
        # by task: stable by tid, row order inside the task
        order = numpy.argsort(self._tid, kind="stable")
        prev = numpy.full(n, -1, dtype=numpy.int32)
        nxt = numpy.full(n, -1, dtype=numpy.int32)
        if n > 1:
            same = self._tid[order[1:]] == self._tid[order[:-1]]
            prev[order[1:][same]] = order[:-1][same]
            nxt[order[:-1][same]] = order[1:][same]
        self._prev_byt, self._next_byt = prev, nxt
        # by queue: arrival order (ties: departure, row); queue 0 by departure (arrivals.pyx:256)
        prev = numpy.full(n, -1, dtype=numpy.int32)
        nxt = numpy.full(n, -1, dtype=numpy.int32)
        self._byq_rows = []
        for q in range(nq):
            rows = idx[self._qid == q]
            if q == 0 or self.net.queues[q].queue_order() == _queues.ARRV_BY_D:
                o = numpy.lexsort((rows, self._a[rows], self._d[rows]))
            else:
                o = numpy.lexsort((rows, self._d[rows], self._a[rows]))
            rows = rows[o].astype(numpy.int32)
            self._byq_rows.append(rows)
            if len(rows) > 1:
                prev[rows[1:]] = rows[:-1]
                nxt[rows[:-1]] = rows[1:]
        self._prev_byq, self._next_byq = prev, nxt
        self._eid2row = None

    def _rows_of_queue(self, q):
        """Rows of queue q following the CURRENT by-queue links."""
        self._flush()
        rows = numpy.nonzero(self._qid == q)[0]
        if len(rows) == 0:
            return rows.astype(numpy.int32)
        heads = rows[self._prev_byq[rows] < 0]
        out = numpy.empty(len(rows), dtype=numpy.int32)
        i = int(heads[0])
        for k in range(len(rows)):
            out[k] = i
            i = int(self._next_byq[i])
        return out

    # ---- service recomputation ------------------------------------------------------------------
    def recompute_all_service(self):
        """Recompute s (and proc) of every event from a, d and the by-queue order (arrivals.pyx:738)."""
        self._flush()
        for q, queue in enumerate(self.net.queues):
            rows = self._rows_of_queue(q)
            if len(rows) == 0:
                continue
            a, d = self._a[rows], self._d[rows]
            if queue.code == _queues.Q_DELAY:  # queues.pyx:771
                self._s[rows] = d - a
                self._proc[rows] = 0
            elif queue.code == _queues.Q_GGK and queue.k == 1:  # queues.pyx:610-612
                dprev = numpy.concatenate(([0.0], d[:-1]))
                self._s[rows] = d - numpy.maximum(a, dprev)
                self._proc[rows] = 0
            elif queue.code == _queues.Q_GGK:  # running k-vector, queues.pyx:654-680
                free = [0.0] * queue.k
                s = numpy.empty(len(rows))
                pr = numpy.empty(len(rows), dtype=numpy.int32)
                for j in range(len(rows)):
                    p = 0
                    m = free[0]
                    for t in range(1, queue.k):
                        if free[t] < m:
                            m, p = free[t], t
                    if j == 0:
                        p, m = 0, 0.0
                    s[j] = d[j] - max(a[j], m)
                    pr[j] = p
                    free[p] = d[j]
                self._s[rows] = s
                self._proc[rows] = pr
            elif queue.code == _queues.Q_PS:  # queues.pyx:1049-1070
                self._s[rows] = ps_service(a, d)
                self._proc[rows] = 0
            elif queue.code == _queues.Q_GG1R:  # queues.pyx:808-816: rows are in departure order
                dprev = numpy.concatenate(([0.0], d[:-1]))
                s = d - numpy.maximum(a, dprev)
                s[0] = d[0] - a[0]
                self._s[rows] = s
                self._proc[rows] = 0
            else:
                raise Exception("unknown queue type for %s" % queue)

    def regenerate_all_caches(self):
        self._flush()

    # ---- the reference's accessors ----------------------------------------------------------------
    def __iter__(self):
        self._flush()
        return (Event._view(self, i) for i in range(self._n))

    def __len__(self):
        self._flush()
        return self._n

    def all_events(self):
        return list(iter(self))

    def event(self, eid):
        self._flush()
        if self._eid2row is None:
            self._eid2row = dict((int(e), i) for i, e in enumerate(self._eid))
        return Event._view(self, self._eid2row[int(eid)])

    def num_events(self):
        self._flush()
        return self._n

    def num_hidden_events(self):  # arrivals.pyx:325: events with unobserved departure
        self._flush()
        return int(numpy.sum(self._obs_d == 0))

    def num_tasks(self):
        self._flush()
        return len(numpy.unique(self._tid))

    def all_tasks(self):
        self._flush()
        return [int(t) for t in numpy.unique(self._tid)]

    def num_queues(self):
        return len(self.net.queues)

    def qnet(self):
        return self.net

    def max_eidx(self):
        self._flush()
        return int(self._eid.max()) if self._n else -1

    def max_tid(self):
        self._flush()
        return int(self._tid.max()) if self._n else -1

    def events_of_task(self, tid):
        self._flush()
        rows = numpy.nonzero(self._tid == tid)[0]
        if len(rows) == 0:
            return []
        i = int(rows[self._prev_byt[rows] < 0][0])
        out = []
        while i >= 0:
            out.append(Event._view(self, i))
            i = int(self._next_byt[i])
        return out

    def events_of_qid(self, qid):
        return [Event._view(self, i) for i in self._rows_of_queue(qid)]

    def events_of_queue(self, q):
        return self.events_of_qid(self.net.qid_of_queue(q))

    def first_arrival_by_qid(self, qid):
        rows = self._rows_of_queue(qid)
        return Event._view(self, rows[0]) if len(rows) else None

    def final_arrival_by_qid(self, qid):
        rows = self._rows_of_queue(qid)
        return Event._view(self, rows[-1]) if len(rows) else None

    def first_arrival_by_queue(self, q): return self.first_arrival_by_qid(self.net.qid_of_queue(q))
    def final_arrival_by_queue(self, q): return self.final_arrival_by_qid(self.net.qid_of_queue(q))

    def is_initial(self, evt): return evt.previous_by_task() is None
    def is_final(self, evt): return evt.next_by_task() is None

    def initial_events(self):
        self._flush()
        return [Event._view(self, i) for i in numpy.nonzero(self._prev_byt < 0)[0]]

    # ---- copies and subsets -----------------------------------------------------------------------
    def duplicate(self):
        self._flush()
        other = Arrivals(self.net, cap=self.cap)
        other._n = self._n
        for k in _INT + _DBL + _FLG + _LNK:
            setattr(other, "_" + k, getattr(self, "_" + k).copy())
        other._byq_rows = [r.copy() for r in self._byq_rows]
        return other

    def subset(self, fn, adapt_fn=None):
        """Copy in which events with fn(arrv, evt) false become unobserved (arrivals.pyx:748-760)."""
        other = self.duplicate()
        for i in range(self._n):
            if not fn(self, Event._view(self, i)):
                other._obs_a[i] = other._obs_d[i] = 0
        return other

    def subset_by_task(self, pct, adapt_fn=None):
        """About pct of the TASKS stay observed; hidden events keep their times (arrivals.pyx:762-768)."""
        self._flush()
        other = self.duplicate()
        tasks = numpy.unique(self._tid)
        hide = numpy.random.rand(len(tasks)) > pct
        hidden = numpy.isin(self._tid, tasks[hide])
        other._obs_a[hidden] = 0
        other._obs_d[hidden] = 0
        return other

    def subset_by_task_fn(self, fn, adapt_fn=None):
        self._flush()
        keep = numpy.array([bool(fn(self, int(t))) for t in self._tid]) if self._n else numpy.zeros(0, bool)
        rows = numpy.nonzero(keep)[0]
        return Arrivals.from_arrays(self.net, self._tid[rows], self._qid[rows], self._state[rows], self._a[rows],
                                    self._d[rows], self._obs_a[rows], self._obs_d[rows], eid=self._eid[rows],
                                    proc=self._proc[rows], aux=self._aux[rows])

    def delete_task(self, tid):
        self._flush()
        rows = numpy.nonzero(self._tid != tid)[0]
        new = Arrivals.from_arrays(self.net, self._tid[rows], self._qid[rows], self._state[rows], self._a[rows],
                                   self._d[rows], self._obs_a[rows], self._obs_d[rows], eid=self._eid[rows],
                                   proc=self._proc[rows], aux=self._aux[rows])
        self.__dict__.update(new.__dict__)

    # ---- flat views for the C-ABI -------------------------------------------------------------------
    def soa(self):
        """Dictionary of the flat arrays (no copies).  This is what ``bq_create`` receives."""
        self._flush()
        return dict((k, getattr(self, "_" + k)) for k in _INT + _DBL + _FLG + _LNK)

    def set_departures(self, d):
        """Overwrite every departure time (and the arrivals they imply), then recompute services."""
        self._flush()
        d = numpy.asarray(d, dtype=numpy.float64)
        assert d.shape == (self._n,)
        self._d[:] = d
        has = self._prev_byt >= 0
        self._a[has] = d[self._prev_byt[has]]
        self.recompute_all_service()

    # ---- validation and output ------------------------------------------------------------------------
    def validate(self):
        """The assertions of arrivals.pyx:802-866 that apply to flat state."""
        self._flush()
        n = self._n
        assert numpy.all(self._s >= 0.0), "Invalid (negative) service time: rows %s" % numpy.nonzero(self._s < 0)[0][:5]
        nx = self._next_byq
        has = nx >= 0
        assert numpy.all(self._prev_byq[nx[has]] == numpy.nonzero(has)[0]), "by-queue links do not match"
        assert numpy.all(self._qid[nx[has]] == self._qid[has]), "by-queue link crosses queues"
        nt = self._next_byt
        has = nt >= 0
        assert numpy.all(self._prev_byt[nt[has]] == numpy.nonzero(has)[0]), "by-task links do not match"
        assert numpy.all(self._tid[nt[has]] == self._tid[has]), "TID pointers don't match, but are linked"
        assert numpy.all(numpy.abs(self._d[has] - self._a[nt[has]]) < 1e-15), "Mismatched arrivals"
        init = self._state == self.net.fsm.initial_state()
        assert numpy.all(self._a[init] == 0), "Initial event with bad arrival"
        chk = self.duplicate()
        chk.recompute_all_service()
        assert numpy.all(numpy.abs(chk._s - self._s) < 1e-8), "Service mismatch"
        for q, queue in enumerate(self.net.queues):
            if queue.code == _queues.Q_GGK and queue.k == 1:
                rows = self._rows_of_queue(q)
                if len(rows) > 1:
                    assert numpy.all(numpy.diff(self._a[rows]) >= 0), "GG1q arrival mismatch in %s" % queue.name
                    assert numpy.all(numpy.diff(self._d[rows]) >= 0), "GG1q departure mismatch in %s" % queue.name
        assert n == len(numpy.unique(self._eid)), "duplicate event ids"

    def as_csv(self):
        ret = "EID TID QID A D S W STATE PROC OBS_A OBS_D CMP\n"
        for evt in self:
            ret += evt.as_csv() + " NA\n"
        return ret

    def write_csv(self, f):
        f.write(self.as_csv())

    def __repr__(self):
        ret = ""
        for qi, q in enumerate(self.net.queues):
            ret += "QUEUE %s (%d)\n" % (q.name, qi)
            for evt in self.events_of_qid(qi):
                ret += str(evt) + "\n"
        return ret

    def report(self):
        print("ARRIVALS NE %d  NT %d" % (self.num_events(), self.num_tasks()))


def ps_service(a, d):
    """Processor-sharing service times: s_e = integral over [a_e, d_e] of dt / N(t), N counting e itself
    (queues.pyx:1049-1070, cninq.pyx:55-58)."""
    a = numpy.asarray(a, dtype=numpy.float64)
    d = numpy.asarray(d, dtype=numpy.float64)
    n = len(a)
    t = numpy.concatenate((a, d))
    delta = numpy.concatenate((numpy.ones(n), -numpy.ones(n)))
    o = numpy.lexsort((-delta, t))  # arrivals before departures at equal times
    t, delta = t[o], delta[o]
    N = numpy.cumsum(delta)
    dt = numpy.diff(t)
    with numpy.errstate(divide="ignore", invalid="ignore"):
        inc = numpy.where(N[:-1] > 0, dt / N[:-1], 0.0)
    V = numpy.concatenate(([0.0], numpy.cumsum(inc)))
    pos = numpy.empty(2 * n, dtype=numpy.int64)
    pos[o] = numpy.arange(2 * n)
    return V[pos[n:]] - V[pos[:n]]


def sort_by_arrival(evtl):
    evtl.sort(key=lambda e: (e.a, e.d))
    return evtl


def sort_by_departure(evtl):
    evtl.sort(key=lambda e: (e.d, e.a))
    return evtl
