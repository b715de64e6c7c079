// Sweep kernels for networks whose queue order is per-chain state: multi-server FIFO queues (QueueGGk
// with k > 1, queues.pyx:444-723) and processor-sharing queues (QueuePS, queues.pyx:1032-1232), mixed
// freely with single-server FIFO queues and delay stations.  MH-within-Gibbs sweeps (Qnet.gibbs_resample,
// qnet.pyx:337) of any network also run here.
//
// What the reference does per evaluation of the conditional (GGkGibbs.inner_dfun, qnet.pyx:1340-1357) is
// build two diff lists -- diffListForDeparture(e, x) in e's queue and diffListForArrival(e', x) in the
// queue of the task's next event -- and sum lpdf(s_new) - lpdf(s_old) over them (queues.pyx:83-95).  Here
// the same sums are produced by forward walks over per-chain arrays that are indexed by POSITION in the
// queue (queue after queue, each in arrival order; queue 0 in departure order, arrivals.pyx:256), so that a
// walk is a stream of consecutive loads with no pointer chasing:
//
//   ord[p], pos[e]   the event at position p / the position of event e
//   rec[p]           one record per position (BQ_REC_*): arrival and departure time of the event there, the k
//                    server-free times it sees, ascending (the reference's Event.d_proc_prev, arrivals.pxd:30, as
//                    a sorted multiset: services depend only on min F, and replacing the minimum by the event's
//                    own departure is the whole recursion of queues.pyx:654-680), and the log of its service time
//                    (Gamma / LogNormal service: the "old" half of every lpdf(s_new) - lpdf(s_old) then costs no
//                    transcendental).  The first 32-byte sector holds what the departure-side scan reads.
//   flag[p]          multi-server FIFO queues: the event arrived before the second server was free -- the only
//                    events a later departure upstream can delay (skip predicate of the scan)
//   dt[p], dord[p], dpos[e], sq[p]   PS queues: departure times sorted (with their events and the inverse
//                    map), and the service time of the event at arrival-position p
//   d[e]             departure times by event: the chain state as the C-ABI exposes it, kept in step
//
// Departure side (ggk_dep_delta): F[p] is the k largest departures before p, so the effect of moving one departure
// on every later position has a closed form and the diff list is an index range found by galloping -- no running
// vector.  Arrival side (ggk_arr_delta): a short walk that starts from the stored vector of the first position whose
// predecessors are unchanged, replays the recursion with the event re-inserted at its new rank (stable sort by
// arrival, queues.pyx:537) and stops past the moved window, where the vectors agree again.  A negative service
// time anywhere gives -inf.
//
// Execution models.  sweep_batch_kernel (default): a warp owns a chain and updates 32 events of it per round,
// every lane its own event on the unchanged state; read / write ranges are recorded, the longest conflict-free
// prefix of the round is committed, the rest redone -- exactly the sequential sweep.  sweep_dynamic_kernel: a
// GROUP of W lanes (W = 4..32; 32/W chains per warp) owns one chain and visits its events one after another; the
// lanes evaluate W consecutive shrink candidates of sampling._slice side by side: a candidate's position depends
// only on where the earlier candidates fell, never on their densities, so the first accepted candidate in order
// is exactly the outcome of the sequential loop.  Sufficient statistics, parameter updates and traces are fused,
// as in the static kernel.
#pragma once
#include "bq_static.cuh"

// The walk and commit routines are __host__ __device__ so that the test harness under tests/emul can run the
// very same code on the CPU (test infrastructure only: the library itself has no host execution path).
#if defined(__CUDA_ARCH__)
#define BQ_SYNCG(m) __syncwarp(m)
#else
#define BQ_SYNCG(m) ((void)(m))
#endif

namespace bq {

// per queue, identical for all chains: discipline, servers, position range [lo, hi), offset of its F block
// Record of one queue position (DynArgs::rec), S doubles:
//   [0] arrival  [1] departure  [2] F[0]  [3] F[1]  |  [4] log service  [5..] F[2..k)
// so the departure-side scan, which needs the two smallest free times and the event's own times, reads one
// 32-byte sector per position.  Networks without multi-server queues keep the log service in slot 3 (S = 4).
#define BQ_REC_AT 0
#define BQ_REC_DQ 1
#define BQ_REC_F 2
// slot of F[j] relative to F[0]
#define BQ_FSLOT(j) ((j) + ((j) >= 2 ? 1 : 0))
BQ_HDF int rec_stride(int kmax) { return (kmax <= 1) ? 4 : ((5 + (kmax - 2) + 3) & ~3); }
BQ_HDF int rec_ls_slot(int S) { return (S == 4) ? 3 : 4; }

struct QInfo {
    int type, k, lo, hi, foff, dist;
};

// one resample-eligible event of the sweep order: the event, the next event of its task, their queues
struct DynRec {
    int e, e1, q0, q1;
};

struct DynArgs {
    int E, Q, P, n_order, n_sweeps, update, flags, trace_base, n_chains, S, mode, scan;
    int chain_lo, chain_hi;  // sweep_batch_kernel_nw: the chains of this launch (a wave), [chain_lo, chain_hi)
    const DynRec *order;  // [n_order]
    const QInfo *qinfo;   // [Q]
    const int *q_poff;    // [Q] offset of the queue's first parameter
    double *state;        // [C][E]  d by event
    int *ord, *pos;       // [C][E]
    double *rec;          // [C][E][S] by position: {arrival, departure, log service, free-time vector F[0..k), pad}
    unsigned char *flag;  // [C][FW] one byte per position: the event arrived before the second server was free
    int FW;               // bytes per chain (a multiple of 32)
    int *dord, *dpos;     // [C][E]  PS
    double *dt, *sq;      // [C][E]  PS
    double *theta;        // [C][P]
    long long *stats;     // [C][BQ_NSTAT]
    double *tr_theta, *tr_wait, *tr_logp;
    unsigned long long seed;
    long long chain_offset;
    unsigned int sweep_base;
};

// ascending multiset of the k server-free times, +inf padded to KMAX so every loop is fully unrolled;
// `f` points at F[0] of a position record (element j at f[BQ_FSLOT(j)])
template <int KMAX>
struct FreeVec {
    double v[KMAX];
    BQ_HDF void load(const double *f, int k) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) v[j] = (j < k) ? f[BQ_FSLOT(j)] : INFINITY;
    }
    // the arriving event takes the server that frees first and holds it until x  (queues.pyx:610-612, 654-680)
    BQ_HDF void replace_min(double x) {
        v[0] = x;
#pragma unroll
        for (int j = 0; j + 1 < KMAX; ++j) {
            const bool sw = v[j] > v[j + 1];  // no NaNs here: one DSETP + selects instead of fmin/fmax
            const double lo = sw ? v[j + 1] : v[j], hi = sw ? v[j] : v[j + 1];
            v[j] = lo; v[j + 1] = hi;
        }
    }
    BQ_HDF bool equals(const double *f, int k) const {
        bool eq = true;
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < k) eq = eq && (v[j] == f[BQ_FSLOT(j)]);
        return eq;
    }
    BQ_HDF void store(double *f, int k) const {
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < k) f[BQ_FSLOT(j)] = v[j];
    }
};

// one chain's arrays + the tables shared by all chains
struct Chain {
    double *d;
    int *ord, *pos;
    double *rec;  // one record of S doubles per position: the neighbourhood of an event is one or two cache lines
    int S, ols;   // ols: slot of the log service
    unsigned char *flag;  // flag[p] = at(p) < F[p][1] ("contended arrival", multi-server FIFO queues; see ggk_dep_delta)
    int *dord, *dpos;
    double *dt, *sq;
    const QInfo *qi;
    const QConst *qc;

    BQ_HDF void bind(const DynArgs &a, int chain, const QInfo *qi_, const QConst *qc_) {
        const size_t c = (size_t)chain;
        d = a.state + c * a.E;
        ord = a.ord + c * a.E;
        pos = a.pos + c * a.E;
        rec = a.rec + c * a.E * a.S;
        S = a.S;
        ols = rec_ls_slot(a.S);
        flag = a.flag + c * a.FW;
        dord = a.dord ? a.dord + c * a.E : nullptr;
        dpos = a.dpos ? a.dpos + c * a.E : nullptr;
        dt = a.dt ? a.dt + c * a.E : nullptr;
        sq = a.sq ? a.sq + c * a.E : nullptr;
        qi = qi_;
        qc = qc_;
    }
    BQ_HDF double &at(int p) const { return rec[(size_t)p * S + BQ_REC_AT]; }  // arrival of the event at position p
    BQ_HDF double &dq(int p) const { return rec[(size_t)p * S + BQ_REC_DQ]; }  // its departure
    BQ_HDF double &ls(int p) const { return rec[(size_t)p * S + ols]; }  // log of its service time (Gamma / LogNormal)
    // server-free times seen by position p of a FIFO queue; the next position's vector is S doubles further
    BQ_HDF double *fvec(const QInfo &, int p) const { return rec + (size_t)p * S + BQ_REC_F; }
    // one byte per position: lanes of a warp that commit different positions never write the same byte
    BQ_HDF void set_contended(int p, bool on) const { flag[p] = on ? 1 : 0; }
    // the flags of positions 8 g .. 8 g + 7, one per byte
    BQ_HDF unsigned long long contended8(int g) const {
        return reinterpret_cast<const unsigned long long *>(flag)[g];
    }
};

// the event being resampled and its task successor, with their positions and current times
struct EventCtx {
    int e, e1, q0, q1, t0, t1, p0, p1;
    double x0, a_e, d1;
};

BQ_HDF void event_load(const Chain &c, const DynRec &rec, EventCtx &ev) {
    ev.e = rec.e; ev.e1 = rec.e1; ev.q0 = rec.q0; ev.q1 = rec.q1;
    ev.t0 = c.qi[ev.q0].type;
    ev.p0 = c.pos[ev.e];
    ev.x0 = c.dq(ev.p0);
    ev.a_e = c.at(ev.p0);
    ev.t1 = -1; ev.p1 = -1; ev.d1 = 0.0;
    if (ev.e1 >= 0) {
        ev.t1 = c.qi[ev.q1].type;
        ev.p1 = c.pos[ev.e1];
        ev.d1 = c.dq(ev.p1);
    }
}

// What one evaluation of the conditional read: walk steps, and the closed ranges of positions it touched in the
// queue of e (lo0..hi0), in the queue of the task's next event (lo1..hi1) and, for PS queues, of slots of the
// sorted-departure arrays (dlo..dhi).  The batch kernel commits an update only if no earlier update of the same
// batch wrote into these ranges.  Ranges are conservative (a superset is always safe).
struct Touch {
    int steps, lo0, hi0, lo1, hi1, dlo, dhi;
    BQ_HDF void reset() { steps = 0; lo0 = lo1 = dlo = 0x7fffffff; hi0 = hi1 = dhi = -1; }
    BQ_HDF void p0(int p) { lo0 = p < lo0 ? p : lo0; hi0 = p > hi0 ? p : hi0; }
    BQ_HDF void p1(int p) { lo1 = p < lo1 ? p : lo1; hi1 = p > hi1 ? p : hi1; }
    BQ_HDF void pa(bool first, int p) { if (first) p0(p); else p1(p); }
    BQ_HDF void pd(int p) { dlo = p < dlo ? p : dlo; dhi = p > dhi ? p : dhi; }
    BQ_HDF void merge(const Touch &o) {
        steps += o.steps;
        lo0 = o.lo0 < lo0 ? o.lo0 : lo0; hi0 = o.hi0 > hi0 ? o.hi0 : hi0;
        lo1 = o.lo1 < lo1 ? o.lo1 : lo1; hi1 = o.hi1 > hi1 ? o.hi1 : hi1;
        dlo = o.dlo < dlo ? o.dlo : dlo; dhi = o.dhi > dhi ? o.dhi : dhi;
    }
};

// first index in [lo, hi) whose value v[i stride] is > t, galloping from `hint`
BQ_HDI int upper_from(const double *v, int lo, int hi, int hint, double t, int stride = 1) {
    if (hint < lo) hint = lo;
    if (hint > hi) hint = hi;
    int l, r;  // invariant: v[(size_t)l * stride] <= t (or l == lo - 1), v[(size_t)r * stride] > t (or r == hi)
    if (hint < hi && v[(size_t)hint * stride] <= t) {
        l = hint;
        int step = 1;
        r = l + step;
        while (r < hi && v[(size_t)r * stride] <= t) { l = r; step <<= 1; r = l + step; }
        if (r > hi) r = hi;
    } else {
        r = hint;
        int step = 1;
        l = r - step;
        while (l >= lo && v[(size_t)l * stride] > t) { r = l; step <<= 1; l = r - step; }
        if (l < lo - 1) l = lo - 1;
    }
    while (r - l > 1) {
        const int m = (l + r) >> 1;
        if (v[(size_t)m * stride] <= t) l = m; else r = m;
    }
    return r;
}

// first index in [lo, hi) whose value is >= t, galloping forward from `hint`
BQ_HDI int lower_from(const double *v, int lo, int hi, int hint, double t) {
    if (hint < lo) hint = lo;
    if (hint > hi) hint = hi;
    int l = lo - 1, r = hint;  // v[l] < t (or l == lo - 1), v[r] >= t (or r == hi)
    if (hint < hi && v[hint] < t) {
        l = hint;
        int step = 1;
        r = l + step;
        while (r < hi && v[r] < t) { l = r; step <<= 1; r = l + step; }
        if (r > hi) r = hi;
    } else {
        int step = 1;
        l = r - step;
        while (l >= lo && v[l] >= t) { r = l; step <<= 1; l = r - step; }
        if (l < lo - 1) l = lo - 1;
    }
    while (r - l > 1) {
        const int m = (l + r) >> 1;
        if (v[m] < t) l = m; else r = m;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// FIFO queue with k servers: likelihoodDelta(diffListForDeparture(e, x))   (queues.pyx:470-511, 83-95)
//
// Every event replaces the smallest free time by its own departure, which is never smaller, so the free-time
// vector F[p] seen at position p is the k largest departures among the positions before p (over k initial zeros),
// F[p][0] is nondecreasing in p, and the value dropped at step p - 1, F[p-1][0], is the largest one outside F[p].
// Moving the departure at p0 from x0 to x therefore changes the server-free time seen at a later position p,
// b_old = F[p][0], independently of the other positions:
//   x > x0:  b_new = min(F[p][1], x)     where x0 <= F[p][0] < x      (x takes the place of x0 or of the minimum)
//   x < x0:  b_new = max(x, F[p-1][0])   where x  <  F[p][0] <= x0    (x0 leaves, x or the last dropped value enters)
// and b_new = b_old everywhere else; past the range the vectors are equal again.  The range is found by galloping
// over F[.][0]; the terms are the ones the reference's diff list holds, added in queue order.
// ---------------------------------------------------------------------------------------------
// first p in [lo, hi) with F[p][0] >= t (strict: > t); f = &F[lo][0], stride k doubles.  *far = furthest position probed
BQ_HDI int free_min_bound(const double *f, int k, int lo, int hi, double t, bool strict, int *far) {
    int l = lo - 1, r = lo, step = 1;  // the predicate is false at l (or l = lo - 1)
    while (r < hi) {
        const double v = f[(size_t)(r - lo) * k];
        if (strict ? (v > t) : (v >= t)) break;
        l = r; step <<= 1; r = l + step;
    }
    if (r > hi) r = hi;
    *far = (r < hi) ? r : hi - 1;
    while (r - l > 1) {
        const int m = (l + r) >> 1;
        const double v = f[(size_t)(m - lo) * k];
        if (strict ? (v > t) : (v >= t)) r = m; else l = m;
    }
    return r;
}

#ifndef BQ_WALK_U
#define BQ_WALK_U 1
#endif
#ifndef BQ_SCAN_LOCAL
#define BQ_SCAN_LOCAL 8  // positions scanned from their records before the flags take over (a multiple of BQ_WALK_U)
#endif
template <int KMAX>
BQ_HDI double ggk_dep_delta(const Chain &c, const EventCtx &ev, double x, Touch &tc) {
    const QInfo &q = c.qi[ev.q0];
    const QConst &qc = c.qc[ev.q0];
    const int k = q.k, p0 = ev.p0;
    const double *f = c.fvec(q, p0);  // F[p0 + i][j] = f[i S + BQ_FSLOT(j)]
    const int S = c.S;
    const double x0 = ev.x0;
    const double st = dmax(ev.a_e, f[0]);
    const double s_new = x - st;
    if (s_new < 0.0) return BQ_NEG_INF;
    const double s_old = x0 - st;
    double acc = 0.0;
    if (s_new != s_old) acc = lpdf(qc, s_new) - lpdf_ls(qc, s_old, c.ls(p0));
    tc.p0(p0);
    if (x == x0 || p0 + 1 >= q.hi) return acc;
    const bool later = x > x0;
    int far;
    int p = free_min_bound(f + S, S, p0 + 1, q.hi, later ? x0 : x, !later, &far);
    tc.p0(far);
    const double stop = later ? x : x0;
    // the positions are independent: fetch BQ_WALK_U of them at once so their cache misses overlap
    const int last = q.hi - 1;
    const bool flagmode = later && k > 1;
    const int p_local = p + BQ_SCAN_LOCAL;
    bool done = false;
    while (p <= last && !done && !(flagmode && p >= p_local)) {
        double bo_[BQ_WALK_U], bx_[BQ_WALK_U], ai_[BQ_WALK_U], di_[BQ_WALK_U];
#pragma unroll
        for (int u = 0; u < BQ_WALK_U; ++u) {
            const int pu = (p + u < last) ? p + u : last;
            const double *fp = f + (size_t)(pu - p0) * S;
            bo_[u] = fp[0];
            bx_[u] = later ? ((k > 1) ? fp[1] : INFINITY) : fp[-S];
            ai_[u] = fp[BQ_REC_AT - BQ_REC_F];
            di_[u] = fp[BQ_REC_DQ - BQ_REC_F];
        }
#pragma unroll
        for (int u = 0; u < BQ_WALK_U; ++u) {
            if (done) break;
            if (p > last) { done = true; break; }
            const double bo = bo_[u];
            if (later ? (bo >= stop) : (bo > stop)) { done = true; break; }  // same free times as before from here on
            ++tc.steps;
            const double bn = later ? ((bx_[u] < x) ? bx_[u] : x) : ((x > bx_[u]) ? x : bx_[u]);
            if (bn != bo) {
                const double ai = ai_[u], di = di_[u];
                const double sn = di - dmax(ai, bn);
                if (sn < 0.0) { tc.p0(p); return BQ_NEG_INF; }
                const double so = di - dmax(ai, bo);
                if (sn != so) acc += lpdf(qc, sn) - lpdf_ls(qc, so, c.ls(p));
            }
            ++p;
        }
    }
    if (flagmode && !done && p <= last) {
        // Only an event that arrived before the second server was free (flag bit) can start later when the minimum
        // free time is taken away: for the others a_p >= F[p][1] >= b_new >= b_old, nothing changes.  The flags are
        // read 8 positions at a time (one record per 32 to find the end of the range), so the long feasible scans
        // of lightly loaded queues cost a few bytes per event instead of a record.  Short scans never get here: the
        // first BQ_SCAN_LOCAL positions are taken from their records, which are next to the event's own.
        const int last_ = q.hi - 1;
        while (p <= last_) {
            const int bbase = p & ~31;
            const int bend = (bbase + 32 < q.hi) ? bbase + 32 : q.hi;   // exclusive
            const bool inside = f[(size_t)(bend - 1 - p0) * S] < x;    // F[.][0] is nondecreasing: whole block in range
            bool ended = false;
            for (int g = p >> 3; g < ((bend + 7) >> 3) && !ended; ++g) {
                unsigned long long w = c.contended8(g);
                const int gb = g << 3;
                if (gb < p) w &= ~0ull << (8 * (p - gb));                       // positions before the range
                if (gb + 8 > bend) w &= (1ull << (8 * (bend - gb))) - 1ull;      // the next queue's positions
                while (w) {
#ifdef __CUDA_ARCH__
                    const int by = (__ffsll((long long)w) - 1) >> 3;
#else
                    const int by = __builtin_ctzll(w) >> 3;
#endif
                    w &= ~(0xffull << (8 * by));
                    const int pp = gb + by;
                    const double *fp = f + (size_t)(pp - p0) * S;
                    const double bo = fp[0];
                    if (bo >= x) { ended = true; break; }  // same free times as before from here on
                    ++tc.steps;
                    const double f1 = fp[1];
                    const double bn = (f1 < x) ? f1 : x;
                    if (bn != bo) {
                        const double ai = fp[BQ_REC_AT - BQ_REC_F], di = fp[BQ_REC_DQ - BQ_REC_F];
                        const double sn = di - dmax(ai, bn);
                        if (sn < 0.0) { tc.p0(pp); return BQ_NEG_INF; }
                        const double so = di - dmax(ai, bo);
                        if (sn != so) acc += lpdf(qc, sn) - lpdf_ls(qc, so, c.ls(pp));
                    }
                }
            }
            tc.p0(bend - 1);
            if (ended || !inside) break;
            p = bend;
        }
        return acc;
    }
    tc.p0(p < q.hi ? p : q.hi - 1);
    return acc;
}

// Position e1 (now at p1) will occupy once its arrival is x: before the first OTHER event whose arrival is > x
// (stable sort by arrival, queues.pyx:537).
BQ_HDF int new_rank(const Chain &c, const QInfo &q, int p1, double x, double a_old, Touch &tc) {
    tc.p1(p1);
    if (x >= a_old) {
        int j = p1;
        while (j + 1 < q.hi && c.at(j + 1) <= x) ++j;
        tc.p1(j + 1 < q.hi ? j + 1 : j);
        return j;
    }
    const int j = upper_from(c.rec + BQ_REC_AT, q.lo, p1, p1 - 1, x, c.S);
    const int far = j - (p1 - j) - 1;  // the backward gallop overshoots by at most the distance covered
    tc.p1(far > q.lo ? far : q.lo);
    return j;
}

// old position of the event that sits at position t once e1 has moved from p1 to j (t != j)
BQ_HDF int moved_src(int t, int j, int p1) {
    if (j >= p1) return (t >= p1 && t < j) ? t + 1 : t;
    return (t > j && t <= p1) ? t - 1 : t;
}

// likelihoodDelta(diffListForArrival(e1, x))   (queues.pyx:514-591, 83-95)
template <int KMAX>
BQ_HDI double ggk_arr_delta(const Chain &c, const EventCtx &ev, double x, Touch &tc) {
    const QInfo &q = c.qi[ev.q1];
    const QConst &qc = c.qc[ev.q1];
    const int k = q.k, p1 = ev.p1;
    const double a_old = ev.x0;
    const int j = new_rank(c, q, p1, x, a_old, tc);
    const int ps = (j < p1) ? j : p1, pe = (j < p1) ? p1 : j;
    FreeVec<KMAX> M;
    M.load(c.fvec(q, ps), k);  // predecessors of position ps are untouched
    tc.p1(ps);
    double acc = 0.0;
    for (int t = ps; t < q.hi; ++t) {
        double ai, ai_old, di;
        const double *fo;  // the vector this event saw before the move
        int sp = p1;       // ... and its position then
        if (t == j) { ai = x; ai_old = a_old; di = ev.d1; fo = c.fvec(q, p1); }
        else {
            sp = moved_src(t, j, p1);
            tc.p1(sp);
            ai = ai_old = c.at(sp); di = c.dq(sp); fo = c.fvec(q, sp);
            if (t > pe && M.equals(fo, k)) break;
        }
        ++tc.steps;
        const double sn = di - dmax(ai, M.v[0]);
        if (sn < 0.0) return BQ_NEG_INF;
        const double so = di - dmax(ai_old, fo[0]);
        if (sn != so) acc += lpdf(qc, sn) - lpdf_ls(qc, so, c.ls(sp));
        M.replace_min(di);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Processor sharing (QueuePS, queues.pyx:1032-1232): s_i = integral over [a_i, d_i] of dt / N(t), with
// N(t) = #{a <= t} - #{d <= t} over the queue (cninq.pyx:55-58).  Moving one arrival or departure of event m from
// t_old to t_new changes N(t) by sgn = +-1 on I = (lo, hi) = (min, max) only, so (diffListForDeparture /
// diffListForArrival, queues.pyx:1149-1183):
//   * m itself gains or loses   sgn * integral over I of dt / max(N_old, N_old + sgn),
//   * every other event i that is in the queue during part of I changes by the integral over I ^ [a_i, d_i] of
//     (1 / (N_old + sgn) - 1 / N_old) dt,
//   * likelihoodDelta adds log(N_old(d_i) + 1) - log(N_new(d_i) + 1) for the events that depart inside I and for
//     m's own departure (queues.pyx:1220-1232).
// The breakpoints inside I are a merge of two contiguous ranges of at[] and dt[]; ranks are found by galloping
// from the moved event's own slots, so no search starts cold.
// ---------------------------------------------------------------------------------------------
struct PsMove {
    int q, pm, dm, lo_q, hi_q, sgn;  // queue, m's arrival position and departure slot
    int ia, id;                      // first arrival / departure strictly after lo
    int n0;                          // N_old just after lo
    double lo, hi;                   // the interval on which N(t) changes
    bool is_dep;
};

// merge cursor over the breakpoints strictly inside (lo, hi)
struct PsCursor {
    int ia, id, n;
    double t0;
    BQ_HDF void start(const PsMove &mv) { ia = mv.ia; id = mv.id; n = mv.n0; t0 = mv.lo; }
    // next piece [t0, t1) with constant N_old = n.  kind: +1 arrival, -1 departure, 0 end of I at t1;
    // `slot` = index of the breakpoint in its array
    BQ_HDF void piece(const Chain &c, const PsMove &mv, double &t1, int &kind, int &slot) const {
        const double ta = (ia < mv.hi_q) ? c.at(ia) : INFINITY, td = (id < mv.hi_q) ? c.dt[id] : INFINITY;
        if (ta <= td) { t1 = ta; kind = 1; slot = ia; } else { t1 = td; kind = -1; slot = id; }
        if (!(t1 < mv.hi)) { t1 = mv.hi; kind = 0; slot = -1; }
    }
    BQ_HDF void advance(double t1, int kind) {
        if (kind > 0) { ++n; ++ia; } else if (kind < 0) { --n; ++id; }
        t0 = t1;
    }
};

BQ_HDI void ps_move_setup(const Chain &c, int q, int pm, int m, bool is_dep, double t_old, double t_new, PsMove &mv,
                          Touch &tc) {
    const QInfo &qi = c.qi[q];
    mv.q = q; mv.pm = pm; mv.dm = c.dpos[m]; mv.is_dep = is_dep;
    mv.lo_q = qi.lo; mv.hi_q = qi.hi;
    mv.lo = (t_new < t_old) ? t_new : t_old;
    mv.hi = (t_new < t_old) ? t_old : t_new;
    // departure later / arrival earlier: m is in the queue longer
    mv.sgn = is_dep ? ((t_new > t_old) ? 1 : -1) : ((t_new < t_old) ? 1 : -1);
    // arrivals <= lo: a_m <= lo always (departure moves stay right of a_m; arrival moves: lo = min(a_old, x));
    // departures <= lo: d_m >= lo always, so the rank is at or before m's own departure slot
    mv.ia = upper_from(c.rec + BQ_REC_AT, mv.lo_q, mv.hi_q, pm, mv.lo, c.S);
    mv.id = upper_from(c.dt, mv.lo_q, mv.hi_q, mv.dm, mv.lo);
    mv.n0 = (mv.ia - mv.lo_q) - (mv.id - mv.lo_q);
    // the gallops probe between the hint and about twice the distance to the answer
    const bool first = is_dep;  // a departure move happens in the queue of e, an arrival move in the queue of e1
    int far = 2 * mv.ia - pm;
    far = far < mv.lo_q ? mv.lo_q : (far >= mv.hi_q ? mv.hi_q - 1 : far);
    tc.pa(first, pm); tc.pa(first, far); tc.pa(first, mv.ia < mv.hi_q ? mv.ia : mv.hi_q - 1);
    if (mv.ia > mv.lo_q) tc.pa(first, mv.ia - 1);
    far = 2 * mv.id - mv.dm;
    far = far < mv.lo_q ? mv.lo_q : (far >= mv.hi_q ? mv.hi_q - 1 : far);
    tc.pd(mv.dm); tc.pd(far); tc.pd(mv.id < mv.hi_q ? mv.id : mv.hi_q - 1);
    if (mv.id > mv.lo_q) tc.pd(mv.id - 1);
}

// change of an event's service: it is in the queue on [b_i, e_i] ^ I; MINE = it is the moved event itself
template <bool MINE>
BQ_HDI double ps_delta_service(const Chain &c, const PsMove &mv, double b_i, double e_i) {
    PsCursor cu;
    cu.start(mv);
    double acc = 0.0;
    for (;;) {  // reads at[] / dt[] only inside the ranges ps_apply_move reports
        double t1; int kind, slot;
        cu.piece(c, mv, t1, kind, slot);
        const double a = (cu.t0 > b_i) ? cu.t0 : b_i, b = (t1 < e_i) ? t1 : e_i;
        if (b > a) {
            const double n = (double)cu.n;
            if (MINE) acc += (b - a) / ((mv.sgn > 0) ? n + 1.0 : n);
            else acc += (b - a) * (1.0 / (n + (double)mv.sgn) - 1.0 / n);
        }
        if (kind == 0 || !(t1 < e_i)) break;
        cu.advance(t1, kind);
    }
    return MINE ? (double)mv.sgn * acc : acc;
}

// Visits every event of the queue whose service changes.  COMMIT = false: returns likelihoodDelta (sum of
// lpdf(s_new) - lpdf(s_old) plus the log N terms).  COMMIT = true: stores the new service times (lane 0), returns 0.
template <bool COMMIT>
BQ_HDI double ps_apply_move(const Chain &c, const PsMove &mv, int lane, Touch &tc) {
    const QConst &qc = c.qc[mv.q];
    double acc = 0.0;
    {   // the moved event
        const double ds = ps_delta_service<true>(c, mv, mv.lo, mv.hi);
        const double so = c.sq[mv.pm];
        double sn = so + ds;
        if (sn < 0.0) sn = 0.0;
        if (COMMIT) { if (lane == 0) c.sq[mv.pm] = sn; }
        else acc += lpdf(qc, sn) - lpdf(qc, so);
    }
    // events in the queue just after lo (other than m): scan back over the arrivals <= lo
    const bool first = mv.is_dep;
    int need = mv.n0 - ((mv.sgn < 0) ? 1 : 0);
    for (int p = mv.ia - 1; p >= mv.lo_q && need > 0; --p) {
        tc.pa(first, p);
        if (p == mv.pm) continue;
        const double di = c.dq(p);
        if (!(di > mv.lo)) continue;
        --need; ++tc.steps;
        const double ds = ps_delta_service<false>(c, mv, mv.lo, di);
        const double so = c.sq[p];
        double sn = so + ds;
        if (sn < 0.0) sn = 0.0;
        if (COMMIT) { if (lane == 0) c.sq[p] = sn; }
        else acc += lpdf(qc, sn) - lpdf(qc, so);
    }
    // events that arrive inside I
    {   // every merge over I reads at[] up to the first arrival >= hi and dt[] up to the first departure >= hi
        PsCursor cu;
        cu.start(mv);
        for (;;) {
            double t1; int kind, slot;
            cu.piece(c, mv, t1, kind, slot);
            if (kind == 0) break;
            cu.advance(t1, kind);
        }
        tc.pa(first, cu.ia < mv.hi_q ? cu.ia : mv.hi_q - 1);
        tc.pd(cu.id < mv.hi_q ? cu.id : mv.hi_q - 1);
    }
    for (int p = mv.ia; p < mv.hi_q && c.at(p) < mv.hi; ++p) {
        if (p == mv.pm) continue;
        ++tc.steps;
        const double ds = ps_delta_service<false>(c, mv, c.at(p), c.dq(p));
        const double so = c.sq[p];
        double sn = so + ds;
        if (sn < 0.0) sn = 0.0;
        if (COMMIT) { if (lane == 0) c.sq[p] = sn; }
        else acc += lpdf(qc, sn) - lpdf(qc, so);
    }
    if (COMMIT) return 0.0;
    // log N terms: departures strictly inside I see one job more (sgn > 0) or less; plus m's own departure
    PsCursor cu;
    cu.start(mv);
    for (;;) {
        double t1; int kind, slot;
        cu.piece(c, mv, t1, kind, slot);
        if (kind == 0) break;
        if (kind < 0 && slot != mv.dm) acc += log_count(cu.n) - log_count(cu.n + mv.sgn);
        cu.advance(t1, kind);
    }
    if (mv.is_dep) {
        // N_old(d_old) + 1 and N_new(x) + 1, m counted (queues.pyx:1225-1230)
        if (mv.sgn > 0) acc += log_count(mv.n0 + 1) - log_count(cu.n + 1);
        else acc += log_count(cu.n) - log_count(mv.n0);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Random-order single server (QueueGG1R, queues.pyx:788-1027).  The queue's list is ordered by DEPARTURE
// (queue_order ARRV_BY_D), s_i = d_i - max(a_i, previous departure) (recompute_service, :808-816), and the
// likelihood carries, per event, - log N(c_i) at its commencement c_i = max(a_i, d_i - s_i) (the server picked it
// among the N jobs present) plus the rule that an event which started on arrival (wait < 1e-10) found the queue
// empty, N(a_i) = 1 (likelihoodDelta, :956-1006).  N(t) = #{a <= t} - #{d <= t} over the queue (cninq.pyx:55-58).
// It lives on the arrays processor sharing already keeps: positions by arrival (rec.at sorted: #{a <= t} is a
// search), dt[] / dord[] / dpos[] the departures sorted, and sq[j] = ARRIVAL time of the event in departure slot j
// (for PS queues sq holds service times by arrival position), so that a diff list is a run of consecutive slots.
//   departure move x0 -> x of m (diffListForDeparture, :866-954): m takes its new place in the departure order;
//     the diff list is the run of slots from the earlier of the old / new predecessor to the later of the old /
//     new successor.  Only m, its old successor and its new successor change service; every slot of the run
//     contributes + log N_old(c_old) [s_old != 0] - log N_new(c_new) [s_new != 0] with
//     N_new(t) = N_old(t) + [x0 <= t] - [x <= t].
//   arrival move a0 -> x of m (diffListForArrival, :859-864): the diff list is m alone; the same terms for m and
//     for every event whose commencement lies in [min(a0, x), max(a0, x)] (c2e_range) -- commencements are
//     nondecreasing along the departure order, so again a run of slots -- with N_new(t) = N_old(t) + [x <= t] - [a0 <= t].
// ---------------------------------------------------------------------------------------------
#ifndef BQ_RSS_PENALTY
#define BQ_RSS_PENALTY 30.0
#endif
struct Gg1rMove {
    double plus, minus;  // N_new(t) = N_old(t) + [plus <= t] - [minus <= t]
    int ha, hd;          // search hints: an arrival position and a departure slot near the last answers
    bool first;          // the queue is the one of e (Touch ranges 0) or of the task's next event (ranges 1)
};

// the positions / slots a galloping search from `hint` that ended at r may have probed
BQ_HDF void gg1r_touch(Touch &tc, bool first, bool slots, int lo, int hi, int hint, int r) {
    int far = 2 * r - hint;
    far = far < lo ? lo : (far >= hi ? hi - 1 : far);
    int h = hint < lo ? lo : (hint >= hi ? hi - 1 : hint);
    int rr = r < hi ? r : hi - 1;
    if (slots) { tc.pd(h); tc.pd(far); tc.pd(rr); if (r > lo) tc.pd(r - 1); }
    else { tc.pa(first, h); tc.pa(first, far); tc.pa(first, rr); if (r > lo) tc.pa(first, r - 1); }
}

BQ_HDI int gg1r_count(const Chain &c, const QInfo &q, Gg1rMove &mv, double t, Touch &tc) {  // N_old(t)
    const int ia = upper_from(c.rec + BQ_REC_AT, q.lo, q.hi, mv.ha, t, c.S);
    gg1r_touch(tc, mv.first, false, q.lo, q.hi, mv.ha, ia);
    const int id = upper_from(c.dt, q.lo, q.hi, mv.hd, t);
    gg1r_touch(tc, mv.first, true, q.lo, q.hi, mv.hd, id);
    mv.ha = ia; mv.hd = id;
    return ia - id;
}
BQ_HDF int gg1r_adjust(const Gg1rMove &mv, double t) { return ((mv.plus <= t) ? 1 : 0) - ((mv.minus <= t) ? 1 : 0); }

// one event of the set whose N(.) terms are re-evaluated; false = the start-on-arrival rule is violated
BQ_HDI bool gg1r_member(const Chain &c, const QInfo &q, Gg1rMove &mv, double a_old, double d_old, double s_old, double a_new,
                        double d_new, double s_new, double &acc, Touch &tc) {
    // QInfo::k of a random-order queue: 1 = the reference's rule (-inf), 2 = the soft rule of the repair sweeps
    // (bq_set_relaxed): a violation costs BQ_RSS_PENALTY nats, counted as the CHANGE of the event's status so that
    // events the move does not concern cancel -- a state that breaks the rule somewhere can then be walked out of
    const bool relaxed = q.k > 1;
    double w = d_new - a_new - s_new;
    if (w < 0.0) w = 0.0;
    if (w < 1e-10 && gg1r_count(c, q, mv, a_new, tc) + gg1r_adjust(mv, a_new) > 1) {
        if (!relaxed) return false;
        acc -= BQ_RSS_PENALTY;
    }
    if (relaxed) {
        double w0 = d_old - a_old - s_old;
        if (w0 < 0.0) w0 = 0.0;
        if (w0 < 1e-10 && gg1r_count(c, q, mv, a_old, tc) > 1) acc += BQ_RSS_PENALTY;
    }
    ++tc.steps;
    if (s_old != 0.0) {
        const double c_old = dmax(a_old, d_old - s_old);
        acc += log_count(gg1r_count(c, q, mv, c_old, tc));
    }
    if (s_new != 0.0) {
        const double c_new = dmax(a_new, d_new - s_new);
        acc -= log_count(gg1r_count(c, q, mv, c_new, tc) + gg1r_adjust(mv, c_new));
    }
    return true;
}

// the departure slot m (now in slot j0) takes when its departure becomes x: `later` = it moves towards the tail.
// later: it sits just before jn, the first slot after j0 with dt >= x; earlier: just after jp, the last slot before
// j0 with dt <= x (the two while loops of queues.pyx:871-880)
BQ_HDI int gg1r_new_slot(const Chain &c, const QInfo &q, int j0, double x0, double x, int &jp, int &jn, Touch &tc) {
    jp = j0 - 1; jn = j0 + 1;
    if (x > x0) {
        jn = lower_from(c.dt, j0 + 1, q.hi, j0 + 1, x);
        gg1r_touch(tc, true, true, q.lo, q.hi, j0 + 1, jn);
        return jn - 1;
    }
    if (x < x0) {
        jp = upper_from(c.dt, q.lo, j0, j0 - 1, x) - 1;
        gg1r_touch(tc, true, true, q.lo, q.hi, j0 - 1, jp + 1);
        return jp + 1;
    }
    return j0;
}

// likelihoodDelta(diffListForDeparture(e, x))
BQ_HDI double gg1r_dep_delta(const Chain &c, const EventCtx &ev, double x, Touch &tc) {
    const QInfo &q = c.qi[ev.q0];
    const QConst &qc = c.qc[ev.q0];
    const int lo = q.lo, hi = q.hi, j0 = c.dpos[ev.e];
    tc.p0(ev.p0); tc.pd(j0);
    const double x0 = ev.x0, a_m = ev.a_e;
    if (x < a_m) return BQ_NEG_INF;  // m's own service x - max(a_m, d_prev) with d_prev <= x; nobody else's can go negative
    int jp, jn;
    const int jnew = gg1r_new_slot(c, q, j0, x0, x, jp, jn, tc);
    const bool later = x > x0;
    const int mlo = later ? (j0 - 1 > lo ? j0 - 1 : lo) : (x < x0 ? (jp > lo ? jp : lo) : (j0 - 1 > lo ? j0 - 1 : lo));
    const int mhi = later ? (jn < hi - 1 ? jn : hi - 1) : (j0 + 1 < hi - 1 ? j0 + 1 : hi - 1);
    tc.pd(mlo); tc.pd(mhi);
    if (mlo > lo) tc.pd(mlo - 1);
    Gg1rMove mv;
    mv.plus = x0; mv.minus = x; mv.ha = ev.p0; mv.hd = j0; mv.first = true;
    const double dprev_m_old = (j0 > lo) ? c.dt[j0 - 1] : 0.0;
    double dprev_m_new = dprev_m_old;
    if (later && jn - 1 > j0) dprev_m_new = c.dt[jn - 1];
    if (x < x0) dprev_m_new = (jp >= lo) ? c.dt[jp] : 0.0;
    const double s_old_m = x0 - dmax(a_m, dprev_m_old), s_new_m = x - dmax(a_m, dprev_m_new);
    if (s_new_m < 0.0) return BQ_NEG_INF;
    double acc = 0.0;
    if (s_new_m != s_old_m) acc = lpdf(qc, s_new_m) - lpdf(qc, s_old_m);
    // the run of slots in the NEW order: m is inserted before old slot `ins` (hi: behind everything; -1: it keeps
    // its slot)
    const int ins = (jnew == j0) ? -1 : (later ? jn : jp + 1);
    for (int j = mlo; j <= mhi; ++j) {
        if (j == j0) {
            if (ins == -1 && !gg1r_member(c, q, mv, a_m, x0, s_old_m, a_m, x, s_new_m, acc, tc)) return BQ_NEG_INF;
            continue;
        }
        if (j == ins && !gg1r_member(c, q, mv, a_m, x0, s_old_m, a_m, x, s_new_m, acc, tc)) return BQ_NEG_INF;
        const double a_j = c.sq[j], d_j = c.dt[j];
        const double dp_old = (j > lo) ? c.dt[j - 1] : 0.0;
        double dp_new = dp_old;
        if (j == ins) dp_new = x;  // m is the new predecessor
        else if (j == j0 + 1) dp_new = (ins == -1) ? x : ((j0 > lo) ? c.dt[j0 - 1] : 0.0);  // m stayed / m left
        const double s_old = d_j - dmax(a_j, dp_old), s_new = d_j - dmax(a_j, dp_new);
        if (s_new < 0.0) return BQ_NEG_INF;
        if (s_new != s_old) acc += lpdf(qc, s_new) - lpdf(qc, s_old);
        if (!gg1r_member(c, q, mv, a_j, d_j, s_old, a_j, d_j, s_new, acc, tc)) return BQ_NEG_INF;
    }
    if (ins > mhi && !gg1r_member(c, q, mv, a_m, x0, s_old_m, a_m, x, s_new_m, acc, tc)) return BQ_NEG_INF;  // m is last
    return acc;
}

// likelihoodDelta(diffListForArrival(e1, x))
BQ_HDI double gg1r_arr_delta(const Chain &c, const EventCtx &ev, double x, Touch &tc) {
    const QInfo &q = c.qi[ev.q1];
    const QConst &qc = c.qc[ev.q1];
    const int lo = q.lo, hi = q.hi, j1 = c.dpos[ev.e1], p1 = ev.p1;
    tc.p1(p1); tc.pd(j1);
    if (j1 > lo) tc.pd(j1 - 1);
    const double a0 = ev.x0, d1 = ev.d1;
    const double dprev = (j1 > lo) ? c.dt[j1 - 1] : 0.0;
    const double s_old = d1 - dmax(a0, dprev), s_new = d1 - dmax(x, dprev);
    if (s_new < 0.0) return BQ_NEG_INF;
    double acc = 0.0;
    if (s_new != s_old) acc = lpdf(qc, s_new) - lpdf(qc, s_old);
    {   // the arrival position e1 will take (the commit re-ranks rec.at): the accepted candidate claims the range
        int r = upper_from(c.rec + BQ_REC_AT, lo, hi, p1, x, c.S);
        gg1r_touch(tc, false, false, lo, hi, p1, r);
    }
    Gg1rMove mv;
    mv.plus = x; mv.minus = a0; mv.ha = p1; mv.hd = j1; mv.first = false;
    const double l = (x < a0) ? x : a0, r = (x < a0) ? a0 : x;
    // slots after e1's: commencements >= d1 >= r, members only on a tie
    for (int j = j1 + 1; j < hi; ++j) {
        tc.pd(j);
        const double a_j = c.sq[j], d_j = c.dt[j];
        const double s_j = d_j - dmax(a_j, c.dt[j - 1]);
        const double c_j = dmax(a_j, d_j - s_j);
        if (c_j > r) break;
        if (c_j >= l && !gg1r_member(c, q, mv, a_j, d_j, s_j, a_j, d_j, s_j, acc, tc)) return BQ_NEG_INF;
    }
    if (!gg1r_member(c, q, mv, a0, d1, s_old, x, d1, s_new, acc, tc)) return BQ_NEG_INF;
    for (int j = j1 - 1; j >= lo; --j) {
        tc.pd(j);
        if (j > lo) tc.pd(j - 1);
        const double a_j = c.sq[j], d_j = c.dt[j];
        const double s_j = d_j - dmax(a_j, (j > lo) ? c.dt[j - 1] : 0.0);
        const double c_j = dmax(a_j, d_j - s_j);
        if (c_j < l) break;
        if (c_j <= r && !gg1r_member(c, q, mv, a_j, d_j, s_j, a_j, d_j, s_j, acc, tc)) return BQ_NEG_INF;
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// the conditional of d_e: inner_dfun(x) - lik0   (qnet.pyx:1340-1357)
// ---------------------------------------------------------------------------------------------
// PS = 0: a build for networks without processor-sharing queues (their code is half of the kernel); 1: with them;
// 2: with the random-order queues as well (their code in the processor-sharing build cost config 4 a quarter of its rate)
template <int KMAX, int PS = 2>
BQ_HDI double dyn_logcond(const Chain &c, const EventCtx &ev, double x, Touch &tc) {
    double g;
    tc.p0(ev.p0);
    if (ev.e1 >= 0) tc.p1(ev.p1);
    if (ev.t0 == Q_GGK) {
        g = ggk_dep_delta<KMAX>(c, ev, x, tc);
    } else if (PS && ev.t0 == Q_PS) {
        if (x < ev.a_e) return BQ_NEG_INF;
        g = 0.0;
        if (x != ev.x0) {
            PsMove mv;
            ps_move_setup(c, ev.q0, ev.p0, ev.e, true, ev.x0, x, mv, tc);
            g = ps_apply_move<false>(c, mv, 0, tc);
        }
    } else if (PS >= 2 && ev.t0 == Q_GG1R) {
        g = gg1r_dep_delta(c, ev, x, tc);
    } else {  // DelayStation: s = d - a, single-event diff list (queues.pyx:745-784)
        const double sn = x - ev.a_e, so = ev.x0 - ev.a_e;
        if (sn < 0.0) return BQ_NEG_INF;
        g = (sn != so) ? lpdf(c.qc[ev.q0], sn) - lpdf_ls(c.qc[ev.q0], so, c.ls(ev.p0)) : 0.0;
    }
    if (!(g > BQ_NEG_INF) || ev.e1 < 0) return g;
    double g1;
    if (ev.t1 == Q_GGK) {
        g1 = ggk_arr_delta<KMAX>(c, ev, x, tc);
    } else if (PS && ev.t1 == Q_PS) {
        if (x > ev.d1) return BQ_NEG_INF;
        g1 = 0.0;
        if (x != ev.x0) {
            PsMove mv;
            ps_move_setup(c, ev.q1, ev.p1, ev.e1, false, ev.x0, x, mv, tc);
            g1 = ps_apply_move<false>(c, mv, 0, tc);
        }
    } else if (PS >= 2 && ev.t1 == Q_GG1R) {
        g1 = gg1r_arr_delta(c, ev, x, tc);
    } else {
        const double sn = ev.d1 - x, so = ev.d1 - ev.x0;
        if (sn < 0.0) return BQ_NEG_INF;
        g1 = (sn != so) ? lpdf(c.qc[ev.q1], sn) - lpdf_ls(c.qc[ev.q1], so, c.ls(ev.p1)) : 0.0;
    }
    return g + g1;
}

// ---------------------------------------------------------------------------------------------
// MH-within-Gibbs (Qnet.gibbs_resample, qnet.pyx:337-514): the proposal is one draw (two slice steps,
// call_sampler thin = 2, qnet.pyx:1292-1300) from h(x) = departureLik(e)(x) + arrivalLik(e')(x)
// (queues.pyx:1239-1256, pwfun.pyx:183-202):
//   FIFO queue:    h_dep(x) = lpdf(x - max(a_e, d_prev)) on [max(a_e, d_prev), inf)      (queues.pyx:414-430, 464-468)
//                  h_arr(x) = lpdf(d_e' - max(x, d_prev'))  on [0, d_e']                  (queues.pyx:394-412, 458-462)
//                  d_prev = last departure of the server the event occupies = min F
//   delay station: h_dep(x) = lpdf(x - a_e) on [a_e, inf),  h_arr(x) = lpdf(d_e' - x) on [0, d_e']   (queues.pyx:750-756)
//   PS has no arrivalLik (queues.pyx:1140: "use slice-only"): MH sweeps are refused for networks with PS queues.
// ---------------------------------------------------------------------------------------------
struct MhProp {
    double b0, d1, dp1, Lh, Uh;
    const QConst *c0, *c1;
    int has1, fifo1;
    BQ_HDF double eval(double x) const {
        if (x < Lh || x > Uh) return BQ_NEG_INF;
        double v = lpdf(*c0, x - b0);
        if (has1) v += lpdf(*c1, d1 - (fifo1 ? dmax(x, dp1) : x));
        return v;
    }
};

BQ_HDI void mh_setup(const Chain &c, const EventCtx &ev, MhProp &pr) {
    pr.c0 = &c.qc[ev.q0];
    pr.c1 = pr.c0;
    pr.b0 = ev.a_e;
    if (ev.t0 == Q_GGK) pr.b0 = dmax(ev.a_e, c.fvec(c.qi[ev.q0], ev.p0)[0]);
    pr.Lh = pr.b0;
    pr.Uh = INFINITY;
    pr.has1 = (ev.e1 >= 0);
    pr.fifo1 = 0; pr.d1 = 0.0; pr.dp1 = 0.0;
    if (pr.has1) {
        pr.c1 = &c.qc[ev.q1];
        pr.d1 = ev.d1;
        pr.Uh = ev.d1;
        if (pr.Lh < 0.0) pr.Lh = 0.0;
        if (ev.t1 == Q_GGK) { pr.fifo1 = 1; pr.dp1 = c.fvec(c.qi[ev.q1], ev.p1)[0]; }
    }
}

// ---------------------------------------------------------------------------------------------
// commit d_e := x   (apply_departure, qnet.pyx:2166-2178 -> applyDiffList, arrivals.pyx:542-604)
// Executed by every lane of the group with identical control flow; lane 0 stores, the shifts are spread over lanes.
// ---------------------------------------------------------------------------------------------

// Slide the entry at slot p of a sorted time array to slot j (its rank under the new time), carrying the parallel
// arrays along: ids ev[] with inverse inv[], and up to two payload arrays (may be null).
template <int W>
BQ_HDI void reslot(double *v, int sv, int *ev, int *inv, double *pay0, int s0, double *pay1, int s1, int p, int j,
                   double t_new, int lane, unsigned gmask) {
#define V_(t) v[(size_t)(t) * sv]
#define P0_(t) pay0[(size_t)(t) * s0]
#define P1_(t) pay1[(size_t)(t) * s1]
    const int m = ev[p];
    const double m0 = pay0 ? P0_(p) : 0.0, m1 = pay1 ? P1_(p) : 0.0;
    BQ_SYNCG(gmask);
    if (j > p) {  // later: entries p+1..j move one slot towards the head
        for (int base = p + 1; base <= j; base += W) {
            const int t = base + lane;
            int i = -1; double x = 0.0, y0 = 0.0, y1 = 0.0;
            if (t <= j) { i = ev[t]; x = V_(t); if (pay0) y0 = P0_(t); if (pay1) y1 = P1_(t); }
            BQ_SYNCG(gmask);
            if (i >= 0) { ev[t - 1] = i; V_(t - 1) = x; inv[i] = t - 1; if (pay0) P0_(t - 1) = y0; if (pay1) P1_(t - 1) = y1; }
            BQ_SYNCG(gmask);
        }
    } else if (j < p) {  // earlier: entries j..p-1 move one slot towards the tail
        for (int top = p - 1; top >= j; top -= W) {
            const int t = top - lane;
            int i = -1; double x = 0.0, y0 = 0.0, y1 = 0.0;
            if (t >= j) { i = ev[t]; x = V_(t); if (pay0) y0 = P0_(t); if (pay1) y1 = P1_(t); }
            BQ_SYNCG(gmask);
            if (i >= 0) { ev[t + 1] = i; V_(t + 1) = x; inv[i] = t + 1; if (pay0) P0_(t + 1) = y0; if (pay1) P1_(t + 1) = y1; }
            BQ_SYNCG(gmask);
        }
    }
    if (lane == 0) { ev[j] = m; V_(j) = t_new; inv[m] = j; if (pay0) P0_(j) = m0; if (pay1) P1_(j) = m1; }
    BQ_SYNCG(gmask);
#undef V_
#undef P0_
#undef P1_
}

template <int W, int KMAX, int PS = 2>
BQ_HDI void dyn_commit(const Chain &c, const EventCtx &ev, double x, int lane, unsigned gmask) {
    // ---- everything that is computed on the old state first ------------------------------------------------
    int j1 = ev.p1;
    if (ev.e1 >= 0 && ev.t1 != Q_DELAY) {
        const QInfo &q1 = c.qi[ev.q1];
        Touch tmp;
        tmp.reset();
        if (ev.t1 == Q_GGK) j1 = new_rank(c, q1, ev.p1, x, ev.x0, tmp);
        else if (PS) {  // PS, GG1R: rank among the other arrivals
            j1 = upper_from(c.rec + BQ_REC_AT, q1.lo, q1.hi, ev.p1, x, c.S);
            if (c.at(ev.p1) <= x) --j1;
        }
    }
    Touch dummy;
    dummy.reset();
    if (PS && ev.t0 == Q_PS) {  // new service times of everything that overlaps the moved interval
        PsMove mv;
        ps_move_setup(c, ev.q0, ev.p0, ev.e, true, ev.x0, x, mv, dummy);
        ps_apply_move<true>(c, mv, lane, dummy);
    }
    if (PS && ev.e1 >= 0 && ev.t1 == Q_PS) {
        PsMove mv;
        ps_move_setup(c, ev.q1, ev.p1, ev.e1, false, ev.x0, x, mv, dummy);
        ps_apply_move<true>(c, mv, lane, dummy);
    }
    // ---- departure side --------------------------------------------------------------------------------------
    const bool log0 = (c.qi[ev.q0].dist != D_EXP);
    if (ev.t0 == Q_GGK) {  // e's own service, then the free-time vectors (and services) downstream of e
        const QInfo &q = c.qi[ev.q0];
        const int k = q.k;
        double *fi = c.fvec(q, ev.p0);
        FreeVec<KMAX> M;
        M.load(fi, k);
        if (log0 && lane == 0) c.ls(ev.p0) = log_d(x - dmax(ev.a_e, M.v[0]));
        M.replace_min(x);
        fi += c.S;
        for (int p = ev.p0 + 1; p < q.hi; ++p, fi += c.S) {
            if (M.equals(fi, k)) break;
            const double di = c.dq(p);
            BQ_SYNCG(gmask);
            if (lane == 0) {
                if (log0 && M.v[0] != fi[0]) c.ls(p) = log_d(di - dmax(c.at(p), M.v[0]));
                M.store(fi, k);
                if (KMAX > 1 && k > 1) c.set_contended(p, c.at(p) < M.v[KMAX > 1 ? 1 : 0]);
            }
            M.replace_min(di);
        }
    } else if (ev.t0 == Q_DELAY) {
        if (log0 && lane == 0) c.ls(ev.p0) = log_d(x - ev.a_e);
    }
    BQ_SYNCG(gmask);
    if (PS && ev.t0 == Q_PS) {  // the departure changes rank among the queue's sorted departures
        const QInfo &q = c.qi[ev.q0];
        const int pd = c.dpos[ev.e];
        int j = upper_from(c.dt, q.lo, q.hi, pd, x);
        if (c.dt[pd] <= x) --j;
        reslot<W>(c.dt, 1, c.dord, c.dpos, nullptr, 1, nullptr, 1, pd, j, x, lane, gmask);
    }
    if (PS >= 2 && ev.t0 == Q_GG1R) {  // the event takes its new place in the departure order, its arrival time with it
        const QInfo &q = c.qi[ev.q0];
        const int pd = c.dpos[ev.e];
        int jp, jn;
        const int j = gg1r_new_slot(c, q, pd, ev.x0, x, jp, jn, dummy);
        reslot<W>(c.dt, 1, c.dord, c.dpos, c.sq, 1, nullptr, 1, pd, j, x, lane, gmask);
    }
    if (lane == 0) { c.d[ev.e] = x; c.dq(ev.p0) = x; }
    BQ_SYNCG(gmask);
    // ---- arrival side ------------------------------------------------------------------------------------------
    if (ev.e1 < 0) return;
    const bool log1 = (c.qi[ev.q1].dist != D_EXP);
    if (ev.t1 == Q_GGK) {  // re-rank e1, then refresh vectors and services from the first changed position
        const QInfo &q = c.qi[ev.q1];
        const int k = q.k, p1 = ev.p1;
        const int ps = (j1 < p1) ? j1 : p1, pe = (j1 < p1) ? p1 : j1;
        reslot<W>(c.rec + BQ_REC_AT, c.S, c.ord, c.pos, c.rec + BQ_REC_DQ, c.S, log1 ? c.rec + c.ols : nullptr, c.S, p1, j1, x,
                  lane, gmask);
        FreeVec<KMAX> M;
        double *fi = c.fvec(q, ps);
        M.load(fi, k);  // predecessors of position ps are untouched
        for (int t = ps; t < q.hi; ++t, fi += c.S) {
            if (t > pe && M.equals(fi, k)) break;
            const double di = c.dq(t);
            BQ_SYNCG(gmask);
            if (lane == 0) {
                if (log1 && (t <= pe || M.v[0] != fi[0])) c.ls(t) = log_d(di - dmax(c.at(t), M.v[0]));
                M.store(fi, k);
                if (KMAX > 1 && k > 1) c.set_contended(t, c.at(t) < M.v[KMAX > 1 ? 1 : 0]);
            }
            M.replace_min(di);
        }
        BQ_SYNCG(gmask);
    } else if (PS && ev.t1 == Q_PS) {
        reslot<W>(c.rec + BQ_REC_AT, c.S, c.ord, c.pos, c.rec + BQ_REC_DQ, c.S, c.sq, 1, ev.p1, j1, x, lane, gmask);
    } else if (PS >= 2 && ev.t1 == Q_GG1R) {  // new rank among the arrivals; the departure order is untouched
        reslot<W>(c.rec + BQ_REC_AT, c.S, c.ord, c.pos, c.rec + BQ_REC_DQ, c.S, nullptr, 1, ev.p1, j1, x, lane, gmask);
        if (lane == 0) c.sq[c.dpos[ev.e1]] = x;
        BQ_SYNCG(gmask);
    } else {
        if (lane == 0) {
            c.at(ev.p1) = x;
            if (log1) c.ls(ev.p1) = log_d(ev.d1 - x);
        }
        BQ_SYNCG(gmask);
    }
}

// ---------------------------------------------------------------------------------------------
// Where the conditional is -inf, known BEFORE the shrink loop (batch kernel).  Two thirds of all evaluations on the
// G/G/3 tandem return -inf -- the bracket [a_e, min(d_e', Tmax)] is far wider than the support -- and they hold 70 %
// of the walk steps (host emulation, 5 000-task tandem).  Each of the four ways a move can make a service negative
// has a WITNESS that a short scan next to the event finds once per update:
//   d_e later:    a position p behind e (F[p][0] >= d_e) that leaves before the second server is free, F[p][1] > d_p:
//                 it would have to wait for min(F[p][1], x) > d_p as soon as x > d_p  ->  -inf for every x > d_p
//   d_e earlier:  only e's own service, x - max(a_e, F[p0][0]) < 0                     ->  -inf for every x < that start
//   a_e' earlier: an event p that e' would overtake (a_p > x) with F[p][1] > d_p and d_e' > d_p: with e' ahead of it
//                 its server is free at min(F[p][1], d_e') > d_p                        ->  -inf for every x < a_p
//   a_e' later:   the event p at which the departures later than d_e' among the predecessors of e' reach k (those in
//                 F[p1] plus the events e' lets pass): every server is busy past d_e'     ->  -inf for every x >= a_p
// A candidate beyond a witness is rejected without a walk; it still counts as an evaluation, and every candidate that
// is not provably outside is evaluated as before, so the update is exactly the sequential sampling._slice (the scans are
// bounded: a missing witness only means less pruning).  Delay stations have their support in closed form; processor
// sharing and random-order queues are left to their evaluations.  The positions scanned are part of the read set.
// ---------------------------------------------------------------------------------------------
#ifndef BQ_SUPPORT_SCAN
#define BQ_SUPPORT_SCAN 96  // default number of positions a witness scan looks at (DynArgs::scan; measured: profiles/r2_support_scan.log)
#endif
struct Support {
    double lo, hi, hi_open;  // x < x0 needs x >= lo; x > x0 needs x <= hi and x < hi_open
    BQ_HDF bool outside(double x, double x0) const { return (x > x0) ? (x > hi || x >= hi_open) : (x < x0 && x < lo); }
};

template <int KMAX>
BQ_HDI void dyn_support(const Chain &c, const EventCtx &ev, double lower, double upper, Support &sp, Touch &tc,
                        int scan = BQ_SUPPORT_SCAN) {
    sp.lo = -INFINITY; sp.hi = INFINITY; sp.hi_open = INFINITY;
    const int S = c.S;
    if (ev.t0 == Q_GGK) {
        const QInfo &q = c.qi[ev.q0];
        const int k = q.k, p0 = ev.p0;
        const double *f = c.fvec(q, p0);
        sp.lo = dmax(ev.a_e, f[0]);
        if (p0 + 1 < q.hi && upper > ev.x0) {
            int far;
            int p = free_min_bound(f + S, S, p0 + 1, q.hi, ev.x0, false, &far);
            tc.p0(far);
            double T = INFINITY;
            bool stop = false;
            for (int n = 0; p < q.hi && n < scan && !stop; p += 4, n += 4) {
                double f0_[4], f1_[4], dp_[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int pu = (p + u < q.hi - 1) ? p + u : q.hi - 1;
                    const double *fp = f + (size_t)(pu - p0) * S;
                    f0_[u] = fp[0]; f1_[u] = (k > 1) ? fp[1] : INFINITY; dp_[u] = fp[BQ_REC_DQ - BQ_REC_F];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {  // only the records the decision looked at belong to the read set
                    if (stop || p + u >= q.hi) break;
                    tc.p0(p + u);
                    if (f0_[u] >= ((T < upper) ? T : upper)) { stop = true; break; }  // nobody further on leaves before T, or is reached at all
                    if (f1_[u] > dp_[u] && dp_[u] < T) T = dp_[u];
                }
            }
            sp.hi = T;
        }
    } else if (ev.t0 == Q_DELAY) {
        sp.lo = ev.a_e;
    }
    if (ev.e1 < 0) return;
    if (ev.t1 == Q_GGK) {
        const QInfo &q = c.qi[ev.q1];
        const int k = q.k, p1 = ev.p1;
        const double d1 = ev.d1;
        if (lower < ev.x0) {
            // four records per trip, fetched before any of them is looked at: their misses overlap
            bool stop = false;
            for (int base = p1 - 1, n = 0; base >= q.lo && n < scan && !stop; base -= 4, n += 4) {
                double ap_[4], f1_[4], dp_[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int pu = (base - u > q.lo) ? base - u : q.lo;
                    const double *r = c.rec + (size_t)pu * S;
                    ap_[u] = r[BQ_REC_AT]; f1_[u] = (k > 1) ? r[BQ_REC_F + 1] : INFINITY; dp_[u] = r[BQ_REC_DQ];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (stop || base - u < q.lo) break;
                    tc.p1(base - u);
                    if (!(ap_[u] > lower)) { stop = true; break; }  // no candidate overtakes this one
                    if (f1_[u] > dp_[u] && d1 > dp_[u]) { sp.lo = dmax(sp.lo, ap_[u]); stop = true; break; }
                }
            }
        }
        if (upper > ev.x0) {
            const double *f1v = c.fvec(q, p1);
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < KMAX; ++j)
                if (j < k && f1v[BQ_FSLOT(j)] > d1) ++cnt;
            if (cnt < k) {
                bool stop = false;
                for (int base = p1 + 1, n = 0; base < q.hi && n < scan && !stop; base += 4, n += 4) {
                    double ap_[4], dp_[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int pu = (base + u < q.hi - 1) ? base + u : q.hi - 1;
                        const double *r = c.rec + (size_t)pu * S;
                        ap_[u] = r[BQ_REC_AT]; dp_[u] = r[BQ_REC_DQ];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (stop || base + u >= q.hi) break;
                        tc.p1(base + u);
                        if (ap_[u] > upper) { stop = true; break; }  // no candidate lets this one pass
                        if (dp_[u] > d1 && ++cnt >= k) { sp.hi_open = ap_[u]; stop = true; break; }
                    }
                }
            }
        }
    } else if (ev.t1 == Q_DELAY) {
        sp.hi = (ev.d1 < sp.hi) ? ev.d1 : sp.hi;
    }
}

// ---------------------------------------------------------------------------------------------
// One lane, one event: the whole update computed WITHOUT writing anything -- one step of sampling._slice on the exact
// conditional (sequential candidates) or one MH-within-Gibbs proposal + accept decision -- together with the ranges
// it read (R) and the ranges the accepted move will write (Wt = what the accepted candidate's evaluation touched).
// Used by sweep_batch_kernel, where the 32 lanes of a warp update 32 different events of one chain at once.
// ---------------------------------------------------------------------------------------------
struct LaneUpdate {
    double x_new;
    int changed, counted, rejected, stuck;
    long long evals;
    Touch R, Wt;
};

struct LaneParams {
    unsigned long long seed, gchain;
    unsigned int gsw;
    int mode, flags;
    double tmax;
    int scan = BQ_SUPPORT_SCAN;
};

template <int KMAX, bool MH = true, bool SLICE = true, int PS = 2>
BQ_HDI void lane_update(const Chain &c, const EventCtx &ev, const LaneParams &P, LaneUpdate &out) {
    out.x_new = ev.x0; out.changed = 0; out.counted = 0; out.rejected = 0; out.stuck = 0; out.evals = 0;
    out.R.reset(); out.Wt.reset();
    out.R.p0(ev.p0);
    if (ev.e1 >= 0) out.R.p1(ev.p1);
    Touch last;
    last.reset();
    auto g = [&](double x) {
        last.reset();
        const double v = dyn_logcond<KMAX, PS>(c, ev, x, last);
        out.R.merge(last);
        return v;
    };
    if (SLICE && (!MH || P.mode != MODE_MH)) {
        const double lower = ev.a_e, upper = (ev.e1 >= 0) ? fmin(ev.d1, P.tmax) : P.tmax;  // qnet.pyx:698-701, 732-734
        Rng rng;
        rng.init(P.seed, P.gchain, P.gsw, (unsigned int)ev.e, TAG_SLICE);
        out.counted = 1;
        Support sp;
        dyn_support<KMAX>(c, ev, lower, upper, sp, out.R, P.scan);
        // sampling._slice with finite bounds (slice_finite_seq, bq_dist.cuh), arranged for a warp whose lanes run it
        // side by side: a lane first passes over the candidates that are provably outside the support -- arithmetic
        // only -- and the lanes then walk their next real candidate TOGETHER, so a round of the warp costs as many
        // walks as its busiest lane needs instead of one sparsely populated walk per candidate index.
        const double x0 = ev.x0;
        // g(x0) is the likelihood delta of not moving: exactly 0 when the two services it would touch are non-negative
        // (no term of a FIFO / delay / PS diff list changes at x = x0), so the walk is only taken when the state is
        // invalid there -- or a random-order queue is involved, whose feasibility rule can fail at x0 itself
        double gx0;
        {
            bool plain = (ev.t0 == Q_GGK || ev.t0 == Q_DELAY) && (ev.e1 < 0 || ev.t1 == Q_GGK || ev.t1 == Q_DELAY);
            if (plain) {
                const double st0 = (ev.t0 == Q_GGK) ? dmax(ev.a_e, c.fvec(c.qi[ev.q0], ev.p0)[0]) : ev.a_e;
                plain = (x0 - st0 >= 0.0);
                if (plain && ev.e1 >= 0) {
                    const double st1 = (ev.t1 == Q_GGK) ? dmax(x0, c.fvec(c.qi[ev.q1], ev.p1)[0]) : x0;
                    plain = (ev.d1 - st1 >= 0.0);
                }
            }
            gx0 = plain ? 0.0 : g(x0);
        }
        ++out.evals;
        const double logy = gx0 - (-log_d(1.0 - rng.uniform()));
        double x1 = x0;
        if (logy > BQ_NEG_INF) {
            double L = lower, R = upper;
            int it = 0;
            for (;;) {
                bool have = false;
                double xc = x0;
                while (it < 1000) {
                    xc = L + (R - L) * (it == 0 ? rng.uniform() : rng.uniform32());  // bq_rng.cuh: candidate uniforms
                    ++it;
                    ++out.evals;
                    if (!sp.outside(xc, x0)) { have = true; break; }
                    if (xc > x0) R = xc; else L = xc;
                    if (R - L < 1e-10) break;
                }
                if (!have) break;
                if (g(xc) >= logy) { x1 = xc; break; }
                if (xc > x0) R = xc; else L = xc;
                if (R - L < 1e-10) break;
            }
        }
        if (x1 != x0) { out.changed = 1; out.x_new = x1; out.Wt = last; }
        else out.stuck = 1;
        return;
    }
    if (!MH) return;
    // MH-within-Gibbs (gibbs_resample_pair / gibbs_resample_final, qnet.pyx:382-498); see mh_update for the streams
    const bool final_evt = (ev.e1 < 0);
    if (final_evt && (P.flags & FLAG_NO_FINAL)) return;
    MhProp pr;
    mh_setup(c, ev, pr);
    if (pr.Uh - pr.Lh < 1e-10) return;
    if (!final_evt && ev.d1 - ev.a_e < 1e-10) return;
    Rng aux;
    aux.init(P.seed, P.gchain, P.gsw, (unsigned int)ev.e, TAG_MH_AUX);
    const double u_acc = aux.uniform();
    (void)aux.uniform();
    const double h0 = pr.eval(ev.x0);
    double x1 = ev.x0;
    if (!(h0 > BQ_NEG_INF)) {
        int tries = 0;
        for (; tries < 25; ++tries) {
            const double u = aux.uniform();
            x1 = final_evt ? pr.b0 + (-log_d(1.0 - u)) : ev.a_e + (ev.d1 - ev.a_e) * u;
            if (pr.eval(x1) > BQ_NEG_INF) break;
        }
        if (tries >= 25) return;
    }
    auto hp = [&](double x) { return pr.eval(x); };
    for (int step = 0; step < 2; ++step) {
        Rng rng;
        rng.init(P.seed, P.gchain, P.gsw, (unsigned int)ev.e, step == 0 ? TAG_MH : TAG_MH2);
        x1 = final_evt ? slice_stepout_seq(hp, x1, pr.Lh, pr.Uh, rng, out.evals)
                       : slice_finite_seq(hp, x1, pr.Lh, pr.Uh, rng, out.evals);
    }
    out.counted = 1;
    const double g_new = g(x1);
    const double ratio = exp(h0 - pr.eval(x1) + g_new);
    if (u_acc < ratio) {
        if (x1 != ev.x0) { out.changed = 1; out.x_new = x1; out.Wt = last; }
    } else {
        out.rejected = 1;
    }
}

// does a lane that read R depend on a lane that writes W?
BQ_HDF bool ranges_overlap(int alo, int ahi, int blo, int bhi) { return alo <= bhi && blo <= ahi; }
BQ_HDF bool touch_conflict(const Touch &r, const Touch &w) {
    return ranges_overlap(r.lo0, r.hi0, w.lo0, w.hi0) || ranges_overlap(r.lo0, r.hi0, w.lo1, w.hi1) ||
           ranges_overlap(r.lo1, r.hi1, w.lo0, w.hi0) || ranges_overlap(r.lo1, r.hi1, w.lo1, w.hi1) ||
           ranges_overlap(r.dlo, r.dhi, w.dlo, w.dhi);
}

// service time of the event at position p of queue q in the current state
BQ_HDF double dyn_service(const Chain &c, const QInfo &q, int p) {
    if (q.type == Q_GGK) return c.dq(p) - dmax(c.at(p), c.fvec(q, p)[0]);
    if (q.type == Q_DELAY) return c.dq(p) - c.at(p);
    if (q.type == Q_GG1R) {  // previous departure of the queue (queues.pyx:808-816)
        const int j = c.dpos[c.ord[p]];
        return c.dq(p) - dmax(c.at(p), (j > q.lo) ? c.dt[j - 1] : 0.0);
    }
    return c.sq[p];
}

// QueueGG1R.allEventLik beyond the service densities (queues.pyx:1009-1024): - log N(commencement) for an event with
// s != 0; *bad when it started on arrival with others present
BQ_HDI double gg1r_event_term(const Chain &c, const QInfo &q, int p, double s, bool *bad) {
    const int j = c.dpos[c.ord[p]];
    const double a = c.at(p), d = c.dq(p);
    double w = d - a - s;
    if (w < 0.0) w = 0.0;
    if (w < 1e-10) {
        const int na = upper_from(c.rec + BQ_REC_AT, q.lo, q.hi, p, a, c.S) - upper_from(c.dt, q.lo, q.hi, j, a);
        if (na > 1) *bad = true;
    }
    if (s == 0.0) return 0.0;
    const double cm = dmax(a, d - s);
    return -log_count(upper_from(c.rec + BQ_REC_AT, q.lo, q.hi, p, cm, c.S) - upper_from(c.dt, q.lo, q.hi, j, cm));
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// group-wide helpers
// ---------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ double group_sum(double x, unsigned gmask) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) x += __shfl_xor_sync(gmask, x, o, W);
    return x;
}
template <int W>
__device__ __forceinline__ double group_max(double x, unsigned gmask) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(gmask, x, o, W));
    return x;
}

// per-queue sufficient statistics of one chain (queues.pyx:1316-1336), every lane returns with ss[] filled
template <int W>
static __device__ BQ_NOINLINE void dyn_suffstats(const Chain &c, int Q, SuffStat *ss, int lane, unsigned gmask, bool want_logp,
                              double *logp_out) {
    double logp_acc = 0.0;
    for (int q = 0; q < Q; ++q) {
        const QInfo &qi = c.qi[q];
        const int dist = qi.dist;
        double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // n, sum, sumlog, nlog, wait, nzero, logp, bad
        for (int p = qi.lo + lane; p < qi.hi; p += W) {
            const double s = dyn_service(c, qi, p);
            const double w = c.dq(p) - c.at(p) - s;
            v[4] += (w > 0.0) ? w : 0.0;
            if (s > 0.0) {
                v[0] += 1.0; v[1] += s;
                if (dist != D_EXP) {
                    if (dist == D_GAMMA || s >= 1e-50) { v[2] += (qi.type >= Q_PS) ? log_d(s) : c.ls(p); v[3] += 1.0; }
                }
            } else if (s == 0.0) v[5] += 1.0;
            if (want_logp) {
                if (s < 0.0) v[7] += 1.0;
                else v[6] += (qi.type >= Q_PS) ? lpdf(c.qc[q], s) : lpdf_ls(c.qc[q], s, c.ls(p));
                if (qi.type == Q_GG1R && s >= 0.0) {
                    bool bad = false;
                    v[6] += gg1r_event_term(c, qi, p, s, &bad);
                    if (bad) v[7] += 1.0;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = group_sum<W>(v[i], gmask);
        double sq = 0.0;
        if (dist == D_LOGNORMAL) {  // second pass, as distributions.pyx:276-299
            const double mean_log = (v[3] > 0.0) ? v[2] / v[3] : 0.0;
            for (int p = qi.lo + lane; p < qi.hi; p += W) {
                const double s = dyn_service(c, qi, p);
                if (s >= 1e-50) { const double z = ((qi.type >= Q_PS) ? log_d(s) : c.ls(p)) - mean_log; sq += z * z; }
            }
            sq = group_sum<W>(sq, gmask);
        }
        if (lane == 0) {
            SuffStat &o = ss[q];
            o.n = v[0]; o.sum = v[1]; o.sumlog = v[2]; o.nlog = v[3]; o.wait = v[4]; o.nzero = v[5];
            o.nall = (double)(qi.hi - qi.lo); o.sumsq = sq;
        }
        logp_acc += (v[7] > 0.0) ? BQ_NEG_INF : v[6];
    }
    __syncwarp(gmask);
    if (logp_out) *logp_out = logp_acc;
}

// ---------------------------------------------------------------------------------------------
// One step of sampling._slice with finite bounds (sampling.pyx:283-298, 344-366), by the W lanes of a group.
// Lane l of round r evaluates candidate c = r W + l (candidate 0 is x0 itself); candidate c uses uniform number c
// of the stream `rng`, uniform 0 draws the slice level.  A candidate's position depends only on where the earlier
// ones fell, so the first accepted candidate in order is the outcome of the sequential loop.
// ---------------------------------------------------------------------------------------------
struct GroupLane {
    int lane;
    unsigned gmask, gshift, gbits;
};

template <int W, class G>
__device__ __forceinline__ bool group_slice(const G &g, double x0, double lower, double upper, const Rng &rng,
                                            const GroupLane &gl, double &xacc, long long &n_evals, long long &n_stuck) {
    double L = lower, R = upper, logy = 0.0;
    for (int cbase = 0;; cbase += W) {
        const unsigned int cidx = (unsigned int)(cbase + gl.lane);
        const double u = cand_uniform(rng, cidx);  // candidate cidx >= 1 (lane 0 of the first round holds x0 itself)
        double myx = x0, myL = L, myR = R;
#pragma unroll 4
        for (int jj = (cbase == 0) ? 1 : 0; jj < W; ++jj) {
            const double uj = __shfl_sync(gl.gmask, u, jj, W);
            const double xj = L + (R - L) * uj;
            if (xj > x0) R = xj; else L = xj;
            if (gl.lane == jj) { myx = xj; myL = L; myR = R; }
        }
        const double gv = g(myx);
        if (cbase == 0) {
            const double g0 = __shfl_sync(gl.gmask, gv, 0, W);
            double u0a, u0b;
            stream_pair(rng, 0u, u0a, u0b);
            logy = g0 - (-log_d(1.0 - u0a));  // rk_exponential = -log(1 - U)   (cdist.c:107-110)
            if (!(logy > BQ_NEG_INF)) { n_stuck += 1; n_evals += 1; return false; }  // sampling.pyx:289-291
        }
        const bool cand = (cidx != 0u);
        const bool ok = cand && (gv >= logy);
        const bool dead = cand && !ok && (myR - myL < 1e-10);  // sampling.pyx:360-362 (we keep x0)
        const unsigned okm = (__ballot_sync(gl.gmask, ok) >> gl.gshift) & gl.gbits;
        const unsigned deadm = (__ballot_sync(gl.gmask, dead) >> gl.gshift) & gl.gbits;
        const int fo = okm ? (__ffs(okm) - 1) : 64, fd = deadm ? (__ffs(deadm) - 1) : 64;
        if (fo < fd) {
            xacc = __shfl_sync(gl.gmask, myx, fo, W);
            n_evals += cbase + fo + 1;
            return true;
        }
        if (fd < 64) { n_stuck += 1; n_evals += cbase + fd + 1; return false; }
        if (cbase + W >= BQ_SLICE_MAX_ITERS) { n_stuck += 1; n_evals += cbase + W; return false; }
    }
}

// One MH-within-Gibbs update of d_e (gibbs_resample_pair / gibbs_resample_final, qnet.pyx:382-498).
// Streams of (chain, sweep, event): TAG_MH and TAG_MH2 feed the two slice steps on the proposal, TAG_MH_AUX holds the
// accept uniform (number 0) and the re-initialisation draws (numbers 2..).
template <int W, int KMAX>
static __device__ BQ_NOINLINE void mh_update(const Chain &c, const EventCtx &ev, const DynArgs &A, unsigned long long gchain,
                          unsigned int gsw, const GroupLane &gl, long long &n_ev, long long &n_evals, long long &n_stuck,
                          long long &n_reject, Touch &tc) {
    const bool final_evt = (ev.e1 < 0);
    if (final_evt && (A.flags & FLAG_NO_FINAL)) return;  // sample_final = False (qnet.pyx:361)
    MhProp pr;
    mh_setup(c, ev, pr);
    if (pr.Uh - pr.Lh < 1e-10) return;                      // qnet.pyx:385, 440: don't sample small intervals
    if (!final_evt && ev.d1 - ev.a_e < 1e-10) return;
    Rng aux;
    aux.init(A.seed, gchain, gsw, (unsigned int)ev.e, TAG_MH_AUX);
    const double u_acc = aux.uniform();
    (void)aux.uniform();  // the spare of block 0 is not used: re-initialisation draws start at block 1
    const double h0 = pr.eval(ev.x0);
    double x_init = ev.x0;
    if (!(h0 > BQ_NEG_INF)) {  // qnet.pyx:389-397, 444-451
        int tries = 0;
        for (; tries < 25; ++tries) {
            const double u = aux.uniform();
            x_init = final_evt ? pr.b0 + (-log_d(1.0 - u)) : ev.a_e + (ev.d1 - ev.a_e) * u;
            if (pr.eval(x_init) > BQ_NEG_INF) break;
        }
        if (tries >= 25) return;
    }
    double x_new = x_init;
    for (int step = 0; step < 2; ++step) {
        Rng rng;
        rng.init(A.seed, gchain, gsw, (unsigned int)ev.e, step == 0 ? TAG_MH : TAG_MH2);
        if (final_evt) {
            x_new = slice_stepout_seq([&](double x) { return pr.eval(x); }, x_new, pr.Lh, pr.Uh, rng, n_evals);
        } else {
            double xa = x_new;
            if (group_slice<W>([&](double x) { return pr.eval(x); }, x_new, pr.Lh, pr.Uh, rng, gl, xa, n_evals, n_stuck))
                x_new = xa;
        }
    }
    n_ev += 1;
    // ratio = exp(h(d_e) - h(d_new)) * prod over the diff lists of exp(lpdf(s_new) - lpdf(s_old))   (qnet.pyx:425-431)
    const double g_new = dyn_logcond<KMAX>(c, ev, x_new, tc);
    const double ratio = exp(h0 - pr.eval(x_new) + g_new);
    if (u_acc < ratio) {
        if (x_new != ev.x0) dyn_commit<W, KMAX>(c, ev, x_new, gl.lane, gl.gmask);
    } else {
        n_reject += 1;
    }
}

// ---------------------------------------------------------------------------------------------
// the sweep kernel
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t dyn_group_smem(int Q, int P) {
    size_t b = (size_t)Q * (sizeof(QConst) + sizeof(SuffStat) + sizeof(QInfo)) + (size_t)((P + 1) & ~1) * sizeof(double);
    return (b + 15) & ~(size_t)15;
}

// the group's slice of shared memory: [QConst Q][SuffStat Q][theta P (padded to even)][QInfo Q]
struct GroupSmem {
    QConst *qc;
    SuffStat *ss;
    double *th;
    QInfo *qi;
    __device__ __forceinline__ void carve(unsigned char *base, int Q, int P) {
        qc = reinterpret_cast<QConst *>(base);
        ss = reinterpret_cast<SuffStat *>(qc + Q);
        th = reinterpret_cast<double *>(ss + Q);
        qi = reinterpret_cast<QInfo *>(th + ((P + 1) & ~1));
    }
};

template <int W, int KMAX>
__global__ void __launch_bounds__(128) sweep_dynamic_kernel(const DynArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = blockDim.x / W;
    const int gib = threadIdx.x / W, lane = threadIdx.x % W;
    const int chain = blockIdx.x * G + gib;
    if (chain >= A.n_chains) return;  // whole groups leave together; no block-wide barrier below
    const unsigned gshift = (threadIdx.x & 31) & ~(W - 1);
    const unsigned gbits = (W == 32) ? 0xffffffffu : ((1u << W) - 1u);
    const unsigned gmask = gbits << gshift;
    const GroupLane gl = {lane, gmask, gshift, gbits};
    GroupSmem sm;
    sm.carve(smem_raw + (size_t)gib * dyn_group_smem(A.Q, A.P), A.Q, A.P);
    QConst *qc = sm.qc;
    SuffStat *ss = sm.ss;
    double *th = sm.th;
    const int Q = A.Q, P = A.P, E = A.E;

    for (int i = lane; i < P; i += W) th[i] = A.theta[(size_t)chain * P + i];
    for (int q = lane; q < Q; q += W) sm.qi[q] = A.qinfo[q];
    __syncwarp(gmask);
    for (int q = lane; q < Q; q += W) {
        const int o = __ldg(&A.q_poff[q]), dist = sm.qi[q].dist;
        qconst_set(qc[q], dist, th[o], (dist == D_EXP) ? 0.0 : th[o + 1]);
    }
    __syncwarp(gmask);
    Chain c;
    c.bind(A, chain, sm.qi, qc);

    const unsigned long long gchain = (unsigned long long)(A.chain_offset + chain);
    long long n_ev = 0, n_evals = 0, n_stuck = 0, n_pevals = 0, n_steps = 0, n_reject = 0;
    double tmax = 0.0;
    const int4 *order4 = reinterpret_cast<const int4 *>(A.order);

    for (int sw = 0; sw < A.n_sweeps; ++sw) {
        const unsigned int gsw = A.sweep_base + (unsigned int)sw;
        if (sw == 0 || (A.flags & FLAG_TMAX_PER_SWEEP)) {  // Tmax = 2 max d   (qnet.pyx:672)
            double m = 0.0;
            for (int i = lane; i < E; i += W) m = fmax(m, c.d[i]);
            tmax = 2.0 * group_max<W>(m, gmask);
        }
        const int n_order = (A.flags & FLAG_PARAMS_ONLY) ? 0 : A.n_order;
        int4 rec_next = make_int4(0, -1, 0, 0);
        const bool rev = (A.flags & FLAG_REVERSE) != 0;  // slice_resample(reverse=True): the scan order backwards
        if (n_order > 0) rec_next = __ldg(&order4[rev ? n_order - 1 : 0]);
        for (int idx = 0; idx < n_order; ++idx) {
            const int4 r4 = rec_next;
            if (idx + 1 < n_order) rec_next = __ldg(&order4[rev ? n_order - 2 - idx : idx + 1]);  // in flight during this event
            const DynRec rec = {r4.x, r4.y, r4.z, r4.w};
            EventCtx ev;
            event_load(c, rec, ev);
            Touch tc;
            tc.reset();
            if (A.mode == MODE_MH) {
                mh_update<W, KMAX>(c, ev, A, gchain, gsw, gl, n_ev, n_evals, n_stuck, n_reject, tc);
                n_steps += tc.steps;
                __syncwarp(gmask);
                continue;
            }
            const double lower = ev.a_e;
            const double upper = (ev.e1 >= 0) ? fmin(ev.d1, tmax) : tmax;  // qnet.pyx:698-701, 732-734
            Rng rng;
            rng.init(A.seed, gchain, gsw, (unsigned int)ev.e, TAG_SLICE);
            n_ev += 1;
            double xacc = ev.x0;
            Support sp;  // candidates provably outside the support need no walk (dyn_support)
            dyn_support<KMAX>(c, ev, lower, upper, sp, tc, A.scan);
            const bool accepted = group_slice<W>([&](double x) { return sp.outside(x, ev.x0) ? BQ_NEG_INF : dyn_logcond<KMAX>(c, ev, x, tc); },
                                                 ev.x0, lower, upper, rng, gl, xacc, n_evals, n_stuck);
            n_steps += tc.steps;
            if (accepted && xacc != ev.x0) dyn_commit<W, KMAX>(c, ev, xacc, lane, gmask);
            __syncwarp(gmask);
        }
        const bool trace = (A.flags & FLAG_TRACE) != 0;
        if (A.update != UPD_NONE || trace) {
            double logp = 0.0;
            dyn_suffstats<W>(c, Q, ss, lane, gmask, trace, &logp);
            const size_t recno = (size_t)(A.trace_base + sw);
            if (trace) {
                for (int q = lane; q < Q; q += W)
                    A.tr_wait[(recno * A.n_chains + chain) * Q + q] = (ss[q].nall > 0.0) ? ss[q].wait / ss[q].nall : 0.0;
                if (lane == 0) A.tr_logp[recno * A.n_chains + chain] = logp;
            }
            if (A.update != UPD_NONE) {
                for (int q = lane; q < Q; q += W) {
                    const int o = __ldg(&A.q_poff[q]), dist = sm.qi[q].dist;
                    double p0 = th[o], p1 = (dist == D_EXP) ? 0.0 : th[o + 1];
                    if (A.update == UPD_BAYES) {
                        Rng prng;
                        prng.init(A.seed, gchain, gsw, (unsigned int)q, TAG_PARAM);
                        sample_params(dist, ss[q], prng, p0, p1, &n_pevals);
                    } else {
                        estimate_params(dist, ss[q], p0, p1);
                    }
                    th[o] = p0;
                    if (dist != D_EXP) th[o + 1] = p1;
                    qconst_set(qc[q], dist, p0, p1);
                }
            }
            __syncwarp(gmask);
            if (trace)
                for (int i = lane; i < P; i += W) A.tr_theta[(recno * A.n_chains + chain) * P + i] = th[i];
        }
    }
    for (int i = lane; i < P; i += W) A.theta[(size_t)chain * P + i] = th[i];
    n_pevals = (long long)group_sum<W>((double)n_pevals, gmask);
    if (lane == 0) {
        long long *st = A.stats + (size_t)chain * BQ_NSTAT;
        st[0] += n_ev; st[1] += n_evals; st[2] += n_pevals; st[3] += n_stuck; st[4] += n_reject; st[5] += n_steps;
    }
}

// ---------------------------------------------------------------------------------------------
// The batch kernel: one WARP owns a chain and its 32 lanes update 32 consecutive events of the sweep order at once.
// The order (built on the host) interleaves each queue's events with a stride of 1/32 of the queue, so the events of a
// batch are far apart.  Every lane computes its update on the unchanged state (lane_update) and the ranges it read and
// would write; writers publish their ranges in a per-chain claim array; the longest prefix of the batch in which no
// lane read a range claimed by an earlier lane is committed, the rest is redone in the next round.  The outcome is
// therefore EXACTLY the sequential sweep in the fixed order -- whatever the batch boundaries -- with 32 updates of a
// chain in flight per warp instead of one.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void claim_range(int *claim, int lo, int hi, int stamp) {
    for (int p = lo; p <= hi; ++p) atomicMax(&claim[p], stamp);
}
// no early exit: the loads of a range are independent of each other and stay in flight together (the read ranges hold
// the witness scans, up to a few hundred positions)
__device__ __forceinline__ bool claimed_before(const int *claim, int lo, int hi, int round_base, int lane) {
    if (lo > hi) return false;  // an unset range is (INT_MAX, -1): nothing below may add to lo
    const int thr = round_base + 31 - lane;  // stamps of this round from earlier lanes are > thr
    bool hit = false;
    int p = lo;
    for (; p <= hi && (reinterpret_cast<size_t>(claim + p) & 15); ++p) hit |= __ldcg(&claim[p]) > thr;  // per-chain rows start anywhere
    for (; p + 3 <= hi; p += 4) {
        const int4 v = __ldcg(reinterpret_cast<const int4 *>(claim + p));
        hit |= (v.x > thr) | (v.y > thr) | (v.z > thr) | (v.w > thr);
    }
    for (; p <= hi; ++p) hit |= __ldcg(&claim[p]) > thr;
    return hit;
}

// the same with the lane's own stamp given (rounds of more than 32 events)
__device__ __forceinline__ bool claimed_before_thr(const int *claim, int lo, int hi, int thr) {
    if (lo > hi) return false;  // an unset range is (INT_MAX, -1): nothing below may add to lo
    // thr = the lane's own stamp: the stamps of earlier events of this round are larger, everything older is smaller
    bool hit = false;
    int p = lo;
    for (; p <= hi && (reinterpret_cast<size_t>(claim + p) & 15); ++p) hit |= __ldcg(&claim[p]) > thr;  // per-chain rows start anywhere
    for (; p + 3 <= hi; p += 4) {
        const int4 v = __ldcg(reinterpret_cast<const int4 *>(claim + p));
        hit |= (v.x > thr) | (v.y > thr) | (v.z > thr) | (v.w > thr);
    }
    for (; p <= hi; ++p) hit |= __ldcg(&claim[p]) > thr;
    return hit;
}

// MINB = CTAs of 128 threads resident per SM: 4 (128 registers per thread, 16 warps/SM) is fastest as long as all
// chains are resident; with more than 16 x 148 chains the 8-CTA build (64 registers, some spilling, 32 warps/SM) avoids
// a second, partly filled wave (measured: +12 % at 4096 chains, -17 % at 2048)
// MHK: the build for MH sweeps; the slice build carries no proposal code (fewer registers, less spilling).
template <int KMAX, int MINB, bool MHK, int PS>
__global__ void __launch_bounds__(128, MINB) sweep_batch_kernel(const DynArgs A, int *claim_all, int *claimd_all) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int W = 32;
    const unsigned FULL = 0xffffffffu;
    const int gib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + gib;
    if (chain >= A.n_chains) return;
    GroupSmem sm;
    sm.carve(smem_raw + (size_t)gib * dyn_group_smem(A.Q, A.P), A.Q, A.P);
    QConst *qc = sm.qc;
    SuffStat *ss = sm.ss;
    double *th = sm.th;
    const int Q = A.Q, P = A.P, E = A.E;
    int *claim = claim_all + (size_t)chain * E;
    int *claimd = claimd_all ? claimd_all + (size_t)chain * E : nullptr;

    for (int i = lane; i < P; i += W) th[i] = A.theta[(size_t)chain * P + i];
    for (int q = lane; q < Q; q += W) sm.qi[q] = A.qinfo[q];
    for (int i = lane; i < E; i += W) { claim[i] = 0; if (claimd) claimd[i] = 0; }
    __syncwarp();
    for (int q = lane; q < Q; q += W) {
        const int o = __ldg(&A.q_poff[q]), dist = sm.qi[q].dist;
        qconst_set(qc[q], dist, th[o], (dist == D_EXP) ? 0.0 : th[o + 1]);
    }
    __syncwarp();
    Chain c;
    c.bind(A, chain, sm.qi, qc);

    LaneParams LP;
    LP.seed = A.seed; LP.gchain = (unsigned long long)(A.chain_offset + chain); LP.mode = A.mode; LP.flags = A.flags;
    LP.scan = A.scan;
    long long n_ev = 0, n_evals = 0, n_stuck = 0, n_pevals = 0, n_steps = 0, n_reject = 0, n_redo = 0;
    const int4 *order4 = reinterpret_cast<const int4 *>(A.order);
    int round = 0;

    for (int sw = 0; sw < A.n_sweeps; ++sw) {
        LP.gsw = A.sweep_base + (unsigned int)sw;
        if (sw == 0 || (A.flags & FLAG_TMAX_PER_SWEEP)) {  // Tmax = 2 max d   (qnet.pyx:672)
            double m = 0.0;
            for (int i = lane; i < E; i += W) m = fmax(m, c.d[i]);
            LP.tmax = 2.0 * group_max<W>(m, FULL);
        }
        const int n_order = (A.flags & FLAG_PARAMS_ONLY) ? 0 : A.n_order;
        for (int s = 0; s < n_order;) {
            ++round;
            const int round_base = round * 32;
            const int idx = s + lane;
            const bool active = idx < n_order;
            EventCtx ev;
            LaneUpdate up;
            up.changed = 0; up.counted = 0; up.rejected = 0; up.stuck = 0; up.evals = 0;
            up.R.reset(); up.Wt.reset();
            if (active) {
                const int4 r4 = __ldg(&order4[(A.flags & FLAG_REVERSE) ? n_order - 1 - idx : idx]);
                const DynRec rec = {r4.x, r4.y, r4.z, r4.w};
                event_load(c, rec, ev);
                lane_update<KMAX, MHK, !MHK, PS>(c, ev, LP, up);
            }
            __syncwarp();
            if (active && up.changed) {  // publish what this lane will write
                const int stamp = round_base + 31 - lane;
                claim_range(claim, up.Wt.lo0, up.Wt.hi0, stamp);
                claim_range(claim, up.Wt.lo1, up.Wt.hi1, stamp);
                if (claimd) claim_range(claimd, up.Wt.dlo, up.Wt.dhi, stamp);
            }
            __syncwarp();
            bool conflict = false;
            if (active && lane > 0) {
                conflict = claimed_before(claim, up.R.lo0, up.R.hi0, round_base, lane) ||
                           claimed_before(claim, up.R.lo1, up.R.hi1, round_base, lane) ||
                           (claimd && claimed_before(claimd, up.R.dlo, up.R.dhi, round_base, lane));
            }
            const unsigned cm = __ballot_sync(FULL, conflict);
            const int nact = (n_order - s < 32) ? n_order - s : 32;
            const int fc = cm ? (__ffs(cm) - 1) : nact;
            if (active && lane < fc) {
                n_ev += up.counted; n_evals += up.evals; n_reject += up.rejected; n_steps += up.R.steps;
                if (LP.mode != MODE_MH) n_stuck += up.stuck;
                if (up.changed) dyn_commit<1, KMAX, PS>(c, ev, up.x_new, 0, 1u << lane);
            }
            n_redo += (lane == 0) ? (nact - fc) : 0;
            __syncwarp();
            s += fc;
        }
        const bool trace = (A.flags & FLAG_TRACE) != 0;
        if (A.update != UPD_NONE || trace) {
            double logp = 0.0;
            dyn_suffstats<W>(c, Q, ss, lane, FULL, trace, &logp);
            const size_t recno = (size_t)(A.trace_base + sw);
            if (trace) {
                for (int q = lane; q < Q; q += W)
                    A.tr_wait[(recno * A.n_chains + chain) * Q + q] = (ss[q].nall > 0.0) ? ss[q].wait / ss[q].nall : 0.0;
                if (lane == 0) A.tr_logp[recno * A.n_chains + chain] = logp;
            }
            if (A.update != UPD_NONE) {
                for (int q = lane; q < Q; q += W) {
                    const int o = __ldg(&A.q_poff[q]), dist = sm.qi[q].dist;
                    double p0 = th[o], p1 = (dist == D_EXP) ? 0.0 : th[o + 1];
                    if (A.update == UPD_BAYES) {
                        Rng prng;
                        prng.init(A.seed, LP.gchain, LP.gsw, (unsigned int)q, TAG_PARAM);
                        sample_params(dist, ss[q], prng, p0, p1, &n_pevals);
                    } else {
                        estimate_params(dist, ss[q], p0, p1);
                    }
                    th[o] = p0;
                    if (dist != D_EXP) th[o + 1] = p1;
                    qconst_set(qc[q], dist, p0, p1);
                }
            }
            __syncwarp();
            if (trace)
                for (int i = lane; i < P; i += W) A.tr_theta[(recno * A.n_chains + chain) * P + i] = th[i];
        }
    }
    for (int i = lane; i < P; i += W) A.theta[(size_t)chain * P + i] = th[i];
    // per-lane counters -> the chain's totals
    double tot[7] = {(double)n_ev, (double)n_evals, (double)n_pevals, (double)n_stuck, (double)n_reject, (double)n_steps,
                     (double)n_redo};
#pragma unroll
    for (int i = 0; i < 7; ++i) tot[i] = group_sum<W>(tot[i], FULL);
    if (lane == 0) {
        long long *st = A.stats + (size_t)chain * BQ_NSTAT;
#pragma unroll
        for (int i = 0; i < 7; ++i) st[i] += (long long)tot[i];
    }
}

// The same with NW warps per chain (round 2; a separate kernel so that the one-warp build above keeps its register
// allocation): with few chains per SM one warp per chain leaves the memory latency uncovered
// (config 4's share of a GPU is 2048 chains = 14 warps per SM), so NW warps of a CTA share a chain and a round
// holds 32 NW events: every warp computes its 32 updates, the claims of all of them are published, the first
// conflict of the whole round is found through shared memory and the prefix is committed by whichever warp holds
// each event -- the same sequential-sweep semantics, three named barriers per round.
template <int KMAX, int MINB, bool MHK, int PS, int NW>
__global__ void __launch_bounds__(128, MINB) sweep_batch_kernel_nw(const DynArgs A, int *claim_all, int *claimd_all) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned s_ballot[4];
    constexpr int W = 32;
    constexpr int RB = 32 * NW;  // events per round
    const unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gib = warp / NW, wic = warp % NW;  // the chain's slot in the CTA, the warp's rank within the chain
    const int chain = A.chain_lo + blockIdx.x * ((blockDim.x >> 5) / NW) + gib;
    if (chain >= A.chain_hi) return;  // all warps of a chain leave together
    auto chain_sync = [&]() {
        if (NW > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + gib), "r"(32 * NW) : "memory");
        else __syncwarp();
    };
    GroupSmem sm;
    sm.carve(smem_raw + (size_t)gib * dyn_group_smem(A.Q, A.P), A.Q, A.P);
    QConst *qc = sm.qc;
    SuffStat *ss = sm.ss;
    double *th = sm.th;
    const int Q = A.Q, P = A.P, E = A.E;
    int *claim = claim_all + (size_t)chain * E;
    int *claimd = claimd_all ? claimd_all + (size_t)chain * E : nullptr;

    if (wic == 0) {
        for (int i = lane; i < P; i += W) th[i] = A.theta[(size_t)chain * P + i];
        for (int q = lane; q < Q; q += W) sm.qi[q] = A.qinfo[q];
    }
    for (int i = wic * W + lane; i < E; i += RB) { claim[i] = 0; if (claimd) claimd[i] = 0; }
    chain_sync();
    if (wic == 0)
        for (int q = lane; q < Q; q += W) {
            const int o = __ldg(&A.q_poff[q]), dist = sm.qi[q].dist;
            qconst_set(qc[q], dist, th[o], (dist == D_EXP) ? 0.0 : th[o + 1]);
        }
    chain_sync();
    Chain c;
    c.bind(A, chain, sm.qi, qc);

    LaneParams LP;
    LP.seed = A.seed; LP.gchain = (unsigned long long)(A.chain_offset + chain); LP.mode = A.mode; LP.flags = A.flags;
    LP.scan = A.scan;
    long long n_ev = 0, n_evals = 0, n_stuck = 0, n_pevals = 0, n_steps = 0, n_reject = 0, n_redo = 0;
    const int4 *order4 = reinterpret_cast<const int4 *>(A.order);
    int round = 0;
    const int i_in = wic * W + lane;  // this lane's place in a round

    for (int sw = 0; sw < A.n_sweeps; ++sw) {
        LP.gsw = A.sweep_base + (unsigned int)sw;
        if (sw == 0 || (A.flags & FLAG_TMAX_PER_SWEEP)) {  // Tmax = 2 max d   (qnet.pyx:672)
            double m = 0.0;
            for (int i = lane; i < E; i += W) m = fmax(m, c.d[i]);
            LP.tmax = 2.0 * group_max<W>(m, FULL);
        }
        const int n_order = (A.flags & FLAG_PARAMS_ONLY) ? 0 : A.n_order;
        for (int s = 0; s < n_order;) {
            ++round;
            const int round_base = round * RB;
            const int idx = s + i_in;
            const bool active = idx < n_order;
            EventCtx ev;
            LaneUpdate up;
            up.changed = 0; up.counted = 0; up.rejected = 0; up.stuck = 0; up.evals = 0;
            up.R.reset(); up.Wt.reset();
            if (active) {
                const int4 r4 = __ldg(&order4[(A.flags & FLAG_REVERSE) ? n_order - 1 - idx : idx]);
                const DynRec rec = {r4.x, r4.y, r4.z, r4.w};
                event_load(c, rec, ev);
                lane_update<KMAX, MHK, !MHK, PS>(c, ev, LP, up);
            }
            __syncwarp();
            const int stamp = round_base + RB - 1 - i_in;  // larger = earlier in the round
            if (active && up.changed) {  // publish what this lane will write
                claim_range(claim, up.Wt.lo0, up.Wt.hi0, stamp);
                claim_range(claim, up.Wt.lo1, up.Wt.hi1, stamp);
                if (claimd) claim_range(claimd, up.Wt.dlo, up.Wt.dhi, stamp);
            }
            chain_sync();
            bool conflict = false;
            if (active && i_in > 0) {
                conflict = claimed_before_thr(claim, up.R.lo0, up.R.hi0, stamp) ||
                           claimed_before_thr(claim, up.R.lo1, up.R.hi1, stamp) ||
                           (claimd && claimed_before_thr(claimd, up.R.dlo, up.R.dhi, stamp));
            }
            const unsigned cm = __ballot_sync(FULL, conflict);
            const int nact = (n_order - s < RB) ? n_order - s : RB;
            int fc = nact;
            if (NW > 1) {
                if (lane == 0) s_ballot[warp] = cm;
                chain_sync();
#pragma unroll
                for (int w = NW - 1; w >= 0; --w) {
                    const unsigned b = s_ballot[gib * NW + w];
                    if (b) fc = w * W + __ffs(b) - 1;
                }
            } else if (cm) fc = __ffs(cm) - 1;
            if (active && i_in < fc) {
                n_ev += up.counted; n_evals += up.evals; n_reject += up.rejected; n_steps += up.R.steps;
                if (LP.mode != MODE_MH) n_stuck += up.stuck;
                if (up.changed) dyn_commit<1, KMAX, PS>(c, ev, up.x_new, 0, 1u << lane);
            }
            n_redo += (i_in == 0) ? (nact - fc) : 0;
            chain_sync();  // the commits are in place (and s_ballot is free) before anybody reads the next round
            s += fc;
        }
        const bool trace = (A.flags & FLAG_TRACE) != 0;
        if (A.update != UPD_NONE || trace) {
            if (wic == 0) {
                double logp = 0.0;
                dyn_suffstats<W>(c, Q, ss, lane, FULL, trace, &logp);
                const size_t recno = (size_t)(A.trace_base + sw);
                if (trace) {
                    for (int q = lane; q < Q; q += W)
                        A.tr_wait[(recno * A.n_chains + chain) * Q + q] = (ss[q].nall > 0.0) ? ss[q].wait / ss[q].nall : 0.0;
                    if (lane == 0) A.tr_logp[recno * A.n_chains + chain] = logp;
                }
                if (A.update != UPD_NONE) {
                    for (int q = lane; q < Q; q += W) {
                        const int o = __ldg(&A.q_poff[q]), dist = sm.qi[q].dist;
                        double p0 = th[o], p1 = (dist == D_EXP) ? 0.0 : th[o + 1];
                        if (A.update == UPD_BAYES) {
                            Rng prng;
                            prng.init(A.seed, LP.gchain, LP.gsw, (unsigned int)q, TAG_PARAM);
                            sample_params(dist, ss[q], prng, p0, p1, &n_pevals);
                        } else {
                            estimate_params(dist, ss[q], p0, p1);
                        }
                        th[o] = p0;
                        if (dist != D_EXP) th[o + 1] = p1;
                        qconst_set(qc[q], dist, p0, p1);
                    }
                }
                __syncwarp();
                if (trace)
                    for (int i = lane; i < P; i += W) A.tr_theta[(recno * A.n_chains + chain) * P + i] = th[i];
            }
            chain_sync();  // the new densities are in shared memory before the next sweep
        }
    }
    if (wic == 0)
        for (int i = lane; i < P; i += W) A.theta[(size_t)chain * P + i] = th[i];
    // per-lane counters -> the chain's totals
    double tot[7] = {(double)n_ev, (double)n_evals, (double)n_pevals, (double)n_stuck, (double)n_reject, (double)n_steps,
                     (double)n_redo};
#pragma unroll
    for (int i = 0; i < 7; ++i) tot[i] = group_sum<W>(tot[i], FULL);
    if (lane == 0) {
        unsigned long long *st = reinterpret_cast<unsigned long long *>(A.stats + (size_t)chain * BQ_NSTAT);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            if (NW > 1) atomicAdd(&st[i], (unsigned long long)(long long)tot[i]);
            else st[i] += (unsigned long long)(long long)tot[i];
        }
    }
}

// ---- parity / reporting kernels ---------------------------------------------------------------

// out[j] = inner_dfun(x[j]) - lik0 for the event of `rec` on one chain (qnet.pyx:1340-1357); one warp.
// what = 1: the MH proposal log-density h(x[j]) instead (queues.pyx:1239-1256)
template <int KMAX>
__global__ void logcond_dynamic_kernel(const DynArgs A, int chain, DynRec rec, const double *x, int n, double *out,
                                       int what) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GroupSmem sm;
    sm.carve(smem_raw, A.Q, A.P);
    const double *th = A.theta + (size_t)chain * A.P;
    for (int q = threadIdx.x; q < A.Q; q += blockDim.x) {
        sm.qi[q] = A.qinfo[q];
        const int o = A.q_poff[q], dist = A.qinfo[q].dist;
        qconst_set(sm.qc[q], dist, th[o], dist == D_EXP ? 0.0 : th[o + 1]);
    }
    __syncthreads();
    Chain c;
    c.bind(A, chain, sm.qi, sm.qc);
    EventCtx ev;
    event_load(c, rec, ev);
    Touch tc;
    tc.reset();
    if (what == 1) {
        MhProp pr;
        mh_setup(c, ev, pr);
        for (int j = threadIdx.x; j < n; j += blockDim.x) out[j] = pr.eval(x[j]);
        return;
    }
    const double g0 = dyn_logcond<KMAX>(c, ev, ev.x0, tc);
    for (int j = threadIdx.x; j < n; j += blockDim.x) out[j] = dyn_logcond<KMAX>(c, ev, x[j], tc) - g0;
}

// Qnet.log_prob (qnet.pyx:1032) and the sufficient statistics, one warp per chain
static __global__ void __launch_bounds__(128) report_dynamic_kernel(const DynArgs A, double *logp_out, double *ss_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int gib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + gib;
    if (chain >= A.n_chains) return;
    GroupSmem sm;
    sm.carve(smem_raw + (size_t)gib * dyn_group_smem(A.Q, A.P), A.Q, A.P);
    const double *th = A.theta + (size_t)chain * A.P;
    for (int q = lane; q < A.Q; q += 32) {
        sm.qi[q] = A.qinfo[q];
        const int o = A.q_poff[q], dist = A.qinfo[q].dist;
        qconst_set(sm.qc[q], dist, th[o], dist == D_EXP ? 0.0 : th[o + 1]);
    }
    __syncwarp();
    Chain c;
    c.bind(A, chain, sm.qi, sm.qc);
    double logp = 0.0;
    dyn_suffstats<32>(c, A.Q, sm.ss, lane, 0xffffffffu, true, &logp);
    if (lane == 0 && logp_out) logp_out[chain] = logp;
    if (ss_out)
        for (int i = lane; i < A.Q * 8; i += 32)
            ss_out[(size_t)chain * A.Q * 8 + i] = reinterpret_cast<double *>(sm.ss)[i];
}

// Position-indexed times and free-time vectors of every queue from the departure times by event and the current
// queue order; one thread per (chain, queue), sequential along the queue (queues.pyx:654-680).  Used when a
// static-order handle enters the dynamic-order kernels (MH sweeps) after slice sweeps have moved the departures.
template <int KMAX>
__global__ void rebuild_positions_kernel(const DynArgs A, const int *ev_pt) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)A.n_chains * A.Q) return;
    const int chain = (int)(t / A.Q), q = (int)(t % A.Q);
    const QInfo qi = A.qinfo[q];
    const double *d = A.state + (size_t)chain * A.E;
    const int *ord = A.ord + (size_t)chain * A.E;
    double *rec = A.rec + (size_t)chain * A.E * A.S;
    FreeVec<KMAX> M;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) M.v[j] = (j < qi.k) ? 0.0 : INFINITY;
    for (int p = qi.lo; p < qi.hi; ++p) {
        const int i = ord[p], pt = ev_pt[i];
        const double di = d[i];
        const double ai = (pt >= 0) ? d[pt] : 0.0;
        double *r = rec + (size_t)p * A.S;
        r[BQ_REC_DQ] = di;
        r[BQ_REC_AT] = ai;
        if (qi.type == Q_GGK) {
            M.store(r + BQ_REC_F, qi.k);
            if (qi.k > 1) A.flag[(size_t)chain * A.FW + p] = (ai < M.v[1]) ? 1 : 0;
            if (qi.dist != D_EXP) r[rec_ls_slot(A.S)] = log_d(di - dmax(ai, M.v[0]));
            M.replace_min(di);
        } else if (qi.type == Q_DELAY && qi.dist != D_EXP) {
            r[rec_ls_slot(A.S)] = log_d(di - ai);
        }
    }
}
#endif  // __CUDACC__

}  // namespace bq
