// Host-side construction of one chain's derived arrays for the dynamic-order kernels (bq_dynamic.cuh): queue
// orders, position-indexed times and free-time vectors, and the Event fields the reference carries (a, s, proc,
// by-queue links).  Used by the library at create / set_state / get_state time and by the test harness under
// tests/emul.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace bq {

struct HostNet {
    int E, Q;
    const int *q_type, *q_k, *q_off, *q_foff;  // [Q], [Q], [Q+1], [Q]: discipline, servers, position ranges, F offsets
    const int *q_dist;                         // [Q] service family (0 = exponential: no log-service cache)
    const int *prev_byt;                       // [E]
};

// one chain's position-indexed arrays (at, dq: all queues; F: FIFO queues; dt, sq, dord, dpos: PS queues, may be null)
struct HostChain {
    int *ord, *pos;
    double *at, *dq, *ls, *F;
    double *dt, *sq;
    int *dord, *dpos;
};

static inline void fv_replace_min(double *v, int k, double x) {
    v[0] = x;
    for (int j = 0; j + 1 < k; ++j)
        if (v[j] > v[j + 1]) std::swap(v[j], v[j + 1]);
}

// s = integral over [a, d] of dt / N(t) on a PS queue's sorted arrival / departure times (n of each;
// queues.pyx:1049-1070)
static double host_ps_service(const double *at, const double *dt, int n, double a, double d) {
    int ia = (int)(std::upper_bound(at, at + n, a) - at), id = (int)(std::upper_bound(dt, dt + n, a) - dt);
    int cnt = ia - id;
    if (d - a < 1e-15) return (d - a) / (1.0 + cnt);
    double t0 = a, s = 0.0;
    for (;;) {
        const double ta = ia < n ? at[ia] : INFINITY, td = id < n ? dt[id] : INFINITY;
        const bool arr = ta <= td;
        const double t1 = arr ? ta : td;
        if (!(t1 < d)) { if (cnt > 0) s += (d - t0) / cnt; break; }
        if (t1 > t0 && cnt > 0) s += (t1 - t0) / cnt;
        if (arr) { ++cnt; ++ia; } else { --cnt; ++id; }
        t0 = t1;
    }
    return s;
}

// c.ord holds the events of each queue (ranges h->q_off) in queue order.  With `resort`, FIFO and PS queues are
// first re-ranked by (arrival, departure) -- queue 0, whose arrivals are all 0, thereby by departure
// (arrivals.pyx:256).  Fills every other array of the chain.
static void host_rebuild(const HostNet *h, const double *d, const HostChain &c, bool resort) {
    auto arr = [&](int i) { int pt = h->prev_byt[i]; return pt >= 0 ? d[pt] : 0.0; };
    std::vector<double> M;
    for (int q = 0; q < h->Q; ++q) {
        const int lo = h->q_off[q], hi = h->q_off[q + 1], type = h->q_type[q];
        if (type == 3 /* Q_GG1R */ && c.dt) {
            // the caller's list is the departure order (queues.pyx:796; equal departures keep the caller's order);
            // positions are by arrival, like everyone's
            for (int p = lo; p < hi; ++p) c.dord[p] = c.ord[p];
            std::stable_sort(c.dord + lo, c.dord + hi, [&](int x, int y) { return d[x] < d[y]; });
        }
        if ((resort && type != 1 /* Q_DELAY */) || type == 3)
            std::stable_sort(c.ord + lo, c.ord + hi, [&](int x, int y) {
                const double ax = arr(x), ay = arr(y);
                if (ax != ay) return ax < ay;
                return d[x] < d[y];
            });
        for (int p = lo; p < hi; ++p) {
            const int i = c.ord[p];
            c.pos[i] = p;
            c.at[p] = arr(i);
            c.dq[p] = d[i];
        }
        const bool logs = h->q_dist[q] != 0;
        if (type == 0 /* Q_GGK */) {  // free-time vectors (queues.pyx:654-680 as a sorted multiset), log services
            const int k = h->q_k[q];
            M.assign(k, 0.0);
            for (int p = lo; p < hi; ++p) {
                memcpy(c.F + h->q_foff[q] + (size_t)(p - lo) * k, M.data(), k * sizeof(double));
                if (logs) c.ls[p] = log(c.dq[p] - std::max(c.at[p], M[0]));
                fv_replace_min(M.data(), k, c.dq[p]);
            }
        }
        if (type == 1 /* Q_DELAY */ && logs)
            for (int p = lo; p < hi; ++p) c.ls[p] = log(c.dq[p] - c.at[p]);
        if (type == 2 /* Q_PS */ && c.dt) {
            for (int p = lo; p < hi; ++p) c.dord[p] = c.ord[p];
            std::stable_sort(c.dord + lo, c.dord + hi, [&](int x, int y) { return d[x] < d[y]; });
            for (int p = lo; p < hi; ++p) { c.dt[p] = d[c.dord[p]]; c.dpos[c.dord[p]] = p; }
            for (int p = lo; p < hi; ++p) c.sq[p] = host_ps_service(c.at + lo, c.dt + lo, hi - lo, c.at[p], c.dq[p]);
        }
        if (type == 3 /* Q_GG1R */ && c.dt)  // departures in order, with the arrival time of the event in each slot
            for (int p = lo; p < hi; ++p) { c.dt[p] = d[c.dord[p]]; c.dpos[c.dord[p]] = p; c.sq[p] = arr(c.dord[p]); }
    }
}

// a, s, proc and the by-queue links of one chain from d and its queue order (what Event carries in the
// reference: arrivals.pxd:11-37; proc = first arg-min of the processor-indexed vector, queues.pyx:1518-1527).
// sq: the PS service times by position; ord0: the initial order, whose links PS queues keep.
static void host_derive(const HostNet *h, const double *d, const int *ord, double *a, double *s, int32_t *proc,
                        int32_t *prev_byq, int32_t *next_byq, const double *sq = nullptr, const int *ord0 = nullptr) {
    const int E = h->E;
    std::vector<double> av(E);
    for (int e = 0; e < E; ++e) { int pt = h->prev_byt[e]; av[e] = pt >= 0 ? d[pt] : 0.0; }
    if (a) memcpy(a, av.data(), E * sizeof(double));
    std::vector<double> fr;
    for (int q = 0; q < h->Q; ++q) {
        const int lo = h->q_off[q], hi = h->q_off[q + 1], k = h->q_k[q], t = h->q_type[q];
        fr.assign(std::max(k, 1), 0.0);
        // the reference never relinks a PS queue (queues.pyx:1149-1183 touch only a, d, s): report the initial links
        const int *lk = (t == 2 /* Q_PS */ && ord0) ? ord0 : ord;
        std::vector<int> bydep;
        if (t == 3 /* Q_GG1R */) {  // its list is the departure order (queues.pyx:796); s = d - max(a, previous departure)
            bydep.assign(ord + lo, ord + hi);
            std::stable_sort(bydep.begin(), bydep.end(), [&](int x, int y) { return d[x] < d[y]; });
            for (int p = lo; p < hi; ++p) {
                const int i = bydep[p - lo];
                if (prev_byq) prev_byq[i] = p > lo ? bydep[p - lo - 1] : -1;
                if (next_byq) next_byq[i] = p + 1 < hi ? bydep[p - lo + 1] : -1;
                if (proc) proc[i] = 0;
                if (s) s[i] = d[i] - std::max(av[i], p > lo ? d[bydep[p - lo - 1]] : 0.0);
            }
            continue;
        }
        for (int p = lo; p < hi; ++p) {
            if (prev_byq) prev_byq[lk[p]] = p > lo ? lk[p - 1] : -1;
            if (next_byq) next_byq[lk[p]] = p + 1 < hi ? lk[p + 1] : -1;
        }
        for (int p = lo; p < hi; ++p) {
            const int i = ord[p];
            int pr = 0;
            double sv = d[i] - av[i];
            if (t == 0 /* Q_GGK */) {
                for (int j = 1; j < k; ++j)
                    if (fr[j] < fr[pr]) pr = j;
                sv = d[i] - std::max(av[i], fr[pr]);
                fr[pr] = d[i];
            }
            if (t == 2 /* Q_PS */ && sq) sv = sq[p];
            if (proc) proc[i] = pr;
            if (s) s[i] = sv;
        }
    }
}

}  // namespace bq
