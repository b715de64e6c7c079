// Host-side construction of one chain's derived arrays for the dynamic-order kernels (bq_dynamic.cuh): queue
// orders, free-time vectors, and the Event fields the reference carries (a, s, proc, by-queue links).  Used by
// the library at create / set_state / get_state time and by the test harness under tests/emul.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace bq {

struct HostNet {
    int E, Q;
    const int *q_type, *q_k, *q_off;  // [Q], [Q], [Q+1]
    const int *ev_f, *prev_byt;       // [E]
};

static inline void fv_replace_min(double *v, int k, double x) {
    v[0] = x;
    for (int j = 0; j + 1 < k; ++j)
        if (v[j] > v[j + 1]) std::swap(v[j], v[j + 1]);
}

// ord: the events of each queue (ranges h->q_off) in queue order.  With `resort`, FIFO and PS queues are first
// re-ranked by (arrival, departure) -- queue 0, whose arrivals are all 0, thereby by departure
// (arrivals.pyx:256).  Fills pos and the free-time vectors F (queues.pyx:654-680 as a sorted multiset).
// PS queues additionally get their sorted arrival / departure times (at, dt), the departure order (dord, dpos)
// and the service times from the definition (queues.pyx:1049-1070); those pointers may be null otherwise.
struct HostPs { double *at, *dt, *svc; int *dord, *dpos; };

// s = integral over [a, d] of dt / N(t) on the queue's sorted arrival / departure times (n of each)
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

static void host_rebuild(const HostNet *h, const double *d, int *ord, int *pos, double *F, bool resort,
                         const HostPs *ps = nullptr) {
    auto arr = [&](int i) { int pt = h->prev_byt[i]; return pt >= 0 ? d[pt] : 0.0; };
    std::vector<double> M;
    for (int q = 0; q < h->Q; ++q) {
        const int lo = h->q_off[q], hi = h->q_off[q + 1];
        if (resort && h->q_type[q] != 1 /* Q_DELAY */)
            std::stable_sort(ord + lo, ord + hi, [&](int x, int y) {
                const double ax = arr(x), ay = arr(y);
                if (ax != ay) return ax < ay;
                return d[x] < d[y];
            });
        for (int p = lo; p < hi; ++p) pos[ord[p]] = p;
        if (h->q_type[q] == 0 /* Q_GGK */) {
            const int k = h->q_k[q];
            M.assign(k, 0.0);
            for (int p = lo; p < hi; ++p) {
                const int i = ord[p];
                memcpy(F + h->ev_f[i], M.data(), k * sizeof(double));
                fv_replace_min(M.data(), k, d[i]);
            }
        }
        if (h->q_type[q] == 2 /* Q_PS */ && ps) {
            for (int p = lo; p < hi; ++p) { ps->at[p] = arr(ord[p]); ps->dord[p] = ord[p]; }
            std::stable_sort(ps->dord + lo, ps->dord + hi, [&](int x, int y) { return d[x] < d[y]; });
            for (int p = lo; p < hi; ++p) { ps->dt[p] = d[ps->dord[p]]; ps->dpos[ps->dord[p]] = p; }
            for (int p = lo; p < hi; ++p)
                ps->svc[ord[p]] = host_ps_service(ps->at + lo, ps->dt + lo, hi - lo, ps->at[p], d[ord[p]]);
        }
    }
}

// a, s, proc and the by-queue links of one chain from d and its queue order (what Event carries in the
// reference: arrivals.pxd:11-37; proc = first arg-min of the processor-indexed vector, queues.pyx:1518-1527)
static void host_derive(const HostNet *h, const double *d, const int *ord, double *a, double *s, int32_t *proc,
                        int32_t *prev_byq, int32_t *next_byq, const double *svc = nullptr,
                        const int *ord0 = nullptr) {
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
            if (t == 2 /* Q_PS */ && svc) sv = svc[i];
            if (proc) proc[i] = pr;
            if (s) s[i] = sv;
        }
    }
}


}  // namespace bq
