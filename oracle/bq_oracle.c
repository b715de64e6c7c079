/*
 * bq_oracle.c -- CPU restatement of the reference's hot path (casutton/bayes-qnet).  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under bayes_qnet_b200/ may include, link, import or execute this file; it exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can check the CUDA kernels against an independent
 * statement of the algorithm.  It shares no code with the product.
 *
 * What is restated, at definition level (every evaluation recomputes the affected queues from scratch
 * instead of building the reference's incremental diff lists -- O(events of the queue) per evaluation, which
 * keeps it obviously right and limits it to small cases):
 *   - service-time semantics: QueueGGk any k (queues.pyx:610-680: s = d - max(a, min_p last departure on p),
 *     proc = first argmin), DelayStation (queues.pyx:771: s = d - a), QueuePS (queues.pyx:1049-1070:
 *     s = integral over [a,d] of dt/N(t), N(t) = #{a<=t} - #{d<=t}, cninq.pyx:55-58), QueueGG1R (queues.pyx:788-1027:
 *     one server taking waiting jobs in random order -- the queue's list is ordered by departure, s = d - max(a, d of
 *     the previous departure); the likelihood carries -log N(commencement) per event and the feasibility rule "started
 *     on arrival => arrived to an empty queue", queues.pyx:956-1027);
 *   - the per-event conditional GGkGibbs.inner_dfun(d) - lik0 (qnet.pyx:1340-1357) =
 *     likelihoodDelta(q0, diffListForDeparture(e, d)) + likelihoodDelta(q1, diffListForArrival(e', d))
 *     (queues.pyx:83-95, 470-591, 1149-1183, 1220-1232), including the re-sort of q1 by arrival time
 *     (queues.pyx:514-556), -inf when any new service is negative, and PS's log N terms;
 *   - one slice update, sampling._slice with finite bounds (sampling.pyx:283-298, 344-366), eligibility and
 *     bounds of Qnet.slice_resample (qnet.pyx:672, 686-702), commit = apply_departure (qnet.pyx:2166-2178);
 *   - one MH-within-Gibbs update, Qnet.gibbs_resample_pair / _final with the TWOS proposal (qnet.pyx:337-498;
 *     queues.pyx:394-430, 458-468, 750-756, 1239-1256);
 *   - Qnet.log_prob (qnet.pyx:1032; queues.pyx:97-117), the log-densities (distributions.pyx:59, 129-133,
 *     266-270), the sufficient statistics and the parameter updates (distributions.pyx:65-74, 92-96,
 *     142-160, 276-320; netutils.py:19-42; queues.pyx:1316-1336).
 *
 * Pinned against the reference itself: tests/golden/ *.npz hold inner_dfun grids, log_prob values, states
 * and lpdf known-answer values produced by the reference's own Cython code (built by oracle/build_ref.py,
 * dumped by oracle/make_golden.py); tests/test_oracle_golden.py checks this file against them.
 *
 * Random numbers: the reference draws from a global MT19937 (randomkit.c); stream parity is not a goal
 * (SURVEY.md section 8a).  So that whole sweeps can be compared with the kernels, this file restates the
 * product's documented stream convention (Philox4x32-7, Salmon et al. 2011; counter = (block, id, sweep,
 * chain lo), key = (seed lo ^ tag*0x9E3779B9, seed hi ^ chain hi); two 53-bit uniforms per block; in a finite-bound
 * slice update block 0 = level + first candidate, every later block = four candidates from one 32-bit word each) --
 * the convention is part of the C-ABI contract (include/bq_capi.h: bq_rng_uniforms, bq_rng_slice_uniforms).
 *
 * Documented deviation shared with the product: when the slice bracket collapses below 1e-10 the reference
 * returns L (sampling.pyx:360-362), which can lie outside the support; both implementations keep x0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { Q_GGK = 0, Q_DELAY = 1, Q_PS = 2, Q_GG1R = 3 };
enum { D_EXP = 0, D_GAMMA = 1, D_LN = 2 };

typedef struct {
    int E, Q;
    const int *q_type, *q_k, *q_dist, *q_poff; /* [Q] */
    const double *theta;                       /* parameter vector (Qnet.parameters order) */
    const int *qid, *prev_byt, *next_byt;      /* [E] static */
    const uint8_t *obs_a, *obs_d;              /* [E] */
    double *a, *d, *s;                         /* [E] chain state */
    int *proc, *prev_byq, *next_byq;           /* [E] chain state */
    /* mixture templates (MixtureTemplate, queues.pyx:1373-1509): NULL aux = none.  Otherwise every event's service
     * density is the one of slot q_slot[qid] + aux[e] (serviceLikelihood, queues.pyx:1455-1457) */
    const int *q_slot, *slot_dist, *slot_poff; /* [Q+1], [NS], [NS] */
    const uint8_t *aux;                        /* [E] component of each event */
} bqo_state;

/* ---------------------------------------------------------------------------------------------- */
/* log-densities (distributions.pyx)                                                               */
/* ---------------------------------------------------------------------------------------------- */
double bqo_lpdf(int dist, double p0, double p1, double x) {
    if (dist == D_EXP) return -log(p0) - x / p0;
    if (dist == D_GAMMA) {
        if (p0 == 1.0) return -x / p1 - p0 * log(p1);
        return (p0 - 1.0) * log(x) - (x / p1) - lgamma(p0) - p0 * log(p1);
    }
    if (x == 0.0) return -INFINITY;
    {
        double lx = log(x);
        return -0.918938533204673 - log(p1) - lx - ((lx - p0) * (lx - p0)) / (2.0 * p1 * p1);
    }
}

static double q_lpdf(const bqo_state *st, int q, int e, double x) {
    int o = st->q_poff[q], dist = st->q_dist[q];
    if (st->aux) {
        int sl = st->q_slot[q] + st->aux[e];
        o = st->slot_poff[sl]; dist = st->slot_dist[sl];
    }
    return bqo_lpdf(dist, st->theta[o], dist == D_EXP ? 0.0 : st->theta[o + 1], x);
}

/* ---------------------------------------------------------------------------------------------- */
/* queue order helpers                                                                              */
/* ---------------------------------------------------------------------------------------------- */
static int queue_list(const bqo_state *st, int q, int *out) {
    int head = -1, n = 0, e;
    for (e = 0; e < st->E; ++e)
        if (st->qid[e] == q && st->prev_byq[e] < 0) { head = e; break; }
    for (e = head; e >= 0; e = st->next_byq[e]) out[n++] = e;
    return n;
}

/* service times (and processors) of the events of one queue listed in `lst` in queue order, for arrival
 * and departure times given by a[] and d[] (indexed by event) */
static void queue_services(const bqo_state *st, int q, const int *lst, int n, const double *a, const double *d,
                           double *s_out, int *proc_out) {
    int type = st->q_type[q], k = st->q_k[q], i, j, p;
    if (type == Q_DELAY) {
        for (i = 0; i < n; ++i) { s_out[i] = d[lst[i]] - a[lst[i]]; proc_out[i] = 0; }
    } else if (type == Q_GGK) {
        double *free_t = (double *)calloc((size_t)k, sizeof(double));
        for (i = 0; i < n; ++i) {
            int e = lst[i];
            double m = free_t[0];
            p = 0;
            for (j = 1; j < k; ++j)
                if (free_t[j] < m) { m = free_t[j]; p = j; }
            if (i == 0) { m = 0.0; p = 0; }
            s_out[i] = d[e] - (a[e] > m ? a[e] : m);
            proc_out[i] = p;
            free_t[p] = d[e];
        }
        free(free_t);
    } else if (type == Q_GG1R) { /* recompute_service, queues.pyx:808-816: the list is in departure order */
        for (i = 0; i < n; ++i) {
            int e = lst[i];
            double dp = i > 0 ? d[lst[i - 1]] : 0.0;
            s_out[i] = (i > 0) ? d[e] - (a[e] > dp ? a[e] : dp) : d[e] - a[e];
            proc_out[i] = 0;
        }
    } else { /* PS: piecewise integration between consecutive breakpoints inside [a_e, d_e] */
        double *pts = (double *)malloc((size_t)(2 * n + 2) * sizeof(double));
        for (i = 0; i < n; ++i) {
            int e = lst[i], np = 0;
            double ae = a[e], de = d[e], acc = 0.0;
            proc_out[i] = 0;
            if (de - ae < 1e-15) {
                int cnt = 0; /* N(a_e) */
                for (j = 0; j < n; ++j) { if (a[lst[j]] <= ae) cnt++; if (d[lst[j]] <= ae) cnt--; }
                s_out[i] = (de - ae) / (1.0 + cnt);
                continue;
            }
            pts[np++] = ae;
            for (j = 0; j < n; ++j) {
                double x = a[lst[j]];
                if (x > ae && x < de) pts[np++] = x;
                x = d[lst[j]];
                if (x > ae && x < de) pts[np++] = x;
            }
            pts[np++] = de;
            /* insertion sort (small) */
            for (j = 1; j < np; ++j) {
                double v = pts[j]; int t = j - 1;
                while (t >= 0 && pts[t] > v) { pts[t + 1] = pts[t]; --t; }
                pts[t + 1] = v;
            }
            for (j = 0; j + 1 < np; ++j) {
                double t0 = pts[j], t1 = pts[j + 1], mid = 0.5 * (t0 + t1);
                int cnt = 0, t;
                if (t1 <= t0) continue;
                for (t = 0; t < n; ++t) if (a[lst[t]] <= mid && d[lst[t]] > mid) cnt++;
                if (cnt > 0) acc += (t1 - t0) / cnt;
            }
            s_out[i] = acc;
        }
        free(pts);
    }
}

/* N(t) = #{a <= t} - #{d <= t} over the events of a queue (cninq.pyx:55-58) */
static int ps_N(const int *lst, int n, const double *a, const double *d, double t) {
    int i, c = 0;
    for (i = 0; i < n; ++i) { if (a[lst[i]] <= t) c++; if (d[lst[i]] <= t) c--; }
    return c;
}

/* recompute a (from by-task links), s and proc of every event from d and the current queue order */
void bqo_recompute(bqo_state *st) {
    int e, q, i;
    int *lst = (int *)malloc((size_t)st->E * sizeof(int));
    double *s = (double *)malloc((size_t)st->E * sizeof(double));
    int *pr = (int *)malloc((size_t)st->E * sizeof(int));
    for (e = 0; e < st->E; ++e) st->a[e] = st->prev_byt[e] >= 0 ? st->d[st->prev_byt[e]] : 0.0;
    for (q = 0; q < st->Q; ++q) {
        int n = queue_list(st, q, lst);
        queue_services(st, q, lst, n, st->a, st->d, s, pr);
        for (i = 0; i < n; ++i) { st->s[lst[i]] = s[i]; st->proc[lst[i]] = pr[i]; }
    }
    free(lst); free(s); free(pr);
}

/* Qnet.log_prob (qnet.pyx:1032): sum of service log-densities, -inf if any s < 0 */
double bqo_log_prob(const bqo_state *st) {
    double r = 0.0;
    int e, q;
    for (e = 0; e < st->E; ++e) {
        if (st->s[e] < 0.0) return -INFINITY;
        r += q_lpdf(st, st->qid[e], e, st->s[e]);
    }
    /* QueueGG1R.allEventLik (queues.pyx:1009-1024): - log N(commencement) per event with s != 0; an event that
     * started on arrival (wait < 1e-10) must have found the queue empty */
    for (q = 0; q < st->Q; ++q) {
        int *lst, n, i;
        if (st->q_type[q] != Q_GG1R) continue;
        lst = (int *)malloc((size_t)st->E * sizeof(int));
        n = queue_list(st, q, lst);
        for (i = 0; i < n; ++i) {
            int ev = lst[i];
            double w = st->d[ev] - st->a[ev] - st->s[ev], c;
            if (w < 0.0) w = 0.0;
            c = st->d[ev] - st->s[ev];
            if (st->a[ev] > c) c = st->a[ev];
            if (st->s[ev] != 0.0) r -= log((double)ps_N(lst, n, st->a, st->d, c));
            if (w < 1e-10 && ps_N(lst, n, st->a, st->d, st->a[ev]) > 1) { free(lst); return -INFINITY; }
        }
        free(lst);
    }
    return r;
}

/* ---------------------------------------------------------------------------------------------- */
/* the conditional: inner_dfun(x) - lik0                                                            */
/* ---------------------------------------------------------------------------------------------- */
/* order of q1 after event e1's arrival becomes x: remove e1, re-insert before the first event whose
 * arrival is > x (stable sort by arrival, queues.pyx:537) */
static int resorted_list(const bqo_state *st, int q1, int e1, double x, const double *a, int *out) {
    int *cur = (int *)malloc((size_t)st->E * sizeof(int));
    int n = queue_list(st, q1, cur), i, m = 0, placed = 0;
    for (i = 0; i < n; ++i) {
        if (cur[i] == e1) continue;
        if (!placed && a[cur[i]] > x) { out[m++] = e1; placed = 1; }
        out[m++] = cur[i];
    }
    if (!placed) out[m++] = e1;
    free(cur);
    return m;
}

static double queue_delta(const bqo_state *st, int q, const int *lst_new, int n, const double *a_new,
                          const double *d_new) {
    /* sum over the queue of lpdf(s_new) - lpdf(s_old); -inf if any s_new < 0 (queues.pyx:83-95);
     * PS adds log(N_old(d_old)+1) - log(N_new(d_new)+1) per event (queues.pyx:1220-1232) */
    double *s = (double *)malloc((size_t)n * sizeof(double));
    int *pr = (int *)malloc((size_t)n * sizeof(int));
    double delta = 0.0;
    int i, bad = 0;
    queue_services(st, q, lst_new, n, a_new, d_new, s, pr);
    for (i = 0; i < n; ++i)
        if (s[i] < 0.0) bad = 1;
    if (!bad) {
        for (i = 0; i < n; ++i) {
            int e = lst_new[i];
            if (s[i] != st->s[e] || st->q_type[q] == Q_PS) delta += q_lpdf(st, q, e, s[i]) - q_lpdf(st, q, e, st->s[e]);
        }
        if (st->q_type[q] == Q_PS) {
            for (i = 0; i < n; ++i) {
                int e = lst_new[i];
                int n_old = ps_N(lst_new, n, st->a, st->d, st->d[e]) + 1;
                int n_new = ps_N(lst_new, n, a_new, d_new, d_new[e]) + 1;
                delta += log((double)n_old) - log((double)n_new);
            }
        }
    }
    free(s); free(pr);
    return bad ? -INFINITY : delta;
}

/* QueueGG1R: likelihoodDelta(diffListForDeparture(m, x)) (is_dep) or likelihoodDelta(diffListForArrival(m, x))
 * (queues.pyx:859-1006).  Departure move: m takes its new place in the departure order (the two while loops of
 * queues.pyx:871-880); the diff list is the run of the OLD list from the earlier of the old / new predecessor to the
 * later of the old / new successor (to the end of the queue if either successor is missing), re-sorted, services
 * recomputed.  Arrival move: the diff list is m alone (queues.pyx:859-864).  On top of the sum of lpdf(s_new) -
 * lpdf(s_old): for every event of the diff list, and after an arrival move every event whose commencement lies in
 * [min(a_old, a_new), max(a_old, a_new)] (c2e_range), + log N_old(c_old) [s_old != 0] - log N_new(c_new) [s_new != 0],
 * and -inf if such an event starts on arrival (wait < 1e-10) while N_new(a) > 1 -- where N(t) = #{a <= t} - #{d <= t}
 * over the queue (cninq.pyx:55-58) with the moved time in N_new.  If new_list is given the new order is returned. */
static double gg1r_delta(const bqo_state *st, int q, int m, int is_dep, double x, int *new_list) {
    int E = st->E, n, i, i0 = -1, g, g2, lo, hi, bad = 0;
    int *old = (int *)malloc((size_t)E * sizeof(int)), *neu = (int *)malloc((size_t)E * sizeof(int));
    int *pr = (int *)malloc((size_t)E * sizeof(int));
    unsigned char *member = (unsigned char *)calloc((size_t)E, 1);
    double *a2 = (double *)malloc((size_t)E * sizeof(double)), *d2 = (double *)malloc((size_t)E * sizeof(double));
    double *sn = (double *)malloc((size_t)E * sizeof(double));
    double delta = 0.0;
    memcpy(a2, st->a, (size_t)E * sizeof(double));
    memcpy(d2, st->d, (size_t)E * sizeof(double));
    n = queue_list(st, q, old);
    for (i = 0; i < n; ++i) if (old[i] == m) i0 = i;
    if (is_dep) {
        /* others = old without m; m sits in gap g (between others[g-1] and others[g]) and moves to gap g2 */
#define OTH(j) (old[(j) < i0 ? (j) : (j) + 1])
        d2[m] = x;
        g = g2 = i0;
        while (g2 > 0 && !(st->d[OTH(g2 - 1)] <= x)) --g2;
        while (g2 < n - 1 && !(x <= st->d[OTH(g2)])) ++g2;
        for (i = 0; i < n - 1; ++i) neu[i < g2 ? i : i + 1] = OTH(i);
        neu[g2] = m;
#undef OTH
        lo = (g < g2 ? g : g2) - 1;
        if (lo < 0) lo = 0;
        hi = (g == n - 1 || g2 == n - 1) ? n - 1 : (g > g2 ? g : g2) + 1;
        for (i = lo; i <= hi; ++i) member[old[i]] = 1;
    } else {
        double l = st->a[m] < x ? st->a[m] : x, r = st->a[m] < x ? x : st->a[m];
        a2[m] = x;
        memcpy(neu, old, (size_t)n * sizeof(int));
        member[m] = 1;
        for (i = 0; i < n; ++i) { /* c2e_range(l, r): current commencements, both ends included */
            int ev = old[i];
            double c = st->d[ev] - st->s[ev];
            if (st->a[ev] > c) c = st->a[ev];
            if (c >= l && c <= r) member[ev] = 2;
        }
        member[m] = 1;
    }
    queue_services(st, q, neu, n, a2, d2, sn, pr);
    /* Queue.likelihoodDelta over the diff list (members marked 1) */
    for (i = 0; i < n; ++i)
        if (member[neu[i]] == 1 && sn[i] < 0.0) bad = 1;
    if (!bad) {
        for (i = 0; i < n; ++i) {
            int ev = neu[i];
            if (member[ev] == 1 && sn[i] != st->s[ev]) delta += q_lpdf(st, q, ev, sn[i]) - q_lpdf(st, q, ev, st->s[ev]);
        }
        for (i = 0; i < n && !bad; ++i) {
            int ev = neu[i];
            double s_new, w, c_old, c_new;
            if (!member[ev]) continue;
            s_new = (member[ev] == 1) ? sn[i] : st->s[ev]; /* c2e_range members are the current events */
            w = d2[ev] - a2[ev] - s_new;
            if (w < 0.0) w = 0.0;
            if (w < 1e-10 && ps_N(neu, n, a2, d2, a2[ev]) > 1) { bad = 1; break; }
            c_old = st->d[ev] - st->s[ev];
            if (st->a[ev] > c_old) c_old = st->a[ev];
            c_new = d2[ev] - s_new;
            if (a2[ev] > c_new) c_new = a2[ev];
            if (st->s[ev] != 0.0) delta += log((double)ps_N(old, n, st->a, st->d, c_old));
            if (s_new != 0.0) delta -= log((double)ps_N(neu, n, a2, d2, c_new));
        }
    }
    if (new_list) memcpy(new_list, neu, (size_t)n * sizeof(int));
    free(old); free(neu); free(pr); free(member); free(a2); free(d2); free(sn);
    return bad ? -INFINITY : delta;
}

double bqo_logcond1(const bqo_state *st, int e, double x) {
    int E = st->E, q0 = st->qid[e], e1 = st->next_byt[e], n;
    double *a2 = (double *)malloc((size_t)E * sizeof(double));
    double *d2 = (double *)malloc((size_t)E * sizeof(double));
    int *lst = (int *)malloc((size_t)E * sizeof(int));
    double r;
    memcpy(a2, st->a, (size_t)E * sizeof(double));
    memcpy(d2, st->d, (size_t)E * sizeof(double));
    d2[e] = x;
    if (st->q_type[q0] == Q_GG1R) r = gg1r_delta(st, q0, e, 1, x, NULL);
    else {
        n = queue_list(st, q0, lst); /* a departure move never reorders its queue (queues.pyx:470-511) */
        r = queue_delta(st, q0, lst, n, st->a, d2);
    }
    if (e1 >= 0) {
        int q1 = st->qid[e1];
        double r1;
        /* diffListForArrival sees the OLD departure times of q1 and the new arrival of e1 */
        a2[e1] = x;
        if (st->q_type[q1] == Q_GG1R) r1 = gg1r_delta(st, q1, e1, 0, x, NULL);
        else {
            if (st->q_type[q1] == Q_GGK) n = resorted_list(st, q1, e1, x, st->a, lst);
            else n = queue_list(st, q1, lst);
            r1 = queue_delta(st, q1, lst, n, a2, st->d);
        }
        r += r1;
    }
    free(a2); free(d2); free(lst);
    return r;
}

void bqo_logcond(const bqo_state *st, int e, const double *x, int n, double *out) {
    int i;
    for (i = 0; i < n; ++i) out[i] = bqo_logcond1(st, e, x[i]);
}

/* commit d_e := x (apply_departure, qnet.pyx:2166-2178 -> applyDiffList, arrivals.pyx:542-604) */
void bqo_apply(bqo_state *st, int e, double x) {
    int e1 = st->next_byt[e], i, n;
    if (st->q_type[st->qid[e]] == Q_GG1R) { /* the departure order changes with the move (queues.pyx:866-940) */
        int *lst = (int *)malloc((size_t)st->E * sizeof(int));
        (void)gg1r_delta(st, st->qid[e], e, 1, x, lst);
        n = 0;
        for (i = 0; i < st->E; ++i) if (st->qid[i] == st->qid[e]) ++n;
        for (i = 0; i < n; ++i) {
            st->prev_byq[lst[i]] = i > 0 ? lst[i - 1] : -1;
            st->next_byq[lst[i]] = i + 1 < n ? lst[i + 1] : -1;
        }
        free(lst);
    }
    st->d[e] = x;
    if (e1 >= 0) {
        int q1 = st->qid[e1];
        if (st->q_type[q1] == Q_GGK) {
            int *lst = (int *)malloc((size_t)st->E * sizeof(int));
            n = resorted_list(st, q1, e1, x, st->a, lst);
            for (i = 0; i < n; ++i) {
                st->prev_byq[lst[i]] = i > 0 ? lst[i - 1] : -1;
                st->next_byq[lst[i]] = i + 1 < n ? lst[i + 1] : -1;
            }
            free(lst);
        }
        st->a[e1] = x;
    }
    bqo_recompute(st);
}

/* ---------------------------------------------------------------------------------------------- */
/* counter-based RNG: the C-ABI's documented stream convention (see header comment)                 */
/* ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t k0, k1, c1, c2, c3, blk; double spare; int has; uint32_t w[4]; int nw; } bqo_rng;

static void philox_r(uint32_t c[4], uint32_t k0, uint32_t k1, int rounds) {
    int r;
    for (r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) { philox_r(c, k0, k1, 7); } /* Philox4x32-7 */
/* one raw block with a chosen round count: known-answer tests against the Random123 vectors */
void bqo_philox_block(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
    uint32_t c[4];
    memcpy(c, ctr, sizeof(c));
    philox_r(c, key[0], key[1], rounds);
    memcpy(out, c, sizeof(c));
}
static void rng_init(bqo_rng *r, uint64_t seed, uint64_t chain, uint32_t sweep, uint32_t id, uint32_t tag) {
    r->k0 = (uint32_t)seed ^ (tag * 0x9E3779B9u);
    r->k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(chain >> 32);
    r->c1 = id; r->c2 = sweep; r->c3 = (uint32_t)chain; r->blk = 0; r->has = 0; r->spare = 0.0; r->nw = 0;
}
static double to_u53(uint32_t lo, uint32_t hi) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}
static double rng_uniform(bqo_rng *r) {
    uint32_t c[4];
    if (r->has) { r->has = 0; return r->spare; }
    c[0] = r->blk++; c[1] = r->c1; c[2] = r->c2; c[3] = r->c3;
    philox(c, r->k0, r->k1);
    r->spare = to_u53(c[2], c[3]);
    r->has = 1;
    return to_u53(c[0], c[1]);
}
/* 32-bit uniform in (0,1), four per block: shrink candidates 2, 3, ... of a finite-bound slice update */
static double rng_uniform32(bqo_rng *r) {
    uint32_t v;
    if (r->nw == 0) {
        r->w[0] = r->blk++; r->w[1] = r->c1; r->w[2] = r->c2; r->w[3] = r->c3;
        philox(r->w, r->k0, r->k1);
        r->nw = 4;
    }
    v = r->w[4 - r->nw];
    r->nw--;
    return ((double)v + 0.5) * (1.0 / 4294967296.0);
}
static uint32_t rng_word(bqo_rng *r) {
    uint32_t c[4];
    c[0] = r->blk++; c[1] = r->c1; c[2] = r->c2; c[3] = r->c3;
    philox(c, r->k0, r->k1);
    return c[0];
}
void bqo_rng_uniforms(uint64_t seed, int64_t chain, uint32_t sweep, uint32_t id, uint32_t tag, int n, double *out) {
    bqo_rng r;
    int i;
    rng_init(&r, seed, (uint64_t)chain, sweep, id, tag);
    for (i = 0; i < n; ++i) out[i] = rng_uniform(&r);
}

void bqo_rng_slice_uniforms(uint64_t seed, int64_t chain, uint32_t sweep, uint32_t id, uint32_t tag, int n,
                            double *out) {
    bqo_rng r;
    int i;
    rng_init(&r, seed, (uint64_t)chain, sweep, id, tag);
    for (i = 0; i < n; ++i) out[i] = (i < 2) ? rng_uniform(&r) : rng_uniform32(&r);
}

/* ---------------------------------------------------------------------------------------------- */
/* one sweep of Qnet.slice_resample over the events in `order` (qnet.pyx:686-709)                    */
/* ---------------------------------------------------------------------------------------------- */
int bqo_eligible(const bqo_state *st, int e) {
    int e1 = st->next_byt[e];
    if (st->obs_d[e]) return 0;
    return e1 < 0 || !st->obs_a[e1];
}

/* returns the number of resamples; evals (if not NULL) accumulates conditional evaluations */
long bqo_slice_sweep(bqo_state *st, const int *order, int n_order, uint64_t seed, int64_t chain, uint32_t sweep,
                     double tmax, int tight, long *evals) {
    long ne = 0;
    int i;
    for (i = 0; i < n_order; ++i) {
        int e = order[i], e1 = st->next_byt[e];
        bqo_rng rng;
        double x0, lower, upper, L, R, g0, logy, x1;
        int ncand;
        if (!bqo_eligible(st, e)) continue;
        x0 = st->d[e];
        lower = st->a[e];
        upper = e1 >= 0 ? st->d[e1] : INFINITY; /* determine_upper, qnet.pyx:732-734 */
        if (upper > tmax) upper = tmax;
        rng_init(&rng, seed, (uint64_t)chain, sweep, (uint32_t)e, 1u);
        (void)tight;
        g0 = bqo_logcond1(st, e, x0);
        if (evals) ++*evals;
        logy = g0 - (-log(1.0 - rng_uniform(&rng)));
        ++ne;
        if (!(logy > -INFINITY)) continue;
        L = lower; R = upper; x1 = x0;
        for (ncand = 0;; ++ncand) {
            double g1;
            x1 = L + (R - L) * (ncand == 0 ? rng_uniform(&rng) : rng_uniform32(&rng));
            g1 = bqo_logcond1(st, e, x1);
            if (evals) ++*evals;
            if (g1 >= logy) break;
            if (x1 > x0) R = x1; else L = x1;
            if (R - L < 1e-10) { x1 = x0; break; }
        }
        if (x1 != x0) bqo_apply(st, e, x1);
    }
    return ne;
}

/* ---------------------------------------------------------------------------------------------- */
/* MH-within-Gibbs sweep: Qnet.gibbs_resample (qnet.pyx:337-498)                                     */
/* ---------------------------------------------------------------------------------------------- */
/* departure of the previous event served by the same processor (QueueGGk.prev_by_proc, queues.pyx:619-624) */
static double prev_by_proc_d(const bqo_state *st, int e) {
    int p = st->prev_byq[e];
    while (p >= 0) {
        if (st->proc[p] == st->proc[e]) return st->d[p];
        p = st->prev_byq[p];
    }
    return 0.0;
}

/* TWOS proposal (queues.pyx:1239-1256): departureLik(e) + arrivalLik(e1) on the intersection of their supports
 * (pwfun.pyx:183-202).  GGk: queues.pyx:394-430, 458-468; DelayStation: queues.pyx:750-756.  -inf outside. */
typedef struct { double b0, d1, dp1, L, U; int q0, q1, has1, fifo1, e0, e1; } mhprop;

static void mh_setup(const bqo_state *st, int e, mhprop *pr) {
    int e1 = st->next_byt[e];
    pr->e0 = e; pr->e1 = e1; pr->q0 = st->qid[e]; pr->q1 = -1; pr->has1 = 0; pr->fifo1 = 0; pr->d1 = 0.0; pr->dp1 = 0.0;
    pr->b0 = st->a[e];
    if (st->q_type[pr->q0] == Q_GGK) { double dp = prev_by_proc_d(st, e); if (dp > pr->b0) pr->b0 = dp; }
    pr->L = pr->b0; pr->U = INFINITY;
    if (e1 >= 0) {
        pr->has1 = 1; pr->q1 = st->qid[e1]; pr->d1 = st->d[e1]; pr->U = st->d[e1];
        if (pr->L < 0.0) pr->L = 0.0;
        if (st->q_type[pr->q1] == Q_GGK) { pr->fifo1 = 1; pr->dp1 = prev_by_proc_d(st, e1); }
    }
}
static double mh_eval(const bqo_state *st, const mhprop *pr, double x) {
    double v;
    if (x < pr->L || x > pr->U) return -INFINITY;
    v = q_lpdf(st, pr->q0, pr->e0, x - pr->b0);
    if (pr->has1) {
        double m = x;
        if (pr->fifo1 && pr->dp1 > m) m = pr->dp1;
        v += q_lpdf(st, pr->q1, pr->e1, pr->d1 - m);
    }
    return v;
}

/* one step of sampling._slice on the proposal (sampling.pyx:267-366); keeps x0 on a collapsed bracket */
static double mh_slice(const bqo_state *st, const mhprop *pr, double x0, bqo_rng *rng, long *evals) {
    double gx0 = mh_eval(st, pr, x0), logy, L, R;
    int it;
    ++*evals;
    logy = gx0 - (-log(1.0 - rng_uniform(rng)));
    if (!(logy > -INFINITY)) return x0;
    if (pr->U < INFINITY) { L = pr->L; R = pr->U; }
    else {
        double w = 1.0, u = w * rng_uniform(rng);
        unsigned long long J, K;
        int guard = 0;
        L = x0 - u; R = x0 + (w - u);
        J = (unsigned long long)(rng_word(rng) % 1025u);
        K = 1023ull - J;
        while (J > 0) { if (L <= pr->L) break; ++*evals; if (mh_eval(st, pr, L) <= logy) break; L -= w; --J; w *= 2.0; }
        while (K > 0 && guard < 1100) { if (R >= pr->U) break; ++*evals; if (mh_eval(st, pr, R) <= logy) break; R += w; --K; w *= 2.0; ++guard; }
        if (L < pr->L) L = pr->L;
        if (R > pr->U) R = pr->U;
    }
    for (it = 0; it < 100000; ++it) {
        /* finite bounds: candidate 1 from the level's block (53 bits), the rest 32-bit; stepping out: all 53-bit */
        double x1 = L + (R - L) * ((pr->U < INFINITY && it > 0) ? rng_uniform32(rng) : rng_uniform(rng));
        ++*evals;
        if (mh_eval(st, pr, x1) >= logy) return x1;
        if (x1 > x0) R = x1; else L = x1;
        if (R - L < 1e-10) return x0;
    }
    return x0;
}

/* Streams of (chain, sweep, event): tag 3 and 4 = the two slice steps of call_sampler(thin = 2) (qnet.pyx:1292-1300),
 * tag 5 = accept uniform (number 0; number 1 unused) and the re-initialisation draws (numbers 2..).
 * Returns the number of proposals made (the reference's nevents); *nreject counts rejections. */
long bqo_mh_sweep(bqo_state *st, const int *order, int n_order, uint64_t seed, int64_t chain, uint32_t sweep,
                  int sample_final, long *nreject, long *evals) {
    long ne = 0, dummy = 0;
    int i;
    if (!evals) evals = &dummy;
    for (i = 0; i < n_order; ++i) {
        int e = order[i], e1 = st->next_byt[e], fin = (e1 < 0), step;
        mhprop pr;
        bqo_rng aux, rng;
        double u_acc, h0, x_init, x_new, g_new, ratio;
        if (!bqo_eligible(st, e)) continue;
        if (fin && !sample_final) continue;
        mh_setup(st, e, &pr);
        if (pr.U - pr.L < 1e-10) continue;                       /* qnet.pyx:385, 440 */
        if (!fin && st->d[e1] - st->a[e] < 1e-10) continue;
        rng_init(&aux, seed, (uint64_t)chain, sweep, (uint32_t)e, 5u);
        u_acc = rng_uniform(&aux);
        (void)rng_uniform(&aux);
        h0 = mh_eval(st, &pr, st->d[e]);
        x_init = st->d[e];
        if (!(h0 > -INFINITY)) {                                  /* qnet.pyx:389-397, 444-451 */
            int tries = 0;
            for (; tries < 25; ++tries) {
                double u = rng_uniform(&aux);
                x_init = fin ? pr.b0 + (-log(1.0 - u)) : st->a[e] + (st->d[e1] - st->a[e]) * u;
                if (mh_eval(st, &pr, x_init) > -INFINITY) break;
            }
            if (tries >= 25) continue;
        }
        x_new = x_init;
        for (step = 0; step < 2; ++step) {
            rng_init(&rng, seed, (uint64_t)chain, sweep, (uint32_t)e, step == 0 ? 3u : 4u);
            x_new = mh_slice(st, &pr, x_new, &rng, evals);
        }
        ++ne;
        g_new = bqo_logcond1(st, e, x_new);                       /* product over both diff lists, qnet.pyx:425-431 */
        ratio = exp(h0 - mh_eval(st, &pr, x_new) + g_new);
        if (u_acc < ratio) { if (x_new != st->d[e]) bqo_apply(st, e, x_new); }
        else if (nreject) ++*nreject;
    }
    return ne;
}

/* the proposal log-density on a grid (parity hook for queues.pyx:1239-1256) */
void bqo_mh_proposal(const bqo_state *st, int e, const double *x, int n, double *out) {
    mhprop pr;
    int i;
    mh_setup(st, e, &pr);
    for (i = 0; i < n; ++i) out[i] = mh_eval(st, &pr, x[i]);
}

/* ---------------------------------------------------------------------------------------------- */
/* exact Gibbs transition density: Qnet.gibbs_kernel (qnet.pyx:741-809)                             */
/* ---------------------------------------------------------------------------------------------- */
typedef struct { const bqo_state *st; int e; double lf; } kern_fn;
static double kern_eval(const kern_fn *k, double x) { return exp(bqo_logcond1(k->st, k->e, x) - k->lf); }

/* adaptive Simpson with Richardson correction; the integrand may jump (edges of the support).  The first 5 levels
 * are always split: with Gamma(2) services the integrand is piecewise polynomial with a kink next to the edge of
 * the support where both pieces vanish, and five samples that miss the short piece fit one cubic exactly. */
static double simpson_rec(const kern_fn *k, double a, double b, double fa, double fm, double fb, double whole,
                          double tol, int depth) {
    double m = 0.5 * (a + b), lm = 0.5 * (a + m), rm = 0.5 * (m + b);
    double flm = kern_eval(k, lm), frm = kern_eval(k, rm);
    double left = (m - a) / 6.0 * (fa + 4.0 * flm + fm), right = (b - m) / 6.0 * (fm + 4.0 * frm + fb);
    double delta = left + right - whole;
    if (depth <= 0 || (depth <= 43 && fabs(delta) <= 15.0 * tol) || (b - a) < 1e-14 * (fabs(a) + fabs(b) + 1e-300))
        return left + right + delta / 15.0;
    return simpson_rec(k, a, m, fa, flm, fm, left, 0.5 * tol, depth - 1) +
           simpson_rec(k, m, b, fm, frm, fb, right, 0.5 * tol, depth - 1);
}
/* rel: tolerance relative to the larger of this piece's first estimate and what has been integrated so far (zref) */
static double simpson(const kern_fn *k, double a, double b, double rel, double zref) {
    /* end points nudged inside: a service of exactly zero has log-density -inf or +inf */
    double w = b - a, a1 = a + 1e-13 * w, b1 = b - 1e-13 * w, m = 0.5 * (a1 + b1);
    double fa = kern_eval(k, a1), fm = kern_eval(k, m), fb = kern_eval(k, b1), whole, tol;
    if (!(fa == fa) || fa > 1e300) fa = 0.0;
    if (!(fb == fb) || fb > 1e300) fb = 0.0;
    whole = (b1 - a1) / 6.0 * (fa + 4.0 * fm + fb);
    tol = rel * fmax(fabs(whole), zref);
    if (!(tol > 1e-300)) tol = 1e-300;
    return simpson_rec(k, a1, b1, fa, fm, fb, whole, tol, 48);
}

static int cmp_double(const void *x, const void *y) {
    double a = *(const double *)x, b = *(const double *)y;
    return (a > b) - (a < b);
}

/* GGkGibbs.integrate (qnet.pyx:1361-1388): the range is cut at the departures of the 3 queue neighbours of e on
 * either side and at the arrivals of the 3 neighbours of its task successor; an infinite upper end (final events)
 * is integrated in doubling steps until the tail no longer matters. */
static double kern_normaliser(const bqo_state *st, int e, double lf, double lower, double upper) {
    double pts[32], z = 0.0;
    int n = 0, i, j, c, e1 = st->next_byt[e];
    kern_fn k;
    k.st = st; k.e = e; k.lf = lf;
    pts[n++] = lower;
    for (c = e, i = 0; i < 3 && c >= 0; ++i, c = st->prev_byq[c])
        if (st->d[c] >= lower && st->d[c] <= upper) pts[n++] = st->d[c];
    for (c = st->next_byq[e], i = 0; i < 3 && c >= 0; ++i, c = st->next_byq[c])
        if (st->d[c] >= lower && st->d[c] <= upper) pts[n++] = st->d[c];
    if (e1 >= 0) {
        for (c = e1, i = 0; i < 3 && c >= 0; ++i, c = st->prev_byq[c])
            if (st->a[c] >= lower && st->a[c] <= upper) pts[n++] = st->a[c];
        for (c = st->next_byq[e1], i = 0; i < 3 && c >= 0; ++i, c = st->next_byq[c])
            if (st->a[c] >= lower && st->a[c] <= upper) pts[n++] = st->a[c];
    }
    if (upper < INFINITY) pts[n++] = upper;
    qsort(pts, (size_t)n, sizeof(double), cmp_double);
    for (i = 0, j = 0; i < n; ++i)
        if (j == 0 || pts[i] > pts[j - 1]) pts[j++] = pts[i];
    n = j;
    for (i = 0; i + 1 < n; ++i) z += simpson(&k, pts[i], pts[i + 1], 1e-12, z);
    if (!(upper < INFINITY)) {
        double t0 = pts[n - 1], w = fmax(1.0, fabs(st->d[e] - lower));
        for (i = 0; i < 200; ++i) {
            double part = simpson(&k, t0, t0 + w, 1e-12, z);
            z += part;
            t0 += w;
            w *= 2.0;
            if (i > 2 && part <= 1e-15 * z) break;
        }
    }
    return z;
}

/* the normaliser of one event's conditional relative to its value at x (parity hook) */
double bqo_kernel_normaliser(const bqo_state *st, int e, double x) {
    int e1 = st->next_byt[e];
    return kern_normaliser(st, e, bqo_logcond1(st, e, x), st->a[e], e1 >= 0 ? st->d[e1] : INFINITY);
}

/* log T(d_to <- st) of one sweep over `order`; st is left at d_to for the events that were moved */
double bqo_gibbs_kernel(bqo_state *st, const int *order, int n_order, const double *d_to, int strict) {
    int *todo = (int *)malloc((size_t)(n_order > 0 ? n_order : 1) * sizeof(int));
    int n = 0, i, progress;
    double prob = 0.0;
    for (i = 0; i < n_order; ++i)
        if (bqo_eligible(st, order[i])) todo[n++] = order[i];
    while (n > 0) {
        int kept = 0;
        progress = 0;
        for (i = 0; i < n; ++i) {
            int e = todo[i], e1 = st->next_byt[e];
            double lower = st->a[e], upper = e1 >= 0 ? st->d[e1] : INFINITY, x = d_to[e], lf, z;
            lf = bqo_logcond1(st, e, x);
            if (x <= lower || upper <= x || !(lf > -INFINITY)) { todo[kept++] = e; continue; } /* deferred */
            z = kern_normaliser(st, e, lf, lower, upper);
            prob -= log(z);
            bqo_apply(st, e, x);
            ++progress;
        }
        n = kept;
        if (!progress || (strict && n > 0)) { prob = -INFINITY; break; } /* strict: the fixed-scan sweep itself */
    }
    free(todo);
    return prob;
}

/* ---------------------------------------------------------------------------------------------- */
/* sufficient statistics and parameter updates                                                      */
/* ---------------------------------------------------------------------------------------------- */
/* out[q*8 + ..] = n(s>0), sum s, sum log s, sum (log s - mean)^2, #log terms, sum wait, n events, #(s==0) */
void bqo_suffstats(const bqo_state *st, double *out) {
    int q, e;
    for (q = 0; q < st->Q; ++q) {
        double *o = out + 8 * q, mean;
        int dist = st->q_dist[q];
        memset(o, 0, 8 * sizeof(double));
        for (e = 0; e < st->E; ++e) {
            double s, w;
            if (st->qid[e] != q) continue;
            s = st->s[e];
            w = st->d[e] - st->a[e] - s;
            o[5] += w > 0.0 ? w : 0.0;
            o[6] += 1.0;
            if (s > 0.0) {
                o[0] += 1.0; o[1] += s;
                if (dist == D_GAMMA) { o[2] += log(s); o[4] += 1.0; }
                else if (dist == D_LN && s >= 1e-50) { o[2] += log(s); o[4] += 1.0; }
            } else if (s == 0.0) o[7] += 1.0;
        }
        if (dist == D_LN && o[4] > 0.0) {
            mean = o[2] / o[4];
            for (e = 0; e < st->E; ++e)
                if (st->qid[e] == q && st->s[e] >= 1e-50) { double z = log(st->s[e]) - mean; o[3] += z * z; }
        }
    }
}

static double digamma_(double x) {
    double r = 0.0, f;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    f = 1.0 / (x * x);
    return r + log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f * (1.0 / 132.0)))));
}
static double trigamma_(double x) {
    double r = 0.0, f;
    while (x < 6.0) { r += 1.0 / (x * x); x += 1.0; }
    f = 1.0 / (x * x);
    return r + 1.0 / x + 0.5 * f + (1.0 / x) * f * (1.0 / 6.0 - f * (1.0 / 30.0 - f * (1.0 / 42.0 - f * (1.0 / 30.0))));
}

/* Qnet.estimate for one queue from its statistics (queues.pyx:1324-1336; distributions.pyx:65-74, 142-160, 276-303) */
void bqo_estimate(int dist, const double *ss, double *p0, double *p1) {
    if (dist == D_EXP) {
        if (ss[0] < 1.0) { *p0 = 1.0; return; }
        *p0 = ss[1] / ss[0];
        if (*p0 < 1e-7) *p0 = 1e-7;
    } else if (dist == D_LN) {
        double sd;
        if (ss[0] < 1.0) { *p0 = 0.0; *p1 = 1.0; return; }
        sd = sqrt(ss[3] / ss[4]);
        if (sd < 1e-10) sd = 1e-10;
        *p0 = ss[2] / ss[4]; *p1 = sd;
    } else {
        double mu, mu_log, a, a_old = -INFINITY;
        int it;
        if (ss[0] < 1.0) return;
        mu = ss[1] / ss[0]; mu_log = ss[2] / ss[0];
        a = 0.5 / (log(mu) - mu_log);
        for (it = 0; it < 200 && fabs(a - a_old) >= 0.01; ++it) {
            a_old = a;
            a = 1.0 / (1.0 / a_old + (mu_log - log(mu) + log(a_old) - digamma_(a_old)) /
                                         (a_old * a_old * (1.0 / a_old - trigamma_(a_old))));
        }
        *p0 = a; *p1 = mu / a;
    }
}

static double draw_normal(bqo_rng *r) {
    double u1 = 1.0 - rng_uniform(r), u2 = rng_uniform(r);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}
static double draw_gamma(bqo_rng *r, double shape) {
    double boost = 1.0, d, c;
    int it;
    if (shape < 1.0) { double u = 1.0 - rng_uniform(r); boost = pow(u, 1.0 / shape); shape += 1.0; }
    d = shape - 1.0 / 3.0; c = 1.0 / sqrt(9.0 * d);
    for (it = 0; it < 1000; ++it) {
        double x = draw_normal(r), v = 1.0 + c * x, u;
        if (v <= 0.0) continue;
        v = v * v * v;
        u = 1.0 - rng_uniform(r);
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
    }
    return boost * d;
}

typedef struct { double logP, q, r, s; } gpost;
static double gp_eval(const gpost *g, int which, double x, double fixed) {
    if (which == 0) return (x - 1.0) * g->logP - g->q / fixed - g->r * lgamma(x) - (x * g->s) * log(fixed);
    return (fixed - 1.0) * g->logP - g->q / x - g->r * lgamma(fixed) - (fixed * g->s) * log(x);
}
/* sampling._slice, stepping-out branch (sampling.pyx:299-366) with lower = 0, upper = inf */
static double slice_stepout(bqo_rng *rng, const gpost *g, int which, double fixed, double x0) {
    double w = 1.0, gx0 = gp_eval(g, which, x0, fixed), logy, u, L, R;
    unsigned long long J, K;
    int guard = 0, it;
    logy = gx0 - (-log(1.0 - rng_uniform(rng)));
    if (logy == -INFINITY || logy != logy) return x0;
    u = w * rng_uniform(rng);
    L = x0 - u; R = x0 + (w - u);
    J = (unsigned long long)(rng_word(rng) % 1025u);
    K = 1023ull - J;
    while (J > 0) { if (L <= 0.0) break; if (gp_eval(g, which, L, fixed) <= logy) break; L -= w; --J; w *= 2.0; }
    while (K > 0 && guard < 1100) { if (gp_eval(g, which, R, fixed) <= logy) break; R += w; --K; w *= 2.0; ++guard; }
    if (L < 0.0) L = 0.0;
    for (it = 0; it < 100000; ++it) {
        double x1 = L + (R - L) * rng_uniform(rng), g1 = gp_eval(g, which, x1, fixed);
        if (g1 >= logy) return x1;
        if (x1 > x0) R = x1; else L = x1;
        if (R - L < 1e-10) return L;
    }
    return x0;
}

/* Qnet.sample_parameters for one queue from its statistics (queues.pyx:1316; distributions.pyx:92-96,
 * 306-320; netutils.py:19-42), drawing from stream (seed, chain, sweep, q, tag 2) */
void bqo_sample_params(int dist, const double *ss, uint64_t seed, int64_t chain, uint32_t sweep, uint32_t q,
                       double *p0, double *p1) {
    bqo_rng rng;
    rng_init(&rng, seed, (uint64_t)chain, sweep, q, 2u);
    if (dist == D_EXP) {
        if (ss[0] < 1.0 || !(ss[1] > 0.0)) return;
        *p0 = 1.0 / (draw_gamma(&rng, ss[0]) / ss[1]);
    } else if (dist == D_LN) {
        double mean = 0.0, sd = 1.0, alpha, beta, tau, sdn;
        if (ss[0] >= 1.0) { mean = ss[2] / ss[4]; sd = sqrt(ss[3] / ss[4]); if (sd < 1e-10) sd = 1e-10; }
        if (sd <= 1e-5) { *p0 = mean; *p1 = sd; return; }
        if (ss[0] < 1.0) { *p0 = 0.0; *p1 = 1.0; return; }
        alpha = (ss[0] + 1.0) / 2.0; beta = 2.0 / (ss[0] * sd * sd);
        tau = draw_gamma(&rng, alpha) * beta;
        sdn = sqrt(1.0 / tau);
        *p0 = mean + sdn * draw_normal(&rng);
        *p1 = sdn;
    } else {
        gpost g;
        double sh = *p0, sc = *p1;
        int t;
        g.logP = 1.0 + ss[2]; g.q = 1.0 + ss[1]; g.r = 1.0 + ss[0]; g.s = 1.0 + ss[0];
        for (t = 0; t < 5; ++t) sh = slice_stepout(&rng, &g, 0, *p1, sh);
        for (t = 0; t < 5; ++t) sc = slice_stepout(&rng, &g, 1, sh, sc);
        *p0 = sh; *p1 = sc;
    }
}
