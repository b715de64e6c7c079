// TEST INFRASTRUCTURE ONLY.  Runs the __host__ __device__ walk / commit routines of the product's dynamic-order
// kernels (bayes_qnet_b200/csrc/bq_dynamic.cuh) on the CPU for ONE chain, so that the CPU test-suite can check the
// kernel logic against oracle/ without a GPU.  Nothing under bayes_qnet_b200/ links or loads this file; the
// product has no host execution path.  The slice loop below restates, sequentially (group width 1), what
// sweep_dynamic_kernel does with W lanes.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../bayes_qnet_b200/csrc/bq_dynamic.cuh"
#include "../../bayes_qnet_b200/csrc/bq_host.h"

using namespace bq;

struct Emul {
    int E, Q, P;
    std::vector<int> q_type, q_k, q_dist, q_poff, q_off, q_foff;
    std::vector<int> qid, prev_byt, next_byt, ord, ord0, pos, dord, dpos;
    std::vector<uint8_t> obs_a, obs_d;
    std::vector<double> d, F, sq, at, dq, ls, dt, theta;
    std::vector<QConst> qc;
    std::vector<QInfo> qinfo;
    std::vector<double> recs;  // at, dq, ls, F interleaved by position: what the kernels work on
    std::vector<unsigned long long> flags;  // contended-arrival flags, one byte per position (8-byte aligned storage)
    int EF, S;
    DynArgs A;
    Chain c;
    long evals, steps;

    void refresh_qc() {
        for (int q = 0; q < Q; ++q) {
            int o = q_poff[q];
            qconst_set(qc[q], q_dist[q], theta[o], q_dist[q] == D_EXP ? 0.0 : theta[o + 1]);
        }
    }
    // at / dq / ls / F (host_rebuild's separate arrays) <-> position records
    void pack() {
        recs.assign((size_t)E * S, 0.0);
        flags.assign((size_t)(((E + 31) & ~31) + 32) / 8, 0ull);
        for (int q = 0; q < Q; ++q)
            for (int p = q_off[q]; p < q_off[q + 1]; ++p) {
                double *r = &recs[(size_t)p * S];
                r[BQ_REC_AT] = at[p]; r[BQ_REC_DQ] = dq[p]; r[rec_ls_slot(S)] = ls[p];
                if (q_type[q] == Q_GGK)
                    for (int j = 0; j < q_k[q]; ++j) r[BQ_REC_F + BQ_FSLOT(j)] = F[q_foff[q] + (size_t)(p - q_off[q]) * q_k[q] + j];
                if (q_type[q] == Q_GGK && q_k[q] > 1 && r[BQ_REC_AT] < r[BQ_REC_F + 1]) reinterpret_cast<unsigned char *>(flags.data())[p] = 1;
            }
    }
    void unpack() {
        for (int q = 0; q < Q; ++q)
            for (int p = q_off[q]; p < q_off[q + 1]; ++p) {
                const double *r = &recs[(size_t)p * S];
                at[p] = r[BQ_REC_AT]; dq[p] = r[BQ_REC_DQ]; ls[p] = r[rec_ls_slot(S)];
                if (q_type[q] == Q_GGK)
                    for (int j = 0; j < q_k[q]; ++j) F[q_foff[q] + (size_t)(p - q_off[q]) * q_k[q] + j] = r[BQ_REC_F + BQ_FSLOT(j)];
            }
    }
    void bind() {
        memset(&A, 0, sizeof(A));
        A.E = E; A.Q = Q; A.P = P; A.n_chains = 1; A.S = S;
        A.qinfo = qinfo.data(); A.q_poff = q_poff.data();
        A.state = d.data(); A.ord = ord.data(); A.pos = pos.data(); A.rec = recs.data(); A.flag = reinterpret_cast<unsigned char *>(flags.data()); A.FW = (int)flags.size() * 8;
        A.dord = dord.data(); A.dpos = dpos.data(); A.dt = dt.data(); A.sq = sq.data();
        A.theta = theta.data();
        c.bind(A, 0, qinfo.data(), qc.data());
    }
    HostNet net() const {
        return HostNet{E, Q, q_type.data(), q_k.data(), q_off.data(), q_foff.data(), q_dist.data(), prev_byt.data()};
    }
    DynRec rec(int e) const {
        const int e1 = next_byt[e];
        return DynRec{e, e1, qid[e], e1 >= 0 ? qid[e1] : -1};
    }
};

extern "C" {

void *bqe_create(int E, int Q, int P, const int *q_type, const int *q_k, const int *q_dist, const int *q_poff,
                 const double *theta, const int *qid, const int *prev_byt, const int *next_byt, const int *prev_byq,
                 const int *next_byq, const uint8_t *obs_a, const uint8_t *obs_d, const double *d) {
    Emul *m = new Emul();
    m->E = E; m->Q = Q; m->P = P;
    m->q_type.assign(q_type, q_type + Q); m->q_k.assign(q_k, q_k + Q); m->q_dist.assign(q_dist, q_dist + Q);
    m->q_poff.assign(q_poff, q_poff + Q);
    m->theta.assign(theta, theta + P);
    m->qid.assign(qid, qid + E); m->prev_byt.assign(prev_byt, prev_byt + E); m->next_byt.assign(next_byt, next_byt + E);
    m->obs_a.assign(obs_a, obs_a + E); m->obs_d.assign(obs_d, obs_d + E);
    m->d.assign(d, d + E);
    m->q_off.assign(Q + 1, 0);
    for (int e = 0; e < E; ++e) m->q_off[qid[e] + 1]++;
    for (int q = 0; q < Q; ++q) m->q_off[q + 1] += m->q_off[q];
    m->ord.assign(E, -1); m->pos.assign(E, -1); m->dord.assign(E, -1); m->dpos.assign(E, -1);
    m->sq.assign(E, 0.0); m->at.assign(E, 0.0); m->dq.assign(E, 0.0); m->ls.assign(E, 0.0); m->dt.assign(E, 0.0);
    for (int q = 0; q < Q; ++q) {
        int p = m->q_off[q];
        for (int e = 0; e < E; ++e)
            if (qid[e] == q && prev_byq[e] < 0)
                for (int x = e; x >= 0; x = next_byq[x]) m->ord[p++] = x;
    }
    m->ord0 = m->ord;
    m->q_foff.assign(Q, 0);
    m->EF = 0;
    m->qinfo.resize(Q);
    for (int q = 0; q < Q; ++q) {
        if (q_type[q] == Q_GGK) { m->q_foff[q] = m->EF; m->EF += q_k[q] * (m->q_off[q + 1] - m->q_off[q]); }
        m->qinfo[q] = QInfo{q_type[q], q_k[q], m->q_off[q], m->q_off[q + 1], m->q_foff[q], q_dist[q]};
    }
    m->F.assign(m->EF + 1, 0.0);
    {
        int kmax = 1;
        for (int q = 0; q < Q; ++q) if (q_type[q] == Q_GGK && q_k[q] > kmax) kmax = q_k[q];
        m->S = rec_stride(kmax);
    }
    m->qc.resize(Q);
    {
        HostNet hn = m->net();
        HostChain hc{m->ord.data(), m->pos.data(), m->at.data(), m->dq.data(), m->ls.data(), m->F.data(), m->dt.data(),
                     m->sq.data(), m->dord.data(), m->dpos.data()};
        host_rebuild(&hn, m->d.data(), hc, false);
    }
    m->pack();
    m->refresh_qc();
    m->bind();
    m->evals = m->steps = 0;
    return m;
}

void bqe_destroy(void *h) { delete (Emul *)h; }

void bqe_set_theta(void *h, const double *theta) {
    Emul *m = (Emul *)h;
    m->theta.assign(theta, theta + m->P);
    m->refresh_qc();
}

void bqe_logcond(void *h, int e, const double *x, int n, double *out) {
    Emul *m = (Emul *)h;
    EventCtx ev;
    event_load(m->c, m->rec(e), ev);
    Touch tc;
    tc.reset();
    const double g0 = dyn_logcond<8>(m->c, ev, ev.x0, tc);
    for (int j = 0; j < n; ++j) out[j] = dyn_logcond<8>(m->c, ev, x[j], tc) - g0;
}

void bqe_apply(void *h, int e, double x) {
    Emul *m = (Emul *)h;
    EventCtx ev;
    event_load(m->c, m->rec(e), ev);
    dyn_commit<1, 8>(m->c, ev, x, 0, 1u);
}

// one sweep over `order`, same stream convention as the kernels: uniform 0 of (chain, sweep, event) is the slice
// level, uniform c >= 1 the c-th shrink candidate (c = 1 from the level's block, 53 bits; c >= 2 one 32-bit word each)
long bqe_slice_sweep(void *h, const int *order, int n_order, uint64_t seed, int64_t chain, uint32_t sweep, double tmax) {
    Emul *m = (Emul *)h;
    long ne = 0;
    for (int idx = 0; idx < n_order; ++idx) {
        EventCtx ev;
        event_load(m->c, m->rec(order[idx]), ev);
        const double lower = ev.a_e, upper = (ev.e1 >= 0) ? fmin(ev.d1, tmax) : tmax;
        Rng rng;
        rng.init(seed, (uint64_t)chain, sweep, (uint32_t)ev.e, TAG_SLICE);
        Touch tc;
        tc.reset();
        const double g0 = dyn_logcond<8>(m->c, ev, ev.x0, tc);
        m->evals++;
        ++ne;
        const double logy = g0 - (-log(1.0 - rng.uniform()));
        if (!(logy > BQ_NEG_INF)) continue;
        double L = lower, R = upper, x1 = ev.x0;
        for (int it = 0; it < BQ_SLICE_MAX_ITERS; ++it) {
            x1 = L + (R - L) * (it == 0 ? rng.uniform() : rng.uniform32());  // candidate uniforms, bq_rng.cuh
            m->evals++;
            if (dyn_logcond<8>(m->c, ev, x1, tc) >= logy) break;
            if (x1 > ev.x0) R = x1; else L = x1;
            if (R - L < 1e-10) { x1 = ev.x0; break; }
        }
        m->steps += tc.steps;
        if (x1 != ev.x0) dyn_commit<1, 8>(m->c, ev, x1, 0, 1u);
    }
    return ne;
}

// one MH-within-Gibbs sweep with the product's proposal / accept routines, sequentially (group width 1)
long bqe_mh_sweep(void *h, const int *order, int n_order, uint64_t seed, int64_t chain, uint32_t sweep, int sample_final,
                  long *nreject) {
    Emul *m = (Emul *)h;
    long ne = 0;
    for (int idx = 0; idx < n_order; ++idx) {
        EventCtx ev;
        event_load(m->c, m->rec(order[idx]), ev);
        const bool fin = ev.e1 < 0;
        if (fin && !sample_final) continue;
        MhProp pr;
        mh_setup(m->c, ev, pr);
        if (pr.Uh - pr.Lh < 1e-10) continue;
        if (!fin && ev.d1 - ev.a_e < 1e-10) continue;
        Rng aux;
        aux.init(seed, (uint64_t)chain, sweep, (uint32_t)ev.e, TAG_MH_AUX);
        const double u_acc = aux.uniform();
        (void)aux.uniform();
        const double h0 = pr.eval(ev.x0);
        double x_init = ev.x0;
        if (!(h0 > BQ_NEG_INF)) {
            int tries = 0;
            for (; tries < 25; ++tries) {
                const double u = aux.uniform();
                x_init = fin ? pr.b0 + (-log(1.0 - u)) : ev.a_e + (ev.d1 - ev.a_e) * u;
                if (pr.eval(x_init) > BQ_NEG_INF) break;
            }
            if (tries >= 25) continue;
        }
        double x_new = x_init;
        long long evals = 0;
        for (int step = 0; step < 2; ++step) {
            Rng rng;
            rng.init(seed, (uint64_t)chain, sweep, (uint32_t)ev.e, step == 0 ? TAG_MH : TAG_MH2);
            if (fin) {
                x_new = slice_stepout_seq([&](double x) { return pr.eval(x); }, x_new, pr.Lh, pr.Uh, rng, evals);
            } else {  // finite bounds: what group_slice does with W lanes, one candidate at a time
                const double x0 = x_new;
                const double logy = pr.eval(x0) - (-log(1.0 - rng.uniform()));
                if (!(logy > BQ_NEG_INF)) continue;
                double L = pr.Lh, R = pr.Uh;
                for (int it = 0; it < BQ_SLICE_MAX_ITERS; ++it) {
                    const double x1 = L + (R - L) * (it == 0 ? rng.uniform() : rng.uniform32());
                    if (pr.eval(x1) >= logy) { x_new = x1; break; }
                    if (x1 > x0) R = x1; else L = x1;
                    if (R - L < 1e-10) break;
                }
            }
        }
        ++ne;
        Touch tc;
        tc.reset();
        const double g_new = dyn_logcond<8>(m->c, ev, x_new, tc);
        const double ratio = exp(h0 - pr.eval(x_new) + g_new);
        if (u_acc < ratio) { if (x_new != ev.x0) dyn_commit<1, 8>(m->c, ev, x_new, 0, 1u); }
        else if (nreject) ++*nreject;
    }
    return ne;
}

void bqe_mh_proposal(void *h, int e, const double *x, int n, double *out) {
    Emul *m = (Emul *)h;
    EventCtx ev;
    event_load(m->c, m->rec(e), ev);
    MhProp pr;
    mh_setup(m->c, ev, pr);
    for (int j = 0; j < n; ++j) out[j] = pr.eval(x[j]);
}

// ---- batch semantics of sweep_batch_kernel, emulated: B lanes evaluate B consecutive events of the order on the same
// (unchanged) state; the longest prefix in which no lane read a range that an earlier lane writes is committed; the
// rest is redone.  Must reproduce the sequential sweep exactly, for any B.
long bqe_sweep_batch(void *h, const int *order, int n_order, uint64_t seed, int64_t chain, uint32_t sweep, double tmax,
                     int B, int mh, int sample_final, long *nreject, long *nredo) {
    Emul *m = (Emul *)h;
    long ne = 0;
    std::vector<EventCtx> evs(B);
    std::vector<LaneUpdate> up(B);
    LaneParams LP;
    LP.seed = seed; LP.gchain = (unsigned long long)chain; LP.gsw = sweep; LP.mode = mh ? MODE_MH : MODE_SLICE;
    LP.flags = sample_final ? 0 : FLAG_NO_FINAL; LP.tmax = tmax;
    for (int s = 0; s < n_order;) {
        const int nb = std::min(B, n_order - s);
        for (int l = 0; l < nb; ++l) {
            event_load(m->c, m->rec(order[s + l]), evs[l]);
            lane_update<8>(m->c, evs[l], LP, up[l]);
        }
        int fc = nb;
        for (int l = 1; l < nb && fc == nb; ++l)
            for (int j = 0; j < l; ++j)
                if (up[j].changed && touch_conflict(up[l].R, up[j].Wt)) { fc = l; break; }
        for (int l = 0; l < fc; ++l) {
            ne += up[l].counted;
            if (up[l].rejected && nreject) ++*nreject;
            if (up[l].changed) dyn_commit<1, 8>(m->c, evs[l], up[l].x_new, 0, 1u);
        }
        if (nredo) *nredo += nb - fc;
        s += fc;
    }
    return ne;
}

void bqe_get(void *h, double *d, int *ord, double *F, double *svc) {
    Emul *m = (Emul *)h;
    if (d) memcpy(d, m->d.data(), m->E * sizeof(double));
    if (ord) memcpy(ord, m->ord.data(), m->E * sizeof(int));
    m->unpack();
    if (F) memcpy(F, m->F.data(), m->EF * sizeof(double));
    if (svc) memcpy(svc, m->sq.data(), m->E * sizeof(double));
}

// a, s, proc, links recomputed from (d, ord) by the library's host code
void bqe_derive(void *h, double *a, double *s, int32_t *proc, int32_t *prev_byq, int32_t *next_byq) {
    Emul *m = (Emul *)h;
    HostNet hn = m->net();
    host_derive(&hn, m->d.data(), m->ord.data(), a, s, proc, prev_byq, next_byq, m->sq.data(), m->ord0.data());
}

// consistency of the incrementally maintained arrays with a from-scratch rebuild: 0 = consistent
int bqe_check(void *h) {
    Emul *m = (Emul *)h;
    m->unpack();
    std::vector<int> ord(m->ord), pos(m->E), dord(m->E, -1), dpos(m->E, -1);
    std::vector<double> F(m->EF + 1, 0.0), at(m->E, 0.0), dq(m->E, 0.0), ls(m->E, 0.0), dt(m->E, 0.0), sq(m->E, 0.0);
    HostNet hn = m->net();
    HostChain hc{ord.data(), pos.data(), at.data(), dq.data(), ls.data(), F.data(), dt.data(), sq.data(), dord.data(),
                 dpos.data()};
    host_rebuild(&hn, m->d.data(), hc, false);
    for (int e = 0; e < m->E; ++e)
        if (pos[e] != m->pos[e]) return 1;
    for (int i = 0; i < m->EF; ++i)
        if (F[i] != m->F[i]) return 2;
    for (int p = 0; p < m->E; ++p) {
        if (at[p] != m->at[p]) return 8;
        if (dq[p] != m->dq[p]) return 9;
    }
    for (int q = 0; q < m->Q; ++q)  // the cached log services of non-exponential FIFO / delay queues
        if (m->q_dist[q] != D_EXP && m->q_type[q] != Q_PS && m->q_type[q] != Q_GG1R)
            for (int p = m->q_off[q]; p < m->q_off[q + 1]; ++p)
                if (ls[p] != m->ls[p] && !(ls[p] != ls[p] && m->ls[p] != m->ls[p])) return 10;
    for (int q = 0; q < m->Q; ++q) {
        if (m->q_type[q] != Q_PS) continue;
        for (int p = m->q_off[q]; p < m->q_off[q + 1]; ++p) {
            if (dt[p] != m->dt[p]) return 5;
            if (dt[p] != m->d[m->dord[p]] || m->dpos[m->dord[p]] != p) return 6;
            if (fabs(sq[p] - m->sq[p]) > 1e-9 * fmax(1.0, fabs(sq[p]))) return 7;
        }
    }
    for (int q = 0; q < m->Q; ++q) {  // random-order queues: departures in order, each slot with its event's arrival
        if (m->q_type[q] != Q_GG1R) continue;
        for (int p = m->q_off[q]; p < m->q_off[q + 1]; ++p) {
            if (dt[p] != m->dt[p]) return 12;
            if (dt[p] != m->d[m->dord[p]] || m->dpos[m->dord[p]] != p) return 13;
            if (m->sq[p] != m->at[m->pos[m->dord[p]]]) return 14;
            if (p > m->q_off[q] && m->dt[p] < m->dt[p - 1]) return 15;
        }
    }
    for (int q = 0; q < m->Q; ++q)  // the contended-arrival bits of multi-server FIFO queues
        if (m->q_type[q] == Q_GGK && m->q_k[q] > 1)
            for (int p = m->q_off[q]; p < m->q_off[q + 1]; ++p) {
                const double f1 = F[m->q_foff[q] + (size_t)(p - m->q_off[q]) * m->q_k[q] + 1];
                const bool want = at[p] < f1, have = reinterpret_cast<const unsigned char *>(m->flags.data())[p] != 0;
                if (want != have) return 11;
            }
    // every non-delay queue must be sorted by arrival
    for (int q = 0; q < m->Q; ++q) {
        if (m->q_type[q] == Q_DELAY) continue;
        for (int p = m->q_off[q] + 1; p < m->q_off[q + 1]; ++p)
            if (m->at[p] < m->at[p - 1]) return 3;
    }
    return 0;
}

long bqe_evals(void *h) { return ((Emul *)h)->evals; }
}
