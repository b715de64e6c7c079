// Sweep kernels for networks whose queue order is invariant under the sampler: every queue is a
// single-server FIFO queue (QueueGGk with k = 1 -- what YAML `GG1` and the default produce,
// qnetu.py:103-110 -- including the source queue, arrivals.pyx:256) or a DelayStation (queues.pyx:745).
// For these the diff lists of queues.pyx:470-591 collapse to the closed form of SURVEY.md A.2/A.4, the
// by-queue links never change and the only per-chain mutable state is d[E].
//
// Execution model (B200): one CTA owns one chain.  The chain's departure times are staged in shared
// memory (E * 8 B; 160 KB for the 5k-task 3-tier benchmark) for the whole launch, all n_sweeps sweeps
// run inside one launch, and within a sweep the resample-eligible events are visited colour class by
// colour class: events of one class have disjoint Markov blankets (no event of the class is read by
// another event's conditional), so the threads of the CTA update them concurrently and the result is
// the same as visiting them one after another -- a fixed-scan Gibbs sweep, just in an order chosen for
// the hardware.  The structure tables are identical for all chains and are
// served from L2.  Sufficient statistics, the service-parameter update and the per-sweep summaries
// are fused into the same kernel.
#pragma once
#include "bq_dist.cuh"

namespace bq {

#ifndef BQ_MAX_THREADS
#define BQ_MAX_THREADS 512
#endif
#define BQ_SLICE_MAX_ITERS 1000
#define BQ_NSTAT 8  // per-chain counters: nevents, slice_evals, param_evals, stuck, nreject, walk steps, 2 spare

enum { UPD_NONE = 0, UPD_BAYES = 1, UPD_STEM = 2 };
enum { MODE_SLICE = 0, MODE_MH = 1 };
enum { FLAG_TMAX_PER_SWEEP = 1, FLAG_NO_FINAL = 2, FLAG_TRACE = 4, FLAG_TIGHT = 8, FLAG_REVERSE = 16, FLAG_PARAMS_ONLY = 1 << 16 };

// Static neighbourhood of one resample-eligible event e (12 x int32 = 3 x int4, stored in schedule
// order so a colour class is one contiguous, coalesced read).  -1 = absent.
struct SchedRec {
    int e;      // the event whose departure is resampled
    int pt_e;   // previous event of e's task: a_e = d[pt_e] (0 when absent: source queue)
    int p;      // previous event in e's queue
    int n;      // next event in e's queue
    int pt_n;   // previous-by-task of n: a_n = d[pt_n]
    int e1;     // next event of e's task (its arrival is d_e)
    int p1;     // previous event in e1's queue
    int pt_p1;  // previous-by-task of p1: a_{p1} (arrival order must be kept)
    int pt_n1;  // previous-by-task of e1's queue successor: a_{n1}
    int q0;     // queue of e
    int q1;     // queue of e1
    int pad;    // the event's id for the RNG stream (= e, except in the cluster kernel's re-encoded records)
};

struct SweepArgs {
    int E, Q, P, n_colors, n_sweeps, update, flags, trace_base, n_chains, refill, mode;
    const int4 *rec;
    const int *color_off;
    const int *ev_q, *ev_p, *ev_pt;  // per event: queue, prev-by-queue, prev-by-task
    const int *q_events, *q_off;     // events grouped by queue
    const int *q_type;               // [Q]
    // Service distributions live in SLOTS: a plain queue owns one slot, a queue with a mixture template
    // (MixtureTemplate, queues.pyx:1373-1509) one per component.  Without mixtures NS == Q and slot == queue.
    int NS;                          // number of slots
    const int *q_dist, *q_poff;      // [NS] family and offset of the first distribution parameter of each SLOT
    const int *q_slot;               // [Q+1] first slot of each queue
    const int *q_mixoff;             // [Q] offset of the queue's mixing weights in theta, or -1
    unsigned char *aux;              // [C][E] mixture component of every event (Event.auxiliaries), or null
    double *state;      // [C][E]
    double *theta;      // [C][P]
    long long *stats;   // [C][BQ_NSTAT]: nevents, slice_evals, param_evals, stuck, ...
    double *tr_theta;   // [rec][C][P]
    double *tr_wait;    // [rec][C][Q]
    double *tr_logp;    // [rec][C]
    unsigned long long seed;
    long long chain_offset;
    unsigned int sweep_base;
};

// max without fmax()'s NaN handling: DSETP + 2 selects instead of ~8 instructions (the hot loop has no NaNs)
__host__ __device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// Local conditional of d_e given its Markov blanket (GGkGibbs.inner_dfun, qnet.pyx:1340-1357, in the
// closed form of SURVEY.md A.2): G(d) = sum of the new-service log-densities; -inf outside the support.
struct Cond {
    double b;            // max(a_e, d_p): when e enters service        (queues.pyx:610-612)
    double a_n, d_n;     // queue successor of e
    double d_e1, d_p1;   // next-by-task event and its queue predecessor's departure
    double flo, fhi;     // feasible interval (all services >= 0, arrival order kept)
    int has_n, has_e1, e1_fifo;
    const QConst *c0, *cn, *c1;  // service density of e, of its queue successor, of the task's next event

    __device__ __forceinline__ double eval(double d) const {
        if (d < flo || d > fhi) return BQ_NEG_INF;
        double v = lpdf(*c0, d - b);
        if (has_n) v += lpdf(*cn, d_n - dmax(a_n, d));
        if (has_e1) v += lpdf(*c1, d_e1 - (e1_fifo ? dmax(d, d_p1) : d));
        return v;
    }
};

// qc is indexed by slot; with mixture templates (aux != null) an event's slot is its queue's first slot plus its
// component, otherwise the queue index itself
// DT: how the chain's departure times are addressed -- a plain pointer (shared or global memory), or a view that spreads
// them over the shared memories of a thread-block cluster (bq_cluster.cuh)
template <class DT>
__device__ __forceinline__ void cond_load(Cond &c, const SchedRec &r, const DT &d, const int *q_type,
                                          const QConst *qc, double &x0, double &lower, double &upper,
                                          double tmax, const int *q_slot = nullptr, const unsigned char *aux = nullptr) {
    x0 = d[r.e];
    double a_e = (r.pt_e >= 0) ? d[r.pt_e] : 0.0;
    int fifo0 = (q_type[r.q0] == Q_GGK);
    c.b = a_e;
    c.c0 = &qc[aux ? q_slot[r.q0] + aux[r.e] : r.q0];
    c.cn = c.c0;
    c.c1 = c.c0;
    double flo = a_e, fhi = INFINITY;
    if (fifo0 && r.p >= 0) { c.b = fmax(a_e, d[r.p]); flo = c.b; }
    c.has_n = (fifo0 && r.n >= 0);
    c.a_n = 0.0; c.d_n = 0.0;
    if (c.has_n) {
        c.d_n = d[r.n];
        c.a_n = (r.pt_n >= 0) ? d[r.pt_n] : 0.0;
        fhi = c.d_n;
        if (aux) c.cn = &qc[q_slot[r.q0] + aux[r.n]];
    }
    c.has_e1 = (r.e1 >= 0);
    c.e1_fifo = 0; c.d_e1 = 0.0; c.d_p1 = 0.0;
    lower = a_e;
    upper = tmax;
    if (c.has_e1) {
        c.d_e1 = d[r.e1];
        c.c1 = &qc[aux ? q_slot[r.q1] + aux[r.e1] : r.q1];
        upper = fmin(c.d_e1, tmax);
        fhi = fmin(fhi, c.d_e1);
        c.e1_fifo = (q_type[r.q1] == Q_GGK);
        if (c.e1_fifo) {
            if (r.p1 >= 0) {
                c.d_p1 = d[r.p1];
                flo = fmax(flo, (r.pt_p1 >= 0) ? d[r.pt_p1] : 0.0);
            }
            if (r.pt_n1 >= 0) fhi = fmin(fhi, d[r.pt_n1]);
        }
    }
    c.flo = flo; c.fhi = fhi;
}

// All-exponential networks (the paper's M/M/1 tiers): lpdf(s) = -log(mu) - s/mu, so up to a constant
// that cancels in every slice comparison G(d) is a sum of three linear pieces -- evaluated branch-free.
struct CondExp {
    double b, a_n, d_n, d_e1, d_p1, flo, fhi, inv0, invn, inv1;

    __device__ __forceinline__ void idle(const QConst *) {
        b = a_n = d_n = d_e1 = d_p1 = inv0 = invn = inv1 = 0.0; flo = 1.0; fhi = 0.0;
    }

    // G(d) - const = -d/mu0 + max(a_n, d)/mu0 [successor] + max(d, d_p1)/mu1 [next-by-task]: the terms of
    // -(d-b)/mu0 - (d_n - max(a_n,d))/mu0 - (d_e1 - max(d,d_p1))/mu1 that do not depend on d are dropped,
    // they cancel in g(x1) >= g(x0) - Exp(1).
    __device__ __forceinline__ double eval(double d) const {
        double v = fma(invn, dmax(a_n, d), fma(inv1, dmax(d, d_p1), -inv0 * d));
        return (d >= flo && d <= fhi) ? v : BQ_NEG_INF;
    }
    template <class DT>
    __device__ __forceinline__ void load(const SchedRec &r, const DT &d, const int *q_type, const QConst *qc,
                                         double &x0, double &lower, double &upper, double tmax, const int *q_slot,
                                         const unsigned char *aux) {
        Cond c;
        cond_load(c, r, d, q_type, qc, x0, lower, upper, tmax, q_slot, aux);
        b = c.b; flo = c.flo; fhi = c.fhi;
        inv0 = c.c0->inv;
        invn = c.has_n ? c.cn->inv : 0.0;
        a_n = c.a_n; d_n = c.d_n;                    // both 0 when absent: the term vanishes
        inv1 = c.has_e1 ? c.c1->inv : 0.0;
        d_e1 = c.d_e1;
        d_p1 = c.e1_fifo ? c.d_p1 : -INFINITY;        // delay station: max(d, -inf) = d
        if (!c.has_e1) d_p1 = 0.0;
    }
};

struct CondGen : Cond {
    // lanes that hold no event still run the branch-free shrink step: give them something harmless to evaluate
    __device__ __forceinline__ void idle(const QConst *qc) {
        b = a_n = d_n = d_e1 = d_p1 = 0.0; flo = 1.0; fhi = 0.0; has_n = has_e1 = e1_fifo = 0; c0 = cn = c1 = qc;
    }
    template <class DT>
    __device__ __forceinline__ void load(const SchedRec &r, const DT &d, const int *q_type, const QConst *qc,
                                         double &x0, double &lower, double &upper, double tmax, const int *q_slot,
                                         const unsigned char *aux) {
        cond_load(*this, r, d, q_type, qc, x0, lower, upper, tmax, q_slot, aux);
    }
};

#ifndef BQ_CAND
#define BQ_CAND 4  // candidates per trip (a multiple of 4: one Philox block feeds four)
#endif
#ifndef BQ_REFILL_THRESH
#define BQ_REFILL_THRESH 20
#endif

// One colour class, one warp's share of it.  Each lane runs the slice sampler of one event at a time as a
// state machine: every trip of the loop evaluates exactly one candidate per active lane; lanes whose
// sampler has finished are refilled from the class's shared work counter as soon as BQ_REFILL_THRESH of
// them are idle, so the data-dependent trip count of sampling._slice (1..~20 shrink steps) costs neither
// idle lanes inside the warp nor idle warps at the class barrier.  The result does not depend on which
// lane handles which event: the RNG stream is keyed by (chain, sweep, event).
// Two 53-bit uniforms of block `blk` of the (chain, sweep, event) stream.
__device__ __forceinline__ void stream_pair(const Rng &r, unsigned int blk, double &ua, double &ub) {
    uint32_t o[4];
    philox4x32(blk, r.c1, r.c2, r.c3, r.k0, r.k1, o);
    ua = u53(o[0], o[1]);
    ub = u53(o[2], o[3]);
}

template <class CondT, class DT>
__device__ __forceinline__ void run_class(const SweepArgs &A, DT d, const int *q_type, const QConst *qc,
                                          int hi, int *ctr, double tmax, unsigned long long gchain,
                                          unsigned int gsw, long long &n_ev, long long &n_evals,
                                          long long &n_stuck, const int *q_slot, const unsigned char *aux) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    bool active = false, exhausted = false;
    CondT cd;
    cd.idle(qc);
    Rng rng;
    rng.init(0ull, 0ull, 0u, 0u, 0u);
    double x0 = 0.0, L = 0.0, R = 0.0, logy = 0.0;
    int e = -1, it = 0;
    unsigned int blk = 0;
    int ev = 0, stuck = 0, nev = 0;
    for (;;) {
        const unsigned need = __ballot_sync(FULL, !active);
        if (!exhausted && (__popc(need) >= A.refill)) {
            const int cnt = __popc(need);
            int base = 0;
            if (lane == 0) base = atomicAdd(ctr, cnt);
            base = __shfl_sync(FULL, base, 0);
            const int my = base + __popc(need & ((1u << lane) - 1u));
            if (base + cnt >= hi) exhausted = true;
            if (!active && my < hi) {
                union { int4 v[3]; SchedRec r; } u;
                u.v[0] = __ldg(&A.rec[3 * my]);
                u.v[1] = __ldg(&A.rec[3 * my + 1]);
                u.v[2] = __ldg(&A.rec[3 * my + 2]);
                double lower, upper, ua, ub;
                cd.load(u.r, d, q_type, qc, x0, lower, upper, tmax, q_slot, aux);
                if (A.flags & FLAG_TIGHT) { lower = fmax(lower, cd.flo); upper = fmin(upper, cd.fhi); }
                e = u.r.e;
                rng.init(A.seed, gchain, gsw, (unsigned int)u.r.pad, TAG_SLICE);
                stream_pair(rng, 0u, ua, ub);
                blk = 1u;
                const double g0 = cd.eval(x0);
                logy = g0 - (-log(1.0 - ua));  // rk_exponential = -log(1 - U)   (cdist.c:107-110)
                nev += 1; ev += 1;
                if (logy > BQ_NEG_INF) {
                    // first candidate comes from the same Philox block as the slice level
                    L = lower; R = upper; it = 0;
                    const double x1 = L + (R - L) * ub;
                    ev += 1;
                    if (cd.eval(x1) >= logy) d[e] = x1;
                    else {
                        if (x1 > x0) R = x1; else L = x1;
                        if (R - L < 1e-10) stuck += 1; else active = true;
                    }
                } else stuck += 1;  // point-mass guard (sampling.pyx:289-291): keep x0
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }
        // BQ_CAND shrink steps of sampling._slice (sampling.pyx:344-364) per trip, BQ_CAND/4 Philox blocks.
        // A candidate's position depends on where the earlier ones fell, never on their densities, so all
        // positions are known up front and the densities are evaluated side by side (instruction-level
        // parallelism for a latency-bound loop); the first accepted one in order wins, which is exactly the
        // outcome of the sequential loop.  Branch-free: idle lanes compute harmlessly, only the store is
        // predicated.
        double u[BQ_CAND], x[BQ_CAND], La[BQ_CAND], Ra[BQ_CAND];
#pragma unroll
        for (int b = 0; b < BQ_CAND / 4; ++b) {  // four 32-bit candidate uniforms per Philox block (bq_rng.cuh)
            uint32_t o[4];
            philox4x32(blk + (unsigned)b, rng.c1, rng.c2, rng.c3, rng.k0, rng.k1, o);
            u[4 * b] = u32(o[0]); u[4 * b + 1] = u32(o[1]); u[4 * b + 2] = u32(o[2]); u[4 * b + 3] = u32(o[3]);
        }
        {
            double Lk = L, Rk = R;
#pragma unroll
            for (int k = 0; k < BQ_CAND; ++k) {
                x[k] = Lk + (Rk - Lk) * u[k];
                const bool up = x[k] > x0;
                Lk = up ? Lk : x[k];
                Rk = up ? x[k] : Rk;
                La[k] = Lk; Ra[k] = Rk;
            }
        }
        bool acc[BQ_CAND];
#pragma unroll
        for (int k = 0; k < BQ_CAND; ++k) acc[k] = cd.eval(x[k]) >= logy;
        bool going = active, done = false, stk = false;
        double xnew = x0;
        int nev_trip = 0;
#pragma unroll
        for (int k = 0; k < BQ_CAND; ++k) {
            nev_trip += going ? 1 : 0;
            const bool a = going && acc[k];
            const bool c = going && !acc[k] && (Ra[k] - La[k] < 1e-10);
            xnew = a ? x[k] : xnew;
            done = done || a;
            stk = stk || c;
            going = going && !a && !c;
        }
        it += BQ_CAND;
        if (going && it >= BQ_SLICE_MAX_ITERS) { stk = true; going = false; }
        if (active) {
            ev += nev_trip;
            if (done) d[e] = xnew;
            if (stk) stuck += 1;
            active = going;
        }
        L = La[BQ_CAND - 1]; R = Ra[BQ_CAND - 1];
        blk += BQ_CAND / 4;
    }
    n_ev += nev; n_evals += ev; n_stuck += stuck;
}

// One colour class of an MH-within-Gibbs sweep (Qnet.gibbs_resample, qnet.pyx:337-498) on a static-order network.
// The TWOS proposal (queues.pyx:1239-1256) is closed form here:
//   h(x) = lpdf_q0(x - b) + [e1] lpdf_q1(d_e1 - max(x, d_p1))   on [b, d_e1]  (b = max(a_e, d_p); no d_p1 for a delay
// station), and so is the exact conditional used in the accept ratio (Cond::eval).  Each lane handles one event at a
// time: two slice steps on h (call_sampler thin = 2, qnet.pyx:1292-1300; stepping out when the event is final and the
// support unbounded), then u < exp(h(d_e) - h(d') + G(d') - G(d_e)) (qnet.pyx:425-431).  Same three counter streams
// per (chain, sweep, event) as the dynamic-order kernel, so both implementations walk the same path.
struct MhPropStatic {
    const Cond *c;
    double Lh, Uh;
    __device__ __forceinline__ double operator()(double x) const {
        if (x < Lh || x > Uh) return BQ_NEG_INF;
        double v = lpdf(*c->c0, x - c->b);
        if (c->has_e1) v += lpdf(*c->c1, c->d_e1 - (c->e1_fifo ? dmax(x, c->d_p1) : x));
        return v;
    }
};

template <class DT>
__device__ __forceinline__ void run_class_mh(const SweepArgs &A, DT d, const int *q_type, const QConst *qc, int hi,
                                             int *ctr, unsigned long long gchain, unsigned int gsw, long long &n_ev,
                                             long long &n_evals, long long &n_reject, const int *q_slot,
                                             const unsigned char *aux) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(ctr, 32);
        base = __shfl_sync(FULL, base, 0);
        if (base >= hi) break;
        const int my = base + lane;
        if (my < hi) {
            union { int4 v[3]; SchedRec r; } u;
            u.v[0] = __ldg(&A.rec[3 * my]);
            u.v[1] = __ldg(&A.rec[3 * my + 1]);
            u.v[2] = __ldg(&A.rec[3 * my + 2]);
            Cond cd;
            double x0, a_e, upper;
            cond_load(cd, u.r, d, q_type, qc, x0, a_e, upper, INFINITY, q_slot, aux);
            const bool final_evt = !cd.has_e1;
            MhPropStatic h;
            h.c = &cd;
            h.Lh = cd.has_e1 ? dmax(cd.b, 0.0) : cd.b;
            h.Uh = cd.has_e1 ? cd.d_e1 : INFINITY;
            const unsigned int e = (unsigned int)u.r.pad;  // stream id
            bool skip = (final_evt && (A.flags & FLAG_NO_FINAL)) || (h.Uh - h.Lh < 1e-10) ||
                        (!final_evt && cd.d_e1 - a_e < 1e-10);  // qnet.pyx:361, 385, 440
            if (!skip) {
                Rng aux;
                aux.init(A.seed, gchain, gsw, e, TAG_MH_AUX);
                const double u_acc = aux.uniform();
                (void)aux.uniform();
                const double h0 = h(x0);
                double x_new = x0;
                if (!(h0 > BQ_NEG_INF)) {  // qnet.pyx:389-397, 444-451
                    int tries = 0;
                    for (; tries < 25; ++tries) {
                        const double uu = aux.uniform();
                        x_new = final_evt ? cd.b + (-log_d(1.0 - uu)) : a_e + (cd.d_e1 - a_e) * uu;
                        if (h(x_new) > BQ_NEG_INF) break;
                    }
                    skip = tries >= 25;
                }
                if (!skip) {
#pragma unroll 1
                    for (int step = 0; step < 2; ++step) {
                        Rng rng;
                        rng.init(A.seed, gchain, gsw, e, step == 0 ? TAG_MH : TAG_MH2);
                        x_new = final_evt ? slice_stepout_seq(h, x_new, h.Lh, h.Uh, rng, n_evals)
                                          : slice_finite_seq(h, x_new, h.Lh, h.Uh, rng, n_evals);
                    }
                    n_ev += 1;
                    const double ratio = exp(h0 - h(x_new) + (cd.eval(x_new) - cd.eval(x0)));
                    if (u_acc < ratio) d[u.r.e] = x_new;
                    else n_reject += 1;
                }
            }
        }
    }
}

template <int N>
__device__ __forceinline__ void block_reduce_sum(double (&v)[N], double *scratch, double *out) {
    // deterministic: warp shuffle tree, then warp 0 adds the per-warp partials in warp order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) scratch[warp * N + i] = x;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double x = 0.0;
        for (int w = 0; w < nwarps; ++w) x += scratch[w * N + threadIdx.x];
        out[threadIdx.x] = x;
    }
    __syncthreads();
}

__device__ __forceinline__ double block_reduce_max(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = scratch[0];
        for (int w = 1; w < nwarps; ++w) m = fmax(m, scratch[w]);
        scratch[32] = m;
    }
    __syncthreads();
    double m = scratch[32];
    __syncthreads();
    return m;
}

// service time of event i under the static order (queues.pyx:610-612 with k = 1; queues.pyx:771)
template <class DT>
__device__ __forceinline__ void event_service(const DT &d, int i, int fifo, int p, int pt, double &a,
                                              double &s) {
    double di = d[i];
    a = (pt >= 0) ? d[pt] : 0.0;
    double c = a;
    if (fifo && p >= 0) c = fmax(a, d[p]);
    s = di - c;
}

// Sufficient statistics per distribution SLOT for the whole chain (a plain queue has one slot; a mixture queue one
// per component, fed by the events currently assigned to it: MixtureTemplate.collect_sobs, queues.pyx:1416-1432);
// result in ss[NS] (shared), logp in *logp_out.  Two passes for LogNormal slots (distributions.pyx:276-299 is two-pass).
__device__ inline void chain_suffstats(const SweepArgs &A, const double *d, const int *q_type,
                                       const int *q_dist, const QConst *qc, SuffStat *ss, double *scratch,
                                       double *red_out, bool want_logp, double *logp_out, const int *q_slot,
                                       const unsigned char *aux) {
    double logp_acc = 0.0;
    for (int q = 0; q < A.Q; ++q) {
        const int lo = A.q_off[q], hi = A.q_off[q + 1];
        const int fifo = (q_type[q] == Q_GGK);
        const int s0 = aux ? q_slot[q] : q, ncomp = aux ? q_slot[q + 1] - q_slot[q] : 1;
        for (int comp = 0; comp < ncomp; ++comp) {
            const int sl = s0 + comp, dist = q_dist[sl];
            const bool need_log = (dist != D_EXP);
            double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // n, sum, sumlog, nlog, wait, nzero, logp, #(s < 0)
            for (int j = lo + threadIdx.x; j < hi; j += blockDim.x) {
                int i = A.q_events[j];
                if (ncomp > 1 && aux[i] != comp) continue;
                double a, s;
                event_service(d, i, fifo, A.ev_p[i], A.ev_pt[i], a, s);
                double w = d[i] - a - s;
                v[4] += (w > 0.0) ? w : 0.0;
                if (s > 0.0) {
                    v[0] += 1.0; v[1] += s;
                    if (need_log) {
                        double ls = log(s);
                        if (dist == D_GAMMA) { v[2] += ls; v[3] += 1.0; }
                        else if (s >= 1e-50) { v[2] += ls; v[3] += 1.0; }
                    }
                } else if (s == 0.0) v[5] += 1.0;
                else v[7] += 1.0;
                if (want_logp && !(s < 0.0)) v[6] += lpdf(qc[sl], s);
            }
            block_reduce_sum<8>(v, scratch, red_out);
            double mean_log = (red_out[3] > 0.0) ? red_out[2] / red_out[3] : 0.0;
            double sq[1] = {0.0};
            if (dist == D_LOGNORMAL) {
                for (int j = lo + threadIdx.x; j < hi; j += blockDim.x) {
                    int i = A.q_events[j];
                    if (ncomp > 1 && aux[i] != comp) continue;
                    double a, s;
                    event_service(d, i, fifo, A.ev_p[i], A.ev_pt[i], a, s);
                    if (s >= 1e-50) { double z = log(s) - mean_log; sq[0] += z * z; }
                }
            }
            if (threadIdx.x == 0) {
                SuffStat &t = ss[sl];
                t.n = red_out[0]; t.sum = red_out[1]; t.sumlog = red_out[2]; t.nlog = red_out[3];
                t.wait = red_out[4]; t.nzero = red_out[5]; t.nall = red_out[0] + red_out[5] + red_out[7]; t.sumsq = 0.0;
                logp_acc += (want_logp && red_out[7] > 0.0) ? BQ_NEG_INF : red_out[6];
            }
            __syncthreads();
            if (dist == D_LOGNORMAL) {
                block_reduce_sum<1>(sq, scratch, red_out);
                if (threadIdx.x == 0) ss[sl].sumsq = red_out[0];
                __syncthreads();
            }
        }
    }
    if (threadIdx.x == 0 && logp_out) *logp_out = logp_acc;
    __syncthreads();
}

// MixtureTemplate.resample_auxiliaries for every event of every mixture queue (qnet.pyx:679-680 -> queues.pyx:1402-1410):
// z_e ~ Cat(pi_i exp(lpdf_i(s_e))), one uniform of the (chain, sweep, event) stream TAG_AUX; when all weights vanish
// (_roll_die_unnorm, sampling.pyx:215-229: Z < 1e-50) a uniformly random component.
__device__ inline void resample_auxiliaries(const SweepArgs &A, const double *d, const int *q_type, const int *q_slot,
                                            const QConst *qc, const double *th, unsigned char *aux,
                                            unsigned long long gchain, unsigned int gsw) {
    for (int q = 0; q < A.Q; ++q) {
        const int ncomp = q_slot[q + 1] - q_slot[q];
        if (ncomp < 2) continue;
        const int lo = A.q_off[q], hi = A.q_off[q + 1], fifo = (q_type[q] == Q_GGK), s0 = q_slot[q];
        const double *mix = th + A.q_mixoff[q];
        for (int j = lo + threadIdx.x; j < hi; j += blockDim.x) {
            const int i = A.q_events[j];
            double a, s;
            event_service(d, i, fifo, A.ev_p[i], A.ev_pt[i], a, s);
            double Z = 0.0;
            for (int c = 0; c < ncomp; ++c) Z += mix[c] * exp(lpdf(qc[s0 + c], s));
            Rng rng;
            rng.init(A.seed, gchain, gsw, (unsigned int)i, TAG_AUX);
            const double u = rng.uniform();
            int pick = ncomp - 1;
            if (Z < 1e-50) {
                pick = (int)(u * ncomp);
                if (pick >= ncomp) pick = ncomp - 1;
            } else {
                const double t = Z * u;
                double S = 0.0;
                for (int c = 0; c < ncomp; ++c) {
                    S += mix[c] * exp(lpdf(qc[s0 + c], s));
                    if (t < S) { pick = c; break; }
                }
            }
            aux[i] = (unsigned char)pick;
        }
    }
    __syncthreads();
}

template <bool SMEM, bool ALL_EXP>
__global__ void __launch_bounds__(BQ_MAX_THREADS, 1) sweep_static_kernel(const SweepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int chain = blockIdx.x;
    const int E = A.E, Q = A.Q, P = A.P;
    double *dg = A.state + (size_t)chain * E;
    // shared layout: [d (E padded to even)] [scratch 34*8] [red_out 8] [theta P] [qc NS] [ss NS] [ints Q + 2 NS + Q + 1]
    const int NS = A.NS;
    double *sm = reinterpret_cast<double *>(smem_raw);
    const int Epad = SMEM ? ((E + 1) & ~1) : 0;
    double *d = SMEM ? sm : dg;
    double *scratch = sm + Epad;            // 16 warps * 8 + 2 slack -> 34*8 is ample
    double *red_out = scratch + 34 * 8;
    double *th = red_out + 8;
    QConst *qc = reinterpret_cast<QConst *>(th + ((P + 1) & ~1));
    SuffStat *ss = reinterpret_cast<SuffStat *>(qc + NS);
    int *q_type = reinterpret_cast<int *>(ss + NS);
    int *q_dist = q_type + Q;      // per slot
    int *q_poff = q_dist + NS;     // per slot
    int *q_slot = q_poff + NS;     // [Q + 1]
    // the all-exponential build never carries mixtures (bq_create): a compile-time null keeps its code as it was
    unsigned char *aux = (!ALL_EXP && A.aux) ? A.aux + (size_t)chain * E : nullptr;
    __shared__ unsigned long long cnt[5];
    __shared__ double logp_sh;
    __shared__ int work_ctr[2];

    for (int i = threadIdx.x; i < P; i += blockDim.x) th[i] = A.theta[(size_t)chain * P + i];
    for (int i = threadIdx.x; i < Q; i += blockDim.x) q_type[i] = A.q_type[i];
    for (int i = threadIdx.x; i <= Q; i += blockDim.x) q_slot[i] = A.q_slot[i];
    for (int i = threadIdx.x; i < NS; i += blockDim.x) { q_dist[i] = A.q_dist[i]; q_poff[i] = A.q_poff[i]; }
    if (threadIdx.x < 5) cnt[threadIdx.x] = 0ull;
    if (SMEM) {
        // coalesced 16-byte loads; each chain's row is 16 B aligned when E is even, else fall back
        if ((E & 1) == 0) {
            const double2 *src = reinterpret_cast<const double2 *>(dg);
            double2 *dst = reinterpret_cast<double2 *>(d);
            for (int i = threadIdx.x; i < E / 2; i += blockDim.x) dst[i] = src[i];
        } else {
            for (int i = threadIdx.x; i < E; i += blockDim.x) d[i] = dg[i];
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < NS; q += blockDim.x) {  // slots may exceed the CTA (bq_create takes up to 1024 queues)
        const int o = q_poff[q];
        qconst_set(qc[q], q_dist[q], th[o], (q_dist[q] == D_EXP) ? 0.0 : th[o + 1]);
    }
    __syncthreads();

    const unsigned long long gchain = (unsigned long long)(A.chain_offset + chain);
    long long n_ev = 0, n_evals = 0, n_stuck = 0, n_pevals = 0, n_reject = 0;
    double tmax = 0.0;

    for (int sw = 0; sw < A.n_sweeps; ++sw) {
        const unsigned int gsw = A.sweep_base + (unsigned int)sw;
        if (sw == 0 || (A.flags & FLAG_TMAX_PER_SWEEP)) {  // Tmax = 2 max d   (qnet.pyx:672)
            double m = 0.0;
            for (int i = threadIdx.x; i < E; i += blockDim.x) m = fmax(m, d[i]);
            tmax = 2.0 * block_reduce_max(m, scratch);
        }
        const int n_colors = (A.flags & FLAG_PARAMS_ONLY) ? 0 : A.n_colors;
        if (aux && n_colors > 0)  // mixture components first (qnet.pyx:679-680)
            resample_auxiliaries(A, d, q_type, q_slot, qc, th, aux, gchain, gsw);
        // reverse scan (slice_resample(reverse=True), qnet.pyx:683-684): the classes in reverse order; the events of a
        // class are conditionally independent, so the order inside a class is immaterial
        const bool rev = (A.flags & FLAG_REVERSE) != 0;
        if (threadIdx.x == 0 && n_colors > 0) work_ctr[0] = A.color_off[rev ? n_colors - 1 : 0];
        __syncthreads();
        for (int c = 0; c < n_colors; ++c) {
            const int cc = rev ? n_colors - 1 - c : c;
            const int hi = A.color_off[cc + 1];
            if (threadIdx.x == 0 && c + 1 < n_colors)  // where the next class starts
                work_ctr[(c + 1) & 1] = A.color_off[rev ? cc - 1 : cc + 1];
            if (A.mode == MODE_MH)
                run_class_mh(A, d, q_type, qc, hi, &work_ctr[c & 1], gchain, gsw, n_ev, n_evals, n_reject, q_slot, aux);
            else if (ALL_EXP)
                run_class<CondExp>(A, d, q_type, qc, hi, &work_ctr[c & 1], tmax, gchain, gsw, n_ev, n_evals, n_stuck,
                                   q_slot, nullptr);
            else
                run_class<CondGen>(A, d, q_type, qc, hi, &work_ctr[c & 1], tmax, gchain, gsw, n_ev, n_evals, n_stuck,
                                   q_slot, aux);
            __syncthreads();
        }
        const bool trace = (A.flags & FLAG_TRACE) != 0;
        if (A.update != UPD_NONE || trace) {
            chain_suffstats(A, d, q_type, q_dist, qc, ss, scratch, red_out, trace, &logp_sh, q_slot, aux);
            const size_t recno = (size_t)(A.trace_base + sw);
            if (trace)
                for (int q = threadIdx.x; q < Q; q += blockDim.x) {
                    double w = 0.0, na = 0.0;
                    for (int sl = q_slot[q]; sl < q_slot[q + 1]; ++sl) { w += ss[sl].wait; na += ss[sl].nall; }
                    A.tr_wait[(recno * A.n_chains + chain) * Q + q] = (na > 0.0) ? w / na : 0.0;
                }
            if (trace && threadIdx.x == 0) A.tr_logp[recno * A.n_chains + chain] = logp_sh;
            for (int q = threadIdx.x; A.update != UPD_NONE && q < NS; q += blockDim.x) {  // one update per slot
                const int o = q_poff[q], dist = q_dist[q];
                double p0 = th[o], p1 = (dist == D_EXP) ? 0.0 : th[o + 1];
                if (A.update == UPD_BAYES) {
                    Rng rng;
                    rng.init(A.seed, gchain, gsw, (unsigned int)q, TAG_PARAM);
                    sample_params(dist, ss[q], rng, p0, p1, &n_pevals);
                } else {
                    estimate_params(dist, ss[q], p0, p1);
                }
                th[o] = p0;
                if (dist != D_EXP) th[o + 1] = p1;
                qconst_set(qc[q], dist, p0, p1);
            }
            // mixing proportions: (1 + #events of the component) / (#components + #events), the same rule in estimate
            // and in sample_parameters (queues.pyx:1416-1451: counts start at 1, N at the number of components)
            for (int q = threadIdx.x; aux && A.update != UPD_NONE && q < Q; q += blockDim.x) {
                const int mo = A.q_mixoff[q], s0 = q_slot[q], nc = q_slot[q + 1] - s0;
                if (mo < 0) continue;
                double N = (double)nc;
                for (int c = 0; c < nc; ++c) N += ss[s0 + c].nall;
                for (int c = 0; c < nc; ++c) th[mo + c] = (1.0 + ss[s0 + c].nall) / N;
            }
            __syncthreads();
            if (trace)
                for (int i = threadIdx.x; i < P; i += blockDim.x)
                    A.tr_theta[(recno * A.n_chains + chain) * P + i] = th[i];
        }
    }

    if (SMEM) {
        __syncthreads();
        if ((E & 1) == 0) {
            double2 *dst = reinterpret_cast<double2 *>(dg);
            const double2 *src = reinterpret_cast<const double2 *>(d);
            for (int i = threadIdx.x; i < E / 2; i += blockDim.x) dst[i] = src[i];
        } else {
            for (int i = threadIdx.x; i < E; i += blockDim.x) dg[i] = d[i];
        }
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x) A.theta[(size_t)chain * P + i] = th[i];
    atomicAdd(&cnt[0], (unsigned long long)n_ev);
    atomicAdd(&cnt[1], (unsigned long long)n_evals);
    atomicAdd(&cnt[2], (unsigned long long)n_pevals);
    atomicAdd(&cnt[3], (unsigned long long)n_stuck);
    atomicAdd(&cnt[4], (unsigned long long)n_reject);
    __syncthreads();
    if (threadIdx.x < 5) A.stats[(size_t)chain * BQ_NSTAT + threadIdx.x] += (long long)cnt[threadIdx.x];
}

// ---- per-chain starting states -------------------------------------------------------------------
// The reference initialises ONE state (gibbs_initialize_via_dfn, qnet.pyx:1454-1554: every hidden task threaded through
// the network with departures drawn from exponentials fitted to the observed data, accepted while all services stay
// non-negative).  Many chains want many starts, and here all chains of a handle share the queue order, so the batched
// version draws, for every chain on its own RNG stream, a state that keeps that order: the resample-eligible events
// are visited in a topological order of the constraint graph (a_e <= d_e, FIFO departure order, FIFO arrival order --
// the lower / upper neighbours are the ones SchedRec lists), each departure is drawn as lower bound + Exp(mean
// service of its queue x the chain's dispersion factor), truncated to the upper bound that the OBSERVED times impose
// on it through its descendants (ub[], the same for all chains, computed once on the host).  One thread per chain.
struct InitArgs {
    int E, Q, P, n_chains, n;      // n = eligible events
    const int4 *rec;               // SchedRec in schedule order
    const int *topo;               // [n] schedule indices in topological order
    const double *ub;              // [n] by schedule index
    const int *q_type, *q_dist, *q_poff, *q_slot;  // q_dist / q_poff per slot (SweepArgs); the first slot of a queue is used
    double *state;                 // [C][E]
    const double *theta;           // [C][P]
    double sigma;                  // log-sd of the per-chain dispersion factor
    unsigned long long seed;
    long long chain_offset;
    unsigned int counter;
};

static __global__ void init_chains_kernel(const InitArgs A) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= A.n_chains) return;
    double *d = A.state + (size_t)chain * A.E;
    const double *th = A.theta + (size_t)chain * A.P;
    const unsigned long long gchain = (unsigned long long)(A.chain_offset + chain);
    Rng rc;
    rc.init(A.seed, gchain, A.counter, 0xffffffffu, TAG_INIT);
    const double kappa = exp(A.sigma * draw_normal(rc));  // one dispersion factor per chain
    for (int i = 0; i < A.n; ++i) {
        const int k = A.topo[i];
        union { int4 v[3]; SchedRec r; } u;
        u.v[0] = __ldg(&A.rec[3 * k]); u.v[1] = __ldg(&A.rec[3 * k + 1]); u.v[2] = __ldg(&A.rec[3 * k + 2]);
        const SchedRec &r = u.r;
        double lb = (r.pt_e >= 0) ? d[r.pt_e] : 0.0;
        if (A.q_type[r.q0] == Q_GGK && r.p >= 0) lb = fmax(lb, d[r.p]);
        if (r.e1 >= 0 && A.q_type[r.q1] == Q_GGK && r.pt_p1 >= 0) lb = fmax(lb, d[r.pt_p1]);
        const double ub = A.ub[k];
        const int sl = A.q_slot[r.q0], o = A.q_poff[sl], dist = A.q_dist[sl];
        double mean = th[o];                                       // Exponential: scale
        if (dist == D_GAMMA) mean = th[o] * th[o + 1];
        else if (dist == D_LOGNORMAL) mean = exp(th[o] + 0.5 * th[o + 1] * th[o + 1]);
        mean *= kappa;
        Rng rng;
        rng.init(A.seed, gchain, A.counter, (unsigned int)r.e, TAG_INIT);
        double x = lb;
        bool ok = false;
        for (int t = 0; t < 8 && !ok; ++t) {
            x = lb + mean * (-log_d(1.0 - rng.uniform()));
            ok = x < ub;
        }
        if (!ok) x = lb + (ub - lb) * rng.uniform();               // tight window: uniform inside it
        d[r.e] = x;
    }
}

// ---- parity / reporting kernels ---------------------------------------------------------------

// out[j] = inner_dfun(x[j]) - lik0 for event described by `r` on chain `chain` (qnet.pyx:1340-1357).
static __global__ void logcond_static_kernel(const SweepArgs A, int chain, SchedRec r, const double *x, int n,
                                      double *out) {
    __shared__ QConst qc[3];  // service densities of e, of its queue successor, of the task's next event
    __shared__ int qt[2];
    const double *d = A.state + (size_t)chain * A.E;
    const double *th = A.theta + (size_t)chain * A.P;
    const unsigned char *aux = A.aux ? A.aux + (size_t)chain * A.E : nullptr;
    if (threadIdx.x == 0) {
        const int ev[3] = {r.e, r.n, r.e1}, qq[3] = {r.q0, r.q0, r.q1};
        for (int k = 0; k < 3; ++k) {
            if (ev[k] < 0) { qc[k] = qc[0]; continue; }
            const int sl = A.q_slot[qq[k]] + (aux ? aux[ev[k]] : 0), o = A.q_poff[sl], dist = A.q_dist[sl];
            qconst_set(qc[k], dist, th[o], dist == D_EXP ? 0.0 : th[o + 1]);
        }
        qt[0] = A.q_type[r.q0];
        qt[1] = (r.e1 >= 0) ? A.q_type[r.q1] : Q_DELAY;
    }
    __syncthreads();
    SchedRec rr = r;
    rr.q0 = 0; rr.q1 = 1;
    double tmax = INFINITY;
    Cond cd;
    double x0, lower, upper;
    cond_load(cd, rr, d, qt, qc, x0, lower, upper, tmax);
    cd.c0 = &qc[0]; cd.cn = &qc[1]; cd.c1 = &qc[2];
    double g0 = cd.eval(x0);
    for (int j = threadIdx.x; j < n; j += blockDim.x) out[j] = cd.eval(x[j]) - g0;
}

// Qnet.log_prob (qnet.pyx:1032) and the fused sufficient statistics (per slot), one CTA per chain.
static __global__ void __launch_bounds__(256) report_static_kernel(const SweepArgs A, double *logp_out,
                                                             double *ss_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int chain = blockIdx.x, Q = A.Q, NS = A.NS;
    const double *d = A.state + (size_t)chain * A.E;
    double *sm = reinterpret_cast<double *>(smem_raw);
    double *scratch = sm;
    double *red_out = scratch + 34 * 8;
    QConst *qc = reinterpret_cast<QConst *>(red_out + 8);
    SuffStat *ss = reinterpret_cast<SuffStat *>(qc + NS);
    int *q_type = reinterpret_cast<int *>(ss + NS);
    int *q_dist = q_type + Q;
    int *q_slot = q_dist + NS;
    const unsigned char *aux = A.aux ? A.aux + (size_t)chain * A.E : nullptr;
    __shared__ double logp_sh;
    for (int q = threadIdx.x; q < Q; q += blockDim.x) q_type[q] = A.q_type[q];
    for (int q = threadIdx.x; q <= Q; q += blockDim.x) q_slot[q] = A.q_slot[q];
    for (int q = threadIdx.x; q < NS; q += blockDim.x) {
        q_dist[q] = A.q_dist[q];
        int o = A.q_poff[q];
        const double *th = A.theta + (size_t)chain * A.P;
        qconst_set(qc[q], q_dist[q], th[o], q_dist[q] == D_EXP ? 0.0 : th[o + 1]);
    }
    __syncthreads();
    chain_suffstats(A, d, q_type, q_dist, qc, ss, scratch, red_out, true, &logp_sh, q_slot, aux);
    if (threadIdx.x == 0 && logp_out) logp_out[chain] = logp_sh;
    if (ss_out)
        for (int i = threadIdx.x; i < NS * 8; i += blockDim.x)
            ss_out[(size_t)chain * NS * 8 + i] = reinterpret_cast<double *>(ss)[i];
}

}  // namespace bq
