// Exact Gibbs transition density: Qnet.gibbs_kernel (qnet.pyx:741-809), the quantity estimation.chib_evidence
// (estimation.py:274-330) averages.  For every chain, log T(d_to <- current state) of one slice sweep in the handle's
// scan order: event by event, minus the log of the normaliser Z_e = integral of exp(g(x) - g(d_to[e])) dx over the
// support of the conditional g (GGkGibbs.integrate, qnet.pyx:1361-1372), then d[e] := d_to[e].  Events whose target
// value is not reachable yet are deferred to a later pass, as the reference does (qnet.pyx:756-805) -- or, with
// `strict`, make the result -inf: that is the density of the fixed-scan sweep the kernels actually run, the one for
// which Chib's identity holds (tests/test_gpu_chib.py; the deferred passes bias the estimate upwards).
//
// Static-order networks (single-server FIFO queues and delay stations): the conditional has the closed form of
// bq_static.cuh (three terms, kinks at a_n and d_p1), so the events of a colour class are independent and are
// integrated side by side; the chain's state is a scratch copy, the handle's own state is not touched.
// Dynamic-order networks (G/G/k, PS): gibbs_kernel_dynamic at the end of this file.
#pragma once
#include "bq_static.cuh"
#include "bq_dynamic.cuh"

namespace bq {

// 7-point Gauss / 15-point Kronrod pair on [-1, 1] (QUADPACK qk15: the rule scipy.integrate.quad starts from)
__constant__ double c_xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                                0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                                0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
__constant__ double c_wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                                0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                                0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
__constant__ double c_wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                               0.381830050505118944950369775488975, 0.417959183673469387755102040816327};

// exp(g(x) - lf) on [a, b] by the 15-point rule; *err = |K15 - G7|
__device__ __forceinline__ double gk15(const Cond &cd, double lf, double a, double b, double *err) {
    const double c = 0.5 * (a + b), h = 0.5 * (b - a);
    const double fc = exp(cd.eval(c) - lf);
    double rk = c_wgk[7] * fc, rg = c_wg[3] * fc;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const double dx = h * c_xgk[j];
        const double f1 = exp(cd.eval(c - dx) - lf), f2 = exp(cd.eval(c + dx) - lf);
        rk += c_wgk[j] * (f1 + f2);
        if (j & 1) rg += c_wg[j >> 1] * (f1 + f2);
    }
    *err = fabs((rk - rg) * h);
    return rk * h;
}

// adaptive bisection of one smooth piece.  An interval is accepted when its error is small against its own value
// or against its share of the whole piece (first estimate, then the running total): the second test stops the
// bisection in the far tails of LogNormal services, which are tiny but never smooth relative to themselves.
__device__ double integrate_piece(const Cond &cd, double lf, double a, double b) {
    const double tol = 1e-11;
    double sa[48], sb[48];
    int top = 0;
    sa[0] = a; sb[0] = b; top = 1;
    double total = 0.0, err0;
    double whole = fabs(gk15(cd, lf, a, b, &err0));
    if (!(whole == whole)) whole = 0.0;
    const double inv_w = 1.0 / (b - a);
    while (top > 0) {
        --top;
        const double lo = sa[top], hi = sb[top];
        double err;
        const double v = gk15(cd, lf, lo, hi, &err);
        const double scale = fmax(whole, total) * (hi - lo) * inv_w;
        const bool fine = err <= tol * fmax(fabs(v), scale) || err < 1e-300 ||
                          (hi - lo) <= 1e-13 * fmax(fabs(lo), fabs(hi)) || !(v == v) || top + 2 > 48;
        if (fine) { if (v == v) total += v; }
        else {
            const double m = 0.5 * (lo + hi);
            sa[top] = m; sb[top] = hi; ++top;  // right half is popped last: left to right accumulation
            sa[top] = lo; sb[top] = m; ++top;
        }
    }
    return total;
}

// Z = integral over the support of exp(g(x) - lf)
__device__ double cond_normaliser(const Cond &cd, double lf) {
    if (!(cd.fhi < INFINITY)) {
        // no successor in the queue and no later visit: g(x) = lpdf(x - b) on [b, inf), a density that integrates to 1
        return exp(-lf);
    }
    double pts[4];
    int n = 0;
    pts[n++] = cd.flo;
    double k0 = cd.has_n ? cd.a_n : cd.flo, k1 = (cd.has_e1 && cd.e1_fifo) ? cd.d_p1 : cd.flo;
    if (k0 > k1) { const double t = k0; k0 = k1; k1 = t; }
    if (k0 > cd.flo && k0 < cd.fhi) pts[n++] = k0;
    if (k1 > pts[n - 1] && k1 < cd.fhi) pts[n++] = k1;
    pts[n++] = cd.fhi;
    double z = 0.0;
    for (int i = 0; i + 1 < n; ++i)
        if (pts[i + 1] > pts[i]) z += integrate_piece(cd, lf, pts[i], pts[i + 1]);
    return z;
}

// One CTA per chain.  work: [C][E] scratch copy of the state; done: [C][n_sched] flags (zeroed by the caller);
// d_to: [n_to][E] with n_to = 1 (one target for all chains) or C; out: [C].
__global__ void __launch_bounds__(256) gibbs_kernel_static(const SweepArgs A, double *work, unsigned char *done_all,
                                                           const double *d_to_all, int n_to, int strict, double *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int chain = blockIdx.x, E = A.E, Q = A.Q, P = A.P;
    double *d = work + (size_t)chain * E;
    const double *d_to = d_to_all + (size_t)(n_to > 1 ? chain : 0) * E;
    double *sm = reinterpret_cast<double *>(smem_raw);
    double *scratch = sm;  // 34 * 8
    double *red_out = scratch + 34 * 8;
    double *th = red_out + 8;
    QConst *qc = reinterpret_cast<QConst *>(th + ((P + 1) & ~1));
    int *q_type = reinterpret_cast<int *>(qc + Q);
    __shared__ int s_progress, s_left;
    for (int i = threadIdx.x; i < P; i += blockDim.x) th[i] = A.theta[(size_t)chain * P + i];
    for (int i = threadIdx.x; i < Q; i += blockDim.x) q_type[i] = A.q_type[i];
    __syncthreads();
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
        const int o = A.q_poff[q], dist = A.q_dist[q];
        qconst_set(qc[q], dist, th[o], (dist == D_EXP) ? 0.0 : th[o + 1]);
    }
    const int n_sched = A.color_off[A.n_colors];
    unsigned char *done = done_all + (size_t)chain * n_sched;
    if (threadIdx.x == 0) { s_left = n_sched; }
    __syncthreads();
    double acc = 0.0;
    bool dead = false;
    for (int pass = 0; pass <= n_sched; ++pass) {
        if (threadIdx.x == 0) s_progress = 0;
        __syncthreads();
        for (int c = 0; c < A.n_colors; ++c) {
            const int lo = A.color_off[c], hi = A.color_off[c + 1];
            // every thread first evaluates its events on the state before the class, then the class is committed
            for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
                if (done[i]) continue;
                const int4 r4 = A.rec[3 * (size_t)i], r4b = A.rec[3 * (size_t)i + 1], r4c = A.rec[3 * (size_t)i + 2];
                SchedRec r;
                r.e = r4.x; r.pt_e = r4.y; r.p = r4.z; r.n = r4.w;
                r.pt_n = r4b.x; r.e1 = r4b.y; r.p1 = r4b.z; r.pt_p1 = r4b.w;
                r.pt_n1 = r4c.x; r.q0 = r4c.y; r.q1 = r4c.z; r.pad = r4c.w;
                Cond cd;
                double x0, lower, upper;
                cond_load(cd, r, d, q_type, qc, x0, lower, upper, INFINITY);
                const double x = d_to[r.e];
                const double lf = cd.eval(x);
                if (x > lower && x < upper && lf > BQ_NEG_INF) {  // reachable now (qnet.pyx:772-781)
                    const double z = cond_normaliser(cd, lf);
                    acc -= log(z);
                    done[i] = 2;  // to be committed at the end of the class
                }
            }
            __syncthreads();
            int mine = 0;
            for (int i = lo + threadIdx.x; i < hi; i += blockDim.x)
                if (done[i] == 2) { const int e = A.rec[3 * (size_t)i].x; d[e] = d_to[e]; done[i] = 1; ++mine; }
            if (mine) atomicAdd(&s_progress, mine);
            __syncthreads();
        }
        if (threadIdx.x == 0) s_left -= s_progress;
        __syncthreads();
        const int left = s_left, prog = s_progress;
        __syncthreads();
        if (left == 0) break;
        // no event of the pass could move: the target is unreachable.  strict: no second pass at all -- an event
        // whose target is outside its support when its turn comes makes the transition impossible
        if (prog == 0 || strict) { dead = true; break; }
    }
    double v[1] = {acc};
    block_reduce_sum<1>(v, scratch, red_out);
    if (threadIdx.x == 0) out[chain] = dead ? BQ_NEG_INF : red_out[0];
}

// ---------------------------------------------------------------------------------------------------------------
// Dynamic-order networks (G/G/k with k > 1, PS): the conditional is only available as a function (dyn_logcond), its
// kinks are not known in advance, and the events of a chain must be taken one after another.  One warp per chain:
// 15 lanes evaluate the 15 Kronrod nodes of an interval side by side, the bisection is driven uniformly by the warp,
// the event is then committed with the sweep kernels' own dyn_commit.  The range is first cut where the reference
// cuts it (the times of the neighbours in both queues, qnet.pyx:1366-1388); the kernel works IN PLACE on the chain's
// arrays -- the caller snapshots and restores them.
// ---------------------------------------------------------------------------------------------------------------
__constant__ double c_x15[15] = {-0.991455371120812639206854697526329, -0.949107912342758524526189684047851,
                                 -0.864864423359769072789712788640926, -0.741531185599394439863864773280788,
                                 -0.586087235467691130294144838258730, -0.405845151377397166906606412076961,
                                 -0.207784955007898467600689403773245, 0.0,
                                 0.207784955007898467600689403773245,  0.405845151377397166906606412076961,
                                 0.586087235467691130294144838258730,  0.741531185599394439863864773280788,
                                 0.864864423359769072789712788640926,  0.949107912342758524526189684047851,
                                 0.991455371120812639206854697526329};
__constant__ double c_wk15[15] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                                  0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                                  0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                  0.204432940075298892414161999234649, 0.209482141084727828012999174891714,
                                  0.204432940075298892414161999234649, 0.190350578064785409913256402421014,
                                  0.169004726639267902826583426598550, 0.140653259715525918745189590510238,
                                  0.104790010322250183839876322541518, 0.063092092629978553290700663189204,
                                  0.022935322010529224963732008058970};
__constant__ double c_wg15[15] = {0.0, 0.129484966168869693270611432679082, 0.0, 0.279705391489276667901467771423780,
                                  0.0, 0.381830050505118944950369775488975, 0.0, 0.417959183673469387755102040816327,
                                  0.0, 0.381830050505118944950369775488975, 0.0, 0.279705391489276667901467771423780,
                                  0.0, 0.129484966168869693270611432679082, 0.0};

// the 15-point rule on [a, b] by the warp: returns the Kronrod value, *err = |K15 - G7|; uniform across the lanes
template <int KMAX>
__device__ __forceinline__ double warp_gk15(const Chain &c, const EventCtx &ev, double lf, double a, double b, int lane,
                                            double *err) {
    const double mid = 0.5 * (a + b), h = 0.5 * (b - a);
    double fk = 0.0, fg = 0.0;
    if (lane < 15) {
        Touch tc;
        tc.reset();
        const double f = exp(dyn_logcond<KMAX>(c, ev, mid + h * c_x15[lane], tc) - lf);
        if (f == f) { fk = c_wk15[lane] * f; fg = c_wg15[lane] * f; }
    }
    fk = group_sum<32>(fk, 0xffffffffu);
    fg = group_sum<32>(fg, 0xffffffffu);
    *err = fabs((fk - fg) * h);
    return fk * h;
}

template <int KMAX>
__device__ double warp_integrate_piece(const Chain &c, const EventCtx &ev, double lf, double a, double b, int lane,
                                       double z_so_far) {
    const double tol = 1e-10;
    double sa[48], sb[48];
    int top = 1;
    sa[0] = a; sb[0] = b;
    double total = 0.0, err0;
    const double whole = warp_gk15<KMAX>(c, ev, lf, a, b, lane, &err0);
    const double inv_w = 1.0 / (b - a);
    while (top > 0) {
        --top;
        const double lo = sa[top], hi = sb[top];
        double err;
        const double v = warp_gk15<KMAX>(c, ev, lf, lo, hi, lane, &err);
        const double scale = fmax(fmax(whole, total), z_so_far) * (hi - lo) * inv_w;
        const bool fine = err <= tol * fmax(fabs(v), scale) || err < 1e-300 ||
                          (hi - lo) <= 1e-13 * fmax(fabs(lo), fabs(hi)) || top + 2 > 48;
        if (fine) total += v;
        else {
            const double m = 0.5 * (lo + hi);
            sa[top] = m; sb[top] = hi; ++top;
            sa[top] = lo; sb[top] = m; ++top;
        }
    }
    return total;
}

template <int KMAX>
__global__ void __launch_bounds__(128) gibbs_kernel_dynamic(const DynArgs A, const double *d_to_all, int n_to,
                                                            int strict, unsigned char *done_all, double *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int gib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + gib;
    if (chain >= A.n_chains) return;
    const unsigned FULL = 0xffffffffu;
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
    const double *d_to = d_to_all + (size_t)(n_to > 1 ? chain : 0) * A.E;
    unsigned char *done = done_all + (size_t)chain * A.n_order;
    const int4 *order4 = reinterpret_cast<const int4 *>(A.order);
    double acc = 0.0;
    bool dead = false;
    int left = A.n_order;
    for (int pass = 0; pass <= A.n_order && left > 0 && !dead; ++pass) {
        int progress = 0;
        for (int idx = 0; idx < A.n_order; ++idx) {
            if (done[idx]) continue;
            const int4 r4 = __ldg(&order4[idx]);
            const DynRec rec = {r4.x, r4.y, r4.z, r4.w};
            EventCtx ev;
            event_load(c, rec, ev);
            const double x = d_to[ev.e], lower = ev.a_e, upper = (ev.e1 >= 0) ? ev.d1 : INFINITY;
            Touch tc;
            tc.reset();
            const double lf = dyn_logcond<KMAX>(c, ev, x, tc);
            if (!(x > lower && x < upper && lf > BQ_NEG_INF)) continue;  // not reachable now (qnet.pyx:772-781)
            // cut points: the neighbours' times in both queues (uniform across the lanes)
            double pts[16];
            int n = 0;
            pts[n++] = lower;
            {
                const QInfo &q0 = c.qi[ev.q0];
                for (int p = ev.p0 - 2; p <= ev.p0 + 3; ++p)
                    if (p >= q0.lo && p < q0.hi) { const double t = c.dq(p); if (t > lower && t < upper) pts[n++] = t; }
                if (ev.e1 >= 0) {
                    const QInfo &q1 = c.qi[ev.q1];
                    for (int p = ev.p1 - 2; p <= ev.p1 + 3; ++p)
                        if (p >= q1.lo && p < q1.hi) { const double t = c.at(p); if (t > lower && t < upper) pts[n++] = t; }
                }
            }
            if (upper < INFINITY) pts[n++] = upper;
            for (int i = 1; i < n; ++i) {  // insertion sort
                const double t = pts[i];
                int j = i - 1;
                while (j >= 0 && pts[j] > t) { pts[j + 1] = pts[j]; --j; }
                pts[j + 1] = t;
            }
            double z = 0.0;
            for (int i = 0; i + 1 < n; ++i)
                if (pts[i + 1] > pts[i]) z += warp_integrate_piece<KMAX>(c, ev, lf, pts[i], pts[i + 1], lane, z);
            if (!(upper < INFINITY)) {  // final event: doubling steps until the tail no longer matters
                double t0 = pts[n - 1], w = fmax(1.0, fabs(ev.x0 - lower));
                for (int i = 0; i < 200; ++i) {
                    const double part = warp_integrate_piece<KMAX>(c, ev, lf, t0, t0 + w, lane, z);
                    z += part;
                    t0 += w;
                    w *= 2.0;
                    if (i > 2 && part <= 1e-15 * z) break;
                }
            }
            acc -= log(z);
            __syncwarp();
            if (x != ev.x0) dyn_commit<32, KMAX>(c, ev, x, lane, FULL);
            __syncwarp();
            if (lane == 0) done[idx] = 1;
            ++progress;
            --left;
        }
        __syncwarp();
        if (left > 0 && (progress == 0 || strict)) dead = true;
    }
    if (lane == 0) out[chain] = dead ? BQ_NEG_INF : acc;
}

}  // namespace bq
