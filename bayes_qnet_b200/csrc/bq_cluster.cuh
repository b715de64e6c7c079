// The chromatic sweep on a THREAD-BLOCK CLUSTER: one chain owned by N CTAs (N = 2, 4, 8) on N SMs of one GPC, its
// departure times partitioned over the N shared memories and read across CTAs through distributed shared memory
// (mapa + generic loads), the colour-class barrier a cluster barrier.  Two uses:
//   * few chains (BASELINE configs[0] and the small end of configs[4]): a single chain no longer runs on 1/148 of the
//     GPU -- its colour class is spread over N x 512 lanes;
//   * large static-order networks: the state of one chain no longer has to fit ONE CTA's shared memory (about 27 000
//     events); with N CTAs it stays on chip up to N times that instead of falling back to the global-memory path.
// Same algorithm, same RNG streams (keyed by chain, sweep, event), same schedule as sweep_static_kernel: the slice
// updates are identical whatever N is; only the order in which the sufficient statistics are summed differs (per
// CTA, then over the CTAs in rank order).  Reuses run_class / run_class_mh / Cond* from bq_static.cuh through the
// state accessor they are templated on.  Networks with mixture templates use the single-CTA kernel.
#pragma once
#include <cooperative_groups.h>

#include "bq_static.cuh"

namespace bq {

namespace cg = cooperative_groups;

// Event indices are re-encoded for this kernel as (owner rank << k) | offset in the owner's partition (host:
// build_cluster_plan), so that an access is a shift, a mask and one mapa.
struct DsmemView {
    double *local;       // this CTA's partition in shared memory
    unsigned k, mask;
    __device__ __forceinline__ double &operator[](int idx) const {
        const unsigned r = (unsigned)idx >> k, o = (unsigned)idx & mask;
        return *cg::this_cluster().map_shared_rank(local + o, r);
    }
};

struct ClusterArgs {
    SweepArgs A;            // rec = re-encoded records sorted by (class, owner rank); q_events / ev_* unused
    int ncta, part, k;      // CTAs per chain, events per partition (even), log2 of the encoding stride
    const int *class_off;   // [n_colors][ncta + 1] record ranges per class and rank
    const int *qe_idx, *qe_p, *qe_pt;  // [E] per queue (ranges q_off): encoded event, queue predecessor, task predecessor
};

// sum over the cluster of v[0..N) held by every thread of every CTA: block reduction, then the per-CTA partials are
// added in rank order by every CTA (identical result everywhere).  part_sh: [8 ranks][N] doubles in shared memory.
template <int N>
__device__ __forceinline__ void cluster_reduce_sum(double (&v)[N], double *scratch, double *red_out, double *part_sh,
                                                   int rank, int ncta) {
    block_reduce_sum<N>(v, scratch, red_out);
    cg::cluster_group cl = cg::this_cluster();
    if (threadIdx.x < N) part_sh[rank * N + threadIdx.x] = red_out[threadIdx.x];  // own slot of the local table
    cl.sync();
    if (threadIdx.x < N) {
        double x = 0.0;
        for (int r = 0; r < ncta; ++r) x += cl.map_shared_rank(part_sh, r)[r * N + threadIdx.x];
        red_out[threadIdx.x] = x;
    }
    cl.sync();  // partials may be overwritten by the next reduction only after every CTA has read them
}

template <bool ALL_EXP>
__global__ void __launch_bounds__(BQ_MAX_THREADS, 1) sweep_cluster_kernel(const ClusterArgs CA) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SweepArgs &A = CA.A;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank(), ncta = CA.ncta;
    const int chain = blockIdx.x / ncta;
    const int E = A.E, Q = A.Q, P = A.P, NS = A.NS;
    double *dg = A.state + (size_t)chain * E;
    // shared layout: [partition of d] [scratch 34*8] [red_out 8] [partials 8*8] [theta P] [qc NS] [ss NS] [ints]
    double *sm = reinterpret_cast<double *>(smem_raw);
    double *dl = sm;
    double *scratch = sm + CA.part;
    double *red_out = scratch + 34 * 8;
    double *part_sh = red_out + 8;
    double *th = part_sh + 64;
    QConst *qc = reinterpret_cast<QConst *>(th + ((P + 1) & ~1));
    SuffStat *ss = reinterpret_cast<SuffStat *>(qc + NS);
    int *q_type = reinterpret_cast<int *>(ss + NS);
    int *q_dist = q_type + Q;
    int *q_poff = q_dist + NS;
    int *q_slot = q_poff + NS;
    __shared__ unsigned long long cnt[5];
    __shared__ int work_ctr[2];

    for (int i = threadIdx.x; i < P; i += blockDim.x) th[i] = A.theta[(size_t)chain * P + i];
    for (int i = threadIdx.x; i < Q; i += blockDim.x) q_type[i] = A.q_type[i];
    for (int i = threadIdx.x; i <= Q; i += blockDim.x) q_slot[i] = A.q_slot[i];
    for (int i = threadIdx.x; i < NS; i += blockDim.x) { q_dist[i] = A.q_dist[i]; q_poff[i] = A.q_poff[i]; }
    if (threadIdx.x < 5) cnt[threadIdx.x] = 0ull;
    const int e_lo = rank * CA.part, e_hi = min(E, e_lo + CA.part);
    for (int i = e_lo + threadIdx.x; i < e_hi; i += blockDim.x) dl[i - e_lo] = dg[i];
    __syncthreads();
    for (int q = threadIdx.x; q < NS; q += blockDim.x) {
        const int o = q_poff[q];
        qconst_set(qc[q], q_dist[q], th[o], (q_dist[q] == D_EXP) ? 0.0 : th[o + 1]);
    }
    cl.sync();  // every partition is loaded before anyone reads across CTAs

    DsmemView d;
    d.local = dl; d.k = (unsigned)CA.k; d.mask = (1u << CA.k) - 1u;
    const unsigned long long gchain = (unsigned long long)(A.chain_offset + chain);
    long long n_ev = 0, n_evals = 0, n_stuck = 0, n_pevals = 0, n_reject = 0;
    double tmax = 0.0;

    for (int sw = 0; sw < A.n_sweeps; ++sw) {
        const unsigned int gsw = A.sweep_base + (unsigned int)sw;
        if (sw == 0 || (A.flags & FLAG_TMAX_PER_SWEEP)) {  // Tmax = 2 max d   (qnet.pyx:672)
            double m = 0.0;
            for (int i = threadIdx.x; i < e_hi - e_lo; i += blockDim.x) m = fmax(m, dl[i]);
            m = block_reduce_max(m, scratch);
            if (threadIdx.x == 0) part_sh[rank] = m;
            cl.sync();
            for (int r = 0; r < ncta; ++r) m = fmax(m, cl.map_shared_rank(part_sh, r)[r]);
            tmax = 2.0 * m;
            cl.sync();
        }
        const int n_colors = (A.flags & FLAG_PARAMS_ONLY) ? 0 : A.n_colors;
        const bool rev = (A.flags & FLAG_REVERSE) != 0;
        for (int c = 0; c < n_colors; ++c) {
            const int cc = rev ? n_colors - 1 - c : c;
            const int lo = CA.class_off[cc * (ncta + 1) + rank], hi = CA.class_off[cc * (ncta + 1) + rank + 1];
            if (threadIdx.x == 0) work_ctr[c & 1] = lo;
            __syncthreads();
            if (A.mode == MODE_MH)
                run_class_mh(A, d, q_type, qc, hi, &work_ctr[c & 1], gchain, gsw, n_ev, n_evals, n_reject, q_slot, nullptr);
            else if (ALL_EXP)
                run_class<CondExp>(A, d, q_type, qc, hi, &work_ctr[c & 1], tmax, gchain, gsw, n_ev, n_evals, n_stuck,
                                   q_slot, nullptr);
            else
                run_class<CondGen>(A, d, q_type, qc, hi, &work_ctr[c & 1], tmax, gchain, gsw, n_ev, n_evals, n_stuck,
                                   q_slot, nullptr);
            cl.sync();  // the class barrier: every CTA's stores are visible cluster-wide before the next class reads
        }
        const bool trace = (A.flags & FLAG_TRACE) != 0;
        if (A.update != UPD_NONE || trace) {
            // sufficient statistics: every CTA takes a share of each queue's events; the per-CTA partials of up to 8
            // queues are exchanged with ONE pair of cluster barriers and summed in rank order by every CTA
            double logp_acc = 0.0;
            const int gtid = rank * blockDim.x + threadIdx.x, gstride = ncta * blockDim.x;
            for (int q0 = 0; q0 < Q; q0 += 8) {
                const int nq = min(8, Q - q0);
                for (int qq = 0; qq < nq; ++qq) {
                    const int q = q0 + qq, lo = A.q_off[q], hi = A.q_off[q + 1];
                    const int fifo = (q_type[q] == Q_GGK), dist = q_dist[q];
                    const bool need_log = (dist != D_EXP);
                    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // n, sum, sumlog, nlog, wait, nzero, logp, #(s < 0)
                    for (int j = lo + gtid; j < hi; j += gstride) {
                        const int i = CA.qe_idx[j];
                        double a, s;
                        event_service(d, i, fifo, CA.qe_p[j], CA.qe_pt[j], a, s);
                        const double w = d[i] - a - s;
                        v[4] += (w > 0.0) ? w : 0.0;
                        if (s > 0.0) {
                            v[0] += 1.0; v[1] += s;
                            if (need_log) {
                                const double ls = log(s);
                                if (dist == D_GAMMA) { v[2] += ls; v[3] += 1.0; }
                                else if (s >= 1e-50) { v[2] += ls; v[3] += 1.0; }
                            }
                        } else if (s == 0.0) v[5] += 1.0;
                        else v[7] += 1.0;
                        if (trace && !(s < 0.0)) v[6] += lpdf(qc[q], s);
                    }
                    block_reduce_sum<8>(v, scratch, red_out);
                    if (threadIdx.x < 8) part_sh[qq * 8 + threadIdx.x] = red_out[threadIdx.x];
                }
                cl.sync();
                if (threadIdx.x < nq * 8) {   // thread t sums value (t & 7) of queue q0 + (t >> 3) over the ranks
                    double x = 0.0;
                    for (int r = 0; r < ncta; ++r) x += cl.map_shared_rank(part_sh, r)[threadIdx.x];
                    reinterpret_cast<double *>(ss + q0)[threadIdx.x] = x;   // staged in ss, unpacked below
                }
                cl.sync();
                if (threadIdx.x == 0) {
                    for (int qq = 0; qq < nq; ++qq) {
                        double r8[8];
                        for (int i = 0; i < 8; ++i) r8[i] = reinterpret_cast<double *>(ss + q0 + qq)[i];
                        SuffStat &t = ss[q0 + qq];
                        t.n = r8[0]; t.sum = r8[1]; t.sumlog = r8[2]; t.nlog = r8[3]; t.wait = r8[4]; t.nzero = r8[5];
                        t.nall = r8[0] + r8[5] + r8[7]; t.sumsq = 0.0;
                        logp_acc += (trace && r8[7] > 0.0) ? BQ_NEG_INF : r8[6];
                    }
                }
                __syncthreads();
                bool any_ln = false;
                for (int qq = 0; qq < nq; ++qq) any_ln = any_ln || (q_dist[q0 + qq] == D_LOGNORMAL);
                if (any_ln) {  // second pass (distributions.pyx:276-299 is two-pass), again one exchange per 8 queues
                    for (int qq = 0; qq < nq; ++qq) {
                        const int q = q0 + qq, lo = A.q_off[q], hi = A.q_off[q + 1];
                        const int fifo = (q_type[q] == Q_GGK);
                        double sq[1] = {0.0};
                        if (q_dist[q] == D_LOGNORMAL) {
                            const double mean_log = (ss[q].nlog > 0.0) ? ss[q].sumlog / ss[q].nlog : 0.0;
                            for (int j = lo + gtid; j < hi; j += gstride) {
                                double a, s;
                                event_service(d, CA.qe_idx[j], fifo, CA.qe_p[j], CA.qe_pt[j], a, s);
                                if (s >= 1e-50) { const double z = log(s) - mean_log; sq[0] += z * z; }
                            }
                        }
                        block_reduce_sum<1>(sq, scratch, red_out);
                        if (threadIdx.x == 0) part_sh[qq] = red_out[0];
                    }
                    cl.sync();
                    if (threadIdx.x < nq) {
                        double x = 0.0;
                        for (int r = 0; r < ncta; ++r) x += cl.map_shared_rank(part_sh, r)[threadIdx.x];
                        ss[q0 + threadIdx.x].sumsq = x;
                    }
                    cl.sync();
                }
            }
            __syncthreads();
            const size_t recno = (size_t)(A.trace_base + sw);
            if (trace && rank == 0) {
                for (int q = threadIdx.x; q < Q; q += blockDim.x)
                    A.tr_wait[(recno * A.n_chains + chain) * Q + q] = (ss[q].nall > 0.0) ? ss[q].wait / ss[q].nall : 0.0;
                if (threadIdx.x == 0) A.tr_logp[recno * A.n_chains + chain] = logp_acc;  // accumulated by thread 0
            }
            // every CTA repeats the (deterministic) parameter update on the same statistics: no broadcast needed
            for (int q = threadIdx.x; A.update != UPD_NONE && q < NS; q += blockDim.x) {
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
            __syncthreads();
            if (trace && rank == 0)
                for (int i = threadIdx.x; i < P; i += blockDim.x)
                    A.tr_theta[(recno * A.n_chains + chain) * P + i] = th[i];
        }
    }

    cl.sync();
    for (int i = e_lo + threadIdx.x; i < e_hi; i += blockDim.x) dg[i] = dl[i - e_lo];
    if (rank == 0)
        for (int i = threadIdx.x; i < P; i += blockDim.x) A.theta[(size_t)chain * P + i] = th[i];
    atomicAdd(&cnt[0], (unsigned long long)n_ev);
    atomicAdd(&cnt[1], (unsigned long long)n_evals);
    atomicAdd(&cnt[2], (unsigned long long)(rank == 0 ? n_pevals : 0));
    atomicAdd(&cnt[3], (unsigned long long)n_stuck);
    atomicAdd(&cnt[4], (unsigned long long)n_reject);
    __syncthreads();
    if (threadIdx.x < 5)
        atomicAdd((unsigned long long *)&A.stats[(size_t)chain * BQ_NSTAT + threadIdx.x], cnt[threadIdx.x]);
    cl.sync();  // no CTA leaves while a peer may still read its shared memory
}

}  // namespace bq
