// libbq_b200.so -- C-ABI of the B200-native bayes-qnet hot path (declared in include/bq_capi.h).
//
// Host side of the library: validates and flattens the network, builds the colour-class sweep
// schedule, owns device memory and the stream, launches the kernels in bq_static.cuh (static queue
// order: single-server FIFO + delay stations) and bq_dynamic.cuh (multi-server FIFO, processor
// sharing).  No torch types, no exceptions across the boundary.  There is no CPU fallback: without a
// CUDA device bq_create fails with BQ_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/bq_capi.h"
#include "bq_static.cuh"
#include "bq_dynamic.cuh"
#include "bq_chib.cuh"
#include "bq_cluster.cuh"
#include "bq_host.h"

// bq_batch_nw.cu: the batch kernel with two warps per chain, and its copy of the log-count table
namespace bq {
int launch_batch_nw2(const DynArgs &A, int *claim, int *claimd, int kmax, bool has_ps, bool dense, int n_chains, int Q, int P,
                     cudaStream_t stream);  // returns the number of kernel launches (waves)
cudaError_t batch_nw_set_log_table(const double *tab256);
}


using namespace bq;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(BQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

struct bq_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int E = 0, Q = 0, P = 0, C = 0;
    long long chain_offset = 0;
    unsigned long long seed = 0;
    unsigned int sweep_counter = 0;
    bool static_order = true;
    bool all_exp = true;  // every service family exponential: branch-free conditional
    int threads = 512;
    int refill = BQ_REFILL_THRESH;  // idle lanes per warp that trigger a refill from the class work counter
    bool use_smem = true;
    size_t smem_bytes = 0;
    // host copies
    std::vector<int> q_type, q_k, q_dist, q_poff;
    std::vector<int> q_wide;  // FIFO queues with k >= #events, run as delay stations
    // distribution slots (mixture templates): per-slot family / parameter offset, first slot and mixing offset per queue
    int NS = 0;
    bool has_mix = false;
    std::vector<int> slot_dist, slot_poff, q_slot, q_mixoff;
    int *d_slot_dist = nullptr, *d_slot_poff = nullptr, *d_q_slot = nullptr, *d_q_mixoff = nullptr;
    unsigned char *d_aux = nullptr;  // [C][E] mixture component of every event
    // thread-block cluster path (bq_cluster.cuh): CTAs per chain (1 = off), partition size, encoding shift, tables
    int ncta = 1, cpart = 0, ck = 0;
    size_t csmem = 0;
    int4 *d_crec = nullptr;
    int *d_class_off = nullptr, *d_qe_idx = nullptr, *d_qe_p = nullptr, *d_qe_pt = nullptr;
    std::vector<int> qid, tid, prev_byt, next_byt, prev_byq, next_byq;
    std::vector<uint8_t> obs_a, obs_d;
    std::vector<int> color_off, order;
    std::vector<SchedRec> recs;
    // device
    int4 *d_rec = nullptr;
    int *d_color_off = nullptr, *d_ev_q = nullptr, *d_ev_p = nullptr, *d_ev_pt = nullptr;
    int *d_q_events = nullptr, *d_q_off = nullptr, *d_q_type = nullptr, *d_q_dist = nullptr,
        *d_q_poff = nullptr;
    double *d_state = nullptr, *d_theta = nullptr;
    long long *d_stats = nullptr;
    double *d_tr_theta = nullptr, *d_tr_wait = nullptr, *d_tr_logp = nullptr;
    int trace_cap = 0, trace_len = 0;
    double *d_scratch = nullptr;  // logcond x/out
    int *d_init_topo = nullptr;   // per-chain initialiser: topological order of the schedule records, upper bounds
    double *d_init_ub = nullptr;
    unsigned int init_counter = 0;
    int scratch_n = 0;
    long long n_sweeps = 0, n_launches = 0;
    float last_ms = 0.f;
    bool timed = false;
    // ---- dynamic-order networks (k > 1, PS): per-chain queue order and free-time vectors ----
    bool has_ps = false;    // a PS or random-order queue: sorted-departure arrays, claim array for their slots
    bool has_gg1r = false;
    int EF = 0, kmax = 1, dyn_w = 32, dyn_threads = 128;
    std::vector<int> q_off, q_foff, ord0;  // queue position ranges, F offsets, initial order
    std::vector<int> dyn_order;          // sweep order of the dynamic-order kernels
    bool dyn_batch = true;               // dynamic-order kernels: warp-per-chain batches (else W-lane groups, BQ_DYN_BATCH=0)
    int order_stride = 0;                // batch order: 0 = lanes 1/32 of the queue apart; r > 0 = windows of 32 r events
    int batch_nw = 1;                    // warps per chain of the batch kernel (rounds of 32 batch_nw events), fixed at create
    int *d_claim = nullptr, *d_claimd = nullptr;
    bool mh_dynamic = false;             // static-order handle: run MH sweeps in the dynamic-order kernel (BQ_MH_DYNAMIC)
    bool dyn_ready = false;              // dynamic arrays allocated
    bool dyn_stale = false;              // ... but the static kernel has moved d since they were built
    DynRec *d_order = nullptr;
    QInfo *d_qinfo = nullptr;
    int *d_ord = nullptr, *d_pos = nullptr, *d_dord = nullptr, *d_dpos = nullptr;
    double *d_prec = nullptr, *d_sq = nullptr, *d_dt = nullptr;  // position records (DynArgs::rec), PS arrays
    unsigned char *d_flag = nullptr;  // contended-arrival flags (DynArgs::flag)
    int FW = 32;                      // bytes per chain
    int recS = 4;  // doubles per position record
    HostNet net() const {
        HostNet n;
        n.E = E; n.Q = Q;
        n.q_type = q_type.data(); n.q_k = q_k.data(); n.q_off = q_off.data(); n.q_foff = q_foff.data();
        n.q_dist = q_dist.data(); n.prev_byt = prev_byt.data();
        return n;
    }
};

extern "C" const char *bq_last_error(void) { return g_err.c_str(); }
extern "C" int bq_version(void) { return 100; }
extern "C" int bq_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(BQ_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    return n;
}

// ---------------------------------------------------------------------------------------------
// schedule: resample-eligible events -> neighbourhood records -> conflict-free colour classes
// ---------------------------------------------------------------------------------------------
static SchedRec make_rec(const bq_handle *h, int e) {
    SchedRec r;
    r.e = e;
    r.pt_e = h->prev_byt[e];
    r.p = h->prev_byq[e];
    r.n = h->next_byq[e];
    r.pt_n = (r.n >= 0) ? h->prev_byt[r.n] : -1;
    r.e1 = h->next_byt[e];
    r.p1 = -1; r.pt_p1 = -1; r.pt_n1 = -1;
    r.q0 = h->qid[e];
    r.q1 = -1;
    r.pad = e;
    if (r.e1 >= 0) {
        r.q1 = h->qid[r.e1];
        r.p1 = h->prev_byq[r.e1];
        if (r.p1 >= 0) r.pt_p1 = h->prev_byt[r.p1];
        int n1 = h->next_byq[r.e1];
        if (n1 >= 0) r.pt_n1 = h->prev_byt[n1];
    }
    return r;
}

// events whose departure time the conditional of r.e reads (besides its own)
static int rec_reads(const bq_handle *h, const SchedRec &r, int out[8]) {
    int n = 0;
    auto add = [&](int x) { if (x >= 0) out[n++] = x; };
    add(r.pt_e);
    if (h->q_type[r.q0] == Q_GGK) { add(r.p); add(r.n); if (r.n >= 0) add(r.pt_n); }
    if (r.e1 >= 0) {
        add(r.e1);
        if (h->q_type[r.q1] == Q_GGK) { add(r.p1); add(r.pt_p1); add(r.pt_n1); }
    }
    return n;
}

static bool eligible(const bq_handle *h, int e) {
    // qnet.pyx:688: not evt.obs_d and (not evt_next or not evt_next.obs_a)
    if (h->obs_d[e]) return false;
    int e1 = h->next_byt[e];
    return e1 < 0 || !h->obs_a[e1];
}

static int build_schedule(bq_handle *h) {
    const int E = h->E;
    std::vector<int> elig;
    for (int e = 0; e < E; ++e)
        if (eligible(h, e)) elig.push_back(e);
    const int n = (int)elig.size();
    std::vector<int> slot(E, -1);
    for (int i = 0; i < n; ++i) slot[elig[i]] = i;
    std::vector<SchedRec> recs(n);
    // adjacency among eligible events: a reads b  =>  a -- b
    std::vector<std::vector<int>> adj(n);
    for (int i = 0; i < n; ++i) {
        recs[i] = make_rec(h, elig[i]);
        int rd[8];
        int nr = rec_reads(h, recs[i], rd);
        for (int k = 0; k < nr; ++k) {
            int j = slot[rd[k]];
            if (j >= 0 && j != i) { adj[i].push_back(j); adj[j].push_back(i); }
        }
    }
    // greedy colouring, smallest available colour
    std::vector<int> color(n, -1);
    int n_colors = 0;
    std::vector<int> mark;
    for (int i = 0; i < n; ++i) {
        mark.assign(n_colors + 1, 0);
        for (int j : adj[i])
            if (color[j] >= 0) mark[color[j]] = 1;
        int c = 0;
        while (c < n_colors && mark[c]) ++c;
        color[i] = c;
        if (c == n_colors) ++n_colors;
    }
    // classes; inside a class sort by (queue of e, queue of e1, e) so warps are homogeneous
    h->color_off.assign(n_colors + 1, 0);
    for (int i = 0; i < n; ++i) h->color_off[color[i] + 1]++;
    for (int c = 0; c < n_colors; ++c) h->color_off[c + 1] += h->color_off[c];
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) {
        if (color[a] != color[b]) return color[a] < color[b];
        if (recs[a].q0 != recs[b].q0) return recs[a].q0 < recs[b].q0;
        if (recs[a].q1 != recs[b].q1) return recs[a].q1 < recs[b].q1;
        return recs[a].e < recs[b].e;
    });
    h->recs.resize(n);
    h->order.resize(n);
    for (int i = 0; i < n; ++i) { h->recs[i] = recs[idx[i]]; h->order[i] = recs[idx[i]].e; }
    return 0;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t upload(T **dst, const void *src, size_t n) {
    cudaError_t e = cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

static size_t static_smem_bytes(const bq_handle *h, bool smem_state) {
    size_t Epad = smem_state ? (size_t)((h->E + 1) & ~1) : 0;
    size_t b = (Epad + 34 * 8 + 8 + ((h->P + 1) & ~1)) * sizeof(double);
    b += (size_t)h->NS * (sizeof(QConst) + sizeof(SuffStat)) + (size_t)(2 * h->Q + 2 * h->NS + 2) * sizeof(int);
    return (b + 15) & ~(size_t)15;
}

static void fill_args(const bq_handle *h, SweepArgs &A) {
    memset(&A, 0, sizeof(A));
    A.E = h->E; A.Q = h->Q; A.P = h->P; A.n_chains = h->C;
    A.n_colors = (int)h->color_off.size() - 1;
    A.rec = h->d_rec; A.color_off = h->d_color_off;
    A.ev_q = h->d_ev_q; A.ev_p = h->d_ev_p; A.ev_pt = h->d_ev_pt;
    A.q_events = h->d_q_events; A.q_off = h->d_q_off;
    A.q_type = h->d_q_type; A.q_dist = h->d_slot_dist; A.q_poff = h->d_slot_poff;
    A.NS = h->NS; A.q_slot = h->d_q_slot; A.q_mixoff = h->d_q_mixoff; A.aux = h->d_aux;
    A.state = h->d_state; A.theta = h->d_theta; A.stats = h->d_stats;
    A.tr_theta = h->d_tr_theta; A.tr_wait = h->d_tr_wait; A.tr_logp = h->d_tr_logp;
    A.seed = h->seed; A.chain_offset = h->chain_offset; A.sweep_base = h->sweep_counter;
    A.refill = h->refill;
}


// ---------------------------------------------------------------------------------------------
// dynamic-order networks (k > 1, PS): host-side construction of one chain's derived arrays
// ---------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t replicate_rows(T *base, size_t row, size_t C, cudaStream_t st) {  // row 0 -> rows 1..C-1
    for (size_t have = 1; have < C; have *= 2) {
        const size_t n = std::min(have, C - have);
        cudaError_t e = cudaMemcpyAsync(base + have * row, base, n * row * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static void fill_dyn_args(const bq_handle *h, DynArgs &A) {
    memset(&A, 0, sizeof(A));
    A.E = h->E; A.Q = h->Q; A.P = h->P; A.n_chains = h->C; A.S = h->recS;
    A.n_order = (int)h->dyn_order.size();
    A.order = h->d_order;
    A.qinfo = h->d_qinfo; A.q_poff = h->d_q_poff;
    A.state = h->d_state; A.ord = h->d_ord; A.pos = h->d_pos; A.rec = h->d_prec; A.flag = h->d_flag; A.FW = h->FW;
    A.dord = h->d_dord; A.dpos = h->d_dpos; A.dt = h->d_dt; A.sq = h->d_sq;
    A.theta = h->d_theta; A.stats = h->d_stats;
    A.tr_theta = h->d_tr_theta; A.tr_wait = h->d_tr_wait; A.tr_logp = h->d_tr_logp;
    A.seed = h->seed; A.chain_offset = h->chain_offset; A.sweep_base = h->sweep_counter;
    A.scan = BQ_SUPPORT_SCAN;
    if (const char *env = getenv("BQ_SUPPORT_SCAN")) A.scan = std::max(0, atoi(env));  // tuning / test hook: 0 = no pruning
}

// host staging of the per-chain derived arrays for `rows` chains
struct HostRows {
    std::vector<int> ord, pos, dord, dpos;
    std::vector<double> at, dq, ls, F, dt, sq;
    std::vector<double> rec;  // at, dq, ls and F interleaved by position, as the device holds them
    std::vector<unsigned char> flag;  // contended-arrival flags
    // one record per position: {at, dq, F[0], F[1], ls, F[2..k)} (bq_dynamic.cuh, BQ_REC_*)
    void pack(const bq_handle *h, size_t rows);
    void alloc(size_t rows, size_t E, size_t EF, bool ps) {
        ord.assign(rows * E, 0); pos.assign(rows * E, 0);
        at.assign(rows * E, 0.0); dq.assign(rows * E, 0.0); ls.assign(rows * E, 0.0);
        F.assign(std::max<size_t>(rows * EF, 1), 0.0);
        if (ps) { dord.assign(rows * E, 0); dpos.assign(rows * E, 0); dt.assign(rows * E, 0.0); sq.assign(rows * E, 0.0); }
    }
    HostChain chain(size_t i, size_t E, size_t EF) {
        HostChain c;
        c.ord = &ord[i * E]; c.pos = &pos[i * E]; c.at = &at[i * E]; c.dq = &dq[i * E]; c.ls = &ls[i * E];
        c.F = &F[i * EF];
        const bool ps = !dt.empty();
        c.dt = ps ? &dt[i * E] : nullptr; c.sq = ps ? &sq[i * E] : nullptr;
        c.dord = ps ? &dord[i * E] : nullptr; c.dpos = ps ? &dpos[i * E] : nullptr;
        return c;
    }
};

void HostRows::pack(const bq_handle *h, size_t rows) {
    const size_t E = h->E, S = h->recS, EF = h->EF;
    rec.assign(rows * E * S, 0.0);
    flag.assign(rows * (size_t)h->FW, 0);
    for (size_t i = 0; i < rows; ++i)
        for (int q = 0; q < h->Q; ++q) {
            const int lo = h->q_off[q], hi = h->q_off[q + 1], k = (h->q_type[q] == Q_GGK) ? h->q_k[q] : 0;
            for (int p = lo; p < hi; ++p) {
                double *r = &rec[(i * E + p) * S];
                r[BQ_REC_AT] = at[i * E + p]; r[BQ_REC_DQ] = dq[i * E + p]; r[rec_ls_slot((int)S)] = ls[i * E + p];
                for (int j = 0; j < k; ++j) r[BQ_REC_F + BQ_FSLOT(j)] = F[i * EF + h->q_foff[q] + (size_t)(p - lo) * k + j];
                if (k > 1 && r[BQ_REC_AT] < r[BQ_REC_F + 1]) flag[i * h->FW + p] = 1;
            }
        }
}

// rows staged on the host -> chains [lo, lo + rows); when rows == 1 < n the row is replicated to n chains
static int upload_rows(bq_handle *h, HostRows &r, int lo, size_t rows, size_t n, bool sync) {
    const size_t E = h->E, o = (size_t)lo * E;
    r.pack(h, rows);
    auto put = [&](auto *dst, const auto &src, size_t row) -> cudaError_t {
        if (!row) return cudaSuccess;
        cudaError_t e = cudaMemcpyAsync(dst, src.data(), rows * row * sizeof(*dst), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess && rows == 1 && n > 1) e = replicate_rows(dst, row, n, h->stream);
        return e;
    };
    CK(put(h->d_ord + o, r.ord, E));
    CK(put(h->d_pos + o, r.pos, E));
    CK(put(h->d_prec + o * h->recS, r.rec, E * h->recS));
    CK(put(h->d_flag + (size_t)lo * h->FW, r.flag, (size_t)h->FW));
    if (h->has_ps) {
        CK(put(h->d_dord + o, r.dord, E));
        CK(put(h->d_dpos + o, r.dpos, E));
        CK(put(h->d_dt + o, r.dt, E));
        CK(put(h->d_sq + o, r.sq, E));
    }
    if (sync) CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}

// initial queue orders from the caller's by-queue links, and the sequential scan order of the dynamic-order kernels:
// queue after queue, each in its initial arrival order (any fixed scan is a valid Gibbs sweep)
static int build_sequential_order(bq_handle *h) {
    const int E = h->E, Q = h->Q;
    h->ord0.assign(E, -1);
    std::vector<int> head(Q, -1);
    for (int e = 0; e < E; ++e)
        if (h->prev_byq[e] < 0) {
            if (head[h->qid[e]] >= 0) return fail(BQ_ERR_INVALID, "a queue has two head events (prev_byq < 0)");
            head[h->qid[e]] = e;
        }
    for (int q = 0; q < Q; ++q) {
        int p = h->q_off[q];
        for (int e = head[q]; e >= 0; e = h->next_byq[e]) {
            if (p >= h->q_off[q + 1]) return fail(BQ_ERR_INVALID, "by-queue links do not form one list per queue");
            h->ord0[p++] = e;
        }
        if (p != h->q_off[q + 1]) return fail(BQ_ERR_INVALID, "by-queue links do not cover the queue");
    }
    h->dyn_order.clear();
    if (const char *env = getenv("BQ_DYN_BATCH")) h->dyn_batch = (env[0] != '0');
    bool batch_order = h->dyn_batch;
    if (const char *env = getenv("BQ_DYN_ORDER")) batch_order = (env[0] == 'b');  // "batch" | "seq": test hook, so that
    if (!batch_order) {                                                            // both kernels can walk one order
        for (int p = 0; p < E; ++p)
            if (eligible(h, h->ord0[p])) h->dyn_order.push_back(h->ord0[p]);
        return BQ_OK;
    }
    // batch kernel: each queue's eligible events interleaved with a stride of 1/32 of the queue, so that any 32
    // consecutive events of the order are far apart in their queue (and rarely touch each other's neighbourhoods)
    // BQ_ORDER_STRIDE = r > 0: windows of 32 r consecutive eligible events, interleaved with stride r INSIDE the window,
    // so the 32 lanes of a round work r positions apart and stay in one window of the queue for r rounds (their
    // records share cache lines and stay in L1 / L2 between rounds); 0 = stride of 1/32 of the whole queue.
    int win_stride = h->order_stride;
    if (const char *env = getenv("BQ_ORDER_STRIDE")) win_stride = atoi(env);
    std::vector<int> lst;
    if (h->batch_nw > 1 && !getenv("BQ_BATCH_NW")) {
        // rounds of 64 events need queues long enough for 64 far-apart events: below ~4096 eligible events per queue the
        // witness scans of neighbouring lanes overlap and most of a round is redone (profiles/r2_batch_nw.log)
        for (int q = 0; q < Q; ++q) {
            int n = 0;
            for (int p = h->q_off[q]; p < h->q_off[q + 1]; ++p) n += eligible(h, h->ord0[p]) ? 1 : 0;
            if (n > 0 && n < 4096) h->batch_nw = 1;
        }
    }
    for (int q = 0; q < Q; ++q) {
        lst.clear();
        for (int p = h->q_off[q]; p < h->q_off[q + 1]; ++p)
            if (eligible(h, h->ord0[p])) lst.push_back(h->ord0[p]);
        const int n = (int)lst.size();
        if (win_stride <= 0) {
            const int rb = 32 * h->batch_nw;  // events per round: consecutive entries of the order are 1/rb of the queue apart
            const int st = (n + rb - 1) / rb;
            for (int r = 0; r < st; ++r)
                for (int i = r; i < n; i += st) h->dyn_order.push_back(lst[i]);
        } else {
            const int win = 32 * win_stride;
            for (int w0 = 0; w0 < n; w0 += win) {
                const int m = std::min(win, n - w0), st = (m + 31) / 32;
                for (int r = 0; r < st; ++r)
                    for (int i = r; i < m; i += st) h->dyn_order.push_back(lst[w0 + i]);
            }
        }
    }
    return BQ_OK;
}

// per-handle tables of the dynamic kernels + the derived arrays of the initial state, replicated to all chains
static int setup_dynamic(bq_handle *h, const double *d0) {
    const int E = h->E, Q = h->Q;
    h->q_foff.assign(Q, 0);
    h->EF = 0; h->kmax = 1;
    for (int q = 0; q < Q; ++q)
        if (h->q_type[q] == Q_GGK) {
            h->kmax = std::max(h->kmax, h->q_k[q]);
            h->q_foff[q] = h->EF;
            h->EF += h->q_k[q] * (h->q_off[q + 1] - h->q_off[q]);
        }
    if (h->kmax > 8)
        return fail(BQ_ERR_UNSUPPORTED, "a FIFO queue with more than 8 processors and more events than processors (queues "
                                        "with at least as many processors as events run as delay stations)");
    {
        int max_optin = 0;
        CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
        if (4 * dyn_group_smem(h->Q, h->P) > (size_t)max_optin)
            return fail(BQ_ERR_UNSUPPORTED, "too many queues for the dynamic-order kernels (per-chain tables of 4 chains must "
                                            "fit the shared memory of a CTA, about 400 queues)");
    }
    h->recS = rec_stride(h->kmax);
    h->FW = ((E + 31) & ~31) + 32;
    HostRows r0;
    r0.alloc(1, E, h->EF, h->has_ps);
    memcpy(r0.ord.data(), h->ord0.data(), E * sizeof(int));
    {
        HostNet hn = h->net();
        host_rebuild(&hn, d0, r0.chain(0, E, h->EF), false);
    }
    if (!h->static_order) {  // the handle's schedule is the sequential scan: one event per class
        h->order = h->dyn_order;
        h->color_off.resize(h->order.size() + 1);
        for (size_t i = 0; i <= h->order.size(); ++i) h->color_off[i] = (int)i;
    }
    std::vector<DynRec> recs(h->dyn_order.size());
    for (size_t i = 0; i < recs.size(); ++i) {
        const int e = h->dyn_order[i], e1 = h->next_byt[e];
        recs[i].e = e; recs[i].e1 = e1; recs[i].q0 = h->qid[e]; recs[i].q1 = e1 >= 0 ? h->qid[e1] : -1;
    }
    std::vector<QInfo> qinfo(Q);
    for (int q = 0; q < Q; ++q) {
        qinfo[q].type = h->q_type[q]; qinfo[q].k = h->q_k[q]; qinfo[q].lo = h->q_off[q]; qinfo[q].hi = h->q_off[q + 1];
        qinfo[q].foff = h->q_foff[q]; qinfo[q].dist = h->q_dist[q];
    }
    static_assert(sizeof(DynRec) == 16, "DynRec is loaded as one int4");

    const size_t C = h->C;
    CK(upload(&h->d_order, recs.data(), recs.size()));
    CK(upload(&h->d_qinfo, qinfo.data(), qinfo.size()));
    CK(cudaMalloc((void **)&h->d_ord, C * E * sizeof(int)));
    CK(cudaMalloc((void **)&h->d_pos, C * E * sizeof(int)));
    CK(cudaMalloc((void **)&h->d_prec, C * E * h->recS * sizeof(double)));
    CK(cudaMalloc((void **)&h->d_flag, C * h->FW));
    CK(cudaMemsetAsync(h->d_flag, 0, C * h->FW, h->stream));
    {   // log of small counts for the PS terms, computed with the host's log()
        double tab[256];
        tab[0] = -INFINITY;
        for (int i = 1; i < 256; ++i) tab[i] = log((double)i);
        CK(cudaMemcpyToSymbol(c_log_count, tab, sizeof(tab)));
        CK(batch_nw_set_log_table(tab));
    }
    if (h->has_ps) {
        CK(cudaMalloc((void **)&h->d_dord, C * E * sizeof(int)));
        CK(cudaMalloc((void **)&h->d_dpos, C * E * sizeof(int)));
        CK(cudaMalloc((void **)&h->d_dt, C * E * sizeof(double)));
        CK(cudaMalloc((void **)&h->d_sq, C * E * sizeof(double)));
    }
    if (h->dyn_batch) {
        CK(cudaMalloc((void **)&h->d_claim, C * E * sizeof(int)));
        if (h->has_ps) CK(cudaMalloc((void **)&h->d_claimd, C * E * sizeof(int)));
    }
    if (int rc = upload_rows(h, r0, 0, 1, C, false)) return rc;
    CK(cudaStreamSynchronize(h->stream));
    // lanes per chain: wide groups hide the latency of a single chain's serial walk when chains are few; with many
    // chains narrower groups waste fewer lanes on speculative candidates (measured on B200, tools/probe_dyn.py)
    h->dyn_w = (h->C >= 12288) ? 4 : (h->C >= 6144 ? 8 : (h->C >= 1536 ? 16 : 32));
    if (const char *env = getenv("BQ_DYN_W")) {
        int w = atoi(env);
        if (w == 4 || w == 8 || w == 16 || w == 32) h->dyn_w = w;
    }
    if (const char *env = getenv("BQ_DYN_THREADS")) {
        int t = atoi(env);
        if (t == 32 || t == 64 || t == 128) h->dyn_threads = t;
    }
    h->dyn_ready = true;
    h->dyn_stale = false;
    return BQ_OK;
}

// new departure times for chains [lo, hi) of a dynamic-order handle: every queue is re-ranked by arrival and
// the derived arrays are rebuilt on the host
static int dynamic_set_rows(bq_handle *h, int lo, int hi, const double *d, bool same_for_all) {
    const size_t E = h->E, n = (size_t)(hi - lo);
    const HostNet hn = h->net();
    if (same_for_all) {
        HostRows r;
        r.alloc(1, E, h->EF, h->has_ps);
        memcpy(r.ord.data(), h->ord0.data(), E * sizeof(int));
        host_rebuild(&hn, d, r.chain(0, E, h->EF), true);
        CK(cudaMemcpyAsync(h->d_state + (size_t)lo * E, d, E * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(replicate_rows(h->d_state + (size_t)lo * E, E, n, h->stream));
        return upload_rows(h, r, lo, 1, n, true);
    }
    // chains with their own states: staged in slices of at most ~256 MB of derived arrays on the host
    const size_t per_row = E * (size_t)(8 * (h->recS + 8) + 24) + (size_t)h->EF * 8 + 64;
    size_t budget = (size_t)256 << 20;
    if (const char *env = getenv("BQ_STAGE_MB")) budget = (size_t)atol(env) << 20;  // test hook: 0 = one chain per slice
    const size_t slice = std::max<size_t>(1, budget / per_row);
    for (size_t i0 = 0; i0 < n; i0 += slice) {
        const size_t rows = std::min(slice, n - i0);
        HostRows r;
        r.alloc(rows, E, h->EF, h->has_ps);
        for (size_t i = 0; i < rows; ++i) {
            memcpy(&r.ord[i * E], h->ord0.data(), E * sizeof(int));
            host_rebuild(&hn, d + (i0 + i) * E, r.chain(i, E, h->EF), true);
        }
        CK(cudaMemcpyAsync(h->d_state + ((size_t)lo + i0) * E, d + i0 * E, rows * E * sizeof(double),
                           cudaMemcpyHostToDevice, h->stream));
        if (int rc = upload_rows(h, r, lo + (int)i0, rows, rows, true)) return rc;  // synchronises: r may go
    }
    return BQ_OK;
}

// dynamic shared memory above the default 48 KB needs an explicit opt-in per kernel (networks with ~90+ queues)
static void smem_optin(const void *fn, size_t sm) {
    if (sm > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
}

template <int W>
static void launch_dynamic_w(bq_handle *h, const DynArgs &A) {
    const int G = h->dyn_threads / W, blocks = (h->C + G - 1) / G;
    const size_t sm = (size_t)G * dyn_group_smem(h->Q, h->P);
    if (sm > 48 * 1024) {  // many narrow groups per CTA: opt in to the large shared-memory carve-out
        cudaFuncSetAttribute((const void *)sweep_dynamic_kernel<W, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cudaFuncSetAttribute((const void *)sweep_dynamic_kernel<W, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    }
    if (h->kmax <= 4) sweep_dynamic_kernel<W, 4><<<blocks, h->dyn_threads, sm, h->stream>>>(A);
    else sweep_dynamic_kernel<W, 8><<<blocks, h->dyn_threads, sm, h->stream>>>(A);
}
static void launch_dynamic(bq_handle *h, const DynArgs &A) {
    if (h->dyn_batch) {  // one warp per chain, 32 events of the chain in flight
        const int G = 4, blocks = (h->C + G - 1) / G;
        const size_t sm = (size_t)G * dyn_group_smem(h->Q, h->P);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const bool dense = h->C > 16 * sms;  // more chains than 16 warps per SM can hold: use the 32-warp build
#ifndef BQ_DENSE_MINB
#define BQ_DENSE_MINB 8
#endif
#define BQ_LAUNCH_BATCH(K, MINB, MHK) \
    do { \
        if (h->has_ps) { smem_optin((const void *)sweep_batch_kernel<K, MINB, MHK, 1>, sm); \
            sweep_batch_kernel<K, MINB, MHK, 1><<<blocks, 32 * G, sm, h->stream>>>(A, h->d_claim, h->d_claimd); } \
        else { smem_optin((const void *)sweep_batch_kernel<K, MINB, MHK, 0>, sm); \
            sweep_batch_kernel<K, MINB, MHK, 0><<<blocks, 32 * G, sm, h->stream>>>(A, h->d_claim, h->d_claimd); } \
    } while (0)
        const bool mh = (A.mode == MODE_MH);
        // two warps per chain (rounds of 64 events) while the chains alone would leave half of the SM's warp slots empty
        const int nw = mh ? 1 : h->batch_nw;  // MH sweeps: one warp per chain on the same order (any batch size is exact)
        if (nw == 2) {  // compiled in its own translation unit (bq_batch_nw.cu): see there
            h->n_launches += launch_batch_nw2(A, h->d_claim, h->d_claimd, h->kmax, h->has_ps, 2 * h->C > 16 * sms, h->C, h->Q, h->P,
                                              h->stream) - 1;  // the caller counts one
            return;
        }
        if (h->has_gg1r) {  // random-order queues: one build (slice only, any k), 16 warps per SM
            smem_optin((const void *)sweep_batch_kernel<8, 4, false, 2>, sm);
            sweep_batch_kernel<8, 4, false, 2><<<blocks, 32 * G, sm, h->stream>>>(A, h->d_claim, h->d_claimd);
            return;
        }
        if (h->kmax <= 4) {
            if (dense) { if (mh) BQ_LAUNCH_BATCH(4, BQ_DENSE_MINB, true); else BQ_LAUNCH_BATCH(4, BQ_DENSE_MINB, false); }
            else { if (mh) BQ_LAUNCH_BATCH(4, 4, true); else BQ_LAUNCH_BATCH(4, 4, false); }
        } else {
            if (dense) { if (mh) BQ_LAUNCH_BATCH(8, BQ_DENSE_MINB, true); else BQ_LAUNCH_BATCH(8, BQ_DENSE_MINB, false); }
            else { if (mh) BQ_LAUNCH_BATCH(8, 4, true); else BQ_LAUNCH_BATCH(8, 4, false); }
        }
#undef BQ_LAUNCH_BATCH
        return;
    }
    if (h->dyn_w == 4) launch_dynamic_w<4>(h, A);
    else if (h->dyn_w == 8) launch_dynamic_w<8>(h, A);
    else if (h->dyn_w == 16) launch_dynamic_w<16>(h, A);
    else launch_dynamic_w<32>(h, A);
}

static size_t cluster_smem_bytes(const bq_handle *h, int part);
static int build_cluster_plan(bq_handle *h, int ncta);

extern "C" int bq_destroy(bq_handle *h) {
    if (!h) return BQ_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void *ptrs[] = {h->d_rec, h->d_color_off, h->d_ev_q, h->d_ev_p, h->d_ev_pt, h->d_q_events, h->d_q_off,
                    h->d_q_type, h->d_q_dist, h->d_q_poff, h->d_state, h->d_theta, h->d_stats,
                    h->d_tr_theta, h->d_tr_wait, h->d_tr_logp, h->d_scratch, h->d_order, h->d_qinfo,
                    h->d_ord, h->d_pos, h->d_dord, h->d_dpos, h->d_prec, h->d_flag, h->d_sq, h->d_dt, h->d_claim,
                    h->d_claimd, h->d_init_topo, h->d_init_ub, h->d_slot_dist, h->d_slot_poff, h->d_q_slot, h->d_q_mixoff,
                    h->d_aux, h->d_crec, h->d_class_off, h->d_qe_idx, h->d_qe_p, h->d_qe_pt};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return BQ_OK;
}

extern "C" int bq_create(const bq_model *m, const bq_arrivals *a, const double *theta0, int32_t n_chains,
                         int64_t chain_offset, uint64_t seed, int32_t device, bq_handle **out) {
    if (!m || !a || !theta0 || !out) return fail(BQ_ERR_INVALID, "bq_create: null argument");
    *out = nullptr;
    if (n_chains < 1) return fail(BQ_ERR_INVALID, "bq_create: n_chains < 1");
    if (m->n_queues < 1 || m->n_queues > 1024) return fail(BQ_ERR_INVALID, "bq_create: bad n_queues");
    if (a->n_events < 1) return fail(BQ_ERR_INVALID, "bq_create: empty arrivals");
    int ndev = bq_device_count();
    if (ndev <= 0) return fail(BQ_ERR_NO_DEVICE, "bq_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(BQ_ERR_INVALID, "bq_create: bad device index");

    bq_handle *h = new bq_handle();
    h->device = device;
    h->E = a->n_events; h->Q = m->n_queues; h->P = m->n_params; h->C = n_chains;
    h->chain_offset = chain_offset; h->seed = seed;
    const int E = h->E, Q = h->Q;
    h->q_type.assign(m->q_type, m->q_type + Q);
    h->q_k.assign(m->q_k, m->q_k + Q);
    h->q_dist.assign(m->q_dist, m->q_dist + Q);
    h->q_poff.assign(m->q_param_off, m->q_param_off + Q);
    h->qid.assign(a->qid, a->qid + E);
    h->tid.assign(a->tid, a->tid + E);
    h->obs_a.assign(a->obs_a, a->obs_a + E);
    h->obs_d.assign(a->obs_d, a->obs_d + E);
    h->prev_byt.assign(a->prev_byt, a->prev_byt + E);
    h->next_byt.assign(a->next_byt, a->next_byt + E);
    h->prev_byq.assign(a->prev_byq, a->prev_byq + E);
    h->next_byq.assign(a->next_byq, a->next_byq + E);

    auto bail = [&](int code, const std::string &msg) { bq_destroy(h); return fail(code, msg); };
    // A FIFO queue with at least as many servers as it has events never makes anyone wait: min F = 0 <= a at every
    // position, s = d - a, no diff list reaches another event -- exactly a delay station (the reference's own check:
    // test_mh_ggk.py:778-784, DELAY == GGk(k = 100)).  Such queues run as delay stations; only `proc` is reported
    // the GGk way (every event takes a fresh server).
    h->q_wide.assign(Q, 0);
    {
        std::vector<int> cnt(Q, 0);
        for (int e = 0; e < E; ++e)
            if (h->qid[e] >= 0 && h->qid[e] < Q) cnt[h->qid[e]]++;
        for (int q = 0; q < Q; ++q)
            if (h->q_type[q] == Q_GGK && h->q_k[q] > 1 && h->q_k[q] >= cnt[q]) {
                h->q_wide[q] = 1; h->q_type[q] = Q_DELAY; h->q_k[q] = 1;
            }
    }
    // distribution slots: one per plain queue, one per component of a mixture template
    h->q_slot.assign(Q + 1, 0);
    if (m->n_slots > 0) {
        if (!m->q_slot_off || !m->slot_dist) return bail(BQ_ERR_INVALID, "n_slots > 0 without q_slot_off / slot_dist");
        h->q_slot.assign(m->q_slot_off, m->q_slot_off + Q + 1);
        h->slot_dist.assign(m->slot_dist, m->slot_dist + m->n_slots);
        if (h->q_slot[0] != 0 || h->q_slot[Q] != m->n_slots) return bail(BQ_ERR_INVALID, "q_slot_off does not cover the slots");
        for (int q = 0; q < Q; ++q) {
            const int nc = h->q_slot[q + 1] - h->q_slot[q];
            if (nc < 1 || nc > 255) return bail(BQ_ERR_INVALID, "a queue needs 1..255 distribution slots");
            h->q_dist[q] = h->slot_dist[h->q_slot[q]];
        }
    } else {
        for (int q = 0; q <= Q; ++q) h->q_slot[q] = q;
        h->slot_dist = h->q_dist;
    }
    h->NS = h->q_slot[Q];
    h->slot_poff.assign(h->NS, 0);
    h->q_mixoff.assign(Q, -1);
    int pcount = 0;
    for (int q = 0; q < Q; ++q) {
        const int t = h->q_type[q], nc = h->q_slot[q + 1] - h->q_slot[q];
        if (t != Q_GGK && t != Q_DELAY && t != Q_PS && t != Q_GG1R) return bail(BQ_ERR_INVALID, "unknown queue type");
        if (h->q_poff[q] != pcount) return bail(BQ_ERR_INVALID, "q_param_off is not the running parameter count");
        if (nc > 1) { h->has_mix = true; h->q_mixoff[q] = pcount; pcount += nc; }
        for (int sl = h->q_slot[q]; sl < h->q_slot[q + 1]; ++sl) {
            const int dd = h->slot_dist[sl];
            if (dd != D_EXP && dd != D_GAMMA && dd != D_LOGNORMAL) return bail(BQ_ERR_INVALID, "unknown service family");
            h->slot_poff[sl] = pcount;
            pcount += (dd == D_EXP) ? 1 : 2;
            if (dd != D_EXP) h->all_exp = false;
        }
        if (t == Q_PS || t == Q_GG1R || (t == Q_GGK && h->q_k[q] != 1)) h->static_order = false;
    }
    if (h->has_mix) h->all_exp = false;  // the branch-free exponential path carries no per-event component
    if (h->has_mix && !h->static_order)
        return bail(BQ_ERR_UNSUPPORTED, "mixture service templates are built for static-order networks (single-server FIFO "
                                        "queues and delay stations), not next to multi-server or PS queues");
    if (pcount != h->P) return bail(BQ_ERR_INVALID, "n_params does not match the service families");
    for (int e = 0; e < E; ++e) {
        if (h->qid[e] < 0 || h->qid[e] >= Q) return bail(BQ_ERR_INVALID, "event with bad qid");
        int lk[4] = {h->prev_byt[e], h->next_byt[e], h->prev_byq[e], h->next_byq[e]};
        for (int k = 0; k < 4; ++k)
            if (lk[k] < -1 || lk[k] >= E) return bail(BQ_ERR_INVALID, "event link out of range");
        if (h->next_byt[e] >= 0 && h->prev_byt[h->next_byt[e]] != e) return bail(BQ_ERR_INVALID, "by-task links inconsistent");
        if (h->next_byq[e] >= 0 && (h->prev_byq[h->next_byq[e]] != e || h->qid[h->next_byq[e]] != h->qid[e]))
            return bail(BQ_ERR_INVALID, "by-queue links inconsistent");
        if (h->next_byt[e] >= 0 && h->qid[h->next_byt[e]] == h->qid[e])
            return bail(BQ_ERR_UNSUPPORTED, "a task visits the same queue twice in a row");
    }
    for (int q = 0; q < Q; ++q) {
        // random-order queues live on the sorted-departure arrays of processor sharing (bq_dynamic.cuh)
        if (h->q_type[q] == Q_PS || h->q_type[q] == Q_GG1R) h->has_ps = true;
        if (h->q_type[q] == Q_GG1R) h->has_gg1r = true;
        if (h->q_k[q] < 1) return bail(BQ_ERR_INVALID, "queue with fewer than one processor");
    }
    if (const char *env = getenv("BQ_FORCE_DYNAMIC"))
        if (env[0] == '1' && !h->has_mix) h->static_order = false;  // run a static-order network through the dynamic kernels

    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return bail(BQ_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
#define CKB(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return bail(BQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
    CKB(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CKB(cudaEventCreate(&h->ev0));
    CKB(cudaEventCreate(&h->ev1));

    build_schedule(h);

    // per-event tables and events grouped by queue (in by-queue order)
    std::vector<int> q_off(Q + 1, 0), q_events(E);
    for (int e = 0; e < E; ++e) q_off[h->qid[e] + 1]++;
    for (int q = 0; q < Q; ++q) q_off[q + 1] += q_off[q];
    h->q_off = q_off;
    {
        std::vector<int> fill(q_off.begin(), q_off.end() - 1);
        for (int e = 0; e < E; ++e) q_events[fill[h->qid[e]]++] = e;
    }
    static_assert(sizeof(SchedRec) == 48, "SchedRec must be 3 x int4");
    CKB(upload(&h->d_rec, h->recs.data(), h->recs.size() * 3));
    CKB(upload(&h->d_color_off, h->color_off.data(), h->color_off.size()));
    CKB(upload(&h->d_ev_q, h->qid.data(), E));
    CKB(upload(&h->d_ev_p, h->prev_byq.data(), E));
    CKB(upload(&h->d_ev_pt, h->prev_byt.data(), E));
    CKB(upload(&h->d_q_events, q_events.data(), E));
    CKB(upload(&h->d_q_off, q_off.data(), Q + 1));
    CKB(upload(&h->d_q_type, h->q_type.data(), Q));
    CKB(upload(&h->d_q_dist, h->q_dist.data(), Q));
    CKB(upload(&h->d_q_poff, h->q_poff.data(), Q));
    CKB(upload(&h->d_slot_dist, h->slot_dist.data(), h->NS));
    CKB(upload(&h->d_slot_poff, h->slot_poff.data(), h->NS));
    CKB(upload(&h->d_q_slot, h->q_slot.data(), Q + 1));
    CKB(upload(&h->d_q_mixoff, h->q_mixoff.data(), Q));
    if (h->has_mix) {
        CKB(cudaMalloc((void **)&h->d_aux, (size_t)n_chains * E));
        CKB(cudaMemset(h->d_aux, 0, (size_t)n_chains * E));
    }

    const size_t C = (size_t)n_chains;
    CKB(cudaMalloc((void **)&h->d_state, C * E * sizeof(double)));
    CKB(cudaMalloc((void **)&h->d_theta, C * h->P * sizeof(double)));
    CKB(cudaMalloc((void **)&h->d_stats, C * BQ_NSTAT * sizeof(long long)));
    CKB(cudaMemset(h->d_stats, 0, C * BQ_NSTAT * sizeof(long long)));
    {
        // replicate the initial state / parameters to every chain
        std::vector<double> rep;
        const size_t chunk = std::max<size_t>(1, std::min<size_t>(C, (64u << 20) / (E * sizeof(double))));
        rep.resize(chunk * E);
        for (size_t i = 0; i < chunk; ++i) memcpy(&rep[i * E], a->d, E * sizeof(double));
        for (size_t c0 = 0; c0 < C; c0 += chunk) {
            size_t nn = std::min(chunk, C - c0);
            CKB(cudaMemcpy(h->d_state + c0 * E, rep.data(), nn * E * sizeof(double), cudaMemcpyHostToDevice));
        }
        std::vector<double> th(C * h->P);
        for (size_t c = 0; c < C; ++c) memcpy(&th[c * h->P], theta0, h->P * sizeof(double));
        CKB(cudaMemcpy(h->d_theta, th.data(), th.size() * sizeof(double), cudaMemcpyHostToDevice));
    }

    {   // two warps per chain (rounds of 64 events) while the chains alone would leave half of the SM's warp slots empty;
        // slice sweeps only -- a handle that may run MH sweeps through these kernels keeps one warp per chain
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        h->batch_nw = (!h->static_order && !h->has_gg1r) ? 2 : 1;  // in waves of at most 8 sms chains; measured: profiles/r2_batch_nw.log
        if (const char *env = getenv("BQ_BATCH_NW")) h->batch_nw = (atoi(env) == 2 && !h->has_gg1r) ? 2 : 1;
    }
    if (int rc = build_sequential_order(h)) { std::string msg = g_err; bq_destroy(h); return fail(rc, msg); }
    if (!h->static_order)
        if (int rc = setup_dynamic(h, a->d)) { std::string msg = g_err; bq_destroy(h); return fail(rc, msg); }

    // shared-memory staging of the chain state when it fits (227 KB per CTA on sm_100a)
    int max_optin = 0;
    CKB(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    size_t need = static_smem_bytes(h, true);
    h->use_smem = need + 1024 <= (size_t)max_optin;
    h->smem_bytes = static_smem_bytes(h, h->use_smem);
    if (const char *env = getenv("BQ_FORCE_GLOBAL"))
        if (env[0] == '1') { h->use_smem = false; h->smem_bytes = static_smem_bytes(h, false); }
    if (const char *env = getenv("BQ_THREADS")) {
        int t = atoi(env);
        if (t >= 32 && t <= BQ_MAX_THREADS && t % 32 == 0) h->threads = t;
    }
    {
        auto set_max = [&](const void *fn) -> cudaError_t {
            cudaFuncAttributes fa;
            cudaError_t e = cudaFuncGetAttributes(&fa, fn);
            if (e != cudaSuccess) return e;
            return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - (int)fa.sharedSizeBytes);
        };
        CKB(set_max((const void *)sweep_static_kernel<true, true>));
        CKB(set_max((const void *)sweep_static_kernel<true, false>));
        CKB(set_max((const void *)sweep_static_kernel<false, true>));
        CKB(set_max((const void *)sweep_static_kernel<false, false>));
    }
    if (getenv("BQ_NO_EXP_FASTPATH")) h->all_exp = false;
    if (const char *env = getenv("BQ_MH_DYNAMIC")) h->mh_dynamic = (env[0] == '1') && !h->has_mix;
    if (const char *env = getenv("BQ_REFILL")) {
        int t = atoi(env);
        if (t >= 1 && t <= 32) h->refill = t;
    }
#undef CKB
    // thread-block clusters (bq_cluster.cuh): when one CTA's shared memory cannot hold the chain, or when there are too
    // few chains to fill the GPU with one CTA each.  BQ_CLUSTER=n forces n CTAs per chain (1 = never).
    if (h->static_order && !h->has_mix) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        // measured on B200 (tools/probe_cluster.py, profiles/r2_probe_cluster.jsonl): a 20 000-event chain sweeps in
        // 134 / 106 / 79 / 65 us on 1 / 2 / 4 / 8 CTAs, so clusters pay only while CTAs would otherwise be idle (and lose
        // against one CTA per chain as soon as the chains fill the SMs, even when the state then has to live in L2);
        // a 400-event chain is slower on a cluster (barrier cost exceeds the work)
        int want = 1;
        if ((int)h->recs.size() >= 2048) {
            if (h->C <= 8) want = 8;
            else if (h->C * 4 <= sms) want = 4;
            else if (h->C * 2 <= sms) want = 2;
        }
        if (const char *env = getenv("BQ_CLUSTER")) {
            const int n = atoi(env);
            if (n == 1 || n == 2 || n == 4 || n == 8) want = n;
        }
        if (getenv("BQ_FORCE_GLOBAL")) want = (getenv("BQ_CLUSTER") ? want : 1);
        if (want > 1 && cluster_smem_bytes(h, (((E + want - 1) / want) + 1) & ~1) + 1024 <= (size_t)max_optin) {
            if (int rc = build_cluster_plan(h, want)) { std::string msg = g_err; bq_destroy(h); return fail(rc, msg); }
        }
    }
    *out = h;
    return BQ_OK;
}

static int check_range(bq_handle *h, int lo, int hi) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (lo < 0 || hi > h->C || lo >= hi) return fail(BQ_ERR_INVALID, "bad chain range");
    return BQ_OK;
}

extern "C" int bq_set_params(bq_handle *h, int32_t lo, int32_t hi, const double *theta) {
    if (int r = check_range(h, lo, hi)) return r;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->d_theta + (size_t)lo * h->P, theta, (size_t)(hi - lo) * h->P * sizeof(double),
                       cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}
extern "C" int bq_get_params(bq_handle *h, int32_t lo, int32_t hi, double *theta) {
    if (int r = check_range(h, lo, hi)) return r;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(theta, h->d_theta + (size_t)lo * h->P, (size_t)(hi - lo) * h->P * sizeof(double),
                       cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}
extern "C" int bq_set_state(bq_handle *h, int32_t lo, int32_t hi, const double *d) {
    if (int r = check_range(h, lo, hi)) return r;
    CK(cudaSetDevice(h->device));
    if (!h->static_order) return dynamic_set_rows(h, lo, hi, d, false);
    h->dyn_stale = true;
    CK(cudaMemcpyAsync(h->d_state + (size_t)lo * h->E, d, (size_t)(hi - lo) * h->E * sizeof(double),
                       cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}

// One chain's departure times from the host, replicated to every chain on the device (E*8 bytes cross PCIe).
extern "C" int bq_broadcast_state(bq_handle *h, const double *d) {
    if (!h || !d) return fail(BQ_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    if (!h->static_order) return dynamic_set_rows(h, 0, h->C, d, true);
    h->dyn_stale = true;
    const size_t E = h->E, C = h->C;
    CK(cudaMemcpyAsync(h->d_state, d, E * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    for (size_t have = 1; have < C; have *= 2) {
        size_t n = std::min(have, C - have);
        CK(cudaMemcpyAsync(h->d_state + have * E, h->d_state, n * E * sizeof(double), cudaMemcpyDeviceToDevice,
                           h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}

// proc of the events of a "wide" FIFO queue (k >= #events, run as a delay station): the i-th event of the queue takes
// server i -- the first arg-min of a vector that still holds a zero (queues.pyx:1518-1527)
static void wide_procs(const bq_handle *h, int32_t *proc, size_t n_chains) {
    for (int q = 0; q < h->Q; ++q) {
        if (!h->q_wide[q]) continue;
        for (size_t c = 0; c < n_chains; ++c)
            for (int p = h->q_off[q]; p < h->q_off[q + 1]; ++p) proc[c * h->E + h->ord0[p]] = p - h->q_off[q];
    }
}

extern "C" int bq_get_state(bq_handle *h, int32_t lo, int32_t hi, double *d, double *a, double *s,
                            int32_t *proc, int32_t *prev_byq, int32_t *next_byq) {
    if (int r = check_range(h, lo, hi)) return r;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)(hi - lo), E = h->E;
    std::vector<double> tmp;
    double *dd = d;
    if (!dd) { tmp.resize(n * E); dd = tmp.data(); }
    CK(cudaMemcpyAsync(dd, h->d_state + (size_t)lo * E, n * E * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (!h->static_order) {
        if (!(a || s || proc || prev_byq || next_byq)) { CK(cudaStreamSynchronize(h->stream)); return BQ_OK; }
        std::vector<int> ord(n * E);
        std::vector<double> svc;
        CK(cudaMemcpyAsync(ord.data(), h->d_ord + (size_t)lo * E, n * E * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        if (h->has_ps) {
            svc.resize(n * E);
            CK(cudaMemcpyAsync(svc.data(), h->d_sq + (size_t)lo * E, n * E * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));
        const HostNet hn = h->net();
        for (size_t c = 0; c < n; ++c)
            host_derive(&hn, dd + c * E, &ord[c * E], a ? a + c * E : nullptr, s ? s + c * E : nullptr,
                        proc ? proc + c * E : nullptr, prev_byq ? prev_byq + c * E : nullptr,
                        next_byq ? next_byq + c * E : nullptr, h->has_ps ? &svc[c * E] : nullptr, h->ord0.data());
        if (proc) wide_procs(h, proc, n);
        return BQ_OK;
    }
    CK(cudaStreamSynchronize(h->stream));
    for (size_t c = 0; c < n; ++c) {
        const double *dc = dd + c * E;
        for (size_t e = 0; e < E; ++e) {
            int pt = h->prev_byt[e], p = h->prev_byq[e];
            double ae = (pt >= 0) ? dc[pt] : 0.0;
            if (a) a[c * E + e] = ae;
            if (s) {
                double cm = ae;
                if (h->q_type[h->qid[e]] == Q_GGK && p >= 0) cm = std::max(ae, dc[p]);
                s[c * E + e] = dc[e] - cm;
            }
            if (proc) proc[c * E + e] = 0;
            if (prev_byq) prev_byq[c * E + e] = h->prev_byq[e];
            if (next_byq) next_byq[c * E + e] = h->next_byq[e];
        }
    }
    if (proc) wide_procs(h, proc, n);
    return BQ_OK;
}

extern "C" int bq_set_aux(bq_handle *h, int32_t lo, int32_t hi, const uint8_t *aux) {
    if (int r = check_range(h, lo, hi)) return r;
    if (!h->d_aux || !aux) return fail(BQ_ERR_INVALID, "bq_set_aux: the handle has no mixture templates");
    CK(cudaSetDevice(h->device));
    std::vector<int> ncomp(h->E);
    for (int e = 0; e < h->E; ++e) ncomp[e] = h->q_slot[h->qid[e] + 1] - h->q_slot[h->qid[e]];
    for (size_t c = 0; c < (size_t)(hi - lo); ++c)
        for (int e = 0; e < h->E; ++e)
            if (aux[c * h->E + e] >= ncomp[e]) return fail(BQ_ERR_INVALID, "bq_set_aux: component index out of range");
    CK(cudaMemcpyAsync(h->d_aux + (size_t)lo * h->E, aux, (size_t)(hi - lo) * h->E, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}
extern "C" int bq_get_aux(bq_handle *h, int32_t lo, int32_t hi, uint8_t *aux) {
    if (int r = check_range(h, lo, hi)) return r;
    if (!h->d_aux || !aux) return fail(BQ_ERR_INVALID, "bq_get_aux: the handle has no mixture templates");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(aux, h->d_aux + (size_t)lo * h->E, (size_t)(hi - lo) * h->E, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}

extern "C" int bq_state_dev(bq_handle *h, double **d_dev, int64_t *chain_stride) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (d_dev) *d_dev = h->d_state;
    if (chain_stride) *chain_stride = h->E;
    return BQ_OK;
}

static int ensure_trace(bq_handle *h, int extra) {
    int need = h->trace_len + extra;
    if (need <= h->trace_cap) return BQ_OK;
    int cap = std::max(need, h->trace_cap * 2);
    const size_t C = h->C;
    double *nt = nullptr, *nw = nullptr, *nl = nullptr;
    CK(cudaMalloc((void **)&nt, (size_t)cap * C * h->P * sizeof(double)));
    CK(cudaMalloc((void **)&nw, (size_t)cap * C * h->Q * sizeof(double)));
    CK(cudaMalloc((void **)&nl, (size_t)cap * C * sizeof(double)));
    if (h->trace_len) {
        CK(cudaMemcpyAsync(nt, h->d_tr_theta, (size_t)h->trace_len * C * h->P * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(nw, h->d_tr_wait, (size_t)h->trace_len * C * h->Q * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(nl, h->d_tr_logp, (size_t)h->trace_len * C * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (h->d_tr_theta) cudaFree(h->d_tr_theta);
    if (h->d_tr_wait) cudaFree(h->d_tr_wait);
    if (h->d_tr_logp) cudaFree(h->d_tr_logp);
    h->d_tr_theta = nt; h->d_tr_wait = nw; h->d_tr_logp = nl;
    h->trace_cap = cap;
    return BQ_OK;
}

// A static-order handle entering the dynamic-order kernels (MH sweeps, proposal hook): build their arrays once, and
// refresh the free-time vectors whenever the static kernel has moved the departure times since.
static int ensure_dynamic(bq_handle *h) {
    if (!h->dyn_ready) {
        std::vector<double> d0(h->E);
        CK(cudaMemcpyAsync(d0.data(), h->d_state, h->E * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (int rc = setup_dynamic(h, d0.data())) return rc;
        h->dyn_stale = true;
    }
    if (h->dyn_stale) {
        DynArgs D;
        fill_dyn_args(h, D);
        const long long n = (long long)h->C * h->Q;
        if (h->kmax <= 4) rebuild_positions_kernel<4><<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(D, h->d_ev_pt);
        else rebuild_positions_kernel<8><<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(D, h->d_ev_pt);
        CK(cudaGetLastError());
        h->n_launches += 1;
        h->dyn_stale = false;
    }
    return BQ_OK;
}

// ---- cluster path: the chain state partitioned over the shared memories of ncta CTAs -------------------------------
static size_t cluster_smem_bytes(const bq_handle *h, int part) {
    size_t b = ((size_t)part + 34 * 8 + 8 + 64 + ((h->P + 1) & ~1)) * sizeof(double);
    b += (size_t)h->NS * (sizeof(QConst) + sizeof(SuffStat)) + (size_t)(2 * h->Q + 2 * h->NS + 2) * sizeof(int);
    return (b + 15) & ~(size_t)15;
}

static int build_cluster_plan(bq_handle *h, int ncta) {
    const int E = h->E, n = (int)h->recs.size(), n_colors = (int)h->color_off.size() - 1;
    int part = (E + ncta - 1) / ncta;
    part = (part + 1) & ~1;
    int k = 0;
    while ((1 << k) < part) ++k;
    auto enc = [&](int e) { return e < 0 ? -1 : (((e / part) << k) | (e % part)); };
    std::vector<SchedRec> crec;
    crec.reserve(n);
    std::vector<int> class_off((size_t)n_colors * (ncta + 1), 0);
    for (int c = 0; c < n_colors; ++c)
        for (int r = 0; r <= ncta; ++r) {
            class_off[(size_t)c * (ncta + 1) + r] = (int)crec.size();
            if (r == ncta) break;
            for (int i = h->color_off[c]; i < h->color_off[c + 1]; ++i) {
                const SchedRec &s = h->recs[i];
                if (s.e / part != r) continue;
                SchedRec t = s;
                t.e = enc(s.e); t.pt_e = enc(s.pt_e); t.p = enc(s.p); t.n = enc(s.n); t.pt_n = enc(s.pt_n);
                t.e1 = enc(s.e1); t.p1 = enc(s.p1); t.pt_p1 = enc(s.pt_p1); t.pt_n1 = enc(s.pt_n1);
                t.pad = s.e;
                crec.push_back(t);
            }
        }
    std::vector<int> qe_idx(E), qe_p(E), qe_pt(E), fill(h->q_off.begin(), h->q_off.end() - 1);
    for (int e = 0; e < E; ++e) {
        const int j = fill[h->qid[e]]++;
        qe_idx[j] = enc(e); qe_p[j] = enc(h->prev_byq[e]); qe_pt[j] = enc(h->prev_byt[e]);
    }
    CK(upload(&h->d_crec, crec.data(), crec.size() * 3));
    CK(upload(&h->d_class_off, class_off.data(), class_off.size()));
    CK(upload(&h->d_qe_idx, qe_idx.data(), (size_t)E));
    CK(upload(&h->d_qe_p, qe_p.data(), (size_t)E));
    CK(upload(&h->d_qe_pt, qe_pt.data(), (size_t)E));
    h->ncta = ncta; h->cpart = part; h->ck = k;
    h->csmem = cluster_smem_bytes(h, part);
    return BQ_OK;
}

static int launch_cluster(bq_handle *h, const SweepArgs &A) {
    ClusterArgs CA;
    CA.A = A;
    CA.A.rec = h->d_crec;
    CA.ncta = h->ncta; CA.part = h->cpart; CA.k = h->ck;
    CA.class_off = h->d_class_off; CA.qe_idx = h->d_qe_idx; CA.qe_p = h->d_qe_p; CA.qe_pt = h->d_qe_pt;
    const void *fn = h->all_exp ? (const void *)sweep_cluster_kernel<true> : (const void *)sweep_cluster_kernel<false>;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->csmem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(h->C * h->ncta), 1, 1);
    cfg.blockDim = dim3((unsigned)h->threads, 1, 1);
    cfg.dynamicSmemBytes = h->csmem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)h->ncta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (h->all_exp) CK(cudaLaunchKernelEx(&cfg, sweep_cluster_kernel<true>, CA));
    else CK(cudaLaunchKernelEx(&cfg, sweep_cluster_kernel<false>, CA));
    return BQ_OK;
}

static int launch_sweep(bq_handle *h, int32_t n_sweeps, int32_t mode, int32_t update, uint32_t flags) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (n_sweeps < 0) return fail(BQ_ERR_INVALID, "n_sweeps < 0");
    if (mode != BQ_MODE_SLICE && mode != BQ_MODE_MH) return fail(BQ_ERR_INVALID, "bad sweep mode");
    if (mode == BQ_MODE_MH && h->has_ps)
        return fail(BQ_ERR_UNSUPPORTED, "MH sweeps need arrivalLik, which PS and GG1R queues do not have (queues.pyx:1140, 856): use BQ_MODE_SLICE");
    if (mode == BQ_MODE_MH && h->has_mix)
        return fail(BQ_ERR_UNSUPPORTED, "MH sweeps: the reference's proposals take a DistributionTemplate only (queues.pyx:394-442), "
                                        "not a mixture: use BQ_MODE_SLICE");
    if (update < 0 || update > 2) return fail(BQ_ERR_INVALID, "bad param_update");
    if (n_sweeps == 0) return BQ_OK;
    CK(cudaSetDevice(h->device));
    if (flags & BQ_FLAG_TRACE)
        if (int r = ensure_trace(h, n_sweeps)) return r;
    SweepArgs A;
    fill_args(h, A);
    A.n_sweeps = n_sweeps; A.update = update; A.flags = (int)flags; A.trace_base = h->trace_len; A.mode = mode;
    // MH sweeps of a static-order network run in the chromatic kernel too (closed-form proposal); BQ_MH_DYNAMIC=1
    // sends them through the sequential dynamic-order kernel instead (cross-check of the two implementations)
    const bool use_dynamic = !h->static_order || (mode == BQ_MODE_MH && !(flags & FLAG_PARAMS_ONLY) && h->mh_dynamic);
    if (use_dynamic && h->static_order)
        if (int rc = ensure_dynamic(h)) return rc;
    CK(cudaEventRecord(h->ev0, h->stream));
    if (use_dynamic) {
        // The batch kernels stamp their claims with round * events-per-round in an int, and a round commits at least one
        // event: a launch may hold at most 2^31 / (64 * n_order) sweeps (699 at config 3's size -- minutes of GPU time).
        // Longer calls run as several launches; Tmax (2 max d, qnet.pyx:672) is then taken at the start of each launch
        // instead of once per call unless BQ_FLAG_TMAX_PER_SWEEP asks for every sweep anyway (estimation.bayes does).
        const long long n_ord = std::max<long long>(1, (long long)h->dyn_order.size());
        int per_launch = (int)std::max<long long>(1, std::min<long long>(n_sweeps, ((1LL << 31) - 1) / (64 * n_ord)));
        if (const char *env = getenv("BQ_MAX_SWEEPS_PER_LAUNCH")) per_launch = std::max(1, std::min(per_launch, atoi(env)));  // test hook
        for (int done = 0; done < n_sweeps; done += per_launch) {
            DynArgs D;
            fill_dyn_args(h, D);
            D.n_sweeps = std::min(per_launch, n_sweeps - done);
            D.update = update; D.flags = (int)flags; D.mode = mode;
            D.trace_base = h->trace_len + ((flags & BQ_FLAG_TRACE) ? done : 0);
            D.sweep_base = h->sweep_counter + (unsigned int)done;
            launch_dynamic(h, D);
            if (done + per_launch < n_sweeps) h->n_launches += 1;
        }
    } else if (h->ncta > 1) {
        h->dyn_stale = true;
        if (int rc = launch_cluster(h, A)) return rc;
    } else if (h->use_smem) {
        h->dyn_stale = true;
        if (h->all_exp) sweep_static_kernel<true, true><<<h->C, h->threads, h->smem_bytes, h->stream>>>(A);
        else sweep_static_kernel<true, false><<<h->C, h->threads, h->smem_bytes, h->stream>>>(A);
    } else {
        h->dyn_stale = true;
        if (h->all_exp) sweep_static_kernel<false, true><<<h->C, h->threads, h->smem_bytes, h->stream>>>(A);
        else sweep_static_kernel<false, false><<<h->C, h->threads, h->smem_bytes, h->stream>>>(A);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->n_launches += 1;
    if (!(flags & FLAG_PARAMS_ONLY)) h->n_sweeps += n_sweeps;
    h->sweep_counter += (unsigned int)n_sweeps;
    if (flags & BQ_FLAG_TRACE) h->trace_len += n_sweeps;
    return BQ_OK;
}

extern "C" int bq_sweep(bq_handle *h, int32_t n_sweeps, int32_t mode, int32_t update, uint32_t flags) {
    return launch_sweep(h, n_sweeps, mode, update, flags & 0xFFFFu);
}

extern "C" int bq_update_params(bq_handle *h, int32_t update) {
    if (update != BQ_UPD_BAYES && update != BQ_UPD_STEM) return fail(BQ_ERR_INVALID, "bad param_update");
    return launch_sweep(h, 1, BQ_MODE_SLICE, update, FLAG_PARAMS_ONLY);
}

// Parameter update from sufficient statistics supplied by the caller instead of the chains' own state: what
// Qnet.estimate / Qnet.sample_parameters do when they are given a LIST of Arrivals (qnet.pyx:844-858, 882-896: the
// events of all of them are pooled per queue).  One thread per (chain, queue).
__global__ void __launch_bounds__(128) update_from_suffstats_kernel(int C, int Q, int P, const int *q_dist, const int *q_poff,
                                                                    const double *ss_in, int n_ss, double *theta, int update,
                                                                    unsigned long long seed, long long chain_offset,
                                                                    unsigned int sweep, long long *stats) {
    // one warp per chain, the statistics staged in shared memory: the same shape as the fused update of the sweep kernels
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int gib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + gib;
    if (chain >= C) return;
    SuffStat *ss = reinterpret_cast<SuffStat *>(smem_raw) + (size_t)gib * Q;
    const double *src = ss_in + (size_t)(n_ss > 1 ? chain : 0) * Q * 8;
    for (int i = lane; i < Q * 8; i += 32) reinterpret_cast<double *>(ss)[i] = src[i];
    __syncwarp();
    double *th = theta + (size_t)chain * P;
    long long n_pevals = 0;
    for (int q = lane; q < Q; q += 32) {
        const int o = q_poff[q], dist = q_dist[q];
        double p0 = th[o], p1 = (dist == D_EXP) ? 0.0 : th[o + 1];
        if (update == UPD_BAYES) {
            Rng prng;
            prng.init(seed, (unsigned long long)(chain_offset + chain), sweep, (unsigned int)q, TAG_PARAM);
            sample_params(dist, ss[q], prng, p0, p1, &n_pevals);
        } else {
            estimate_params(dist, ss[q], p0, p1);
        }
        th[o] = p0;
        if (dist != D_EXP) th[o + 1] = p1;
    }
    n_pevals = (long long)group_sum<32>((double)n_pevals, 0xffffffffu);
    if (lane == 0) stats[(size_t)chain * BQ_NSTAT + 2] += n_pevals;
}

extern "C" int bq_update_params_from(bq_handle *h, int32_t update, const double *suffstats, int32_t n_ss) {
    if (!h || !suffstats) return fail(BQ_ERR_INVALID, "null argument");
    if (update != BQ_UPD_BAYES && update != BQ_UPD_STEM) return fail(BQ_ERR_INVALID, "bad param_update");
    if (n_ss != 1 && n_ss != h->C) return fail(BQ_ERR_INVALID, "n_ss must be 1 or the number of chains");
    if (h->has_mix) return fail(BQ_ERR_UNSUPPORTED, "bq_update_params_from: mixture templates");
    CK(cudaSetDevice(h->device));
    double *d_ss = nullptr;
    const size_t n = (size_t)n_ss * h->Q * 8;
    CK(cudaMalloc((void **)&d_ss, n * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_ss, suffstats, n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        const size_t sm = 4 * (size_t)h->Q * sizeof(SuffStat);
        smem_optin((const void *)update_from_suffstats_kernel, sm);
        update_from_suffstats_kernel<<<(h->C + 3) / 4, 128, sm, h->stream>>>(
            h->C, h->Q, h->P, h->d_q_dist, h->d_q_poff, d_ss, n_ss, h->d_theta, update, h->seed, h->chain_offset,
            h->sweep_counter, h->d_stats);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_ss);
    if (e != cudaSuccess) return fail(BQ_ERR_CUDA, cudaGetErrorString(e));
    h->n_launches += 1;
    h->sweep_counter += 1;
    return BQ_OK;
}

// Plan of the per-chain initialiser (init_chains_kernel): a topological order of the resample-eligible events under
// "lower neighbour before upper neighbour" and, for each, the upper bound the fixed (observed / determined) times
// impose on it through its upper neighbours.  d0: the handle's current chain-0 state (only fixed entries are read).
static int build_init_plan(bq_handle *h, const double *d0) {
    const int n = (int)h->recs.size();
    std::vector<int> slot(h->E, -1);
    for (int i = 0; i < n; ++i) slot[h->recs[i].e] = i;
    auto lowers = [&](const SchedRec &r, int out[3]) {
        int m = 0;
        if (r.pt_e >= 0) out[m++] = r.pt_e;
        if (h->q_type[r.q0] == Q_GGK && r.p >= 0) out[m++] = r.p;
        if (r.e1 >= 0 && h->q_type[r.q1] == Q_GGK && r.pt_p1 >= 0) out[m++] = r.pt_p1;
        return m;
    };
    auto uppers = [&](const SchedRec &r, int out[3]) {
        int m = 0;
        if (h->q_type[r.q0] == Q_GGK && r.n >= 0) out[m++] = r.n;
        if (r.e1 >= 0) out[m++] = r.e1;
        if (r.e1 >= 0 && h->q_type[r.q1] == Q_GGK && r.pt_n1 >= 0) out[m++] = r.pt_n1;
        return m;
    };
    std::vector<int> indeg(n, 0), topo;
    std::vector<std::vector<int>> succ(n);
    for (int i = 0; i < n; ++i) {
        int lo[3];
        const int m = lowers(h->recs[i], lo);
        for (int k = 0; k < m; ++k)
            if (slot[lo[k]] >= 0 && slot[lo[k]] != i) { succ[slot[lo[k]]].push_back(i); indeg[i]++; }
    }
    topo.reserve(n);
    std::vector<int> stack;
    for (int i = n - 1; i >= 0; --i)
        if (!indeg[i]) stack.push_back(i);
    while (!stack.empty()) {
        const int i = stack.back();
        stack.pop_back();
        topo.push_back(i);
        for (int j : succ[i])
            if (--indeg[j] == 0) stack.push_back(j);
    }
    if ((int)topo.size() != n) return fail(BQ_ERR_STATE, "bq_init_chains: the ordering constraints of the state are cyclic");
    std::vector<double> ub(n, INFINITY);
    for (int t = n - 1; t >= 0; --t) {
        const int i = topo[t];
        int up[3];
        const int m = uppers(h->recs[i], up);
        double b = INFINITY;
        for (int k = 0; k < m; ++k) b = std::min(b, slot[up[k]] >= 0 ? ub[slot[up[k]]] : d0[up[k]]);
        ub[i] = b;
    }
    CK(upload(&h->d_init_topo, topo.data(), topo.size()));
    CK(upload(&h->d_init_ub, ub.data(), ub.size()));
    return BQ_OK;
}

extern "C" int bq_init_chains(bq_handle *h, double sigma) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (!(sigma >= 0.0)) return fail(BQ_ERR_INVALID, "bq_init_chains: sigma must be >= 0");
    if (!h->static_order)
        return fail(BQ_ERR_UNSUPPORTED, "bq_init_chains: per-chain starts are built for static-order networks (single-server "
                                        "FIFO queues and delay stations); multi-server and PS networks share one start");
    CK(cudaSetDevice(h->device));
    if (!h->d_init_topo) {
        std::vector<double> d0(h->E);
        CK(cudaMemcpyAsync(d0.data(), h->d_state, (size_t)h->E * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (int rc = build_init_plan(h, d0.data())) return rc;
    }
    InitArgs I;
    memset(&I, 0, sizeof(I));
    I.E = h->E; I.Q = h->Q; I.P = h->P; I.n_chains = h->C; I.n = (int)h->recs.size();
    I.rec = h->d_rec; I.topo = h->d_init_topo; I.ub = h->d_init_ub;
    I.q_type = h->d_q_type; I.q_dist = h->d_slot_dist; I.q_poff = h->d_slot_poff; I.q_slot = h->d_q_slot;
    I.state = h->d_state; I.theta = h->d_theta; I.sigma = sigma;
    I.seed = h->seed; I.chain_offset = h->chain_offset; I.counter = h->init_counter++;
    init_chains_kernel<<<(h->C + 63) / 64, 64, 0, h->stream>>>(I);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->n_launches += 1;
    h->dyn_stale = true;
    return BQ_OK;
}

extern "C" int bq_synchronize(bq_handle *h) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}

// The sweep index that the next bq_sweep call will use as its first counter value: together with the seed, the chain
// ids, the departure times and the parameters it determines everything that follows (checkpoint / resume).
// The soft start-on-arrival rule of random-order queues for repair sweeps (gg1r_member, bq_dynamic.cuh): the per-queue
// table goes to the device again with k = 2 for those queues.
extern "C" int bq_set_relaxed(bq_handle *h, int32_t on) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (!h->has_gg1r) return fail(BQ_ERR_UNSUPPORTED, "bq_set_relaxed: the network has no random-order (GG1R) queue");
    CK(cudaSetDevice(h->device));
    std::vector<QInfo> qinfo(h->Q);
    for (int q = 0; q < h->Q; ++q) {
        qinfo[q].type = h->q_type[q]; qinfo[q].k = h->q_k[q]; qinfo[q].lo = h->q_off[q]; qinfo[q].hi = h->q_off[q + 1];
        qinfo[q].foff = h->q_foff[q]; qinfo[q].dist = h->q_dist[q];
        if (h->q_type[q] == Q_GG1R) qinfo[q].k = on ? 2 : 1;
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(h->d_qinfo, qinfo.data(), qinfo.size() * sizeof(QInfo), cudaMemcpyHostToDevice));
    return BQ_OK;
}

extern "C" int bq_get_sweep_counter(bq_handle *h, uint32_t *counter) {
    if (!h || !counter) return fail(BQ_ERR_INVALID, "null argument");
    *counter = h->sweep_counter;
    return BQ_OK;
}
extern "C" int bq_set_sweep_counter(bq_handle *h, uint32_t counter) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    h->sweep_counter = counter;
    return BQ_OK;
}

extern "C" int bq_last_sweep_ms(bq_handle *h, float *ms) {
    if (!h || !ms) return fail(BQ_ERR_INVALID, "null argument");
    if (!h->timed) return fail(BQ_ERR_INVALID, "no sweep has been launched");
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return BQ_OK;
}

extern "C" int bq_stream(bq_handle *h, void **stream) {
    if (!h || !stream) return fail(BQ_ERR_INVALID, "null argument");
    *stream = (void *)h->stream;
    return BQ_OK;
}

static int eval_hook(bq_handle *h, int32_t chain, int32_t eid, const double *x, int32_t n, double *out, int what) {
    if (int r = check_range(h, chain, chain + 1)) return r;
    if (what == 1 && h->has_mix)
        return fail(BQ_ERR_UNSUPPORTED, "bq_mh_proposal: not built for mixture templates");
    if (what == 1 && h->has_ps)
        return fail(BQ_ERR_UNSUPPORTED, "PS queues have no arrivalLik (queues.pyx:1140): no MH proposal");
    if (eid < 0 || eid >= h->E || n < 0) return fail(BQ_ERR_INVALID, "bad event id / n");
    if (n == 0) return BQ_OK;
    CK(cudaSetDevice(h->device));
    if (h->scratch_n < 2 * n) {
        if (h->d_scratch) cudaFree(h->d_scratch);
        h->d_scratch = nullptr;
        CK(cudaMalloc((void **)&h->d_scratch, (size_t)2 * n * sizeof(double)));
        h->scratch_n = 2 * n;
    }
    CK(cudaMemcpyAsync(h->d_scratch, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (what == 1 && h->static_order)
        if (int rc = ensure_dynamic(h)) return rc;
    if (!h->static_order || what == 1) {
        DynArgs D;
        fill_dyn_args(h, D);
        const size_t sm = dyn_group_smem(h->Q, h->P);
        const int e1 = h->next_byt[eid];
        const DynRec rec = {eid, e1, h->qid[eid], e1 >= 0 ? h->qid[e1] : -1};
        smem_optin((const void *)logcond_dynamic_kernel<4>, sm);
        smem_optin((const void *)logcond_dynamic_kernel<8>, sm);
        if (h->kmax <= 4)
            logcond_dynamic_kernel<4><<<1, 32, sm, h->stream>>>(D, chain, rec, h->d_scratch, n, h->d_scratch + n, what);
        else
            logcond_dynamic_kernel<8><<<1, 32, sm, h->stream>>>(D, chain, rec, h->d_scratch, n, h->d_scratch + n, what);
    } else {
        SweepArgs A;
        fill_args(h, A);
        SchedRec r = make_rec(h, eid);
        logcond_static_kernel<<<1, 128, 0, h->stream>>>(A, chain, r, h->d_scratch, n, h->d_scratch + n);
    }
    CK(cudaGetLastError());
    h->n_launches += 1;
    CK(cudaMemcpyAsync(out, h->d_scratch + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}

extern "C" int bq_logcond(bq_handle *h, int32_t chain, int32_t eid, const double *x, int32_t n, double *out) {
    return eval_hook(h, chain, eid, x, n, out, 0);
}
extern "C" int bq_mh_proposal(bq_handle *h, int32_t chain, int32_t eid, const double *x, int32_t n, double *out) {
    return eval_hook(h, chain, eid, x, n, out, 1);
}

static int run_report(bq_handle *h, double *logp_host, double *ss_host) {
    CK(cudaSetDevice(h->device));
    const size_t C = h->C;
    double *d_lp = nullptr, *d_ss = nullptr;
    CK(cudaMalloc((void **)&d_lp, C * sizeof(double)));
    cudaError_t e = cudaMalloc((void **)&d_ss, C * h->NS * 8 * sizeof(double));
    if (e != cudaSuccess) { cudaFree(d_lp); return fail(BQ_ERR_CUDA, cudaGetErrorString(e)); }
    if (!h->static_order) {
        DynArgs D;
        fill_dyn_args(h, D);
        smem_optin((const void *)report_dynamic_kernel, 4 * dyn_group_smem(h->Q, h->P));
        report_dynamic_kernel<<<(h->C + 3) / 4, 128, 4 * dyn_group_smem(h->Q, h->P), h->stream>>>(D, d_lp, d_ss);
    } else {
        SweepArgs A;
        fill_args(h, A);
        size_t sm = static_smem_bytes(h, false);
        report_static_kernel<<<h->C, 256, sm, h->stream>>>(A, d_lp, d_ss);
    }
    e = cudaGetLastError();
    h->n_launches += 1;
    if (e == cudaSuccess && logp_host)
        e = cudaMemcpyAsync(logp_host, d_lp, C * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && ss_host)
        e = cudaMemcpyAsync(ss_host, d_ss, C * h->NS * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_lp);
    cudaFree(d_ss);
    if (e != cudaSuccess) return fail(BQ_ERR_CUDA, cudaGetErrorString(e));
    return BQ_OK;
}

extern "C" int bq_log_prob(bq_handle *h, double *out) {
    if (!h || !out) return fail(BQ_ERR_INVALID, "null argument");
    return run_report(h, out, nullptr);
}
extern "C" int bq_suffstats(bq_handle *h, double *out) {
    if (!h || !out) return fail(BQ_ERR_INVALID, "null argument");
    return run_report(h, nullptr, out);
}

// dynamic-order networks: the kernel commits with the sweep kernels' own code, in place; every per-chain array is
// saved before and put back after
static int gibbs_kernel_dynamic_path(bq_handle *h, const double *d_to, int32_t n_to, int32_t strict, double *out) {
    const size_t E = h->E, C = h->C, n_order = h->dyn_order.size();
    struct Saved { void *dev; void *copy; size_t bytes; };
    std::vector<Saved> saved = {{h->d_state, nullptr, C * E * sizeof(double)},
                                {h->d_prec, nullptr, C * E * h->recS * sizeof(double)},
                                {h->d_ord, nullptr, C * E * sizeof(int)},
                                {h->d_pos, nullptr, C * E * sizeof(int)},
                                {h->d_flag, nullptr, C * (size_t)h->FW}};
    if (h->has_ps) {
        saved.push_back({h->d_dord, nullptr, C * E * sizeof(int)});
        saved.push_back({h->d_dpos, nullptr, C * E * sizeof(int)});
        saved.push_back({h->d_dt, nullptr, C * E * sizeof(double)});
        saved.push_back({h->d_sq, nullptr, C * E * sizeof(double)});
    }
    double *to = nullptr, *res = nullptr;
    unsigned char *done = nullptr;
    bool dirty = false;  // the kernel has (possibly) overwritten the live arrays and they are not restored yet
    auto cleanup = [&]() {
        if (dirty) {  // error after the launch: put the chains back before the copies go (best effort)
            cudaGetLastError();
            for (auto &sv : saved)
                if (sv.copy) cudaMemcpyAsync(sv.dev, sv.copy, sv.bytes, cudaMemcpyDeviceToDevice, h->stream);
            cudaStreamSynchronize(h->stream);
        }
        for (auto &sv : saved) cudaFree(sv.copy);
        cudaFree(to); cudaFree(res); cudaFree(done);
    };
#define CKG(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::string m_ = cudaGetErrorString(e_); cleanup(); return fail(BQ_ERR_CUDA, m_); } } while (0)
    for (auto &sv : saved) {
        CKG(cudaMalloc(&sv.copy, sv.bytes));
        CKG(cudaMemcpyAsync(sv.copy, sv.dev, sv.bytes, cudaMemcpyDeviceToDevice, h->stream));
    }
    CKG(cudaMalloc((void **)&to, (size_t)n_to * E * sizeof(double)));
    CKG(cudaMalloc((void **)&res, C * sizeof(double)));
    CKG(cudaMalloc((void **)&done, std::max<size_t>(C * n_order, 1)));
    CKG(cudaMemcpyAsync(to, d_to, (size_t)n_to * E * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CKG(cudaMemsetAsync(done, 0, std::max<size_t>(C * n_order, 1), h->stream));
    DynArgs D;
    fill_dyn_args(h, D);
    const int G = 4;
    const size_t sm = (size_t)G * dyn_group_smem(h->Q, h->P);
    const unsigned blocks = (unsigned)((C + G - 1) / G);
    smem_optin((const void *)gibbs_kernel_dynamic<4>, sm);
    smem_optin((const void *)gibbs_kernel_dynamic<8>, sm);
    if (h->kmax <= 4) gibbs_kernel_dynamic<4><<<blocks, 32 * G, sm, h->stream>>>(D, to, n_to, strict ? 1 : 0, done, res);
    else gibbs_kernel_dynamic<8><<<blocks, 32 * G, sm, h->stream>>>(D, to, n_to, strict ? 1 : 0, done, res);
    dirty = true;
    CKG(cudaGetLastError());
    for (auto &sv : saved) CKG(cudaMemcpyAsync(sv.dev, sv.copy, sv.bytes, cudaMemcpyDeviceToDevice, h->stream));
    CKG(cudaMemcpyAsync(out, res, C * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CKG(cudaStreamSynchronize(h->stream));
    dirty = false;
#undef CKG
    cleanup();
    h->n_launches += 1;
    return BQ_OK;
}

// Qnet.gibbs_kernel (qnet.pyx:741-809) for every chain: log T(d_to <- state of the chain) of one slice sweep
extern "C" int bq_gibbs_kernel(bq_handle *h, const double *d_to, int32_t n_to, int32_t strict, double *out) {
    if (!h || !d_to || !out) return fail(BQ_ERR_INVALID, "null argument");
    if (n_to != 1 && n_to != h->C) return fail(BQ_ERR_INVALID, "n_to must be 1 or the number of chains");
    if (h->has_mix) return fail(BQ_ERR_UNSUPPORTED, "bq_gibbs_kernel: mixture templates (the reference's parameter_kernel raises too, queues.pyx:1453)");
    if (h->has_gg1r) return fail(BQ_ERR_UNSUPPORTED, "bq_gibbs_kernel: not built for random-order (GG1R) queues");
    CK(cudaSetDevice(h->device));
    if (!h->static_order) return gibbs_kernel_dynamic_path(h, d_to, n_to, strict, out);
    const size_t E = h->E, C = h->C, n_sched = h->order.size();
    double *work = nullptr, *to = nullptr, *res = nullptr;
    unsigned char *done = nullptr;
    int rc = BQ_OK;
    auto cleanup = [&]() { cudaFree(work); cudaFree(to); cudaFree(res); cudaFree(done); };
#define CKG(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); return fail(BQ_ERR_CUDA, cudaGetErrorString(e_)); } } while (0)
    CKG(cudaMalloc((void **)&work, C * E * sizeof(double)));
    CKG(cudaMalloc((void **)&to, (size_t)n_to * E * sizeof(double)));
    CKG(cudaMalloc((void **)&res, C * sizeof(double)));
    CKG(cudaMalloc((void **)&done, std::max<size_t>(C * n_sched, 1)));
    CKG(cudaMemcpyAsync(work, h->d_state, C * E * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CKG(cudaMemcpyAsync(to, d_to, (size_t)n_to * E * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CKG(cudaMemsetAsync(done, 0, std::max<size_t>(C * n_sched, 1), h->stream));
    SweepArgs A;
    fill_args(h, A);
    const size_t sm = (34 * 8 + 8 + ((h->P + 1) & ~1)) * sizeof(double) + h->Q * (sizeof(QConst) + sizeof(int)) + 16;
    gibbs_kernel_static<<<(unsigned)C, 256, sm, h->stream>>>(A, work, done, to, n_to, strict ? 1 : 0, res);
    CKG(cudaGetLastError());
    CKG(cudaMemcpyAsync(out, res, C * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CKG(cudaStreamSynchronize(h->stream));
#undef CKG
    cleanup();
    h->n_launches += 1;
    return rc;
}

extern "C" int bq_stats(bq_handle *h, bq_stats_t *out) {
    if (!h || !out) return fail(BQ_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    std::vector<long long> st((size_t)h->C * BQ_NSTAT);
    CK(cudaMemcpyAsync(st.data(), h->d_stats, st.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memset(out, 0, sizeof(*out));
    for (int c = 0; c < h->C; ++c) {
        const long long *r = &st[(size_t)c * BQ_NSTAT];
        out->nevents += r[0];
        out->slice_evals += r[1];
        out->param_evals += r[2];
        out->n_stuck += r[3];
        out->nreject += r[4];
        out->diff_len += r[5];
        out->n_redone += r[6];
    }
    out->slice_calls = out->nevents;
    out->num_diff = out->slice_evals;
    out->n_sweeps = h->n_sweeps;
    out->n_launches = h->n_launches;
    return BQ_OK;
}
extern "C" int bq_reset_stats(bq_handle *h) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->d_stats, 0, (size_t)h->C * BQ_NSTAT * sizeof(long long), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->n_sweeps = 0;
    return BQ_OK;
}

extern "C" int bq_trace_len(bq_handle *h, int32_t *n_rec) {
    if (!h || !n_rec) return fail(BQ_ERR_INVALID, "null argument");
    *n_rec = h->trace_len;
    return BQ_OK;
}
extern "C" int bq_get_trace(bq_handle *h, double *theta, double *mean_wait, double *logp) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (!h->trace_len) return BQ_OK;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->trace_len * h->C;
    if (theta) CK(cudaMemcpyAsync(theta, h->d_tr_theta, n * h->P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (mean_wait) CK(cudaMemcpyAsync(mean_wait, h->d_tr_wait, n * h->Q * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (logp) CK(cudaMemcpyAsync(logp, h->d_tr_logp, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return BQ_OK;
}
extern "C" int bq_clear_trace(bq_handle *h) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    h->trace_len = 0;
    return BQ_OK;
}
extern "C" int bq_trace_dev(bq_handle *h, double **theta_dev, double **mean_wait_dev, double **logp_dev) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (theta_dev) *theta_dev = h->d_tr_theta;
    if (mean_wait_dev) *mean_wait_dev = h->d_tr_wait;
    if (logp_dev) *logp_dev = h->d_tr_logp;
    return BQ_OK;
}

extern "C" int bq_get_schedule(bq_handle *h, int32_t *n_colors, int32_t *n_eligible, int32_t *color_off,
                               int32_t *order) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (n_colors) *n_colors = (int)h->color_off.size() - 1;
    if (n_eligible) *n_eligible = (int)h->order.size();
    if (color_off) memcpy(color_off, h->color_off.data(), h->color_off.size() * sizeof(int));
    if (order) memcpy(order, h->order.data(), h->order.size() * sizeof(int));
    return BQ_OK;
}

extern "C" int bq_get_sequential_order(bq_handle *h, int32_t *n_eligible, int32_t *order) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (n_eligible) *n_eligible = (int)h->dyn_order.size();
    if (order) memcpy(order, h->dyn_order.data(), h->dyn_order.size() * sizeof(int));
    return BQ_OK;
}

extern "C" int bq_get_mh_order(bq_handle *h, int32_t *n_eligible, int32_t *order) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    const std::vector<int> &o = (h->static_order && !h->mh_dynamic) ? h->order : h->dyn_order;
    if (n_eligible) *n_eligible = (int)o.size();
    if (order) memcpy(order, o.data(), o.size() * sizeof(int));
    return BQ_OK;
}

// Pure host function (no device needed): the uniforms exactly as the kernels draw them.
extern "C" int bq_rng_uniforms(uint64_t seed, int64_t chain, uint32_t sweep, uint32_t id, uint32_t tag,
                               int32_t n, double *out) {
    if (!out || n < 0) return fail(BQ_ERR_INVALID, "bad argument");
    Rng r;
    r.init(seed, (uint64_t)chain, sweep, id, tag);
    for (int i = 0; i < n; ++i) out[i] = r.uniform();
    return BQ_OK;
}

extern "C" int bq_rng_slice_uniforms(uint64_t seed, int64_t chain, uint32_t sweep, uint32_t id, uint32_t tag,
                                     int32_t n, double *out) {
    if (!out || n < 0) return fail(BQ_ERR_INVALID, "bad argument");
    Rng r;
    r.init(seed, (uint64_t)chain, sweep, id, tag);
    for (int i = 0; i < n; ++i) out[i] = (i < 2) ? r.uniform() : r.uniform32();
    return BQ_OK;
}

extern "C" int bq_philox_block(const uint32_t ctr[4], const uint32_t key[2], int32_t rounds, uint32_t out[4]) {
    if (!ctr || !key || !out) return fail(BQ_ERR_INVALID, "null argument");
    if (rounds == 7) philox4x32<7>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
    else if (rounds == 10) philox4x32<10>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
    else return fail(BQ_ERR_INVALID, "bq_philox_block: rounds must be 7 or 10");
    return BQ_OK;
}

extern "C" int bq_cluster_size(bq_handle *h, int32_t *ncta) {
    if (!h || !ncta) return fail(BQ_ERR_INVALID, "null argument");
    *ncta = h->ncta;
    return BQ_OK;
}

extern "C" int bq_kernel_info(bq_handle *h, int32_t *threads, int32_t *smem_bytes, int32_t *state_in_smem) {
    if (!h) return fail(BQ_ERR_INVALID, "null handle");
    if (threads) *threads = h->threads;
    if (smem_bytes) *smem_bytes = (int)h->smem_bytes;
    if (state_in_smem) *state_in_smem = h->use_smem ? 1 : 0;
    return BQ_OK;
}
