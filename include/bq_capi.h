/*
 * bq_capi.h — C ABI of the B200-native bayes-qnet hot path (libbq_b200.so).
 *
 * The reference (casutton/bayes-qnet) has no FFI layer: its hot path sits behind Cython `cdef`
 * virtual interfaces.  Each entry point below replaces the reference interface cited beside it
 * (paths relative to the reference's src/).  All buffers are caller-owned HOST memory unless the
 * name ends in `_dev`; the library copies in/out.  Every function returns 0 on success or a
 * negative bq_status; the message is available from bq_last_error().  No exception crosses the
 * boundary.  One handle = one CUDA device + one stream; a handle is not thread-safe, different
 * handles are independent.
 *
 * Everything numeric is fp64, ids/links are int32, flags are uint8.
 */
#ifndef BQ_CAPI_H
#define BQ_CAPI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bq_handle bq_handle;

typedef enum {
    BQ_OK = 0,
    BQ_ERR_INVALID = -1,     /* bad argument / inconsistent model                */
    BQ_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError)   */
    BQ_ERR_UNSUPPORTED = -3, /* model feature outside the kernels' scope         */
    BQ_ERR_NO_DEVICE = -4,   /* no CUDA device: there is NO CPU fallback         */
    BQ_ERR_STATE = -5        /* chain state invalid (negative service etc.)      */
} bq_status;

/* queue disciplines: queues.pyx:444 (QueueGGk; YAML GG1 -> k=1, qnetu.py:109), :745 (DelayStation),
 * :1032 (QueuePS), :788 (QueueGG1R: one server, waiting jobs taken in random order; its by-queue list is ordered by
 * departure, and so are the links bq_create takes and bq_get_state returns for it) */
enum { BQ_Q_GGK = 0, BQ_Q_DELAY = 1, BQ_Q_PS = 2, BQ_Q_GG1R = 3 };
/* service families: distributions.pyx:52 (Exponential), :107 (Gamma), :245 (LogNormal) */
enum { BQ_D_EXP = 0, BQ_D_GAMMA = 1, BQ_D_LOGNORMAL = 2 };
/* sweep kinds: qnet.pyx:659 slice_resample, qnet.pyx:337 gibbs_resample (MH-within-Gibbs) */
enum { BQ_MODE_SLICE = 0, BQ_MODE_MH = 1 };
/* parameter update after each sweep: none (estimation.py:171 sample_departures),
 * posterior draw (estimation.py:104 bayes -> qnet.pyx:882), ML (estimation.py:29 sem -> qnet.pyx:844) */
enum { BQ_UPD_NONE = 0, BQ_UPD_BAYES = 1, BQ_UPD_STEM = 2 };
/* flags for bq_sweep */
enum {
    BQ_FLAG_TMAX_PER_SWEEP = 1, /* recompute Tmax = 2 max d before every sweep (one slice_resample call per
                                   sweep, as estimation.bayes does); otherwise once per bq_sweep call
                                   (qnet.pyx:672) -- or, on dynamic-order networks, once per launch when a call
                                   holds more sweeps than one launch may (2^31 / (64 x resampled events)) */
    BQ_FLAG_NO_FINAL = 2,       /* MH only: sample_final=False (qnet.pyx:337)                              */
    BQ_FLAG_TRACE = 4,          /* record per-sweep theta / mean wait / log-prob traces                    */
    BQ_FLAG_TIGHT = 8,          /* start the slice bracket at the support of the conditional instead of
                                   [a_e, min(d_next, Tmax)] (qnet.pyx:698-701): same stationary law, fewer
                                   rejected candidates; off by default (reference-faithful bracket)        */
    BQ_FLAG_REVERSE = 16        /* visit the events in the reverse of the scan order: slice_resample(reverse=True)
                                   (qnet.pyx:683-684), the backward kernel of estimation.chib_evidence        */
};

/* Flattened network: replaces the Queue vtable (queues.pxd:19-37), FactorTemplate (queues.pxd:12-17)
 * and Distribution (distributions.pxd:1-7) objects. */
typedef struct {
    int32_t n_queues;
    int32_t n_params;           /* total length of the parameter vector (Qnet.parameters, qnet.pyx:898) */
    const int32_t *q_type;      /* [n_queues] BQ_Q_*                                                    */
    const int32_t *q_k;         /* [n_queues] processors (1 for DELAY / PS)                             */
    const int32_t *q_dist;      /* [n_queues] BQ_D_*                                                    */
    const int32_t *q_param_off; /* [n_queues] offset of the queue's first parameter                     */
    /* Mixture service templates (MixtureTemplate, queues.pyx:1373-1509; YAML [MIX, w0, [..], w1, [..], ...]).  n_slots = 0:
     * none, q_dist holds each queue's family.  Otherwise every queue owns the distribution slots q_slot_off[q] ..
     * q_slot_off[q+1]-1 (one for a plain queue, one per component for a mixture), slot_dist gives their families and
     * q_dist is ignored.  Parameters of a mixture queue: its mixing weights, then each component's parameters
     * (MixtureTemplate.getParameters, queues.pyx:1470-1474).  Static-order networks only. */
    int32_t n_slots;
    const int32_t *q_slot_off;  /* [n_queues + 1] */
    const int32_t *slot_dist;   /* [n_slots] BQ_D_* */
} bq_model;

/* Flattened Arrivals (arrivals.pxd:11-37 Event fields, arrivals.pyx:240 Arrivals).  Event index =
 * position in these arrays.  a[e] is implied: d[prev_byt[e]], or 0 for a task's first event
 * (arrivals.pyx:848).  prev/next_byq give the INITIAL order inside each queue (arrival order; queue 0
 * by departure, arrivals.pyx:256). */
typedef struct {
    int32_t n_events;
    const int32_t *qid;      /* [E] */
    const int32_t *tid;      /* [E] */
    const uint8_t *obs_a;    /* [E] */
    const uint8_t *obs_d;    /* [E] */
    const int32_t *prev_byt; /* [E] -1 = none */
    const int32_t *next_byt; /* [E] -1 = none */
    const int32_t *prev_byq; /* [E] -1 = none */
    const int32_t *next_byq; /* [E] -1 = none */
    const double *d;         /* [E] initial departure times, replicated to every chain */
} bq_arrivals;

/* counters mirroring qnet.stats() (qnet.pyx:115-132) and sampling.statistics() (sampling.pyx:37-52) */
typedef struct {
    int64_t nevents;     /* resamples (qnet.pyx:416,481,709) — the unit of the headline metric */
    int64_t nreject;     /* MH rejections                                                      */
    int64_t num_diff;    /* conditional-density evaluations (qnet.pyx:1351)                    */
    int64_t diff_len;    /* events touched by those evaluations                                */
    int64_t slice_calls; /* sampling.pyx:276 */
    int64_t slice_evals; /* sampling.pyx:280,347 */
    int64_t n_sweeps;    /* sweeps completed per chain since create / reset */
    int64_t n_launches;  /* CUDA kernels launched by this handle */
    int64_t param_evals; /* log-target evaluations inside the Gamma posterior slice sampler (netutils.py:19) */
    int64_t n_stuck;     /* slice updates that kept x0: point mass / bracket < 1e-10 (sampling.pyx:289,360) */
    int64_t n_redone;    /* batch kernel: updates recomputed because an earlier update of the batch wrote what
                            they had read (the result is the sequential sweep either way)                 */
} bq_stats_t;

const char *bq_last_error(void);
int bq_version(void);
/* number of visible CUDA devices, or a negative bq_status */
int bq_device_count(void);

/* Build device state for n_chains independent chains (global chain ids chain_offset ..
 * chain_offset+n_chains-1: the RNG counter uses the GLOBAL id, so a sharded run reproduces the
 * unsharded one).  theta0[n_params] is replicated to every chain. */
int bq_create(const bq_model *model, const bq_arrivals *arrv, const double *theta0, int32_t n_chains,
              int64_t chain_offset, uint64_t seed, int32_t device, bq_handle **out);
int bq_destroy(bq_handle *h);

/* Qnet.parameters get/set (qnet.pyx:898-915), per chain: theta[(chain_hi-chain_lo) * n_params] */
int bq_set_params(bq_handle *h, int32_t chain_lo, int32_t chain_hi, const double *theta);
int bq_get_params(bq_handle *h, int32_t chain_lo, int32_t chain_hi, double *theta);

/* Chain state in/out (Arrivals.duplicate / Event fields).  Any output pointer may be NULL.
 * d,a,s: [(chain_hi-chain_lo) * E] fp64; proc, prev_byq, next_byq: same shape int32. */
int bq_set_state(bq_handle *h, int32_t chain_lo, int32_t chain_hi, const double *d);
/* Arrivals.duplicate for every chain: one state d[E] from the host, replicated on the device. */
int bq_broadcast_state(bq_handle *h, const double *d);
int bq_get_state(bq_handle *h, int32_t chain_lo, int32_t chain_hi, double *d, double *a, double *s,
                 int32_t *proc, int32_t *prev_byq, int32_t *next_byq);
/* Mixture component of every event (Event.auxiliaries / Event.variable, arrivals.pxd:36): aux[(chain_hi-chain_lo) * E],
 * 0 for events of plain queues.  Only handles created with mixture templates hold them (BQ_ERR_INVALID otherwise).
 * New handles start with component 0 everywhere; every sweep resamples them first (qnet.pyx:679-680). */
int bq_set_aux(bq_handle *h, int32_t chain_lo, int32_t chain_hi, const uint8_t *aux);
int bq_get_aux(bq_handle *h, int32_t chain_lo, int32_t chain_hi, uint8_t *aux);
/* Same, but d is DEVICE memory laid out [n_chains][E] (zero-copy access for torch callers). */
int bq_state_dev(bq_handle *h, double **d_dev, int64_t *chain_stride);

/* A starting state of its own for every chain (replaces one call of Qnet.gibbs_initialize, qnet.pyx:531 / 1454-1554, per
 * chain): every resample-eligible departure is redrawn, in a topological order of the ordering constraints, as its lower
 * bound plus an exponential with the mean service of its queue (from the chain's current parameters) times a per-chain
 * factor exp(sigma N(0,1)), truncated to the upper bound the observed times impose; the queue order of the handle and
 * every observed time are kept, every service stays >= 0.  sigma = 0: same scale for all chains (still different
 * draws).  Static-order networks (single-server FIFO queues, delay stations); BQ_ERR_UNSUPPORTED otherwise. */
int bq_init_chains(bq_handle *h, double sigma);

/* n_sweeps full Gibbs sweeps over every chain, each followed by the requested parameter update:
 * Qnet.slice_resample (qnet.pyx:659) / Qnet.gibbs_resample (qnet.pyx:337) + Qnet.sample_parameters
 * (qnet.pyx:882) / Qnet.estimate (qnet.pyx:844).  Asynchronous on the handle's stream;
 * bq_synchronize() waits. */
int bq_sweep(bq_handle *h, int32_t n_sweeps, int32_t mode, int32_t param_update, uint32_t flags);
int bq_synchronize(bq_handle *h);
/* The parameter update alone, from the chains' current state: Qnet.sample_parameters (qnet.pyx:882,
 * BQ_UPD_BAYES) or Qnet.estimate (qnet.pyx:844, BQ_UPD_STEM) as separate calls. */
int bq_update_params(bq_handle *h, int32_t param_update);

/* The same update from sufficient statistics the CALLER supplies (layout of bq_suffstats; n_ss = 1: one set for every
 * chain, n_ss = n_chains: one per chain): Qnet.estimate / Qnet.sample_parameters given a LIST of Arrivals pool the
 * events of all of them per queue (qnet.pyx:844-858, 882-896) -- the host adds the per-Arrivals statistics and passes
 * the sum here. */
int bq_update_params_from(bq_handle *h, int32_t param_update, const double *suffstats, int32_t n_ss);

/* Parity hook: GGkGibbs.inner_dfun(x) - lik0 (qnet.pyx:1340-1357) for event `eid` of chain `chain`
 * at n points. */
int bq_logcond(bq_handle *h, int32_t chain, int32_t eid, const double *x, int32_t n, double *out);
/* Parity hook for the MH path: the TWOS proposal log-density departureLik(e) + arrivalLik(e') at n points
 * (queues.pair_proposal / departure_proposal, queues.pyx:1239-1283; -inf outside its support). */
int bq_mh_proposal(bq_handle *h, int32_t chain, int32_t eid, const double *x, int32_t n, double *out);
/* Qnet.gibbs_kernel (qnet.pyx:741-809), the transition density estimation.chib_evidence (estimation.py:274-330)
 * averages: out[c] = log T(d_to <- state of chain c) of one slice sweep in the handle's scan order = minus the sum over
 * the resampled events of log integral exp(g(x) - g(d_to[e])) dx (GGkGibbs.integrate, qnet.pyx:1361-1372), each event
 * moved to its target before the next one is taken; events whose target is not reachable yet are retried in later
 * passes and -inf is returned when a pass moves nothing (qnet.pyx:756-805).  strict != 0: no later passes -- an
 * unreachable target gives -inf, the density of the fixed-scan sweep itself (what Chib's identity needs).  d_to:
 * n_to * n_events doubles, n_to = 1 (one target for all chains) or n_chains.  The chains' states are not modified
 * (networks with multi-server or PS queues are evaluated in place and restored).  Every supported discipline. */
int bq_gibbs_kernel(bq_handle *h, const double *d_to, int32_t n_to, int32_t strict, double *out);
/* Qnet.log_prob (qnet.pyx:1032): out[n_chains] */
int bq_log_prob(bq_handle *h, double *out);
/* Per chain per queue (with mixture templates: per distribution slot, n_slots of them) sufficient statistics over
 * events with s>0 (queues.pyx:1317):
 * out[n_chains * n_queues * 8] = {N, sum s, sum log s, sum (log s - mean log s)^2, #log terms, sum wait,
 * n_events, #(s == 0)} */
int bq_suffstats(bq_handle *h, double *out);
int bq_stats(bq_handle *h, bq_stats_t *out);
int bq_reset_stats(bq_handle *h);

/* Traces recorded by sweeps run with BQ_FLAG_TRACE since the last bq_clear_trace():
 * theta[n_rec * n_chains * n_params], mean_wait[n_rec * n_chains * n_queues], logp[n_rec * n_chains].
 * bq_trace_len returns n_rec.  Any pointer may be NULL. */
int bq_trace_len(bq_handle *h, int32_t *n_rec);
int bq_get_trace(bq_handle *h, double *theta, double *mean_wait, double *logp);
int bq_clear_trace(bq_handle *h);
/* Device pointers to the same traces (for NCCL gathers without a host bounce). */
int bq_trace_dev(bq_handle *h, double **theta_dev, double **mean_wait_dev, double **logp_dev);

/* The sweep schedule the kernels use: resample-eligible events grouped into conflict-free colour
 * classes (any fixed scan is a valid Gibbs sampler).
 * n_colors+1 offsets into `order`; order has n_eligible entries.  Pointers may be NULL to query sizes. */
int bq_get_schedule(bq_handle *h, int32_t *n_colors, int32_t *n_eligible, int32_t *color_off,
                    int32_t *order);

/* The scan order of the sequential (dynamic-order) kernels: the resample-eligible events queue after queue, each
 * queue in its initial arrival order. */
int bq_get_sequential_order(bq_handle *h, int32_t *n_eligible, int32_t *order);
/* The order in which BQ_MODE_MH sweeps visit the events (Qnet.gibbs_resample, qnet.pyx:354-368): the schedule of
 * bq_get_schedule for static-order networks, the sequential order otherwise. */
int bq_get_mh_order(bq_handle *h, int32_t *n_eligible, int32_t *order);

/* Raw counter-based RNG stream (Philox4x32 with SEVEN rounds, csrc/bq_rng.cuh) exactly as the kernels consume it:
 * n 53-bit uniforms in [0,1), two per block, for (seed, global chain id, sweep index, event or queue id, stream tag:
 * 1 slice, 2 parameters, 3/4 the two slice steps of an MH proposal, 5 MH accept + re-initialisation).
 * Pure host function, needs no device.  Replaces rk_double (randomkit.c:264). */
int bq_rng_uniforms(uint64_t seed, int64_t chain, uint32_t sweep, uint32_t id, uint32_t tag, int32_t n,
                    double *out);
/* The uniforms of one finite-bound slice update (sampling._slice, sampling.pyx:283-298, 344-366) in the order it
 * consumes them: out[0] the slice level's uniform and out[1] the first shrink candidate (block 0, 53 bits each), then
 * out[c], c >= 2, candidate c from one 32-bit word, (word + 0.5) 2^-32, four per block.  Same stream keys as above. */
int bq_rng_slice_uniforms(uint64_t seed, int64_t chain, uint32_t sweep, uint32_t id, uint32_t tag, int32_t n,
                          double *out);
/* One raw Philox4x32 block with `rounds` rounds (7 = what the kernels run, 10 = the customary count): the hook the
 * known-answer test against the Random123 vectors goes through. */
int bq_philox_block(const uint32_t ctr[4], const uint32_t key[2], int32_t rounds, uint32_t out[4]);

/* Launch geometry the handle chose: threads per CTA, dynamic shared memory per CTA, and whether the
 * chain state is staged in shared memory for the whole launch. */
int bq_kernel_info(bq_handle *h, int32_t *threads, int32_t *smem_bytes, int32_t *state_in_smem);

/* CTAs per chain of the chromatic sweep: 1, or 2 / 4 / 8 when the handle runs it on thread-block clusters with the chain
 * state spread over their shared memories (few chains, or a chain too large for one CTA's shared memory). */
int bq_cluster_size(bq_handle *h, int32_t *ncta);

/* Checkpoint / resume (the reference restarts from 5-decimal CSV snapshots, gibbs.py:137-160, arrivals.pyx:97 -- lossy;
 * here a run resumes exactly): the index of the next sweep, which with the seed, the global chain ids, the departure
 * times (bq_get_state / bq_set_state) and the parameters fixes every random number that follows. */
/* Random-order queues only (round 2): on != 0 replaces the rule "an event that started on arrival found the queue empty"
 * (-inf in QueueGG1R.likelihoodDelta, queues.pyx:984-989) by a penalty per violation, so that slice sweeps can walk a
 * state that breaks it -- what gibbs_initialize leaves when the departure that made an observed job wait is hidden --
 * into the support; on = 0 restores the reference's rule.  Not a sampler of the posterior while on. */
int bq_set_relaxed(bq_handle *h, int32_t on);
int bq_get_sweep_counter(bq_handle *h, uint32_t *counter);
int bq_set_sweep_counter(bq_handle *h, uint32_t counter);

/* CUDA-event timing of the last bq_sweep call on the handle's stream, in milliseconds. */
int bq_last_sweep_ms(bq_handle *h, float *ms);
/* The handle's cudaStream_t (as void*) so callers can order their own work after the sweeps. */
int bq_stream(bq_handle *h, void **stream);

#ifdef __cplusplus
}
#endif
#endif /* BQ_CAPI_H */
