// Service-time families: log-density with per-chain precomputed constants, posterior draws and ML updates.
//
// Follows the reference's distributions.pyx: Exponential (:52-105), Gamma (:107-207), LogNormal (:245-345),
// netutils.sample_gamma_posterior (netutils.py:19-42) and the stepping-out branch of sampling._slice
// (sampling.pyx:299-338) that the Gamma posterior sampler goes through.
#pragma once
#include <math.h>
#include "bq_rng.cuh"

// Heavy routines are kept out of line in device code: the sweep kernels call lpdf (an fp64 log) from a dozen places,
// and inlining every copy grows the kernel past the instruction cache.
#if defined(__CUDA_ARCH__)
#define BQ_NOINLINE __noinline__
#else
#define BQ_NOINLINE
#endif

#define BQ_HDF __host__ __device__ __forceinline__
#define BQ_HDI __host__ __device__ inline
#define BQ_HDN __host__ __device__ BQ_NOINLINE inline  // one copy in device code

namespace bq {

// libdevice's fp64 log / lgamma are 100+ / 1000+ instructions and get inlined at every call site: one out-of-line
// copy each keeps the sweep kernels inside the instruction cache
__host__ __device__ BQ_NOINLINE inline double log_d(double x) { return log(x); }
__host__ __device__ BQ_NOINLINE inline double lgamma_d(double x) { return lgamma(x); }

// log of a small positive integer (the N(t) + 1 of processor sharing, queues.pyx:1225-1230): table in constant
// memory on the device, filled by the library with the host's log(); plain log() on the host
#if defined(__CUDACC__)
static __constant__ double c_log_count[256];  // one copy per translation unit (bq_lib.cu, bq_batch_nw.cu)
#endif
__host__ __device__ __forceinline__ double log_count(int n) {
#if defined(__CUDA_ARCH__)
    return (n >= 0 && n < 256) ? c_log_count[n] : log_d((double)n);
#else
    return log((double)n);
#endif
}

enum { Q_GGK = 0, Q_DELAY = 1, Q_PS = 2, Q_GG1R = 3 };
enum { D_EXP = 0, D_GAMMA = 1, D_LOGNORMAL = 2 };

#define BQ_NEG_INF (-INFINITY)
#define BQ_HALF_LOG_2PI (-0.918938533204673)  // distributions.pyx:243 (sic: this is -log(sqrt(2 pi)))

// Per chain, per queue: the distribution family with constants folded so that one lpdf costs
// 0 (Exp), or 1 (Gamma, LogNormal) fp64 log.
struct QConst {
    double c0, c1, inv;
    int dist;
    int pad;
};

__host__ __device__ __forceinline__ void qconst_set(QConst &qc, int dist, double p0, double p1) {
    qc.dist = dist;
    qc.pad = 0;
    if (dist == D_EXP) {  // -log(scale) - x/scale                      (distributions.pyx:59)
        qc.c0 = -log_d(p0); qc.c1 = 0.0; qc.inv = 1.0 / p0;
    } else if (dist == D_GAMMA) {  // shape p0, scale p1                 (distributions.pyx:129-133)
        qc.c1 = p0 - 1.0;
        qc.inv = 1.0 / p1;
        qc.c0 = (p0 == 1.0) ? -p0 * log_d(p1) : -lgamma_d(p0) - p0 * log_d(p1);
    } else {  // LogNormal meanlog p0, sdlog p1                          (distributions.pyx:266-270)
        qc.c0 = BQ_HALF_LOG_2PI - log_d(p1);
        qc.c1 = p0;
        qc.inv = 1.0 / (2.0 * p1 * p1);
    }
}

// log-density of service time s >= 0; the Gamma / LogNormal part (one fp64 log) is a real call
__host__ __device__ BQ_NOINLINE inline double lpdf_log(const QConst &qc, double s) {
    if (qc.dist == D_GAMMA) return qc.c1 * log_d(s) - s * qc.inv + qc.c0;
    const double ls = log_d(s), z = ls - qc.c1;
    return qc.c0 - ls - z * z * qc.inv;
}
// the same with log(s) supplied by the caller (the dynamic-order kernels cache it per position): no transcendental
__host__ __device__ __forceinline__ double lpdf_ls(const QConst &qc, double s, double ls) {
    if (qc.dist == D_EXP) return qc.c0 - s * qc.inv;
    if (qc.dist == D_GAMMA) {
        if (qc.c1 == 0.0) return qc.c0 - s * qc.inv;
        return qc.c1 * ls - s * qc.inv + qc.c0;
    }
    if (s == 0.0) return BQ_NEG_INF;
    const double z = ls - qc.c1;
    return qc.c0 - ls - z * z * qc.inv;
}
__host__ __device__ __forceinline__ double lpdf(const QConst &qc, double s) {
    if (qc.dist == D_EXP) return qc.c0 - s * qc.inv;
    if (qc.dist == D_GAMMA) {
        if (qc.c1 == 0.0) return qc.c0 - s * qc.inv;  // shape == 1 special case handles s == 0
        return lpdf_log(qc, s);
    }
    if (s == 0.0) return BQ_NEG_INF;
    return lpdf_log(qc, s);
}

// ---------------------------------------------------------------------------------------------
// random variates for the parameter updates (numpy.random.gamma / normal in the reference)
// ---------------------------------------------------------------------------------------------
__device__ inline double draw_normal(Rng &r) {  // Box-Muller
    double u1 = 1.0 - r.uniform();         // (0,1]
    double u2 = r.uniform();
    return sqrt(-2.0 * log_d(u1)) * cospi(2.0 * u2);
}

__device__ inline double draw_gamma(Rng &r, double shape) {  // Marsaglia-Tsang, scale 1
    double boost = 1.0;
    if (shape < 1.0) {
        double u = 1.0 - r.uniform();
        boost = pow(u, 1.0 / shape);
        shape += 1.0;
    }
    double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 1000; ++it) {
        double x = draw_normal(r), v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = 1.0 - r.uniform();
        if (log_d(u) < 0.5 * x * x + d - d * v + d * log_d(v)) return boost * d * v;
    }
    return boost * d;
}

// digamma / trigamma for the Minka fixed point (distributions.pyx:142-160)
__device__ inline double digamma_d(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    double f = 1.0 / (x * x);
    return r + log_d(x) - 0.5 / x -
           f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f * (1.0 / 132.0)))));
}
__device__ inline double trigamma_d(double x) {
    double r = 0.0;
    while (x < 6.0) { r += 1.0 / (x * x); x += 1.0; }
    double f = 1.0 / (x * x);
    return r + 1.0 / x + 0.5 * f +
           (1.0 / x) * f * (1.0 / 6.0 - f * (1.0 / 30.0 - f * (1.0 / 42.0 - f * (1.0 / 30.0))));
}

// log target of the Gamma posterior conditionals (netutils.py:29,37); fixed = the other parameter
struct GammaPost {
    double logP, q, r, s;  // 1+sum log x, 1+sum x, 1+N, 1+N
    __device__ double shape_lnf(double al, double scale0) const {
        return (al - 1.0) * logP - q / scale0 - r * lgamma_d(al) - (al * s) * log_d(scale0);
    }
    __device__ double scale_lnf(double beta, double shape) const {
        return (shape - 1.0) * logP - q / beta - r * lgamma_d(shape) - (shape * s) * log_d(beta);
    }
};

// sampling._slice with lower = 0, upper = +inf: stepping out, w0 = 1 doubling, m = 1024
// (sampling.pyx:264-366).  which = 0: shape conditional, 1: scale conditional.
__device__ inline double slice_stepout(Rng &rng, const GammaPost &gp, int which, double fixed, double x0,
                                       long long *evals) {
    const double lower = 0.0;
    double w = 1.0;
#define BQ_GP(x) ((which == 0) ? gp.shape_lnf((x), fixed) : gp.scale_lnf((x), fixed))
    double gx0 = BQ_GP(x0);
    ++*evals;
    double logy = gx0 - (-log_d(1.0 - rng.uniform()));
    if (logy == BQ_NEG_INF || logy != logy) return x0;
    double u = w * rng.uniform();
    double L = x0 - u, R = x0 + (w - u);
    // rk_interval(m): uniform integer in [0, m]; K = (m-1) - J in unsigned arithmetic
    unsigned long long J = (unsigned long long)(rng.word() % 1025u);
    unsigned long long K = (unsigned long long)(1023ull - J);
    while (J > 0) {
        if (L <= lower) break;
        ++*evals;
        if (BQ_GP(L) <= logy) break;
        L -= w; --J; w *= 2.0;
    }
    int guard = 0;
    while (K > 0 && guard < 1100) {
        ++*evals;
        if (BQ_GP(R) <= logy) break;
        R += w; --K; w *= 2.0; ++guard;
    }
    if (L < lower) L = lower;
    for (int it = 0; it < 100000; ++it) {
        double x1 = L + (R - L) * rng.uniform();
        ++*evals;
        double gx1 = BQ_GP(x1);
        if (gx1 >= logy) return x1;
        if (x1 > x0) R = x1; else L = x1;
        if (R - L < 1e-10) return L;
    }
#undef BQ_GP
    return x0;
}

// ---- one-lane slice sampling on an arbitrary log-density (the MH proposal) ------------------------------------
// sampling._slice when a bound is infinite: stepping out with w0 = 1, doubling, m = 1024 (sampling.pyx:299-341),
// then shrinkage.  Sequential; used for final events of MH sweeps (upper = inf).  On a collapsed bracket we keep x0
// (the documented deviation from sampling.pyx:360-362).
template <class G>
BQ_HDI double slice_stepout_seq(const G &g, double x0, double lower, double upper, Rng &rng, long long &evals) {
    double w = 1.0;
    const double gx0 = g(x0);
    ++evals;
    const double logy = gx0 - (-log_d(1.0 - rng.uniform()));
    if (!(logy > BQ_NEG_INF)) return x0;
    const double u = w * rng.uniform();
    double L = x0 - u, R = x0 + (w - u);
    unsigned long long J = (unsigned long long)(rng.word() % 1025u);  // rk_interval(m): uniform on [0, m]
    unsigned long long K = 1023ull - J;
    while (J > 0) {
        if (L <= lower) break;
        ++evals;
        if (g(L) <= logy) break;
        L -= w; --J; w *= 2.0;
    }
    int guard = 0;
    while (K > 0 && guard < 1100) {
        if (R >= upper) break;
        ++evals;
        if (g(R) <= logy) break;
        R += w; --K; w *= 2.0; ++guard;
    }
    if (L < lower) L = lower;
    if (R > upper) R = upper;
    for (int it = 0; it < 100000; ++it) {
        const double x1 = L + (R - L) * rng.uniform();
        ++evals;
        if (g(x1) >= logy) return x1;
        if (x1 > x0) R = x1; else L = x1;
        if (R - L < 1e-10) return x0;
    }
    return x0;
}

// sampling._slice with finite bounds (sampling.pyx:283-298, 344-366), one lane, candidates in sequence: uniform 0 of
// the stream draws the slice level, uniform c the c-th shrink candidate -- the same mapping the W-lane groups of the
// dynamic-order kernel use.  Keeps x0 on a collapsed bracket.
template <class G>
BQ_HDI double slice_finite_seq(const G &g, double x0, double lower, double upper, Rng &rng, long long &evals) {
    const double gx0 = g(x0);
    ++evals;
    const double logy = gx0 - (-log_d(1.0 - rng.uniform()));
    if (!(logy > BQ_NEG_INF)) return x0;
    double L = lower, R = upper;
    for (int it = 0; it < 1000; ++it) {
        const double x1 = L + (R - L) * (it == 0 ? rng.uniform() : rng.uniform32());  // bq_rng.cuh: candidate uniforms
        ++evals;
        if (g(x1) >= logy) return x1;
        if (x1 > x0) R = x1; else L = x1;
        if (R - L < 1e-10) return x0;
    }
    return x0;
}

// Sufficient statistics of one queue over its events with s > 0 (queues.pyx:1317) plus bookkeeping.
struct SuffStat {
    double n;       // #{s > 0}
    double sum;     // sum s
    double sumlog;  // sum log s   (s >= 1e-50 for LogNormal's second moment, distributions.pyx:283)
    double sumsq;   // sum (log s - mean log)^2 over s >= 1e-50
    double nlog;    // #{s >= 1e-50}
    double wait;    // sum max(0, d - a - s) over all events (arrivals.pyx:187)
    double nall;    // # events
    double nzero;   // #{s == 0}
};

// Posterior draw of one queue's parameters (DistributionTemplate.sample_parameters, queues.pyx:1316).
static __device__ BQ_NOINLINE void sample_params(int dist, const SuffStat &st, Rng &rng, double &p0, double &p1,
                                     long long *evals) {
    if (dist == D_EXP) {  // distributions.pyx:92-96: scale = 1 / Gamma(N, 1/sum)
        if (st.n < 1.0 || !(st.sum > 0.0)) return;
        double x = draw_gamma(rng, st.n) / st.sum;
        p0 = 1.0 / x;
    } else if (dist == D_LOGNORMAL) {  // distributions.pyx:306-320
        double mean = 0.0, sd = 1.0;
        if (st.n >= 1.0) {
            mean = st.sumlog / st.nlog;
            sd = sqrt(st.sumsq / st.nlog);
            if (sd < 1e-10) sd = 1e-10;
        }
        if (sd <= 1e-5) { p0 = mean; p1 = sd; return; }
        if (st.n < 1.0) { p0 = 0.0; p1 = 1.0; return; }
        double alpha = (st.n + 1.0) / 2.0;  // true division: Py3 oracle
        double beta = 2.0 / (st.n * sd * sd);
        double tau = draw_gamma(rng, alpha) * beta;
        double sd_new = sqrt(1.0 / tau);
        p0 = mean + sd_new * draw_normal(rng);
        p1 = sd_new;
    } else {  // Gamma: slice-within-Gibbs, thin 5 each (netutils.py:19-42)
        GammaPost gp;
        gp.logP = 1.0 + st.sumlog; gp.q = 1.0 + st.sum; gp.r = 1.0 + st.n; gp.s = 1.0 + st.n;
        double shape = p0, scale = p1;
        double sh = shape;
        for (int t = 0; t < 5; ++t) sh = slice_stepout(rng, gp, 0, scale, sh, evals);
        double sc = scale;
        for (int t = 0; t < 5; ++t) sc = slice_stepout(rng, gp, 1, sh, sc, evals);
        p0 = sh; p1 = sc;
    }
}

// ML update of one queue's parameters (DistributionTemplate.estimate, queues.pyx:1324).
static __device__ BQ_NOINLINE void estimate_params(int dist, const SuffStat &st, double &p0, double &p1) {
    if (dist == D_EXP) {  // distributions.pyx:65-74
        if (st.n < 1.0) { p0 = 1.0; return; }
        double mu = st.sum / st.n;
        p0 = (mu < 1e-7) ? 1e-7 : mu;
    } else if (dist == D_LOGNORMAL) {  // distributions.pyx:276-303
        if (st.n < 1.0) { p0 = 0.0; p1 = 1.0; return; }
        double mean = st.sumlog / st.nlog, sd = sqrt(st.sumsq / st.nlog);
        if (sd < 1e-10) sd = 1e-10;
        p0 = mean; p1 = sd;
    } else {  // Minka fixed point, distributions.pyx:142-160
        if (st.n < 1.0) return;
        double mu = st.sum / st.n, mu_log = st.sumlog / st.n;
        double a = 0.5 / (log_d(mu) - mu_log), a_old = -INFINITY;
        for (int it = 0; it < 200 && fabs(a - a_old) >= 0.01; ++it) {
            a_old = a;
            a = 1.0 / (1.0 / a_old + (mu_log - log_d(mu) + log_d(a_old) - digamma_d(a_old)) /
                                         (a_old * a_old * (1.0 / a_old - trigamma_d(a_old))));
        }
        p0 = a; p1 = mu / a;
    }
}

}  // namespace bq
