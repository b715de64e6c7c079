// Counter-based per-chain RNG (Philox4x32-7) shared by the device kernels and the host test hook.
//
// Replaces the reference's global MT19937 stream (randomkit.c:185-269, cdist.c:107-121).  Stream parity
// with MT19937 is not a goal; what matters is that every (chain, sweep, event,
// purpose) owns an independent stream, so the result does not depend on how chains are sharded over
// GPUs or on the order the kernel visits events.
//
//   counter = (block index, event id or queue id, sweep index, global chain id low 32)
//   key     = (seed low 32 ^ tag * 0x9E3779B9, seed high 32 ^ global chain id high 32)
//
// One Philox call yields 4 x 32 bits = two 53-bit uniforms in [0,1), or four 32-bit uniforms in (0,1).  SEVEN rounds
// (BQ_PHILOX_ROUNDS): the smallest round count of Philox4x32 that Salmon et al. (SC'11, table 2) report as passing
// BigCrush ("Crush-resistant"); the customary 10 only adds margin, and the generator was ~25% of the static sweep
// kernel's instructions.  The round function is pinned against the Random123 known-answer vectors for 7 and 10
// rounds by the test suite (through the host hook bq_philox_block).
//
// Uniforms of a finite-bound slice update (sampling._slice, sampling.pyx:283-298, 344-366), identical in every kernel
// (and in the CPU checker the tests use): block 0 gives the slice level (words 0,1: 53 bits -- it goes through a log) and the first shrink
// candidate (words 2,3: 53 bits); block b >= 1 gives the four candidates 2 + 4 (b - 1) + w, w = 0..3, from one 32-bit
// word each, u = (word + 0.5) 2^-32.  A candidate only places a point inside the current bracket, so 2^-32 of the
// bracket is ample resolution, and one Philox block then feeds four candidates instead of two.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BQ_HD __host__ __device__ __forceinline__
#else
#define BQ_HD inline
#endif

#ifndef BQ_PHILOX_ROUNDS
#define BQ_PHILOX_ROUNDS 7
#endif

namespace bq {

enum : uint32_t {
    TAG_SLICE = 1,   // hidden-time slice updates (qnet.pyx:701)
    TAG_PARAM = 2,   // service-parameter draws (qnet.pyx:882)
    TAG_MH = 3,      // MH-within-Gibbs: first slice step on the proposal (qnet.pyx:337, 1292-1300)
    TAG_MH2 = 4,     //   second slice step (thin = 2)
    TAG_MH_AUX = 5,  //   accept uniform (number 0) and re-initialisation draws (qnet.pyx:389-397, 444-451)
    TAG_AUX = 6,     // mixture components (MixtureTemplate.resample_auxiliaries, queues.pyx:1402)
    TAG_TEST = 7,
    TAG_INIT = 8     // per-chain starting states (bq_init_chains)
};

BQ_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

template <int ROUNDS = BQ_PHILOX_ROUNDS>
BQ_HD void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                         uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

BQ_HD double u53(uint32_t lo, uint32_t hi) {
    uint64_t v = ((uint64_t)hi << 32) | (uint64_t)lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}
// 32-bit uniform in (0,1): the shrink candidates after the first
BQ_HD double u32(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

struct Rng {
    uint32_t k0, k1, c1, c2, c3, blk;
    double spare;
    int has_spare;
    uint32_t wbuf[4];
    int nwords;

    BQ_HD void init(uint64_t seed, uint64_t chain, uint32_t sweep, uint32_t id, uint32_t tag) {
        k0 = (uint32_t)seed ^ (tag * 0x9E3779B9u);
        k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(chain >> 32);
        c1 = id; c2 = sweep; c3 = (uint32_t)chain;
        blk = 0; has_spare = 0; spare = 0.0; nwords = 0;
    }
    // uniform in [0,1), 53 bits
    BQ_HD double uniform() {
        if (has_spare) { has_spare = 0; return spare; }
        uint32_t o[4];
        philox4x32(blk++, c1, c2, c3, k0, k1, o);
        spare = u53(o[2], o[3]);
        has_spare = 1;
        return u53(o[0], o[1]);
    }
    // next 32-bit uniform in (0,1): four per Philox block (shrink candidates 2, 3, ... of a finite-bound slice update)
    BQ_HD double uniform32() {
        if (nwords == 0) { philox4x32(blk++, c1, c2, c3, k0, k1, wbuf); nwords = 4; }
        const uint32_t w = (nwords == 4) ? wbuf[0] : (nwords == 3) ? wbuf[1] : (nwords == 2) ? wbuf[2] : wbuf[3];
        --nwords;
        return u32(w);
    }
    // raw 32-bit word (consumes half a uniform's worth; used for rk_interval-style integers)
    BQ_HD uint32_t word() {
        uint32_t o[4];
        philox4x32(blk++, c1, c2, c3, k0, k1, o);
        return o[0];
    }
};

// uniform of shrink candidate c >= 1 of a finite-bound slice update (see the header comment)
BQ_HD double cand_uniform(const Rng &r, unsigned int c) {
    uint32_t o[4];
    if (c <= 1u) {
        philox4x32(0u, r.c1, r.c2, r.c3, r.k0, r.k1, o);
        return u53(o[2], o[3]);
    }
    philox4x32(1u + ((c - 2u) >> 2), r.c1, r.c2, r.c3, r.k0, r.k1, o);
    const unsigned int w = (c - 2u) & 3u;
    return u32(w == 0u ? o[0] : w == 1u ? o[1] : w == 2u ? o[2] : o[3]);
}

}  // namespace bq
