// The batch kernel with two warps per chain (sweep_batch_kernel_nw, bq_dynamic.cuh), in a translation unit of its own.
// Compiled together with the one-warp kernels it changed their inlining: the processor-sharing build of
// sweep_batch_kernel went from 272 to 432 bytes of stack and config 4 from 7.7e8 to 5.5e8 resamples/s, with not a line
// of that kernel touched.  Here the two sets of instantiations cannot influence each other.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "bq_dynamic.cuh"

namespace bq {

#ifndef BQ_DENSE_MINB
#define BQ_DENSE_MINB 8
#endif

cudaError_t batch_nw_set_log_table(const double *tab256) {
    return cudaMemcpyToSymbol(c_log_count, tab256, 256 * sizeof(double));
}

template <int K, int MINB, int PSV>
static void launch_one(const DynArgs &A, int *claim, int *claimd, int blocks, size_t sm, cudaStream_t stream) {
    if (sm > 48 * 1024)
        cudaFuncSetAttribute((const void *)sweep_batch_kernel_nw<K, MINB, false, PSV, 2>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    sweep_batch_kernel_nw<K, MINB, false, PSV, 2><<<blocks, 128, sm, stream>>>(A, claim, claimd);
}

// minb = CTAs of 128 threads per SM the build is compiled for: 4 (128 registers) or 8 (64).  Measured on configs 3 / 4
// (profiles/r2_batch_nw.log): 3 CTAs (168 registers) 5.10e8 / 7.26e8, 4 CTAs 5.63e8 / 8.24e8, 5 CTAs (96) 3.80e8 / 7.14e8,
// 6 CTAs (80) 3.58e8 / 6.69e8, 8 CTAs (64) 4.40e8 / 5.94e8 resamples/s: spills cost more than the extra warps bring.
template <int K, int PSV>
static void launch_m(const DynArgs &A, int *claim, int *claimd, int minb, int blocks, size_t sm, cudaStream_t stream) {
    if (minb >= 8) launch_one<K, BQ_DENSE_MINB, PSV>(A, claim, claimd, blocks, sm, stream);
    else launch_one<K, 4, PSV>(A, claim, claimd, blocks, sm, stream);
}
template <int K>
static void launch_k(const DynArgs &A, int *claim, int *claimd, bool has_ps, int minb, int blocks, size_t sm,
                     cudaStream_t stream) {
    if (has_ps) launch_m<K, 1>(A, claim, claimd, minb, blocks, sm, stream);
    else launch_m<K, 0>(A, claim, claimd, minb, blocks, sm, stream);
}

// slice sweeps of a dynamic-order network, two chains per CTA of four warps
int launch_batch_nw2(const DynArgs &A, int *claim, int *claimd, int kmax, bool has_ps, bool dense, int n_chains, int Q, int P,
                     cudaStream_t stream) {
    // Waves: two warps per chain pay while all of them fit the 16-warps-per-SM (128-register) build, so more chains
    // than that run as equal waves, one launch after the other on the handle's stream -- every chain does all the
    // sweeps of the call inside its wave.  `dense` (the caller's wish for the 64-register build) is kept for the test
    // hook BQ_BATCH_WAVE=0.
    const int G = 2;
    const size_t sm = (size_t)G * dyn_group_smem(Q, P);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int minb = 4;  // tuning hook BQ_NW_MINB: more CTAs per SM with fewer registers each
    if (const char *env = getenv("BQ_NW_MINB")) minb = atoi(env);
    if (minb != 8) minb = 4;
    int cap = (4 * minb * sms) / 2;  // chains resident at once
    if (const char *env = getenv("BQ_BATCH_WAVE")) cap = atoi(env) > 0 ? atoi(env) : n_chains;
    const int waves = (n_chains + cap - 1) / cap;
    const int per = (n_chains + waves - 1) / waves;
    for (int lo = 0; lo < n_chains; lo += per) {
        DynArgs W = A;
        W.chain_lo = lo;
        W.chain_hi = (lo + per < n_chains) ? lo + per : n_chains;
        const int nc = W.chain_hi - W.chain_lo, blocks = (nc + G - 1) / G;
        const int mb = (dense && 2 * nc > 4 * minb * sms) ? 8 : minb;
        if (kmax <= 4) launch_k<4>(W, claim, claimd, has_ps, mb, blocks, sm, stream);
        else launch_k<8>(W, claim, claimd, has_ps, mb, blocks, sm, stream);
    }
    return waves;
}

}  // namespace bq
