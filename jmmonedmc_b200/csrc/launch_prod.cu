// Launches of the many-chain production kernels (prod.cuh).
#include <math.h>

#include "handle.h"
#include "prod.cuh"

using namespace jmm;

template <int POT, int ARITH, int UNROLL>
static cudaError_t launch_step_prod_u(jmm_handle *h, const StepArgs &a) {
    auto kern = a.accept_log ? k_chains_step_prod<POT, ARITH, true, UNROLL> : k_chains_step_prod<POT, ARITH, false, UNROLL>;
    auto sliced = a.accept_log ? k_chains_step_prod_sliced<POT, ARITH, true, UNROLL> : k_chains_step_prod_sliced<POT, ARITH, false, UNROLL>;
    cudaError_t e;
    const size_t tile_bytes = h->smem;                    // [N][32] doubles
    const unsigned ntiles = nblk(h->S.nchains, kTile);
    // warps (= tiles) per CTA: as many tiles as fit in the 227 KB of one CTA, at most kProdMaxWarps (registers)
    int nw = (int) std::min<size_t>(kProdMaxWarps<ARITH>, (227 * 1024 - 1024) / tile_bytes);
    if (const char *ev = getenv("JMM_PROD_WARPS")) nw = std::max(1, std::min(nw, atoi(ev)));
    nw = std::max(1, std::min<int>(nw, (int) ntiles));
    const size_t smem = tile_bytes * nw;
    if (smem > 48 * 1024) {
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(sliced, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess) return e;
    }
    // how many warps of the sliced kernel are co-resident on this device
    int per_sm = 0, nsm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sliced, nw * kTile, smem)) != cudaSuccess) return e;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->cfg.device);
    const unsigned ctas = (unsigned) std::max(1, per_sm * nsm);
    const unsigned slots = ctas * (unsigned) nw;
    const double waves = (double) ntiles / slots;
    // (histogram bins are only touched by REDs and L2 reads, so a chain may change SM between chunks)
    const bool slice = getenv("JMM_FORCE_SLICE") ||
                       (!getenv("JMM_NO_SLICE") && ntiles > slots && (waves - floor(waves)) < 0.85 && a.nsteps >= 16);
    if (!slice) {
        kern<<<nblk(ntiles, nw), nw * kTile, smem, h->stream>>>(h->S, a, h->H, ntiles);
        h->launches++;
        return cudaGetLastError();
    }
    // ~12+ chunks per launch bounds the imbalance to one chunk in twelve; at least 8 steps per chunk
    uint32_t chunk = (uint32_t) std::max<uint64_t>(8, (a.nsteps + 11) / 12);
    if (const char *ev = getenv("JMM_SLICE_CHUNK")) chunk = (uint32_t) std::max(1, atoi(ev));
    const uint32_t nchunks = (uint32_t) ((a.nsteps + chunk - 1) / chunk);
    if (h->work_words < (size_t) ntiles + 1) {
        if (h->d_work) cudaFree(h->d_work);
        h->d_work = nullptr; h->work_words = 0;
        if ((e = cudaMalloc((void **) &h->d_work, ((size_t) ntiles + 1) * sizeof(unsigned int))) != cudaSuccess) return e;
        h->work_words = (size_t) ntiles + 1;
    }
    if ((e = cudaMemsetAsync(h->d_work, 0, ((size_t) ntiles + 1) * sizeof(unsigned int), h->stream)) != cudaSuccess) return e;
    const unsigned grid = std::min(ctas, nblk((uint64_t) ntiles * nchunks, nw));
    sliced<<<grid, nw * kTile, smem, h->stream>>>(h->S, a, h->H, chunk, ntiles, nchunks, h->d_work, h->d_work + 1);
    h->launches++;
    return cudaGetLastError();
}

template <int POT, int ARITH>
static cudaError_t launch_step_prod(jmm_handle *h, const StepArgs &a) {
    if constexpr (ARITH == kArithFast) {
        // partners in flight per thread: 8 (C4: 7.74e9 trials/s) or 4 (7.54e9)
        const char *e = getenv("JMM_PROD_UNROLL");
        if (!(e && atoi(e) == 4)) return launch_step_prod_u<POT, ARITH, 8>(h, a);
    }
    return launch_step_prod_u<POT, ARITH, 4>(h, a);
}

template <int POT>
static cudaError_t launch_step_prod_pot(jmm_handle *h, const StepArgs &a) {
    if constexpr (POT != kPotHarmonic) {
        if (h->cfg.arith == JMM_ARITH_FAST) return launch_step_prod<POT, kArithFast>(h, a);
    }
    return launch_step_prod<POT, kArithReference>(h, a);
}

cudaError_t jmm_launch_prod(jmm_handle *h, const StepArgs &a) {
    switch (h->cfg.pot) {
        case JMM_POT_LJ: return launch_step_prod_pot<kPotLJ>(h, a);
        case JMM_POT_LJCUT: return launch_step_prod_pot<kPotLJcut>(h, a);
        default: return launch_step_prod_pot<kPotHarmonic>(h, a);
    }
}
