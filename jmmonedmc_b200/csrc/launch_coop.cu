// Launches of the few-chain kernels: coop.cuh (G lanes per chain, any potential) and bond.cuh (HARMONIC, NBN 1).
#include <math.h>
#include <cmath>

#include "handle.h"
#include "coop.cuh"
#include "bond.cuh"
#include "solo.cuh"

using namespace jmm;

static_assert(kCoopScratchRows == kCoopChunk, "jmm_create sizes the scratch of coop.cuh");

template <int POT, int G>
static cudaError_t launch_step_coop_g(jmm_handle *h, const StepArgs &a) {
    auto kern = a.accept_log ? k_chains_step_coop<POT, G, true> : k_chains_step_coop<POT, G, false>;
    if (h->coop_smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->coop_smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned per_block = 128 / G;
    kern<<<nblk(h->S.nchains, per_block), 128, h->coop_smem, h->stream>>>(h->S, a, h->coop_npad);
    h->launches++;
    return cudaGetLastError();
}

static cudaError_t launch_step_bond(jmm_handle *h, const StepArgs &a) {
    const bool inf = std::isinf(h->S.cutoff);
    // threads per CTA: 128 (a warp per sub-partition); JMM_BOND_BLOCK = 32, 64 or 96 for wave-quantisation experiments
    unsigned threads = 128;
    if (const char *e = getenv("JMM_BOND_BLOCK")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 96 || v == 128) threads = (unsigned) v; }
    const unsigned per_block = threads / kB2G;
    if (h->bond == 3) {
        // k_chains_step_solo: one chain per thread, one warp per CTA (4096 chains = 128 CTAs: a warp per SM)
        const bool ten = h->S.N == 10;
        void (*kern)(ChainsDev, StepArgs);
        if (a.accept_log) kern = inf ? (ten ? k_chains_step_solo<10, true, true> : k_chains_step_solo<0, true, true>)
                                     : (ten ? k_chains_step_solo<10, true, false> : k_chains_step_solo<0, true, false>);
        else kern = inf ? (ten ? k_chains_step_solo<10, false, true> : k_chains_step_solo<0, false, true>)
                        : (ten ? k_chains_step_solo<10, false, false> : k_chains_step_solo<0, false, false>);
        kern<<<nblk(h->S.nchains, 32), 32, (size_t) h->S.N * 32 * sizeof(double), h->stream>>>(h->S, a);
        h->launches++;
        return cudaGetLastError();
    }
    if (h->bond == 4 || h->bond == 5) {
        // solo.cuh, a CTA per 32 chains: k_chains_step_trio (trials | Philox | virial + ECheck + sums) or, with volume trials in
        // the deck, k_chains_step_crew (displacement | volume | Philox | virial + sums | ECheck).  A CTA that meets an energy
        // discrepancy stores nothing and raises its word in `redo`; k_chains_step_bond then repeats the launch for its chains.
        const bool ten = h->S.N == 10;
        const bool crew = h->bond == 5 && (uint64_t) h->S.numTrialTypes > h->S.N;
        void (*kern)(ChainsDev, StepArgs, unsigned int *, int);
#define JMM_PICK(K) (a.accept_log ? (inf ? (ten ? K<10, true, true> : K<0, true, true>) : (ten ? K<10, true, false> : K<0, true, false>)) \
                                  : (inf ? (ten ? K<10, false, true> : K<0, false, true>) : (ten ? K<10, false, false> : K<0, false, false>)))
        if (crew) kern = JMM_PICK(k_chains_step_crew); else kern = JMM_PICK(k_chains_step_trio);
#undef JMM_PICK
        cudaError_t e;
        const unsigned nctas = nblk(h->S.nchains, 32);
        const size_t smem = crew ? sizeof(CrewShared) + (size_t) 4 * h->S.N * 32 * sizeof(double)   // live, V's copy, W's copy, zeros
                                 : sizeof(TrioRings) + (size_t) 2 * h->S.N * 32 * sizeof(double);
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess) return e;
        if (h->work_words < (size_t) nctas) {
            if (h->d_work) cudaFree(h->d_work);
            h->d_work = nullptr; h->work_words = 0;
            if ((e = cudaMalloc((void **) &h->d_work, (size_t) nctas * sizeof(unsigned int))) != cudaSuccess) return e;
            h->work_words = nctas;
        }
        if ((e = cudaMemsetAsync(h->d_work, 0, (size_t) nctas * sizeof(unsigned int), h->stream)) != cudaSuccess) return e;
        const char *fr = getenv("JMM_SOLO_FORCE_REDO");
        kern<<<nctas, crew ? 160 : 96, smem, h->stream>>>(h->S, a, h->d_work, (fr && atoi(fr) != 0) ? 1 : 0);
        h->launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        int npad = (int) h->S.N;
        npad += (npad & 1) ? 0 : 1;
        auto again = a.accept_log ? (inf ? k_chains_step_bond<true, true> : k_chains_step_bond<true, false>)
                                  : (inf ? k_chains_step_bond<false, true> : k_chains_step_bond<false, false>);
        again<<<nblk(h->S.nchains, 128 / kB2G), 128, (size_t) (128 / kB2G) * npad * sizeof(double), h->stream>>>(h->S, a, npad, h->d_work);
        h->launches++;
        return cudaGetLastError();
    }
    if (h->bond == 2) {
        // k_chains_step_bond2: the shared row only parks the positions for the rare paths; the thermo ring follows it
        int npad = (int) ((h->S.N + 1) & ~1ull) + kThermoRing * kThermoSlots;
        npad += (npad & 1) ? 0 : 1;
        const bool every = a.eci == 1;
        void (*kern)(ChainsDev, StepArgs, int);
        if (a.accept_log) kern = inf ? (every ? k_chains_step_bond2<true, true, true> : k_chains_step_bond2<true, true, false>)
                                     : (every ? k_chains_step_bond2<true, false, true> : k_chains_step_bond2<true, false, false>);
        else kern = inf ? (every ? k_chains_step_bond2<false, true, true> : k_chains_step_bond2<false, true, false>)
                        : (every ? k_chains_step_bond2<false, false, true> : k_chains_step_bond2<false, false, false>);
        kern<<<nblk(h->S.nchains, per_block), threads, (size_t) per_block * npad * sizeof(double), h->stream>>>(h->S, a, npad);
        h->launches++;
        return cudaGetLastError();
    }
    int npad = (int) h->S.N;
    npad += (npad & 1) ? 0 : 1;                                  // odd row length: the groups of a warp hit different banks
    auto kern = a.accept_log ? (inf ? k_chains_step_bond<true, true> : k_chains_step_bond<true, false>)
                             : (inf ? k_chains_step_bond<false, true> : k_chains_step_bond<false, false>);
    kern<<<nblk(h->S.nchains, per_block), threads, (size_t) per_block * npad * sizeof(double), h->stream>>>(h->S, a, npad, nullptr);
    h->launches++;
    return cudaGetLastError();
}

template <int POT>
static cudaError_t launch_step_coop(jmm_handle *h, const StepArgs &a) {
    if constexpr (POT == kPotHarmonic) {
        if (h->bond) return launch_step_bond(h, a);
    }
    switch (h->coop_g) {
        case 8: return launch_step_coop_g<POT, 8>(h, a);
        case 16: return launch_step_coop_g<POT, 16>(h, a);
        default: return launch_step_coop_g<POT, 32>(h, a);
    }
}

cudaError_t jmm_launch_coop(jmm_handle *h, const StepArgs &a) {
    switch (h->cfg.pot) {
        case JMM_POT_LJ: return launch_step_coop<kPotLJ>(h, a);
        case JMM_POT_LJCUT: return launch_step_coop<kPotLJcut>(h, a);
        default: return launch_step_coop<kPotHarmonic>(h, a);
    }
}
