// Launches of the G-lanes-per-chain fast-arithmetic kernel (lanes.cuh).
#include <math.h>

#include "handle.h"
#include "lanes.cuh"
#include "team.cuh"

using namespace jmm;

template <int POT, int G, int NPL>
static cudaError_t launch_lanes_npl(jmm_handle *h, const StepArgs &a) {
    auto kern = a.accept_log ? k_chains_step_lanes<POT, G, NPL, true> : k_chains_step_lanes<POT, G, NPL, false>;
    cudaError_t e;
    constexpr unsigned CPW = 32 / G;
    const unsigned ntiles = nblk(h->S.nchains, CPW);
    const size_t warp_bytes = (size_t) CPW * h->lanes_stride * sizeof(double);
    int nw = (int) std::min<size_t>(kLanesMaxWarps, (227 * 1024 - 1024) / warp_bytes);
    if (const char *ev = getenv("JMM_LANES_WARPS")) nw = std::max(1, std::min(nw, atoi(ev)));
    nw = std::max(1, std::min<int>(nw, (int) ntiles));
    const size_t smem = warp_bytes * nw;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess) return e;
    int per_sm = 0, nsm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nw * 32, smem)) != cudaSuccess) return e;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->cfg.device);
    const unsigned ctas = (unsigned) std::max(1, per_sm * nsm);
    const unsigned slots = ctas * (unsigned) nw;
    // One chunk when the tiles fill the warp slots (almost) evenly; otherwise ~12 chunks per tile, so that every warp
    // slot ends up with the same number of items to within one in twelve.
    const double waves = (double) ntiles / slots;
    const double fill = waves / ceil(waves);
    uint32_t chunk = (uint32_t) a.nsteps;
    if ((getenv("JMM_FORCE_SLICE") || (fill < 0.93 && ntiles > (unsigned) nsm)) && !getenv("JMM_NO_SLICE") && a.nsteps >= 16)
        chunk = (uint32_t) std::max<uint64_t>(8, (a.nsteps + 11) / 12);
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
    kern<<<grid, nw * 32, smem, h->stream>>>(h->S, a, chunk, ntiles, nchunks, (uint32_t) h->lanes_npad, (uint32_t) h->lanes_stride,
                                             h->d_work, h->d_work + 1);
    h->launches++;
    return cudaGetLastError();
}

// team.cuh: eight loop warps + one bookkeeper warp per 32 chains (N in 73 ... 80, NBN < 0)
template <int POT>
static cudaError_t launch_team(jmm_handle *h, const StepArgs &a) {
    constexpr int NT = 2;                                        // teams per CTA (18 warps, one CTA per SM)
    auto kern = a.accept_log ? k_chains_step_team<POT, NT, true> : k_chains_step_team<POT, NT, false>;
    cudaError_t e;
    const unsigned ntiles = nblk(h->S.nchains, kTeamChains);
    const size_t smem = NT * sizeof(TeamShared);
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess) return e;
    int per_sm = 0, nsm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT * kTeamWarps * 32, smem)) != cudaSuccess) return e;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->cfg.device);
    const unsigned ctas = (unsigned) std::max(1, per_sm * nsm);
    const unsigned slots = ctas * NT;
    const double waves = (double) ntiles / slots;
    const double fill = waves / ceil(waves);
    uint32_t chunk = (uint32_t) a.nsteps;
    if ((getenv("JMM_FORCE_SLICE") || (fill < 0.93 && ntiles > (unsigned) nsm)) && !getenv("JMM_NO_SLICE") && a.nsteps >= 16)
        chunk = (uint32_t) std::max<uint64_t>(8, (a.nsteps + 11) / 12);
    if (const char *ev = getenv("JMM_SLICE_CHUNK")) chunk = (uint32_t) std::max(1, atoi(ev));
    const uint32_t nchunks = (uint32_t) ((a.nsteps + chunk - 1) / chunk);
    if (h->work_words < (size_t) ntiles + 1) {
        if (h->d_work) cudaFree(h->d_work);
        h->d_work = nullptr; h->work_words = 0;
        if ((e = cudaMalloc((void **) &h->d_work, ((size_t) ntiles + 1) * sizeof(unsigned int))) != cudaSuccess) return e;
        h->work_words = (size_t) ntiles + 1;
    }
    if ((e = cudaMemsetAsync(h->d_work, 0, ((size_t) ntiles + 1) * sizeof(unsigned int), h->stream)) != cudaSuccess) return e;
    const unsigned grid = std::min(ctas, nblk((uint64_t) ntiles * nchunks, NT));
    kern<<<grid, NT * kTeamWarps * 32, smem, h->stream>>>(h->S, a, chunk, ntiles, nchunks, h->d_work, h->d_work + 1);
    h->launches++;
    return cudaGetLastError();
}

template <int POT, int G>
static cudaError_t launch_lanes_g(jmm_handle *h, const StepArgs &a) {
    // the fully unrolled partner loop exists for one row length per G (N in (G (NPL-1), G NPL], NBN < 0): the
    // RunJobs shape N = 80; every other shape takes the run-time loop
    constexpr int NPL = G == 2 ? 40 : G == 4 ? 20 : G == 8 ? 10 : G == 16 ? 5 : 0;
    if constexpr (NPL > 0) {
        if (h->lanes_npl == NPL) return launch_lanes_npl<POT, G, NPL>(h, a);
    }
    return launch_lanes_npl<POT, G, 0>(h, a);
}

template <int POT>
static cudaError_t launch_lanes_pot(jmm_handle *h, const StepArgs &a) {
    switch (h->lanes_g) {
        case 2: return launch_lanes_g<POT, 2>(h, a);
        case 4: return launch_lanes_g<POT, 4>(h, a);
        case 8: return launch_lanes_g<POT, 8>(h, a);
        case 16: return launch_lanes_g<POT, 16>(h, a);
        default: return launch_lanes_g<POT, 32>(h, a);
    }
}

// shared-memory shape of a handle served by lanes.cuh: row length (positions + pads), stride between groups
void jmm_lanes_shape(jmm_handle *h, int g) {
    const int N = (int) h->S.N;
    const int npl_t = g == 2 ? 40 : g == 4 ? 20 : g == 8 ? 10 : g == 16 ? 5 : 0;
    const bool unrolled = npl_t > 0 && h->cfg.nbn < 0 && N <= g * npl_t && N > g * (npl_t - 1) && !getenv("JMM_LANES_GENERIC");
    h->lanes_g = g;
    h->lanes_npl = unrolled ? npl_t : 0;
    h->lanes_npad = unrolled ? g * npl_t : N;
    int stride = h->lanes_npad + g * 9 + kThermoRing * kThermoSlots;   // positions, [G][NC] scratch, thermo ring
    while (stride % 16 != g % 16) ++stride;                  // the groups of a half-warp start G (mod 16) banks apart
    h->lanes_stride = stride;
}

cudaError_t jmm_launch_lanes(jmm_handle *h, const StepArgs &a) {
    // warp-specialised teams (team.cuh) for the shape they are built for; JMM_TEAM=0/1 overrides
    bool team = h->lanes_g == 8 && h->lanes_npl == kTeamNPL;
    if (const char *e = getenv("JMM_TEAM")) team = team && atoi(e) != 0; else team = false;   // (opt-in until measured)
    if (team) return h->cfg.pot == JMM_POT_LJ ? launch_team<kPotLJ>(h, a) : launch_team<kPotLJcut>(h, a);
    if (h->cfg.pot == JMM_POT_LJ) return launch_lanes_pot<kPotLJ>(h, a);
    return launch_lanes_pot<kPotLJcut>(h, a);
}
