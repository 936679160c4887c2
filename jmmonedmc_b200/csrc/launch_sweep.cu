// Launches of the colour-decomposed sweep kernel (sweep.cuh: k_sweep).
#include "handle.h"

using namespace jmm;

template <int POT, int G>
static cudaError_t launch_sweep_inst(jmm_handle *h, const SweepShape &s, const SweepDev &W, uint64_t step0, int nsub, unsigned ntiles) {
    dim3 grid(ntiles, (unsigned) h->S.nchains);
    cudaError_t e = cudaFuncSetAttribute(k_sweep<POT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s.smem);
    if (e != cudaSuccess) return e;
    k_sweep<POT, G><<<grid, s.threads, s.smem, h->stream>>>(W, step0, nsub, s.tile, s.halo, h->d_partial, h->cb_counts);
    h->launches++;
    return cudaGetLastError();
}

template <int POT, int G, int NB = 0>
static cudaError_t launch_sweep_fast_inst(jmm_handle *h, const SweepShape &s, const SweepDev &W, uint64_t step0, int nsub, unsigned ntiles) {
    dim3 grid(ntiles, (unsigned) h->S.nchains);
    cudaError_t e = cudaFuncSetAttribute(k_sweep_fast<POT, G, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s.smem);
    if (e != cudaSuccess) return e;
    k_sweep_fast<POT, G, NB><<<grid, s.threads, s.smem, h->stream>>>(W, philox_keys((uint32_t) W.seed, (uint32_t)(W.seed >> 32)), step0, nsub, s.tile, s.halo, s.rounds, s.rad, h->d_partial,
                                                                   h->cb_counts, h->cb_tile_done, h->cb_tot, h->cb_acc);
    h->launches++;
    return cudaGetLastError();
}

template <int POT>
static cudaError_t launch_sweep(jmm_handle *h, const SweepShape &s, const SweepDev &W, uint64_t step0, int nsub, unsigned ntiles) {
    if constexpr (POT != kPotHarmonic) {
        if (s.fast) {
            switch (s.G) {
                case 1:
                    // compile-time NBN for the short-range decks (fully unrolled partner loop)
                    switch (getenv("JMM_SWEEP_NOUNROLL") ? 0 : h->cfg.nbn) {
                        case 1: return launch_sweep_fast_inst<POT, 1, 1>(h, s, W, step0, nsub, ntiles);
                        case 2: return launch_sweep_fast_inst<POT, 1, 2>(h, s, W, step0, nsub, ntiles);
                        case 3: return launch_sweep_fast_inst<POT, 1, 3>(h, s, W, step0, nsub, ntiles);
                        case 4: return launch_sweep_fast_inst<POT, 1, 4>(h, s, W, step0, nsub, ntiles);
                        case 6: return launch_sweep_fast_inst<POT, 1, 6>(h, s, W, step0, nsub, ntiles);
                        case 8: return launch_sweep_fast_inst<POT, 1, 8>(h, s, W, step0, nsub, ntiles);
                        default: return launch_sweep_fast_inst<POT, 1>(h, s, W, step0, nsub, ntiles);
                    }
                case 2: return launch_sweep_fast_inst<POT, 2>(h, s, W, step0, nsub, ntiles);
                case 4: return launch_sweep_fast_inst<POT, 4>(h, s, W, step0, nsub, ntiles);
                case 8: return launch_sweep_fast_inst<POT, 8>(h, s, W, step0, nsub, ntiles);
                case 16: return launch_sweep_fast_inst<POT, 16>(h, s, W, step0, nsub, ntiles);
                default: return launch_sweep_fast_inst<POT, 32>(h, s, W, step0, nsub, ntiles);
            }
        }
    }
    if (s.G == 1) return launch_sweep_inst<POT, 1>(h, s, W, step0, nsub, ntiles);
    return launch_sweep_inst<POT, 32>(h, s, W, step0, nsub, ntiles);
}

cudaError_t jmm_launch_sweep(jmm_handle *h, const SweepShape &s, const SweepDev &W, uint64_t step0, int nsub, unsigned ntiles) {
    switch (h->cfg.pot) {
        case JMM_POT_LJ: return launch_sweep<kPotLJ>(h, s, W, step0, nsub, ntiles);
        case JMM_POT_LJCUT: return launch_sweep<kPotLJcut>(h, s, W, step0, nsub, ntiles);
        default: return launch_sweep<kPotHarmonic>(h, s, W, step0, nsub, ntiles);
    }
}
