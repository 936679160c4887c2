// Long-chain kernels (configs C3/C5): colour-decomposed ("checkerboard") displacement sweeps and the
// parallel configuration-totals reduction.
//
// No reference counterpart: the reference makes one qad2 trial per Step (src/jmmMCState.cpp:1758-1811)
// and keeps O(N^2) pair tables (:308-328,442-456), so N = 2^20 cannot even be allocated there.  What IS
// the reference's is each individual trial: md, wall test, partner set |i-nm| <= NBN, the (acc-old)+new
// sums and the Metropolis rule are exactly qad2 (:1182-1377) with distances taken from positions.
//
// Decomposition.  With NBN = k two particles interact iff |i-j| <= k (:1217,1312), so the particles of
// one colour (i mod (k+1)) are mutually independent: their trials commute and are made concurrently.
// One "half-sweep" = draw a colour uniformly (Philox), try every particle of that colour once.
//
// Tiling.  A thread block owns TILE consecutive particles and stages them plus HALO = nsub*NBN particles
// on either side in shared memory with one TMA bulk copy (cp.async.bulk, 1-D).  Because every random
// number is a pure function of (seed, chain, half-sweep, particle), neighbouring blocks recompute each
// other's halo trials bit-identically, so a block can run nsub half-sweeps from shared memory with no
// inter-block communication: after half-sweep t the outer (t+1)*NBN halo particles are stale and are no
// longer used.  HBM traffic per launch: (TILE+2*HALO)*8 B read + TILE*8 B written per tile, for
// TILE*nsub/(NBN+1) trials.  Positions are double-buffered in HBM (r_in -> r_out) so halos read by a
// neighbour are never overwritten mid-launch.
#pragma once
#include <math.h>
#include "pot.cuh"
#include "fastlj.cuh"

// build-time shape of k_sweep_fast (defaults = the measured best)
#ifndef JMM_SWEEP_UNROLL
#define JMM_SWEEP_UNROLL 2        // iterations of the run-time partner loop in flight (two pair terms each)
#endif
#ifndef JMM_SWEEP_NS_MAX
#define JMM_SWEEP_NS_MAX 160      // longest pause (ns) between two polls of the neighbour hand-shake; 0 = spin
#endif
#ifndef JMM_SWEEP_MAXT
#define JMM_SWEEP_MAXT 768        // __launch_bounds__ of k_sweep_fast (register cap = 65536 / MAXT)
#endif

namespace jmm {

constexpr int kSweepUnroll = JMM_SWEEP_UNROLL;

struct SweepDev {
    uint64_t nchains, N;
    int nbn, ncol;
    double cutoff;
    const double *r_in;   // [nchains][N]
    double *r_out;        // [nchains][N]
    const double *l, *T, *maxStep;     // [nchains]
    uint64_t seed, chain_id0;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

// The neighbour hand-shake of k_sweep_fast as release/acquire operations of the PTX memory model (CTA scope, shared
// memory): a warp publishes "half-sweep t done" with a release store after its position writes, a neighbour polls with
// acquire loads; what the publisher wrote before the release is visible after the acquire that reads it.  (Round 1 used a
// volatile int + __threadfence_block(), correct on the hardware but defined only by observed behaviour.)
__device__ __forceinline__ void st_release_cta(int *p, int v) {
    asm volatile("st.release.cta.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta(const int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

__device__ __forceinline__ int colour_of(uint64_t seed, uint32_t chain, uint64_t step, int ncol) {
    const Philox4 b = philox4x32_10((uint32_t) step, (uint32_t)(step >> 32), 0xFFFFFFFFu, kTagColour | chain,
                                    (uint32_t) seed, (uint32_t)(seed >> 32));
    return (int) (((uint64_t) b.w[0] * (uint64_t) ncol) >> 32);
}

// G lanes cooperate on one particle's trial.  G = 1: reference summation order; G = 32: one warp per
// particle, lane-strided partners + xor butterfly, for wide neighbour sets (only 1 and 32 are
// instantiated: a group must be a whole warp for the full-mask shuffles below).
template <int POT, int G>
__global__ void __launch_bounds__(512, 1) k_sweep(SweepDev S, uint64_t step0, int nsub, int tile, int halo,
                                                 double *partial /*[nchains][nsub][ntiles][9]*/,
                                                 unsigned long long *counts /*[nchains][2] accepted, trials*/) {
    constexpr int NC = PotTraits<POT>::NC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nwarps = blockDim.x >> 5;
    double *w = reinterpret_cast<double *>(smem_raw);                 // window of positions
    const int64_t N = (int64_t) S.N;
    const int chain = blockIdx.y;
    const int64_t tile_lo = (int64_t) blockIdx.x * tile;
    const int64_t tile_hi = min(tile_lo + tile, N);
    const int64_t g0 = max((int64_t) 0, tile_lo - halo);              // window = [g0, g1)
    const int64_t g1 = min(N, tile_hi + halo);
    const int wlen = (int) (g1 - g0);
    const int wcap = tile + 2 * halo;
    double *red = w + wcap;                                           // [2][nwarps][9]
    int *colours = reinterpret_cast<int *>(red + 2 * nwarps * 9);     // [nsub]
    int *firsts = colours + nsub;                                     // [nsub] window index of the first particle to try
    __shared__ __align__(8) unsigned long long mbar;

    const double *src = S.r_in + (uint64_t) chain * S.N + g0;
    // ---- stage the window: one TMA bulk copy for the 16-byte-aligned body, plain loads for the rest
    const int body = ((((uintptr_t) src) & 15) == 0) ? (wlen & ~1) : 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && body > 0) {
        const uint32_t bytes = (uint32_t) body * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(w)), "l"(src), "r"(bytes), "r"(smem_u32(&mbar)) : "memory");
    }
    for (int i = body + threadIdx.x; i < wlen; i += blockDim.x) w[i] = src[i];
    for (int t = threadIdx.x; t < nsub; t += blockDim.x) {
        const int col = colour_of(S.seed, (uint32_t)(S.chain_id0 + chain), step0 + t, S.ncol);
        colours[t] = col;
        // first particle of that colour whose whole neighbourhood is still valid in this window at half-sweep t
        const int64_t ulo = (g0 == 0) ? 0 : g0 + (int64_t)(t + 1) * S.nbn;
        firsts[t] = (int) (ulo + (((int64_t) col - ulo % S.ncol) + S.ncol) % S.ncol - g0);
    }
    if (body > 0) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        }
    }
    __syncthreads();

    const double lbox = S.l[chain], T = S.T[chain], maxStep = S.maxStep[chain], cutoff = S.cutoff;
    const double invT = 1.0 / T;
    const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32);
    const uint32_t tag = kTagParticle | (uint32_t)(S.chain_id0 + chain);
    const int nbn = S.nbn, ncol = S.ncol;
    const int group = threadIdx.x / G, lane = threadIdx.x % G, ngroups = blockDim.x / G;
    const int warp = threadIdx.x >> 5;
    unsigned long long n_acc = 0, n_try = 0;

    for (int t = 0; t < nsub; ++t) {
        // particles whose whole neighbourhood is still valid in this window
        const int64_t uhi = (g1 == N) ? N : g1 - (int64_t)(t + 1) * nbn;
        const int64_t first = g0 + firsts[t];
        constexpr int NA = NC;
        double dacc[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) dacc[k] = 0;
        // trial index within the half-sweep (particle = colour + j ncol): one Philox block serves the trials 2m, 2m+1
        int64_t j = first / ncol + group;
        for (int64_t g = first + (int64_t) group * ncol; g < uhi; g += (int64_t) ngroups * ncol, j += ngroups) {
            uint32_t w0, w1;
            if (G == 1 || lane == 0) {
                const Philox4 b = philox4x32_10((uint32_t)(step0 + t), (uint32_t)((step0 + t) >> 32), (uint32_t)(j >> 1), tag, k0, k1);
                w0 = b.w[2 * (j & 1)]; w1 = b.w[2 * (j & 1) + 1];
            }
            if constexpr (G > 1) { w0 = __shfl_sync(0xffffffffu, w0, 0); w1 = __shfl_sync(0xffffffffu, w1, 0); }
            const double rn = u01(w0), ran = u01(w1);
            const int x = (int) (g - g0);
            const double rnm = w[x];
            const double md = (rn - 0.5) * 2 * maxStep;                               // qad2 :1182
            const double rT = rnm + md;                                               // :1183
            const bool owned = (g >= tile_lo) && (g < tile_hi);
            if (owned && lane == 0) ++n_try;
            if (fabs(rT) > lbox / 2.0) continue;                                      // :1188 (group-uniform)
            const int lo = (int) max((int64_t) 0, g - nbn) - (int) g0, hi = (int) min(N - 1, g + nbn) - (int) g0;
            double d[NA];
            bool accept;
            if constexpr (G == 1) {
                // every lane walks nbn left partners (ascending index) then nbn right partners: uniform trip
                // counts, slots outside the chain are skipped; left and right sums apart (:1277, :1354)
                double dsum[NC], dleft[NC], po[NC], pn[NC];
#pragma unroll
                for (int k = 0; k < NC; ++k) { dsum[k] = 0; dleft[k] = 0; }
                for (int p = x - nbn; p < x; ++p) {
                    if (p < lo) continue;
                    const double rp = w[p];
                    phi<POT, true>(rnm - rp, cutoff, lbox, po);
                    phi<POT, true>(rT - rp, cutoff, lbox, pn);
#pragma unroll
                    for (int k = 0; k < NC; ++k) dleft[k] = dleft[k] - po[k] + pn[k];     // :1244
                }
                for (int p = x + 1; p <= x + nbn; ++p) {
                    if (p > hi) break;
                    const double rp = w[p];
                    phi<POT, true>(rp - rnm, cutoff, lbox, po);
                    phi<POT, true>(rp - rT, cutoff, lbox, pn);
#pragma unroll
                    for (int k = 0; k < NC; ++k) dsum[k] = dsum[k] - po[k] + pn[k];       // :1339
                }
#pragma unroll
                for (int k = 0; k < NC; ++k) d[k] = dleft[k] + dsum[k];                   // :1354
                accept = metropolis_accept(d[0], T, invT, ran);                           // :1367-1377
            } else {
                // one warp per particle: lane-strided partners; only dE is combined across the lanes (the
                // decision needs it), the other components stay lane-local until the block-wide sum below
                double po[NC], pn[NC];
#pragma unroll
                for (int k = 0; k < NC; ++k) d[k] = 0;
                for (int p = lo + lane; p <= hi; p += G) {
                    if (p == x) continue;
                    const bool left = p < x;
                    const double rp = w[p];
                    phi<POT, true>(left ? rnm - rp : rp - rnm, cutoff, lbox, po);
                    phi<POT, true>(left ? rT - rp : rp - rT, cutoff, lbox, pn);
#pragma unroll
                    for (int k = 0; k < NC; ++k) d[k] = d[k] - po[k] + pn[k];
                }
                double dE = d[0];
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1) dE += __shfl_xor_sync(0xffffffffu, dE, off, G);
                accept = metropolis_accept(dE, T, invT, ran);
            }
            if (accept) {
                if (G > 1) __syncwarp();
                if (lane == 0) { w[x] = rT; if (owned) ++n_acc; }
                if (owned) {
#pragma unroll
                    for (int k = 0; k < NA; ++k) dacc[k] += d[k];
                }
            }
        }
        // block-wide sum of this half-sweep's deltas over the owned particles -> partial[chain][t][tile][:]
#pragma unroll
        for (int k = 0; k < NA; ++k) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) dacc[k] += __shfl_xor_sync(0xffffffffu, dacc[k], off);
        }
        double *rbuf = red + (t & 1) * nwarps * 9;
        if ((threadIdx.x & 31) == 0)
#pragma unroll
            for (int k = 0; k < NA; ++k) rbuf[warp * 9 + k] = dacc[k];
        __syncthreads();                       // also orders this half-sweep's position writes before the next reads
        if (threadIdx.x < 9) {
            double s = 0;
            if (threadIdx.x < NC) {
                for (int wv = 0; wv < nwarps; ++wv) s += rbuf[wv * 9 + threadIdx.x];
            }
            partial[(((uint64_t) chain * nsub + t) * gridDim.x + blockIdx.x) * 9 + threadIdx.x] = s;
        }
    }

    // ---- write back the owned particles
    double *dst = S.r_out + (uint64_t) chain * S.N;
    for (int64_t g = tile_lo + threadIdx.x; g < tile_hi; g += blockDim.x) dst[g] = w[g - g0];
    // counters
    if (G > 1 && lane != 0) { n_acc = 0; n_try = 0; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
        n_try += __shfl_xor_sync(0xffffffffu, n_try, off);
    }
    if ((threadIdx.x & 31) == 0 && (n_acc | n_try)) {
        atomicAdd(&counts[2 * chain], n_acc);
        atomicAdd(&counts[2 * chain + 1], n_try);
    }
}

// After a k_sweep launch: fold the per-tile deltas into the running totals half-sweep by half-sweep
// (fixed summation order -> reproducible) and sample the twelve sums once per half-sweep
// (updateThermo :1941-1961 with l constant).
__device__ __forceinline__ void cb_sample(double (&a)[12], const double *cur, double N, double lbox) {
    const double rho = N / lbox, E = cur[0], Vir = cur[1], HV = cur[6];
    a[0] += rho; a[1] += rho * rho; a[2] += lbox; a[3] += lbox * lbox;
    a[4] += E; a[5] += E * E; a[6] += lbox * E; a[7] += Vir; a[8] += Vir * Vir; a[9] += E * Vir;
    a[10] += HV; a[11] += HV * HV;
}


// One warp: running totals over the half-sweeps of a launch (ts[t][9] = the nine deltas of half-sweep t) and one
// sample of the twelve sums per half-sweep.  presample != 0: one extra sample of the current totals first (the
// updateThermo of src/Main.cpp:96).  Lane k < 9 carries component k, lane 0 the twelve sums.
__device__ __forceinline__ void sweep_finish_warp(const double *ts, int nsub, uint64_t N, double lbox, double *tot /*[9]*/,
                                                  double *acc /*[12]*/, int presample, int lane) {
    double mine = lane < 9 ? tot[lane] : 0.0;
    double a[12];
    if (lane == 0)
        for (int q = 0; q < 12; ++q) a[q] = acc[q];
    double c3[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (presample) {
        c3[0] = __shfl_sync(0xffffffffu, mine, 0); c3[1] = __shfl_sync(0xffffffffu, mine, 1); c3[6] = __shfl_sync(0xffffffffu, mine, 6);
        if (lane == 0) cb_sample(a, c3, (double) N, lbox);
    }
    for (int t = 0; t < nsub; ++t) {
        if (lane < 9) mine += ts[t * 9 + lane];
        c3[0] = __shfl_sync(0xffffffffu, mine, 0); c3[1] = __shfl_sync(0xffffffffu, mine, 1); c3[6] = __shfl_sync(0xffffffffu, mine, 6);
        if (lane == 0) cb_sample(a, c3, (double) N, lbox);
    }
    if (lane < 9) tot[lane] = mine;
    if (lane == 0)
        for (int q = 0; q < 12; ++q) acc[q] = a[q];
}

// JMM_ARITH_FAST variant of k_sweep for the LJ family (fastlj.cuh: one reciprocal per partner, 18 fp64-pipe
// instructions for the old and the new pair term together; only s6 = sum(b^-6 - a^-6) and s12 are carried).
// Same staging, same tiling, same random numbers, same trials as k_sweep; what differs:
//   * G lanes (1, 2, 4, 8, 16, 32) share one particle's trial; lane j takes the partners at index distance
//     q = j+1, j+1+G, ... on BOTH sides, so one loop iteration holds two independent pair terms and the loop
//     has no self test, no clamps and (for interior particles) a trip count known to the whole group;
//   * the orientation of a distance is known statically (left: r[nm]-r[p], right: r[p]-r[nm]), which is all
//     LJcut's signed `d <= cutOff` test needs (src/pot.cpp:53);
//   * every lane of a group evaluates the same Philox block (no broadcast shuffles, no divergence);
//   * the first/last NBN particles of the chain take a bounds-checked copy of the loop;
//   * NO block-wide barrier between half-sweeps.  Warp w owns the trials j in [wK, (w+1)K) of every half-sweep
//     (K = rounds * 32/G consecutive same-colour particles = a contiguous stretch of K*ncol particles), so what it
//     reads and writes can only collide with warps at most `rad` away (host: rad = 1 + floor((nsub*NBN + ncol +
//     2 NBN) / (K ncol)), the drift of the colour offsets over the launch).  A warp starts half-sweep t once the
//     warps within `rad` have published half-sweep t-1 as done (a counter per warp in shared memory, release /
//     acquire by __threadfence_block): read-after-write and write-after-read are both covered, and a slow warp
//     only holds up its neighbours.  With the __syncthreads version "barrier" was the second largest stall reason
//     (profiles/r01_c3_k_sweep_fast.txt).
//   * the reductions are part of the kernel (no follow-up launches): per half-sweep every warp leaves its two sums
//     in shared memory; at the end the CTA adds them over its warps -> partial[chain][t][tile][2], and the LAST CTA
//     of a chain to finish (a counter per chain) adds the tiles, expands the nine deltas and runs the
//     totals/twelve-sums recurrence of k_sweep_finish.  Every sum has a fixed order, so results do not depend on
//     scheduling.
// G, rounds and rad are chosen by the host (jmm_gpu.cu: sweep_shape).
// NB > 0: NBN known at compile time (G = 1 only): the partner loop is unrolled completely, 2 NB independent pair terms.
template <int POT, int G, int NB>
__global__ void __launch_bounds__(JMM_SWEEP_MAXT, 1) k_sweep_fast(const __grid_constant__ SweepDev S, const __grid_constant__ PhiloxKeys RK, uint64_t step0, int nsub, int tile, int halo, int rounds, int rad,
                                                       double *partial /*[nchains][nsub][ntiles][2]*/,
                                                       unsigned long long *counts /*[nchains][2] accepted, trials*/,
                                                       unsigned int *tile_done /*[nchains], zero between launches*/,
                                                       double *tot /*[nchains][9]*/, double *acc /*[nchains][12]*/) {
    static_assert(POT != kPotHarmonic, "fast arithmetic is an LJ-family optimisation");
    constexpr bool CUT = (POT == kPotLJcut);
    constexpr int GPW = 32 / G;                                       // groups (trials in flight) per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nwarps = blockDim.x >> 5;
    double *w = reinterpret_cast<double *>(smem_raw);                 // window of positions
    const int64_t N = (int64_t) S.N;
    const int chain = blockIdx.y;
    const int64_t tile_lo = (int64_t) blockIdx.x * tile;
    const int64_t tile_hi = min(tile_lo + tile, N);
    const int64_t g0 = max((int64_t) 0, tile_lo - halo);              // window = [g0, g1)
    const int64_t g1 = min(N, tile_hi + halo);
    const int wlen = (int) (g1 - g0);
    const int wcap = tile + 2 * halo;
    int *firsts = reinterpret_cast<int *>(w + wcap);                  // [nsub] window index of the first particle to try
    int *bases = firsts + nsub;                                       // [nsub] the same, moved down to an EVEN trial index
    uint32_t *jbs = reinterpret_cast<uint32_t *>(bases + nsub);       // [nsub] that (even) trial index within the half-sweep
    int *done = bases + 2 * nsub;                                     // [nwarps] half-sweeps completed by each warp (release/acquire)
    double2 *wsum = reinterpret_cast<double2 *>(w + wcap + ((3 * nsub + nwarps + 3) >> 2) * 2);   // [nsub][nwarps] (s12, s6)
    double *ts = reinterpret_cast<double *>(wsum + nsub * nwarps);    // [nsub][9], last CTA of a chain only
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ int is_last;

    const double *src = S.r_in + (uint64_t) chain * S.N + g0;
    const int body = ((((uintptr_t) src) & 15) == 0) ? (wlen & ~1) : 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < nwarps) done[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0 && body > 0) {
        const uint32_t bytes = (uint32_t) body * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(w)), "l"(src), "r"(bytes), "r"(smem_u32(&mbar)) : "memory");
    }
    for (int i = body + threadIdx.x; i < wlen; i += blockDim.x) w[i] = src[i];
    for (int t = threadIdx.x; t < nsub; t += blockDim.x) {
        const int col = colour_of(S.seed, (uint32_t)(S.chain_id0 + chain), step0 + t, S.ncol);
        const int64_t ulo = (g0 == 0) ? 0 : g0 + (int64_t)(t + 1) * S.nbn;
        const int64_t first = ulo + (((int64_t) col - ulo % S.ncol) + S.ncol) % S.ncol;
        const int64_t jf = first / S.ncol;                            // trial index of `first` in its half-sweep
        firsts[t] = (int) (first - g0);
        bases[t] = (int) (first - (jf & 1) * S.ncol - g0);            // trials 2m, 2m+1 share a Philox block
        jbs[t] = (uint32_t) (jf & ~(int64_t) 1);
    }
    if (body > 0) {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        }
    }
    __syncthreads();

    const double half_l = S.l[chain] / 2.0, T = S.T[chain], step2 = 2 * S.maxStep[chain];
    const double invT = 1.0 / T;
    const double cb = S.cutoff;
    const uint32_t tag = kTagParticle | (uint32_t)(S.chain_id0 + chain);
    const int nbn = NB > 0 ? NB : S.nbn, ncol = nbn + 1;
    const int warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
    const int gi = lane32 / G, lane = lane32 % G;
    // the groups of a warp reject at the wall independently: shuffle within the group only
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane32 & ~(G - 1)));
    // range tests as one unsigned compare each: (unsigned)(x - lo) < span
    const int x_lo = (int) (tile_lo - g0);                                       // owned window range [x_lo, x_lo + own_span)
    const unsigned own_span = (unsigned) (tile_hi - tile_lo);
    const int x_first_interior = (int) max((int64_t) 0, (int64_t) nbn - g0);   // x >= this: all left partners exist
    const int x_last_interior = (int) (min(g1, N - nbn) - g0) - 1;             // x <= this: all right partners exist
    const unsigned interior_span = (unsigned) max(0, x_last_interior - x_first_interior + 1);
    const int x_end0 = (int) (((g1 == N) ? N : g1 - nbn) - g0), x_end_step = (g1 == N) ? 0 : nbn;   // x_end of half-sweep t
    const int w_lo = max(0, warp - rad), w_hi = min(nwarps - 1, warp + rad);
    uint32_t n_acc = 0, n_try = 0;

    // Interior first.  A trial at offset o of this warp's stretch (o = 0 .. K-1, K = rounds * GPW) touches particles
    // within NBN of its own; the stretch of half-sweep t is shifted against that of t-1 by -NBN-ncol .. 2 NBN+ncol
    // particles (the halo shrinks by NBN — not in the first tile of a chain —, the colour changes by less than ncol,
    // and the base moves down by one trial when the first trial index is odd).  So for 2 <= o <= K-4 everything the
    // trial reads or overwrites was last touched, at t-1, by THIS warp; only o = 0, 1, K-3, K-2 and K-1 can collide
    // with a neighbour's stretch (and never with a stretch further away, however far that warp lags: it is at least
    // (d-1)K+2 trials off after d half-sweeps of drift of at most 2 NBN + ncol each).  The trials are therefore taken
    // in the order o = 2, 3, ..., K-1, 0, 1 and the wait for the two neighbours comes right before the first round
    // that holds one of the last five: the hand-shake latency hides behind the interior rounds.
    // (K < 6, or a stretch so short that the host asks for rad > 1: wait before round 0, as a barrier would.)
    const int K = rounds * GPW;
    const bool high = lane32 >= 16;                                   // G = 1: lanes 16-31 take the odd trials of a round
    const int lane_o = 2 * (lane32 & 15) + (lane32 >> 4);
    const bool interior_first = rad == 1 && K >= 6;
    const int r_wait = interior_first ? (K - 5) / GPW : 0, rot = interior_first ? 2 : 0;
    for (int t = 0; t < nsub; ++t) {
        const int x_end = x_end0 - t * x_end_step;
        const uint32_t s_lo = (uint32_t)(step0 + t), s_hi = (uint32_t)((step0 + t) >> 32);
        double acc6 = 0, acc12 = 0;
        const int x_valid = firsts[t], x_base = bases[t] + warp * K * ncol;
        const unsigned valid_span = (unsigned) max(0, x_end - x_valid);
        const uint32_t j_base = jbs[t] + (uint32_t) (warp * K);       // trial index of offset 0 (even for G = 1: K is)
        Philox4 ahead{};
        for (int r = 0; r < rounds; ++r) {
#ifndef JMM_ABL_NOSYNC
            if (t > 0 && r == r_wait) {
                // every warp whose stretch can touch ours must have finished half-sweep t-1: lane j watches warp
                // w_lo + j, one vote per poll, the pause doubling up to 160 ns
                const int v = min(w_lo + lane32, w_hi);
                unsigned ns = 20;
                while (!__all_sync(0xffffffffu, ld_acquire_cta(done + v) >= t)) {
                    if (JMM_SWEEP_NS_MAX > 0) { __nanosleep(ns); ns = min(ns * 2, (unsigned) JMM_SWEEP_NS_MAX); }
                }
            }
#endif
            // One Philox block serves the trials 2m and 2m+1 of a half-sweep (words 0,1 and 2,3).
            uint32_t w0, w1;
            int o;
            if constexpr (G == 1) {
                // Two rounds = 64 consecutive offsets = 32 blocks, one per lane, evaluated in the even round.  Lanes
                // 0-15 take the even trials of a round and lanes 16-31 the odd ones, so that in the even round
                // (blocks 0-15) a low lane owns its words and a high lane fetches words 2,3 from lane - 16, and in the
                // odd round (blocks 16-31) a high lane owns its words and a low lane fetches words 0,1 from lane + 16.
                const int o_pair = (r >> 1) * 64 + rot;                                   // first offset of the round pair
                if ((r & 1) == 0) {
                    int op = o_pair + 2 * lane32;
                    if (op >= K) op -= K;
                    ahead = philox4x32_10(s_lo, s_hi, (j_base + (uint32_t) op) >> 1, tag, RK);
                }
                const bool odd_round = r & 1;
                const uint32_t give0 = odd_round ? ahead.w[0] : ahead.w[2], give1 = odd_round ? ahead.w[1] : ahead.w[3];
                const uint32_t got0 = __shfl_xor_sync(0xffffffffu, give0, 16), got1 = __shfl_xor_sync(0xffffffffu, give1, 16);
                const bool own = high == odd_round;
                w0 = own ? (high ? ahead.w[2] : ahead.w[0]) : got0;
                w1 = own ? (high ? ahead.w[3] : ahead.w[1]) : got1;
                o = o_pair + 32 * (r & 1) + lane_o;
                if (o >= K) o -= K;
            } else {
                // the G lanes of a group need the same two words; evaluated by all of them the block is G-1 times
                // redundant.  Instead lane j evaluates the block of the group's trial j rounds ahead, once every G
                // rounds, and each round fetches its words from the lane that holds them.
                o = r * GPW + gi + rot;
                if (o >= K) o -= K;
                const int rb = r % G;
                if (rb == 0) {
                    int oj = (r + lane) * GPW + gi + rot;
                    if (oj >= K) oj -= K;                    // (rounds past the last one: an unused block)
                    const uint32_t jj = j_base + (uint32_t) oj;
                    const Philox4 b4 = philox4x32_10(s_lo, s_hi, jj >> 1, tag, RK);
                    ahead.w[0] = b4.w[2 * (jj & 1)]; ahead.w[1] = b4.w[2 * (jj & 1) + 1];
                }
                w0 = __shfl_sync(gmask, ahead.w[0], rb, G);
                w1 = __shfl_sync(gmask, ahead.w[1], rb, G);
            }
            const int x = x_base + o * ncol;
            if ((unsigned) (x - x_valid) >= valid_span) continue;
#ifdef JMM_ABL_NOPHILOX
            w0 = (s_lo * 2654435761u) ^ ((uint32_t)(g0 + x) * 2246822519u); w1 = w0 * 3266489917u + tag;
#endif
            const double ran = u01_shifted(w1, 1.0);                                  // = w1 / 2^32 exactly
            const double rnm = w[x];
            const double md = u01_shifted(w0, 1.5) * step2;                           // qad2 :1182 ((rn-.5)*2*maxStep; rn-.5 and 2*maxStep exact)
            const double rT = rnm + md;                                               // :1183
            const bool owned = (unsigned) (x - x_lo) < own_span;
            if (owned && lane == 0) ++n_try;
            // :1188: a move through the wall is rejected whatever its energy; only the two ends of a chain can get
            // there, so it is a predicate on the decision, not a branch around the pair terms
            const bool inside = !(fabs(rT) > half_l);
            double s6 = 0, s12 = 0;
            if ((unsigned) (x - x_first_interior) < interior_span) {
                if constexpr (NB > 0) {
                    static_assert(NB == 0 || G == 1, "compile-time NBN is a G = 1 specialisation");
                    double t6 = 0, t12 = 0;                   // right partners apart: two independent accumulation chains
#pragma unroll
                    for (int q = 1; q <= NB; ++q) {
                        const double rl = w[x - q], rr = w[x + q];
                        lj_partner<CUT>(rnm - rl, rT - rl, cb, s6, s12);
                        lj_partner<CUT>(rr - rnm, rr - rT, cb, t6, t12);
                    }
                    s6 += t6; s12 += t12;
                } else {
                    const double *wl = w + x - 1 - lane, *wr = w + x + 1 + lane;
#pragma unroll (kSweepUnroll)
                    for (int q = lane; q < nbn; q += G, wl -= G, wr += G) {
                        const double rl = *wl, rr = *wr;
                        lj_partner<CUT>(rnm - rl, rT - rl, cb, s6, s12);
                        lj_partner<CUT>(rr - rnm, rr - rT, cb, s6, s12);
                    }
                }
            } else {
                for (int q = lane + 1; q <= nbn; q += G) {
                    if (g0 + x - q >= 0) { const double rl = w[x - q]; lj_partner<CUT>(rnm - rl, rT - rl, cb, s6, s12); }
                    if (g0 + x + q < N) { const double rr = w[x + q]; lj_partner<CUT>(rr - rnm, rr - rT, cb, s6, s12); }
                }
            }
            const double m6 = s6, m12 = s12;                 // this lane's share
#pragma unroll
            for (int off = 1; off < G; off <<= 1) {
                s6 += __shfl_xor_sync(gmask, s6, off);
                s12 += __shfl_xor_sync(gmask, s12, off);
            }
            // Metropolis rule :1367-1377 (see metropolis_accept, pot.cuh) without the early-out branches: decided by
            // ran - ea unless that lies within the band; dE <= 0 gives ea >= 1 > ran.  A NaN fails both tests and is
            // rejected by the exact one.
            const double dE = 4 * s12 - 4 * s6;
#ifdef JMM_ABL_NOMET
            bool accept = dE * invT < ran * 2.0;
#else
            const double gap = ran - (double) exp_neg_approx(dE * invT);
            bool accept = gap < -kMetropolisBand;
            if (fabs(gap) <= kMetropolisBand) accept = dE <= 0 || exp(-dE / T) > ran;
#endif
            if (accept && inside) {
                if (lane == 0) w[x] = rT;
                if (owned) { acc6 += m6; acc12 += m12; if (lane == 0) ++n_acc; }
            }
        }
        __syncwarp();
        // this warp's half-sweep is complete: publish it (the position writes above first), then its two sums
        if (lane32 == 0) st_release_cta(done + warp, t + 1);      // (the __syncwarp above orders the other lanes' writes before it)
        // two sums over the warp with one value per lane after the first exchange: even lanes carry s12, odd lanes s6
        {
            const bool odd = lane32 & 1;
            const double give = odd ? acc12 : acc6, keep = odd ? acc6 : acc12;
            double v = keep + __shfl_xor_sync(0xffffffffu, give, 1);
#pragma unroll
            for (int off = 2; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            const double other = __shfl_sync(0xffffffffu, v, 1);             // lane 1: the s6 total
            acc12 = v; acc6 = other;
        }
        if (lane32 == 0) wsum[t * nwarps + warp] = make_double2(acc12, acc6);
    }
    __syncthreads();

    // ---- write back the owned particles
    double *dst = S.r_out + (uint64_t) chain * S.N;
    for (int64_t g = tile_lo + threadIdx.x; g < tile_hi; g += blockDim.x) dst[g] = w[g - g0];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
        n_try += __shfl_xor_sync(0xffffffffu, n_try, off);
    }
    if ((threadIdx.x & 31) == 0 && (n_acc | n_try)) {
        atomicAdd(&counts[2 * chain], (unsigned long long) n_acc);
        atomicAdd(&counts[2 * chain + 1], (unsigned long long) n_try);
    }

    // ---- this tile's two sums per half-sweep (warps added in index order) -> partial[chain][t][tile]
    const unsigned ntiles = gridDim.x;
    double2 *part2 = reinterpret_cast<double2 *>(partial) + (uint64_t) chain * nsub * ntiles;
    for (int t = threadIdx.x; t < nsub; t += blockDim.x) {
        double s12 = 0, s6 = 0;
        for (int v = 0; v < nwarps; ++v) { const double2 q = wsum[t * nwarps + v]; s12 += q.x; s6 += q.y; }
        part2[(size_t) t * ntiles + blockIdx.x] = make_double2(s12, s6);
    }
    // ---- the last tile of the chain to get here folds the launch into the totals and the twelve sums
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&tile_done[chain], 1u);
        is_last = (prev == ntiles - 1);
        if (is_last) tile_done[chain] = 0;                 // ready for the next launch
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int t = warp; t < nsub; t += nwarps) {            // a warp per half-sweep: lanes strided over the tiles + butterfly
        double s12 = 0, s6 = 0;
        for (unsigned b = lane32; b < ntiles; b += 32) {
            const double2 q = __ldcg(&part2[(size_t) t * ntiles + b]);
            s12 += q.x; s6 += q.y;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s12 += __shfl_xor_sync(0xffffffffu, s12, off);
            s6 += __shfl_xor_sync(0xffffffffu, s6, off);
        }
        if (lane32 == 0) {
            double v[9];
            lj_nine(s6, s12, v);
#pragma unroll
            for (int k = 0; k < 9; ++k) ts[t * 9 + k] = v[k];
        }
    }
    __syncthreads();
    if (warp == 0) sweep_finish_warp(ts, nsub, S.N, S.l[chain], tot + chain * 9, acc + chain * 12, 0, lane32);
}

// Step 1, grid (nsub, nchains), 9 warps: warp k adds component k of one half-sweep over the tiles (lanes strided,
// fixed butterfly: the result does not depend on scheduling) -> tsum[chain][t][k].  All half-sweeps of a launch
// are reduced concurrently; only the running sum over t below is serial.
static __global__ void __launch_bounds__(288) k_sweep_reduce(const double *partial, int nsub, int ntiles, double *tsum) {
    const int t = blockIdx.x, chain = blockIdx.y;
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double *p = partial + (((uint64_t) chain * nsub + t) * ntiles) * 9 + k;
    double s = 0;
    for (int b = lane; b < ntiles; b += 32) s += p[(uint64_t) b * 9];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) tsum[((uint64_t) chain * nsub + t) * 9 + k] = s;
}

// Step 2, one warp per chain: running totals over the half-sweeps and one sample of the twelve sums per half-sweep.
// presample != 0: one extra sample of the current totals first (the updateThermo of src/Main.cpp:96).
static __global__ void __launch_bounds__(32) k_sweep_finish(const double *tsum, int nsub, uint64_t N, const double *l,
                                                            double *tot /*[nchains][9]*/, double *acc /*[nchains][12]*/, int presample) {
    extern __shared__ double ts[];                        // [nsub][9]
    const int chain = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < nsub * 9; i += 32) ts[i] = tsum[(uint64_t) chain * nsub * 9 + i];
    __syncwarp();
    sweep_finish_warp(ts, nsub, N, l[chain], tot + chain * 9, acc + chain * 12, presample, lane);
}

// Parallel configuration totals (SURVEY §3.3 loop): grid (nblocks, nchains), rows i strided over the
// grid, warp-shuffle + shared-memory reduction, per-block partials summed in fixed order by k_totals_finish.
// Works for both layouts through (ps, cs) = (particle stride, chain stride).
template <int POT>
__global__ void __launch_bounds__(256) k_totals_partial(const double *__restrict__ r, uint64_t ps, uint64_t cs, uint64_t N, int nbn,
                                                        double cutoff, const double *__restrict__ l, int nblocks,
                                                        double *partial /*[nchains][nblocks][9]*/) {
    constexpr int NC = PotTraits<POT>::NC;
    const uint64_t chain = blockIdx.x / nblocks;      // grid.x = nchains * nblocks (grid.y is capped at 65535)
    const uint32_t blk = blockIdx.x % nblocks;
    const double *rc = r + chain * cs;
    const double lbox = l[chain];
    double s[NC], p[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) s[k] = 0;
    for (uint64_t i = (uint64_t) blk * blockDim.x + threadIdx.x; i + 1 < N; i += (uint64_t) nblocks * blockDim.x) {
        const uint64_t jmax = (nbn < 0 || i + (uint64_t) nbn > N - 1) ? N - 1 : i + (uint64_t) nbn;
        const double ri = rc[i * ps];
        for (uint64_t j = i + 1; j <= jmax; ++j) {
            phi<POT, true>(rc[j * ps] - ri, cutoff, lbox, p);
#pragma unroll
            for (int k = 0; k < NC; ++k) s[k] += p[k];
        }
    }
    __shared__ double red[8][9];
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < NC; ++k) red[threadIdx.x >> 5][k] = s[k];
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0;
        if (threadIdx.x < NC)
            for (int wv = 0; wv < (int) (blockDim.x >> 5); ++wv) v += red[wv][threadIdx.x];
        partial[(chain * nblocks + blk) * 9 + threadIdx.x] = v;
    }
}

static __global__ void k_totals_finish(const double *partial, int nblocks, uint64_t nchains, double *out, uint64_t ks, uint64_t cs) {
    const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchains * 9) return;
    const uint64_t chain = t / 9, k = t % 9;
    double s = 0;
    for (int b = 0; b < nblocks; ++b) s += partial[(chain * nblocks + b) * 9 + k];
    out[k * ks + chain * cs] = s;
}

}  // namespace jmm
