// Production many-chain kernel for PLENTIFUL chains (configs C4-like: >= ~16k chains): one chain per
// thread, positions in an [N][32] shared-memory tile addressed with compile-time strides, Philox stream,
// positions-only arithmetic.  Same trial moves as chains.cuh (qad2 :1160-1464 etc.), but the partner loop
// is written for issue efficiency, because this regime is bound by the fp64 pipe, not by latency:
//   * shared-memory loads with a constant 256-byte stride instead of generic 64-bit address arithmetic
//     (the generic kernel spent 48 IMAD per partner on it, profiles/r01_c4_k_chains_step.txt);
//   * JMM_ARITH_REFERENCE: every pair term and every sum exactly as the reference/oracle (bit-identical);
//   * JMM_ARITH_FAST (LJ, LJcut): per partner ONE division serves the old and the new term
//     (1/(A*B) with A = d_old^6, B = d_new^6), only the r^-6 and r^-12 differences are accumulated, and the
//     nine deltas are derived from them by the exact ratios of src/pot.cpp:56-66
//     (Vir12 = 12 E12, Vir6 = 6 E6, HV12 = 144 E12, HV6 = 36 E6).  ~30 fp64 instructions per partner
//     instead of ~76.  Totals differ from the reference-order sums by rounding only (<= 1e-12 relative);
//     an accept/reject decision can differ only if exp(-dE/T) and ran agree to ~1e-15.
#pragma once
#include "chains.cuh"
#include "fastlj.cuh"

namespace jmm {

constexpr int kArithReference = 0, kArithFast = 1;
constexpr int kTile = 32;                        // chains per block = stride between particles in the tile

template <int POT>
__device__ __forceinline__ uint8_t prod_displacement_ref(Chain<POT> &ch, const double *rs /*shared tile column*/,
                                                         double *rs_w, uint32_t nm, double rn, double ran) {
    constexpr int NC = PotTraits<POT>::NC;
    const double md = (rn - 0.5) * 2 * ch.maxStep;
    const double rnm = rs[nm * kTile];
    const double rT = rnm + md;
    if (fabs(rT) > ch.l / 2.0) { ch.cnt[1]++; return kLogWall; }
    const uint32_t N = ch.N;
    const uint32_t lo = (ch.nbn < 0 || (uint32_t) ch.nbn > nm) ? 0u : nm - (uint32_t) ch.nbn;
    const uint32_t hi = (ch.nbn < 0 || nm + (uint32_t) ch.nbn > N - 1) ? N - 1 : nm + (uint32_t) ch.nbn;
    double dsum[NC], dleft[NC], po[NC], pn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { dsum[k] = 0; dleft[k] = 0; }
    const double *rp = rs + lo * kTile;
    for (uint32_t p = lo; p <= hi; ++p, rp += kTile) {
        if (p == nm) {
#pragma unroll
            for (int k = 0; k < NC; ++k) { dleft[k] = dsum[k]; dsum[k] = 0; }
            continue;
        }
        const bool left = p < nm;
        const double r = *rp;
        phi<POT, true>(left ? rnm - r : r - rnm, ch.cutoff, ch.l, po);
        phi<POT, true>(left ? rT - r : r - rT, ch.cutoff, ch.l, pn);
#pragma unroll
        for (int k = 0; k < NC; ++k) dsum[k] = dsum[k] - po[k] + pn[k];
    }
    double d[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) d[k] = dleft[k] + dsum[k];
    if (!metropolis_accept(d[0], ch.T, ch.invT, ran)) { ch.cnt[1]++; return 0; }
    ch.cnt[0]++;
#pragma unroll
    for (int k = 0; k < NC; ++k) ch.tot[k] += d[k];
    rs_w[nm * kTile] = rT;
    return kLogAccepted;
}

// LJ / LJcut only: fastlj.cuh arithmetic, one chain per thread.
//
// The moved particle itself is taken out of the loop without a per-partner test: its slot in the tile is
// overwritten with a far-away sentinel for the duration of the loop.  LJcut masks the sentinel "pair" to exactly
// zero (distance > cutoff); LJ adds ~6|md|/1e56 to s6, far below one ulp of any partner term.
//
// Measured and dropped: G = 2 or 4 lanes per chain (lane-strided partners, xor-butterfly of the two sums, [N][32/G]
// tile so that G times more warps are resident).  C4, trial moves/s: G = 1 7.25e9, G = 2 6.22e9, G = 4 4.35e9 — the
// per-step work every lane of a group repeats and the 128/80-register budgets cost more than the extra warps give.
constexpr double kFarAway = 1.0e8;

// UNROLL = partners in flight per thread (independent reciprocal chains)
template <int POT, int UNROLL>
__device__ __forceinline__ uint8_t prod_displacement_fast(Chain<POT> &ch, double *rs, uint32_t nm, double rn, double ran) {
    static_assert(POT != kPotHarmonic, "fast arithmetic is an LJ-family optimisation");
    const double md = (rn - 0.5) * 2 * ch.maxStep;
    const double rnm = rs[nm * kTile];
    const double rT = rnm + md;
    if (fabs(rT) > ch.l / 2.0) { ch.cnt[1]++; return kLogWall; }
    const uint32_t N = ch.N;
    const uint32_t lo = (ch.nbn < 0 || (uint32_t) ch.nbn > nm) ? 0u : nm - (uint32_t) ch.nbn;
    const uint32_t hi = (ch.nbn < 0 || nm + (uint32_t) ch.nbn > N - 1) ? N - 1 : nm + (uint32_t) ch.nbn;
    rs[nm * kTile] = kFarAway;
    const double cb = ch.cutoff;
    double s6 = 0, s12 = 0;
    const double *rp = rs + lo * kTile;
#pragma unroll (UNROLL)
    for (uint32_t p = lo; p <= hi; ++p, rp += kTile) {
        const double r = *rp;
        if constexpr (POT == kPotLJcut) {
            // signed distances in the reference's orientation: `d <= cutOff` is a test on the signed d (src/pot.cpp:53)
            const bool left = p < nm;
            lj_partner<true>(left ? rnm - r : r - rnm, left ? rT - r : r - rT, cb, s6, s12);
        } else {
            lj_partner<false>(r - rnm, r - rT, cb, s6, s12);      // even powers only: the orientation does not matter
        }
    }
    const double dE12 = 4 * s12, dE6 = 4 * s6;
    const double dE = dE12 - dE6;
    if (!metropolis_accept(dE, ch.T, ch.invT, ran)) { rs[nm * kTile] = rnm; ch.cnt[1]++; return 0; }
    ch.cnt[0]++;
    const double dV12 = 12 * dE12, dV6 = 6 * dE6, dH12 = 144 * dE12, dH6 = 36 * dE6;
    ch.tot[0] += dE;  ch.tot[2] += dE12; ch.tot[4] += dE6;
    ch.tot[1] += dV12 - dV6; ch.tot[3] += dV12; ch.tot[5] += dV6;
    ch.tot[6] += dH12 - dH6; ch.tot[7] += dH12; ch.tot[8] += dH6;
    rs[nm * kTile] = rT;
    return kLogAccepted;
}

// One tile (32 chains) advanced by `count` steps starting after step sn0 (log rows from log_row0).
// own = false: a surplus lane of a ragged last tile; it shadows the last chain and stores nothing, so that the
// warp-wide histogram refill (hist_after_trial) always sees a whole warp.
// COHERENT: chain state is read with ld.global.cg (L2) because another SM may just have written it
// (persistent time-sliced launch below).
template <int POT, int ARITH, bool LOG, bool COHERENT, int UNROLL>
__device__ __forceinline__ void prod_run_tile(const ChainsDev &S, const StepArgs &a, const HistDev &H, uint64_t c, double *smem,
                                              uint64_t sn0, uint32_t count, uint64_t log_row0, bool own = true) {
    const bool lead = own;                                   // the lane that owns the chain's global state
    Chain<POT> ch;
    const uint32_t lane = threadIdx.x & 31;                  // a tile belongs to ONE warp; a CTA holds several tiles
    load_chain<POT, COHERENT>(ch, S, c, smem + lane, kTile);   // ch.r -> shared tile (rare paths use it generically)
    double *rs = smem + lane;                                // same column, known to be shared memory

    Rng<kRngPhilox> rng;
    rng.k0 = (uint32_t) S.seed; rng.k1 = (uint32_t)(S.seed >> 32); rng.chain = (uint32_t)(S.chain_id0 + c);
    const uint32_t ntt = (uint32_t) S.numTrialTypes;
    const uint32_t scale = 0xffffffffu / ntt;
    const bool scaling_volume = (POT == kPotLJ) && S.nbn < 0;
    uint64_t sn = sn0;
    auto until = [&](uint64_t every) -> uint32_t {
        if (!every) return 0xffffffffu;
        const uint64_t left = every - sn % every;
        return left > 0xfffffffeull ? 0xffffffffu : (uint32_t) left;
    };
    uint32_t eci_left = until(a.eci);
    uint32_t mdai_left = a.adapt_device ? until(a.mdai) : 0xffffffffu;
    uint32_t mvai_left = a.adapt_device ? until(a.mvai) : 0xffffffffu;
    uint32_t relax_left = (a.adapt_device && S.relax > 0 && S.ensemble == kEnsNPT) ? until(10000) : 0xffffffffu;
    const uint32_t eci32 = a.eci > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.eci;
    const uint32_t mdai32 = a.mdai > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.mdai;
    const uint32_t mvai32 = a.mvai > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.mvai;
    double l_seen = ch.l, rho = (double) ch.N / ch.l;         // N/l only changes with l
    uint64_t hist_u = H.ucount ? ld_state<COHERENT>(H.ucount + c) : 0;

    for (uint32_t s = 0; s < count; ++s) {
        ++sn;
        rng.begin(sn);
        const uint32_t nm = rng.trial_type(ntt, scale);
        const double rn = rng.rn();
        const double maxStep_used = ch.maxStep;
        uint8_t flags;
        double vscale = 0.0;                                  // != 0: an accepted volume move still has to scale the positions
        if (nm < ch.N) {
            if constexpr (ARITH == kArithFast) flags = prod_displacement_fast<POT, UNROLL>(ch, rs, nm, rn, rng.ran());
            else flags = prod_displacement_ref<POT>(ch, rs, rs, nm, rn, rng.ran());
        } else {
            if constexpr (POT == kPotLJ) {
                flags = scaling_volume ? volume_trial_scaling<POT, false>(ch, rn, rng, &vscale)
                                       : volume_trial_full<POT, false>(ch, rn, rng, &vscale);
            } else flags = volume_trial_full<POT, false>(ch, rn, rng, &vscale);
        }
        {   // r *= s of an accepted volume move (qavLJ :1692, fav :2264-2266), done by the whole warp: a third of the
            // steps have a volume trial in some lane, and one lane walking its N positions alone (N dependent
            // load-multiply-store round trips while 31 lanes idle) was 10 % of the stall samples of C4
            // (profiles/r01m_c4_k_chains_step_prod_sliced_fast.txt).  Same products, so the same positions.
            __syncwarp();
            unsigned pend = __ballot_sync(0xffffffffu, vscale != 0.0);
            while (pend) {
                const int src = __ffs(pend) - 1;
                pend &= pend - 1;
                const double f = __shfl_sync(0xffffffffu, vscale, src);
                double *col = smem + src;
                for (uint32_t i = lane; i < ch.N; i += kTile) col[i * kTile] = col[i * kTile] * f;
            }
            __syncwarp();
        }
        if (H.ucount) hist_after_trial<POT, false>(H, c, ch, nm, rn, maxStep_used, flags, scaling_volume, hist_u, lead, 0xffffffffu);
        if (--eci_left == 0) { energy_check<POT, false>(ch); eci_left = eci32; }
        if (ch.l != l_seen) { l_seen = ch.l; rho = (double) ch.N / ch.l; }
        {   // updateThermo :1941-1961 with the cached N/l
            const double E = ch.tot[0], Vir = ch.tot[1];
            ch.acc[0] = ch.acc[0] + rho;        ch.acc[1] = ch.acc[1] + rho * rho;
            ch.acc[2] = ch.acc[2] + ch.l;       ch.acc[3] = ch.acc[3] + ch.l * ch.l;
            ch.acc[4] = ch.acc[4] + E;          ch.acc[5] = ch.acc[5] + E * E;
            ch.acc[6] = ch.acc[6] + ch.l * E;   ch.acc[7] = ch.acc[7] + Vir;
            ch.acc[8] = ch.acc[8] + Vir * Vir;  ch.acc[9] = ch.acc[9] + E * Vir;
            if constexpr (PotTraits<POT>::NC > 6) {
                const double HV = ch.tot[6];
                ch.acc[10] = ch.acc[10] + HV;   ch.acc[11] = ch.acc[11] + HV * HV;
            }
        }
        if (LOG && lead) a.accept_log[(log_row0 + s) * S.nchains + c] = flags;
        if (a.adapt_device) {
            if (--mdai_left == 0) { adjust_max_step(ch, a.log_ideal); mdai_left = mdai32; }
            if (--mvai_left == 0) { adjust_max_dl(ch, a.log_ideal); mvai_left = mvai32; }
            if (--relax_left == 0) { if (sn < 1000000ull) relax_volume<POT, false>(ch); relax_left = 10000; }
        }
    }
    if (lead) {
        if (H.ucount) H.ucount[c] = hist_u;
        store_chain(ch, S, c, true);
    }
}

// One CTA per SM, one tile per WARP: the [80][32] tile of C4 is 20 KB, and ten one-warp CTAs were all that fit
// (every CTA also reserves 1 KB of shared memory), but one CTA of eleven warps holds eleven tiles in 220 KB of its
// 227 KB.  The warps of a CTA never synchronise with each other.  Fast arithmetic: <= 12 warps (65536 / (12 x 32) =
// 170 registers, it needs 164); reference arithmetic: <= 10 warps (204 registers, as with the one-warp CTAs).
template <int ARITH> constexpr int kProdMaxWarps = ARITH == kArithFast ? 12 : 10;

template <int POT, int ARITH, bool LOG, int UNROLL>
__global__ void __launch_bounds__(kProdMaxWarps<ARITH> * kTile, 1) k_chains_step_prod(ChainsDev S, StepArgs a, HistDev H, uint32_t ntiles) {
    extern __shared__ double smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const uint32_t tile = blockIdx.x * nwarps + warp;
    if (tile >= ntiles) return;                              // (a whole warp)
    const uint64_t c = (uint64_t) tile * kTile + lane;
    const bool own = c < S.nchains;
    prod_run_tile<POT, ARITH, LOG, false, UNROLL>(S, a, H, own ? c : S.nchains - 1, smem + (size_t) warp * S.N * kTile, a.sn0,
                                                  (uint32_t) a.nsteps, 0, own);
}

// Persistent, time-sliced variant for launches that would otherwise need a fractional number of waves
// (65 536 chains = 2048 tiles on 1628 resident warps = 1.26 waves: the last 0.26 wave leaves most SMs idle).
// The grid is exactly the number of co-resident CTAs.  Work items are (chunk, tile) pairs, chunk-major, handed
// out to the WARPS by an atomic counter; a tile's chunk k+1 waits for its chunk k through a per-tile progress word
// (release/acquire).  Every earlier item is held by a running warp, so the wait always ends.  Chain state goes
// through L2 between chunks (8N+256 B per chain per chunk: negligible next to `chunk` steps of work).
template <int POT, int ARITH, bool LOG, int UNROLL>
__global__ void __launch_bounds__(kProdMaxWarps<ARITH> * kTile, 1) k_chains_step_prod_sliced(ChainsDev S, StepArgs a, HistDev H, uint32_t chunk, uint32_t ntiles,
                                                                   uint32_t nchunks, unsigned int *work, unsigned int *progress) {
    extern __shared__ double smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *tile_smem = smem + (size_t) warp * S.N * kTile;
    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(work, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= ntiles * nchunks) return;
        const uint32_t tile = item % ntiles, k = item / ntiles;
        if (lane == 0) {
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(progress + tile) : "memory");
                if (seen < k) __nanosleep(200);
            } while (seen < k);
        }
        __syncwarp();
        const uint64_t c = (uint64_t) tile * kTile + lane;
        const uint32_t s0 = k * chunk;
        const uint32_t count = min(chunk, (uint32_t) a.nsteps - s0);
        const bool own = c < S.nchains;
        prod_run_tile<POT, ARITH, LOG, true, UNROLL>(S, a, H, own ? c : S.nchains - 1, tile_smem, a.sn0 + s0, count, s0, own);
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(progress + tile), "r"(k + 1) : "memory");
    }
}

}  // namespace jmm
