// Many-chain production kernel for a MODERATE number of chains (a state-point sweep sharded over several GPUs:
// 65 536 chains / 8 GPUs = 8192 chains of N = 80 per GPU, scripts/RunJobs.bash:14-27): G lanes of a warp share one
// chain and split the partner loop of qad2 (:1160-1464) between them, with the JMM_ARITH_FAST arithmetic of
// fastlj.cuh.
//
// Why a third many-chain kernel.  prod.cuh (one chain per thread) needs >= ~50 000 chains to fill 148 SMs: 8192
// chains are 256 warps on 592 sub-partitions.  coop.cuh (G lanes per chain) fills the machine but keeps the
// reference's arithmetic and order of addition through a shared-memory scratch: ~76 fp64 instructions per partner.
// Here the partner loop is prod.cuh's — 18 fp64-pipe instructions per partner, the moved particle taken out by a
// far-away sentinel, NO per-partner index test — strided over the G lanes (80 positions = 10 per lane at G = 8, fully
// unrolled: ten independent reciprocal chains in flight per lane), the two sums s6, s12 closed by an xor butterfly
// (identical bits in every lane), and everything that is not the partner loop (Philox, the Metropolis rule, the
// nine totals, the twelve running sums) is group-uniform.  The per-warp instruction count of the partner loop is
// the same as prod.cuh's (N/G iterations serve 32/G chains), so the fp64 pipe sees the same work per trial; what is
// added is the non-loop work, now per 32/G chains instead of per 32.
//
// Launch: persistent, time-sliced like k_chains_step_prod_sliced — one CTA per SM, work items (chunk, tile) handed to
// the WARPS by an atomic counter, a tile (32/G chains) per warp — so that any number of chains spreads evenly over
// the 592 sub-partitions (2048 tiles on 148 x 16 = 2368 warp slots would otherwise leave 20 SMs idle).
//
// Rare paths (volume trials, ECheck, relaxVolume, step-size adaptation) are coop.cuh's functions with the lean
// ordered sums (reference order of addition: relaxVolume feeds the positions, so its sums must be the oracle's).
// Positions and accept/reject sequences are bit-identical to the oracle; totals agree to <= 1e-12 (as prod.cuh fast).
#pragma once
#include "coop.cuh"
#include "fastlj.cuh"
#include "prod.cuh"   // kFarAway

namespace jmm {

constexpr int kLanesMaxWarps = 16;

// qad2 with the fast arithmetic, partners strided over the G lanes of the group.
// NPL > 0: every lane visits exactly NPL slots p = lane + G i of a row padded to NPL*G positions (pads hold kFarAway
// for ever); needs NBN < 0.  NPL == 0: run-time bounds (NBN >= 0 or an unusual N).
template <int POT, int G, int NPL>
__device__ __forceinline__ uint8_t lanes_displacement(Coop<POT, G> &c, uint32_t nm, double rn, double ran) {
    static_assert(POT != kPotHarmonic, "fast arithmetic is an LJ-family optimisation");
    const double md = (rn - 0.5) * 2 * c.maxStep;
    const double rnm = c.r[nm];
    const double rT = rnm + md;
    if (fabs(rT) > c.half_l) { c.cnt[1]++; return kLogWall; }
    c.sync();                                     // every lane has read r[nm]
    if (c.lane == 0) c.r[nm] = kFarAway;          // the moved particle leaves the loop without a per-partner test
    c.sync();
    double s6 = 0, s12 = 0;
    if constexpr (NPL > 0) {
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const uint32_t p = c.lane + G * i;
            const double r = c.r[p];
            if constexpr (POT == kPotLJcut) {
                const bool left = p < nm;
                lj_partner<true>(left ? rnm - r : r - rnm, left ? rT - r : r - rT, c.cutoff, s6, s12);
            } else {
                lj_partner<false>(r - rnm, r - rT, c.cutoff, s6, s12);
            }
        }
    } else {
        const uint32_t N = c.N;
        const uint32_t lo = (c.nbn < 0 || (uint32_t) c.nbn > nm) ? 0u : nm - (uint32_t) c.nbn;
        const uint32_t hi = (c.nbn < 0 || nm + (uint32_t) c.nbn > N - 1) ? N - 1 : nm + (uint32_t) c.nbn;
#pragma unroll 4
        for (uint32_t p = lo + c.lane; p <= hi; p += G) {
            const double r = c.r[p];
            if constexpr (POT == kPotLJcut) {
                const bool left = p < nm;
                lj_partner<true>(left ? rnm - r : r - rnm, left ? rT - r : r - rT, c.cutoff, s6, s12);
            } else {
                lj_partner<false>(r - rnm, r - rT, c.cutoff, s6, s12);
            }
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {         // commutative additions: the same bits in every lane of the group
        s6 += __shfl_xor_sync(c.gmask, s6, o, G);
        s12 += __shfl_xor_sync(c.gmask, s12, o, G);
    }
    const double dE12 = 4 * s12, dE6 = 4 * s6;
    const double dE = dE12 - dE6;
    const bool acc = metropolis_accept(dE, c.T, c.invT, ran);
    if (c.lane == 0) c.r[nm] = acc ? rT : rnm;    // (every lane is past its loads: the butterfly is a rendezvous)
    c.sync();
    if (!acc) { c.cnt[1]++; return 0; }
    c.cnt[0]++;
    const double dV12 = 12 * dE12, dV6 = 6 * dE6, dH12 = 144 * dE12, dH6 = 36 * dE6;
    c.tot[0] += dE;  c.tot[2] += dE12; c.tot[4] += dE6;
    c.tot[1] += dV12 - dV6; c.tot[3] += dV12; c.tot[5] += dV6;
    c.tot[6] += dH12 - dH6; c.tot[7] += dH12; c.tot[8] += dH6;
    return kLogAccepted;
}

// One chain (the G lanes of its group) advanced by `count` steps starting after step sn0.
template <int POT, int G, int NPL, bool LOG>
__device__ __forceinline__ void lanes_run_chain(const ChainsDev &S, const StepArgs &a, uint64_t chain, double *row, uint32_t npad,
                                                uint64_t sn0, uint32_t count, uint64_t log_row0) {
    constexpr int NC = PotTraits<POT>::NC;
    const uint64_t C = S.nchains;
    Coop<POT, G> c;
    c.lane = threadIdx.x % G;
    c.gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    c.chain_of_bonds = false;
    c.lean = true;
    c.r = row;
    c.sc = row + npad;
    c.N = (uint32_t) S.N; c.nbn = S.nbn; c.cutoff = S.cutoff;
    // (ld.global.cg: another SM may have written this chain's state a moment ago, in the previous chunk)
    c.P = __ldcg(S.P + chain); c.T = __ldcg(S.T + chain); c.maxStep = __ldcg(S.maxStep + chain); c.maxdl = __ldcg(S.maxdl + chain);
    c.invT = 1.0 / c.T;
    c.set_l(__ldcg(S.l + chain));
#pragma unroll
    for (int k = 0; k < NC; ++k) c.tot[k] = __ldcg(S.tot + k * C + chain);
#pragma unroll
    for (int k = 0; k < kNAcc; ++k) c.acc[k] = __ldcg(S.acc + k * C + chain);
#pragma unroll
    for (int k = 0; k < kNCnt; ++k) c.cnt[k] = __ldcg(S.cnt + k * C + chain);
    c.vAErr = __ldcg(S.vAErr + chain); c.echecks = __ldcg(S.echeck + chain); c.discrepancies = __ldcg(S.echeck + C + chain);
    for (uint32_t i = c.lane; i < npad; i += G) row[i] = i < c.N ? __ldcg(S.r + (uint64_t) i * C + chain) : kFarAway;
    c.sync();

    const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32), cid = (uint32_t)(S.chain_id0 + chain);
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

    uint32_t my_nm = 0, my_w1 = 0, my_w2 = 0;                 // this lane's share of the Philox batch
    uint32_t batch_pos = G;                                   // G = empty

    for (uint32_t s = 0; s < count; ++s) {
        ++sn;
        if (batch_pos == G) {                                 // lane j draws the block of step sn + j
            const uint64_t mine = sn + c.lane;
            const Philox4 b = philox4x32_10((uint32_t) mine, (uint32_t)(mine >> 32), cid, kTagTrial, k0, k1);
            uint32_t k = b.w[0] / scale;                      // gsl_rng_uniform_int rule, see Rng<kRngPhilox>
            if (k >= ntt) { k = b.w[3] / scale; if (k >= ntt) k = mulhi32(b.w[3], ntt); }
            my_nm = k; my_w1 = b.w[1]; my_w2 = b.w[2];
            batch_pos = 0;
        }
        const uint32_t nm = __shfl_sync(c.gmask, my_nm, batch_pos, G);
        const double rn = u01(__shfl_sync(c.gmask, my_w1, batch_pos, G));
        const double ran = u01(__shfl_sync(c.gmask, my_w2, batch_pos, G));
        ++batch_pos;

        uint8_t flags;
        if (nm < c.N) flags = lanes_displacement<POT, G, NPL>(c, nm, rn, ran);
        else {
            if constexpr (POT == kPotLJ) {
                flags = scaling_volume ? coop_volume_scaling(c, rn, ran) : coop_volume_full(c, rn, ran);
            } else flags = coop_volume_full(c, rn, ran);
        }
        if (--eci_left == 0) { coop_energy_check(c); eci_left = eci32; }
        coop_update_thermo(c);
        if (LOG && c.lane == 0) a.accept_log[(log_row0 + s) * C + chain] = flags;
        if (a.adapt_device) {
            if (--mdai_left == 0) { coop_adjust_max_step(c, a.log_ideal); mdai_left = mdai32; }
            if (--mvai_left == 0) { coop_adjust_max_dl(c, a.log_ideal); mvai_left = mvai32; }
            if (--relax_left == 0) { if (sn < 1000000ull) coop_relax_volume(c); relax_left = 10000; }
        }
    }

    c.sync();
    for (uint32_t i = c.lane; i < c.N; i += G) S.r[(uint64_t) i * C + chain] = row[i];
    if (c.lane == 0) {
        S.l[chain] = c.l; S.maxStep[chain] = c.maxStep; S.maxdl[chain] = c.maxdl;
#pragma unroll
        for (int k = 0; k < NC; ++k) S.tot[k * C + chain] = c.tot[k];
#pragma unroll
        for (int k = 0; k < kNAcc; ++k) S.acc[k * C + chain] = c.acc[k];
#pragma unroll
        for (int k = 0; k < kNCnt; ++k) S.cnt[k * C + chain] = c.cnt[k];
        S.vAErr[chain] = c.vAErr; S.echeck[chain] = c.echecks; S.echeck[C + chain] = c.discrepancies;
    }
}

// Persistent grid: warps take (chunk, tile) items, chunk-major, from an atomic counter; chunk k+1 of a tile waits for
// its chunk k through a per-tile progress word (release/acquire); state travels through L2 between chunks.
// `stride` = doubles per group in shared memory (row of npad positions + the [G][NC] scratch, padded so that the
// groups of a half-warp start G banks apart).
template <int POT, int G, int NPL, bool LOG>
__global__ void __launch_bounds__(kLanesMaxWarps * 32, 1) k_chains_step_lanes(ChainsDev S, StepArgs a, uint32_t chunk, uint32_t ntiles,
                                                                              uint32_t nchunks, uint32_t npad, uint32_t stride,
                                                                              unsigned int *work, unsigned int *progress) {
    extern __shared__ double smem[];
    constexpr uint32_t CPW = 32 / G;                          // chains per warp = per tile
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *row = smem + ((size_t) warp * CPW + lane / G) * stride;
    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(work, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= ntiles * nchunks) return;
        const uint32_t tile = item % ntiles, k = item / ntiles;
        if (lane == 0 && k > 0) {
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(progress + tile) : "memory");
                if (seen < k) __nanosleep(200);
            } while (seen < k);
        }
        __syncwarp();
        const uint64_t chain = (uint64_t) tile * CPW + lane / G;
        const uint32_t s0 = k * chunk;
        const uint32_t count = min(chunk, (uint32_t) a.nsteps - s0);
        if (chain < S.nchains) lanes_run_chain<POT, G, NPL, LOG>(S, a, chain, row, npad, a.sn0 + s0, count, s0);
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(progress + tile), "r"(k + 1) : "memory");
    }
}

}  // namespace jmm
