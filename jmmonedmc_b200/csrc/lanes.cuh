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

#ifndef JMM_LANES_MAXW
#define JMM_LANES_MAXW 12
#endif
constexpr int kLanesMaxWarps = JMM_LANES_MAXW;   // 12 warps per CTA = 168 registers per thread (16 = 128: spills)

// Two sums over the G lanes with one value per lane after the first exchange: even lanes carry s12, odd lanes s6
// (log2 G + 1 shuffles of ONE double instead of log2 G of two).  Commutative additions: every lane of the group ends up
// with the same bits.  Full-warp mask: the callers keep the warp converged here.
template <int G>
__device__ __forceinline__ void lanes_butterfly(uint32_t lane, double &s6, double &s12) {
    if constexpr (G == 1) return;
    const bool odd = lane & 1;
    const double give = odd ? s12 : s6, keep = odd ? s6 : s12;
    double v = keep + __shfl_xor_sync(0xffffffffu, give, 1, G);
#pragma unroll
    for (int o = 2; o < G; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    const double other = __shfl_xor_sync(0xffffffffu, v, 1, G);
    s12 = odd ? other : v;
    s6 = odd ? v : other;
}

// This lane's share of the partner sums of one displacement trial: s6 = sum(b^-6 - a^-6), s12 = sum(b^-12 - a^-12) over
// the NPL slots p = lane + G i of the (padded) row held in rr[].  The slot of the moved particle holds kFarAway while
// this runs, so there is no per-partner index test.  (Also the loop warps of team.cuh.)
template <int POT, int G, int NPL>
__device__ __forceinline__ void lanes_row_sums(const Coop<POT, G> &c, const double (&rr)[NPL], uint32_t nm, double rnm, double rT,
                                               double &s6, double &s12) {
    double a[NPL], b[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const double r = rr[i];
        if constexpr (POT == kPotLJcut) {
            const bool left = c.lane + G * i < nm;
            a[i] = left ? rnm - r : r - rnm; b[i] = left ? rT - r : r - rT;
        } else { a[i] = r - rnm; b[i] = r - rT; }
    }
    s6 = 0; s12 = 0;
    constexpr int H = NPL > 10 ? NPL / 2 : NPL;               // at most ten chains in flight (registers)
    if constexpr (H == NPL) lj_partners<POT == kPotLJcut, NPL>(a, b, c.cutoff, s6, s12);
    else {
        double a1[H], b1[H], a2[NPL - H], b2[NPL - H];
#pragma unroll
        for (int i = 0; i < H; ++i) { a1[i] = a[i]; b1[i] = b[i]; }
#pragma unroll
        for (int i = H; i < NPL; ++i) { a2[i - H] = a[i]; b2[i - H] = b[i]; }
        lj_partners<POT == kPotLJcut, H>(a1, b1, c.cutoff, s6, s12);
        lj_partners<POT == kPotLJcut, NPL - H>(a2, b2, c.cutoff, s6, s12);
    }
}

// The partner sums of one displacement trial, strided over the G lanes and closed by the butterfly (identical bits in
// every lane of the group).  NPL > 0: every lane visits exactly NPL slots of a row padded to NPL*G positions (pads hold
// kFarAway for ever); needs NBN < 0.  NPL == 0: run-time bounds (NBN >= 0 or an unusual N).
template <int POT, int G, int NPL>
__device__ __forceinline__ void lanes_partner_sums(const Coop<POT, G> &c, uint32_t nm, double rnm, double rT, double &s6, double &s12) {
    static_assert(POT != kPotHarmonic, "fast arithmetic is an LJ-family optimisation");
    if constexpr (NPL > 0) {
        double rr[NPL];
#pragma unroll
        for (int i = 0; i < NPL; ++i) rr[i] = c.r[c.lane + G * i];
        lanes_row_sums<POT, G, NPL>(c, rr, nm, rnm, rT, s6, s12);
    } else {
        s6 = 0; s12 = 0;
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
        __syncwarp();                             // (run-time partner bounds: the groups may have left the loop apart)
    }
    lanes_butterfly<G>(c.lane, s6, s12);
}

// One chain (the G lanes of its group) advanced by `count` steps starting after step sn0.
//
// The step loop is software-pipelined, because with ~2000 warps on the machine this kernel is bound by the LATENCY
// of one step, not by instruction throughput (first version: 3000 cycles per step, 62 % of the stall samples outside
// the partner loop, profiles/r2a_c4lanes_*):
//   * the random numbers of step t+1 are shuffled out of the Philox batch at the top of step t, so that their
//     latency hides behind the partner loop of step t;
//   * the end of a trial and the beginning of the next one are ONE shared-memory transaction by lane 0 — write
//     r[nm] (new or old position), read r[nm'] of the next trial, put the far-away sentinel there — followed by one
//     shuffle of that position and one group rendezvous, instead of three rendezvous per step;
//   * the Metropolis rule is evaluated without branches (the exact expression only inside the approximation band);
//   * the four interval countdowns (ECheck, the two adjustments, relaxVolume) are one counter; a step on which any
//     of them is due takes the slow path, which first makes the positions consistent again.
template <int POT, int G, int NPL, bool LOG>
__device__ __forceinline__ void lanes_run_chain(const ChainsDev &S, const StepArgs &a, uint64_t chain, bool own, double *row, uint32_t npad,
                                                uint64_t sn0, uint32_t count, uint64_t log_row0) {
    constexpr int NC = PotTraits<POT>::NC;
    const uint64_t C = S.nchains;
    Coop<POT, G> c;
    c.lane = threadIdx.x % G;
    c.gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    c.chain_of_bonds = false;
    c.lean = true;
    c.consistent_virial = (S.flags & 1) != 0;
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
    for (int k = 0; k < kNCnt; ++k) c.cnt[k] = __ldcg(S.cnt + k * C + chain);
    c.vAErr = __ldcg(S.vAErr + chain); c.echecks = __ldcg(S.echeck + chain); c.discrepancies = __ldcg(S.echeck + C + chain);
    for (uint32_t i = c.lane; i < npad; i += G) row[i] = i < c.N ? __ldcg(S.r + (uint64_t) i * C + chain) : kFarAway;
    ThermoLanes<G> th;                            // (c.acc is not used by this kernel)
    th.init(c.sc + G * NC, c.lane, S.acc, C, chain);
    c.sync();

    const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32), cid = (uint32_t)(S.chain_id0 + chain);
    const uint32_t ntt = (uint32_t) S.numTrialTypes;
    const uint32_t scale = 0xffffffffu / ntt;
    const bool scaling_volume = (POT == kPotLJ) && S.nbn < 0;
    const bool relax_on = a.adapt_device && S.relax > 0 && S.ensemble == kEnsNPT;
    uint64_t sn = sn0;
    // steps until the next one on which ECheck, an adjustment or a relaxation is due (a launch is far shorter than 2^32 steps)
    auto until_event = [&]() -> uint32_t {
        uint64_t left = 0xffffffffull;
        auto upd = [&](uint64_t every) { if (every) left = min(left, every - sn % every); };
        upd(a.eci);
        if (a.adapt_device) { upd(a.mdai); upd(a.mvai); }
        if (relax_on && sn < 1000000ull) upd(10000);
        return (uint32_t) left;
    };
    uint32_t ev_left = until_event();

    uint32_t my_nm = 0, my_w1 = 0, my_w2 = 0;                 // this lane's share of the Philox batch
    uint32_t batch_pos = G;                                   // G = empty
    // trial type and the two words of step `step` (called with consecutive step numbers): lane j of the group holds the
    // block of the j-th step of the current batch, so one pass of Philox serves G steps
    auto draw = [&](uint64_t step, uint32_t &nm_o, uint32_t &w1_o, uint32_t &w2_o) {
        if (batch_pos == G) {
            const uint64_t mine = step + c.lane;
            const Philox4 b = philox4x32_10((uint32_t) mine, (uint32_t)(mine >> 32), cid, kTagTrial, k0, k1);
            uint32_t k = b.w[0] / scale;                      // gsl_rng_uniform_int rule, see Rng<kRngPhilox>
            if (k >= ntt) { k = b.w[3] / scale; if (k >= ntt) k = mulhi32(b.w[3], ntt); }
            my_nm = k; my_w1 = b.w[1]; my_w2 = b.w[2];
            batch_pos = 0;
        }
        nm_o = __shfl_sync(0xffffffffu, my_nm, batch_pos, G);
        w1_o = __shfl_sync(0xffffffffu, my_w1, batch_pos, G);
        w2_o = __shfl_sync(0xffffffffu, my_w2, batch_pos, G);
        ++batch_pos;
    };
    // lane 0: [write `val` to r[nm_w]] [read r[nm_r], leave the sentinel there]; everyone gets that position
    auto handover = [&](bool wr, uint32_t nm_w, double val, bool rd, uint32_t nm_r) -> double {
        double got = 0.0;
        if (c.lane == 0) {
            if (wr) c.r[nm_w] = val;
            if (rd) { got = c.r[nm_r]; c.r[nm_r] = kFarAway; }
        }
        got = __shfl_sync(0xffffffffu, got, 0, G);
        __syncwarp();
        return got;
    };

    uint32_t n_acc = 0, n_rej = 0;
    uint32_t nm, w1, w2;
    __syncwarp();
    draw(sn + 1, nm, w1, w2);
    bool disp = nm < c.N;
    double rnm = handover(false, 0, 0.0, disp, nm);

    for (uint32_t s = 0; s < count; ++s) {
        ++sn;                                                 // (the warp is converged here: every iteration ends in handover's rendezvous)
        const bool more = s + 1 < count;
        uint32_t nm1 = 0, w11 = 0, w21 = 0;
        if (more) draw(sn + 1, nm1, w11, w21);                // (uniform in the warp: every group runs `count` steps)
        const bool disp1 = more && nm1 < c.N;

        // The whole warp runs the displacement code CONVERGED (full-mask shuffles: a per-group mask costs a MATCH + REDUX +
        // VOTE per shuffle): a group whose trial is a volume trial computes a discarded dummy, a move through the wall is
        // a predicate on the decision (:1188), not a branch around the pair terms.
        uint8_t flags = 0;
        bool acc = false;
        double rT;
        {                                                     // qad2 :1160-1464
            const double md = u01_shifted(w1, 1.5) * 2 * c.maxStep;     // (rn - 0.5) * 2 * maxStep, :1182
            rT = rnm + md;
            const bool wall = fabs(rT) > c.half_l;            // :1188
            double s6, s12;
            lanes_partner_sums<POT, G, NPL>(c, disp ? nm : 0u, rnm, rT, s6, s12);
            const double dE12 = 4 * s12, dE6 = 4 * s6;
            const double dE = dE12 - dE6;
            // Metropolis rule :1367-1377 through the band of metropolis_accept(), without early-out branches
            const double ran = u01_shifted(w2, 1.0);
            const double ea = (double) exp_neg_approx(dE * c.invT);
            const bool down = dE <= 0;
            const bool acc_b = ran < ea - kMetropolisBand, rej_b = ran > ea + kMetropolisBand;
            acc = down | acc_b;
            if (disp && !wall && !(down | acc_b | rej_b)) acc = metropolis_exact(dE, c.T, ran);
            acc = acc && disp && !wall;
            if (acc) {
                const double dV12 = 12 * dE12, dV6 = 6 * dE6, dH12 = 144 * dE12, dH6 = 36 * dE6;
                c.tot[0] += dE;  c.tot[2] += dE12; c.tot[4] += dE6;
                c.tot[1] += dV12 - dV6; c.tot[3] += dV12; c.tot[5] += dV6;
                c.tot[6] += dH12 - dH6; c.tot[7] += dH12; c.tot[8] += dH6;
            }
            if (LOG) flags = (disp && wall) ? kLogWall : (acc ? kLogAccepted : 0);
            n_acc += acc ? 1u : 0u;                           // (32-bit tallies, folded into the 64-bit counters at events)
            n_rej += (disp && !acc) ? 1u : 0u;
        }
        if (!disp) {                                          // volume trial: no sentinel is out, the row is consistent
            const double rn = u01(w1), ran = u01(w2);
            th.flush(c);                                      // (fav's ordered sums use the scratch next to the ring; rare anyway)
            if constexpr (POT == kPotLJ) {
                flags = scaling_volume ? coop_volume_scaling(c, rn, ran) : coop_volume_full(c, rn, ran);
            } else flags = coop_volume_full(c, rn, ran);
        }
        __syncwarp();                                         // converged again

        double rnm1;
        if (--ev_left != 0) {
            th.push(c);                                       // updateThermo :1805 (the state of this step; the sums follow in flush)
            rnm1 = handover(disp, nm, acc ? rT : rnm, disp1, nm1);
            if (th.fill == kThermoRing) th.flush(c);
        } else {
            handover(disp, nm, acc ? rT : rnm, false, 0);     // positions consistent for whatever is due now
            c.cnt[0] += n_acc; c.cnt[1] += n_rej; n_acc = n_rej = 0;
            th.flush(c);
            if (a.eci && sn % a.eci == 0) coop_energy_check(c);                        // Step :1800
            th.push(c); th.flush(c);                                                     // :1805
            if (a.adapt_device) {                                                        // src/Main.cpp:145-176
                if (a.mdai && sn % a.mdai == 0) coop_adjust_max_step(c, a.log_ideal);
                if (a.mvai && sn % a.mvai == 0) coop_adjust_max_dl(c, a.log_ideal);
                if (relax_on && sn % 10000 == 0 && sn < 1000000ull) coop_relax_volume(c);
            }
            ev_left = until_event();
            rnm1 = handover(false, 0, 0.0, disp1, nm1);
        }
        if (LOG && own && c.lane == 0) a.accept_log[(log_row0 + s) * C + chain] = flags;
        nm = nm1; w1 = w11; w2 = w21; rnm = rnm1; disp = disp1;
    }

    th.flush(c);
    c.cnt[0] += n_acc; c.cnt[1] += n_rej;
    if (!own) return;                                         // a surplus group of a ragged last tile shadows the last chain
    th.store(c.lane, S.acc, C, chain);
    for (uint32_t i = c.lane; i < c.N; i += G) S.r[(uint64_t) i * C + chain] = row[i];
    if (c.lane == 0) {
        S.l[chain] = c.l; S.maxStep[chain] = c.maxStep; S.maxdl[chain] = c.maxdl;
#pragma unroll
        for (int k = 0; k < NC; ++k) S.tot[k * C + chain] = c.tot[k];
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
        const bool own = chain < S.nchains;
        lanes_run_chain<POT, G, NPL, LOG>(S, a, own ? chain : S.nchains - 1, own, row, npad, a.sn0 + s0, count, s0);
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(progress + tile), "r"(k + 1) : "memory");
    }
}

}  // namespace jmm
