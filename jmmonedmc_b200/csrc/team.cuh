// k_chains_step_team — the moderate-chain-count regime (8192 ... 16 384 chains of N <= 80 per GPU: a sweep sharded over
// 4-8 GPUs) with the ACCOUNTING of a step pooled over 28 chains and taken off the step's critical path.
//
// lanes.cuh gives a chain G = 8 lanes and runs four chains per warp; the partner loop then costs what prod.cuh's costs
// per trial, but everything else of a step is executed once per FOUR chains instead of once per 28-32, and that holds
// the kernel at 44 % of the fp64 pipe.  Here warps are specialised:
//   * a TEAM = seven LOOP warps + one BOOKKEEPER warp serves 28 chains;
//   * loop warp w owns chains 4w ... 4w+3 in lanes.cuh's layout (8 lanes per chain, 10 partners per lane, stage-wise
//     evaluation, butterfly, Philox batched over the lanes) and runs the TRIALS of its chains on its own: draw, move,
//     partner sums, Metropolis decision, commit of the position, sentinel of the next trial.  It does NOT keep totals,
//     thermodynamic sums or counters: per step it leaves (nm, accepted, wall, s6, s12) of each chain in a ring in
//     shared memory and goes on;
//   * the bookkeeper holds ONE CHAIN PER LANE (scalars in registers, as prod.cuh) and trails the loop warps through the
//     ring: nine totals, updateThermo's twelve sums, the counters, the accept log — once per 28 chains — with the
//     reference's operation order per chain, so the results are what lanes.cuh gives;
//   * the steps that need the totals or touch a whole row are done BY the bookkeeper with chains.cuh's per-thread
//     functions on the shared rows: volume trials (qavLJ / fav), ECheck, the two adjustments, relaxVolume.  A loop
//     warp that meets one (a volume trial of one of its chains: 1 step in 81 per chain; an interval step: all warps)
//     finishes the step, publishes it, and waits until the bookkeeper has booked that step; the other loop warps run on.
//   * flow control = three kinds of words in shared memory, written with st.release and polled with ld.acquire:
//     tail[w] (loop warp w has finished steps 1..tail[w] of the chunk), booked (the bookkeeper has accounted steps
//     1..booked), and the ring depth: a loop warp starts step s only when s <= booked + kTeamRing.
// (The first version of this file — profiles/r2t_c4team_lockstep_barriers.txt — had the bookkeeper also draw, set up
// and decide every trial, with two named barriers per step: 34 % fewer instructions than lanes.cuh, but ~475 serial
// bookkeeper instructions on every step's critical path; 5.63e9 against 6.08e9 trial moves/s.)
// Arithmetic of a trial = lanes.cuh's (same row sums, same butterfly, same decision code): decisions and positions are
// bit-identical to the oracle, totals to <= 1e-12.
#pragma once
#include "lanes.cuh"

namespace jmm {

// Seven loop warps + the bookkeeper = 8 warps and 28 chains per team (bookkeeper lanes 28-31 idle): two teams per SM are 16
// warps = 128 registers per thread (nine-warp teams are allocated as 20 warps: 96 registers, and the partner loop spills),
// and 8192 chains are 293 tiles on 148 x 2 = 296 team slots: one wave, nothing to time-slice.
constexpr int kTeamLoopWarps = 7;
constexpr int kTeamWarps = kTeamLoopWarps + 1;
constexpr int kTeamChains = 4 * kTeamLoopWarps;
constexpr int kTeamG = 8, kTeamNPL = 10, kTeamRow = 88;   // lanes per chain, slots per lane, doubles per row (80 + pad, = 8 mod 16)
#ifndef JMM_TEAM_NS
#define JMM_TEAM_NS 40                                    // pause between two looks of an idle bookkeeper at the loop warps' progress
#endif
constexpr int kTeamRing = 16;                             // steps a loop warp may be ahead of the bookkeeper
constexpr uint32_t kTeamAccepted = 1u << 16, kTeamWall = 1u << 17;   // result word = nm | flags

struct TeamShared {                                       // one per team, in dynamic shared memory
    double row[32][kTeamRow];                             // (32 rows: the idle bookkeeper lanes have a scratch row of their own)
    double s6[kTeamRing][32], s12[kTeamRing][32];         // results of a step (loop warps -> bookkeeper)
    uint32_t what[kTeamRing][32];                         //   nm | kTeamAccepted | kTeamWall
    uint32_t w1[kTeamRing][32], w2[kTeamRing][32];        //   the step's random words (written for volume trials only)
    double maxStep[32], half_l[32], invT[32], T[32];      // what a trial needs of the chain's scalars (bookkeeper -> loop warps)
    uint32_t tail[kTeamWarps];                            // loop warp w has published steps 1 .. tail[w] of the chunk
    uint32_t booked;                                      // the bookkeeper has accounted steps 1 .. booked
    uint32_t next_event;                                  // first step > booked on which ECheck / an adjustment / relaxVolume is due
    uint32_t count;                                       // steps of the chunk; 0 = no more work
    uint32_t sn0_lo, sn0_hi;                              // step number before the chunk's first step
    uint32_t cid0;                                        // Philox chain id of the tile's first chain
    uint32_t nvalid;                                      // chains of the tile that exist (the rest re-run the last one, unsaved)
};

__device__ __forceinline__ void team_bar_sync(int id, int n) { asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }   // (non-aligned: reached from different places by the two roles)
__device__ __forceinline__ void team_st_release(uint32_t *p, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.b32 [%0], %1;" ::"r"((uint32_t) __cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t team_ld_acquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(v) : "r"((uint32_t) __cvta_generic_to_shared(p)) : "memory");
    return v;
}
// every lane polls the same word (one broadcast load per try); a wait that cannot end is a protocol bug: trap, do not hang
__device__ __forceinline__ void team_wait_ge(const uint32_t *p, uint32_t v) {
    uint32_t spins = 0;
    while (team_ld_acquire(p) < v) {
        if (++spins > (1u << 25)) __trap();
    }
}

template <int POT>
__device__ __forceinline__ void team_loop_warp(TeamShared &T, const ChainsDev &S, int w, int go, int done) {
    constexpr int G = kTeamG, NPL = kTeamNPL;
    const uint32_t lane32 = threadIdx.x & 31, g = lane32 / G, j = lane32 % G;
    const uint32_t chain = 4 * w + g;
    Coop<POT, G> c;                                       // (only lane and cutoff are used by lanes_row_sums)
    c.lane = j;
    c.cutoff = S.cutoff;
    const uint32_t N = (uint32_t) S.N;
    const uint32_t ntt = (uint32_t) S.numTrialTypes, scale = 0xffffffffu / ntt;
    const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32);
    double *row = T.row[chain];
    for (;;) {
        team_bar_sync(go, kTeamWarps * 32);
        const uint32_t count = T.count;
        if (count == 0) return;
        const uint64_t sn0 = ((uint64_t) T.sn0_hi << 32) | T.sn0_lo;
        const uint32_t cid = T.cid0 + min(chain, T.nvalid - 1);
        const double invT = T.invT[chain], temp = T.T[chain];
        double maxStep = T.maxStep[chain], half_l = T.half_l[chain];
        uint32_t next_event = T.next_event;

        uint32_t my_nm = 0, my_w1 = 0, my_w2 = 0;         // this lane's share of the Philox batch (lanes.cuh)
        uint32_t batch_pos = G;
        auto draw = [&](uint64_t step, uint32_t &nm_o, uint32_t &w1_o, uint32_t &w2_o) {
            if (batch_pos == G) {
                const uint64_t mine = step + j;
                const Philox4 b = philox4x32_10((uint32_t) mine, (uint32_t)(mine >> 32), cid, kTagTrial, k0, k1);
                uint32_t k = b.w[0] / scale;              // gsl_rng_uniform_int rule, see Rng<kRngPhilox>
                if (k >= ntt) { k = b.w[3] / scale; if (k >= ntt) k = mulhi32(b.w[3], ntt); }
                my_nm = k; my_w1 = b.w[1]; my_w2 = b.w[2];
                batch_pos = 0;
            }
            nm_o = __shfl_sync(0xffffffffu, my_nm, batch_pos, G);
            w1_o = __shfl_sync(0xffffffffu, my_w1, batch_pos, G);
            w2_o = __shfl_sync(0xffffffffu, my_w2, batch_pos, G);
            ++batch_pos;
        };
        // lane 0 of the group takes the particle of the coming trial out of the row (far-away sentinel); all get its position
        auto take = [&](bool disp, uint32_t nm) -> double {
            double got = 0.0;
            if (j == 0 && disp) { got = row[nm]; row[nm] = kFarAway; }
            got = __shfl_sync(0xffffffffu, got, 0, G);
            __syncwarp();
            return got;
        };

        uint32_t nm, w1, w2;
        draw(sn0 + 1, nm, w1, w2);
        bool disp = nm < N;
        double rnm = take(disp, nm);
        for (uint32_t s = 1; s <= count; ++s) {
            const bool more = s < count;
            uint32_t nm1 = 0, w11 = 0, w21 = 0;
            if (more) draw(sn0 + s + 1, nm1, w11, w21);
            if (s > kTeamRing) team_wait_ge(&T.booked, s - kTeamRing);          // the ring slot of this step is free
            // the displacement, converged over the warp as in lanes.cuh (a group on a volume trial computes a discarded dummy)
            const double rT = rnm + u01_shifted(w1, 1.5) * 2 * maxStep;         // (rn - 0.5) * 2 * maxStep, :1182
            const bool wall = fabs(rT) > half_l;                                  // :1188
            double rr[NPL];
#pragma unroll
            for (int i = 0; i < NPL; ++i) rr[i] = row[j + G * i];
            double s6, s12;
            lanes_row_sums<POT, G, NPL>(c, rr, disp ? nm : 0u, rnm, rT, s6, s12);
            lanes_butterfly<G>(j, s6, s12);
            const double dE = 4 * s12 - 4 * s6;
            // Metropolis rule :1367-1377 through the band of metropolis_accept(), without early-out branches
            const double ran = u01_shifted(w2, 1.0);
            const double ea = (double) exp_neg_approx(dE * invT);
            const bool down = dE <= 0;
            const bool acc_b = ran < ea - kMetropolisBand, rej_b = ran > ea + kMetropolisBand;
            bool acc = down | acc_b;
            if (disp && !wall && !(down | acc_b | rej_b)) acc = metropolis_exact(dE, temp, ran);
            acc = acc && disp && !wall;
            if (j == 0) {
                const uint32_t slot = s % kTeamRing;
                if (disp) row[nm] = acc ? rT : rnm;                                // the sentinel goes, the particle is back
                T.s6[slot][chain] = s6; T.s12[slot][chain] = s12;
                T.what[slot][chain] = nm | (acc ? kTeamAccepted : 0u) | ((disp && wall) ? kTeamWall : 0u);
                if (!disp) { T.w1[slot][chain] = w1; T.w2[slot][chain] = w2; }
            }
            __syncwarp();
            if (lane32 == 0) team_st_release(&T.tail[w], s);
            // a step the bookkeeper has to finish (volume trial of one of the four chains, or an interval step)
            if (__any_sync(0xffffffffu, !disp) || s == next_event) {
                team_wait_ge(&T.booked, s);
                maxStep = T.maxStep[chain]; half_l = T.half_l[chain];
                next_event = T.next_event;
            }
            nm = nm1; w1 = w11; w2 = w21;
            disp = more && nm < N;
            if (more) rnm = take(disp, nm);
        }
        team_bar_sync(done, kTeamWarps * 32);
    }
}

// The bookkeeper warp: one chain per lane.  `own` = this lane has a chain (a ragged last tile leaves lanes idle; they
// take part in the warp-collective position scaling only).
template <int POT, bool LOG>
__device__ __forceinline__ void team_bookkeeper(TeamShared &T, const ChainsDev &S, const StepArgs &a, uint32_t chunk, uint32_t ntiles,
                                                uint32_t nchunks, unsigned int *work, unsigned int *progress, int go, int done) {
    const uint32_t lane = threadIdx.x & 31;
    const bool scaling_volume = (POT == kPotLJ) && S.nbn < 0;
    const bool relax_on = a.adapt_device && S.relax > 0 && S.ensemble == kEnsNPT;
    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(work, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= ntiles * nchunks) break;
        const uint32_t tile = item % ntiles, k = item / ntiles;
        if (lane == 0 && k > 0) {
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(progress + tile) : "memory");
                if (seen < k) __nanosleep(200);
            } while (seen < k);
        }
        __syncwarp();
        const uint64_t c0 = (uint64_t) tile * kTeamChains;
        const uint32_t nvalid = (uint32_t) min((uint64_t) kTeamChains, S.nchains - c0);
        const bool own = lane < nvalid;
        const uint64_t c = c0 + min(lane, nvalid - 1);       // (idle lanes shadow the tile's last chain, unsaved)
        const uint32_t s0 = k * chunk, count = min(chunk, (uint32_t) a.nsteps - s0);
        uint64_t sn = a.sn0 + s0;

        Chain<POT> ch;
        double *row = T.row[lane];
        load_chain<POT, true>(ch, S, c, row, 1);
        for (uint32_t i = ch.N; i < kTeamG * kTeamNPL; ++i) row[i] = kFarAway;           // pads of the unrolled partner loop
        double rho = (double) ch.N / ch.l;
        Rng<kRngPhilox> rng;                                  // (volume trials take their acceptance number from b.w[2])
        rng.k0 = (uint32_t) S.seed; rng.k1 = (uint32_t)(S.seed >> 32); rng.chain = (uint32_t)(S.chain_id0 + c);
        auto until_event = [&]() -> uint32_t {
            uint64_t left = 0xffffffffull;
            auto upd = [&](uint64_t every) { if (every) left = min(left, every - sn % every); };
            upd(a.eci);
            if (a.adapt_device) { upd(a.mdai); upd(a.mvai); }
            if (relax_on && sn < 1000000ull) upd(10000);
            return (uint32_t) left;
        };
        uint32_t ev_left = until_event();
        T.maxStep[lane] = ch.maxStep; T.half_l[lane] = ch.l / 2.0; T.invT[lane] = ch.invT; T.T[lane] = ch.T;
        if (lane < (uint32_t) kTeamWarps) T.tail[lane] = 0;
        if (lane == 0) {
            T.booked = 0; T.next_event = ev_left; T.count = count;
            T.sn0_lo = (uint32_t) sn; T.sn0_hi = (uint32_t)(sn >> 32);
            T.cid0 = (uint32_t)(S.chain_id0 + c0); T.nvalid = nvalid;
        }
        __syncwarp();
        team_bar_sync(go, kTeamWarps * 32);

        uint32_t n_acc = 0, n_rej = 0;
        for (uint32_t s = 1; s <= count; ++s) {
            {                                                 // until every loop warp has published step s
                uint32_t spins = 0;
                for (;;) {
                    const uint32_t t = lane < (uint32_t) kTeamLoopWarps ? team_ld_acquire(&T.tail[lane]) : 0xffffffffu;
                    if (__all_sync(0xffffffffu, t >= s)) break;
                    if (++spins > (1u << 25)) __trap();
                    if (JMM_TEAM_NS > 0) __nanosleep(JMM_TEAM_NS);
                }
            }
            ++sn;
            const uint32_t slot = s % kTeamRing;
            const uint32_t what = lane < (uint32_t) kTeamChains ? T.what[slot][lane] : 0u;   // (idle lanes: a rejected move of particle 0)
            const uint32_t nm = what & 0xffffu;
            const bool disp = nm < ch.N, acc = (what & kTeamAccepted) != 0;
            uint8_t flags = 0;
            if (LOG) flags = (what & kTeamWall) ? kLogWall : (acc ? kLogAccepted : 0);
            if (acc) {                                        // the nine totals, as lanes.cuh
                const double dE12 = 4 * T.s12[slot][lane], dE6 = 4 * T.s6[slot][lane];
                const double dV12 = 12 * dE12, dV6 = 6 * dE6, dH12 = 144 * dE12, dH6 = 36 * dE6;
                ch.tot[0] += dE12 - dE6;  ch.tot[2] += dE12; ch.tot[4] += dE6;
                ch.tot[1] += dV12 - dV6; ch.tot[3] += dV12; ch.tot[5] += dV6;
                ch.tot[6] += dH12 - dH6; ch.tot[7] += dH12; ch.tot[8] += dH6;
            }
            n_acc += acc ? 1u : 0u;
            n_rej += (disp && !acc) ? 1u : 0u;
            // volume trials (qavLJ :1648-1730 / fav :2161-2293): in the lane; r *= s by the whole warp (as prod.cuh).
            // The loop warp of such a chain is waiting for `booked`; its row has no sentinel out.
            const bool any_volume = __any_sync(0xffffffffu, !disp);      // (uniform: the scaling below is warp-collective)
            if (any_volume) {
                double vscale = 0.0;
                if (own && !disp) {
                    rng.b.w[2] = T.w2[slot][lane];
                    const double rn = u01(T.w1[slot][lane]);
                    if constexpr (POT == kPotLJ) {
                        flags = scaling_volume ? volume_trial_scaling<POT, false>(ch, rn, rng, &vscale) : volume_trial_full<POT, false>(ch, rn, rng, &vscale);
                    } else flags = volume_trial_full<POT, false>(ch, rn, rng, &vscale);
                }
                __syncwarp();
                unsigned pend = __ballot_sync(0xffffffffu, vscale != 0.0);
                while (pend) {
                    const int src = __ffs(pend) - 1;
                    pend &= pend - 1;
                    const double f = __shfl_sync(0xffffffffu, vscale, src);
                    double *col = T.row[src];
                    for (uint32_t i = lane; i < ch.N; i += 32) col[i] = col[i] * f;
                }
                if (vscale != 0.0) { T.half_l[lane] = ch.l / 2.0; rho = (double) ch.N / ch.l; }
            }
            const bool event = --ev_left == 0;
            if (event) {
                ch.cnt[0] += n_acc; ch.cnt[1] += n_rej; n_acc = n_rej = 0;
                if (own && a.eci && sn % a.eci == 0) energy_check<POT, false>(ch);           // Step :1800
            }
            {                                                 // updateThermo :1941-1961 with the cached N/l, Step :1805
                const double E = ch.tot[0], Vir = ch.tot[1], HV = ch.tot[6];
                ch.acc[0] = ch.acc[0] + rho;        ch.acc[1] = ch.acc[1] + rho * rho;
                ch.acc[2] = ch.acc[2] + ch.l;       ch.acc[3] = ch.acc[3] + ch.l * ch.l;
                ch.acc[4] = ch.acc[4] + E;          ch.acc[5] = ch.acc[5] + E * E;
                ch.acc[6] = ch.acc[6] + ch.l * E;   ch.acc[7] = ch.acc[7] + Vir;
                ch.acc[8] = ch.acc[8] + Vir * Vir;  ch.acc[9] = ch.acc[9] + E * Vir;
                ch.acc[10] = ch.acc[10] + HV;       ch.acc[11] = ch.acc[11] + HV * HV;
            }
            if (event) {
                if (own && a.adapt_device) {                                                 // src/Main.cpp:145-176
                    if (a.mdai && sn % a.mdai == 0) adjust_max_step(ch, a.log_ideal);
                    if (a.mvai && sn % a.mvai == 0) adjust_max_dl(ch, a.log_ideal);
                    if (relax_on && sn % 10000 == 0 && sn < 1000000ull) relax_volume<POT, false>(ch);
                }
                rho = (double) ch.N / ch.l;
                T.maxStep[lane] = ch.maxStep; T.half_l[lane] = ch.l / 2.0;
                ev_left = until_event();
                if (lane == 0) T.next_event = s + ev_left;
            }
            if (LOG && own) a.accept_log[(uint64_t)(s0 + s - 1) * S.nchains + c] = flags;
            __syncwarp();
            if (lane == 0) team_st_release(&T.booked, s);
        }
        team_bar_sync(done, kTeamWarps * 32);
        ch.cnt[0] += n_acc; ch.cnt[1] += n_rej;
        if (own) store_chain(ch, S, c, true);
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(progress + tile), "r"(k + 1) : "memory");
    }
    if (lane == 0) T.count = 0;                             // no more work: let the loop warps go
    __syncwarp();
    team_bar_sync(go, kTeamWarps * 32);
}

// grid = co-resident CTAs (persistent); a CTA holds NT teams of eight warps; work items = (chunk, tile of 28 chains)
template <int POT, int NT, bool LOG>
__global__ void __launch_bounds__(NT * kTeamWarps * 32, 1) k_chains_step_team(ChainsDev S, StepArgs a, uint32_t chunk, uint32_t ntiles,
                                                                               uint32_t nchunks, unsigned int *work, unsigned int *progress) {
    extern __shared__ __align__(16) unsigned char team_smem[];
    const int warp = threadIdx.x >> 5;
    const int team = warp / kTeamWarps, w = warp % kTeamWarps;
    TeamShared &T = reinterpret_cast<TeamShared *>(team_smem)[team];
    const int go = 1 + 2 * team, done = 2 + 2 * team;     // named barriers 1 ... 2 NT (0 is __syncthreads)
    if (w < kTeamLoopWarps) team_loop_warp<POT>(T, S, w, go, done);
    else team_bookkeeper<POT, LOG>(T, S, a, chunk, ntiles, nchunks, work, progress, go, done);
}

}  // namespace jmm
