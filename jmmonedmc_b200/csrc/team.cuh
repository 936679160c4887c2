// k_chains_step_team — the moderate-chain-count regime (8192 ... 16 384 chains of N <= 80 per GPU: a sweep sharded over
// 4-8 GPUs) with the bookkeeping of a step POOLED over 32 chains.
//
// lanes.cuh gives a chain G = 8 lanes and runs four chains per warp; the partner loop then costs what prod.cuh's costs
// per trial, but everything else of a step (Philox, trial set-up, Metropolis, nine totals, twelve sums, hand-over:
// ~250 of 444 warp instructions) is executed once per FOUR chains instead of once per 28-32, and that is what holds the
// kernel at 44 % of the fp64 pipe.  Here warps are specialised:
//   * a TEAM = seven LOOP warps + one BOOKKEEPER warp serves 28 chains;
//   * loop warp w owns chains 4w ... 4w+3 in lanes.cuh's layout (8 lanes per chain, 10 partners per lane, stage-wise
//     evaluation, butterfly) and does nothing but: wait for the trial descriptors, read the row, sum, write (s6, s12);
//   * the bookkeeper holds ONE CHAIN PER LANE (scalars in registers, as prod.cuh): Philox, trial type, md, wall test,
//     sentinel, Metropolis, commit, totals, sums, counters, volume trials, ECheck / adjustments / relaxVolume — once
//     per 32 chains, with chains.cuh's per-thread functions on the shared rows;
//   * two named barriers per team and step: "go" (bookkeeper arrives, loop warps wait: descriptors and rows are
//     ready) and "done" (loop warps arrive, bookkeeper waits: the sums are ready).  bar.arrive / bar.sync order the
//     shared-memory traffic (PTX ISA, producer-consumer use of named barriers);
//   * whatever the next partner loop does not need (totals, thermo sums, counters, the Philox block of the trial after
//     next) is done by the bookkeeper AFTER it has released the loop warps, i.e. concurrently with their loop.
// Arithmetic of a trial = lanes.cuh's (same row sums, same butterfly): decisions and positions are bit-identical to
// the oracle, totals to <= 1e-12.  Rows, descriptors and results live in shared memory; a row is only written
// between "done" and "go", while the loop warps are parked.
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

struct TeamShared {                                       // one per team, in dynamic shared memory
    double row[32][kTeamRow];                             // (32 rows: the idle bookkeeper lanes have a scratch row of their own)
    double rnm[32], rT[32];                               // trial descriptors (bookkeeper -> loop warps)
    double s6[32], s12[32];                               // partner sums (loop warps -> bookkeeper)
    uint32_t nm[32];
    int cmd;                                              // 1 = a trial is published, 0 = leave
    int pad_;
};

__device__ __forceinline__ void team_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void team_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int POT>
__device__ __forceinline__ void team_loop_warp(TeamShared &T, int w, int go, int done) {
    constexpr int G = kTeamG, NPL = kTeamNPL;
    const uint32_t lane32 = threadIdx.x & 31, g = lane32 / G, j = lane32 % G;
    const uint32_t chain = 4 * w + g;
    Coop<POT, G> c;                                       // (only lane, cutoff are used by lanes_row_sums)
    c.lane = j;
    const double *row = T.row[chain];
    for (;;) {
        team_bar_sync(go, kTeamWarps * 32);
        if (T.cmd == 0) return;
        const double rnm = T.rnm[chain], rT = T.rT[chain];
        const uint32_t nm = T.nm[chain];
        double rr[NPL];
#pragma unroll
        for (int i = 0; i < NPL; ++i) rr[i] = row[j + G * i];
        double s6, s12;
        lanes_row_sums<POT, G, NPL>(c, rr, nm, rnm, rT, s6, s12);
        lanes_butterfly<G>(j, s6, s12);
        if (j == 0) { T.s6[chain] = s6; T.s12[chain] = s12; }
        __syncwarp();
        team_bar_arrive(done, kTeamWarps * 32);
    }
}

// The bookkeeper warp: one chain per lane.  `own` = this lane has a chain (a ragged last tile leaves lanes idle; they
// take part in the warp-collective position scaling only).
template <int POT, bool LOG>
__device__ __forceinline__ void team_bookkeeper(TeamShared &T, const ChainsDev &S, const StepArgs &a, uint32_t chunk, uint32_t ntiles,
                                                uint32_t nchunks, unsigned int *work, unsigned int *progress, int go, int done) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ntt = (uint32_t) S.numTrialTypes, scale = 0xffffffffu / ntt;
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
        const uint64_t c = (uint64_t) tile * kTeamChains + lane;
        const bool own = lane < (uint32_t) kTeamChains && c < S.nchains;
        const uint32_t s0 = k * chunk, count = min(chunk, (uint32_t) a.nsteps - s0);
        uint64_t sn = a.sn0 + s0;

        Chain<POT> ch;
        double *row = T.row[lane];
        load_chain<POT, true>(ch, S, own ? c : S.nchains - 1, row, 1);
        for (uint32_t i = ch.N; i < kTeamG * kTeamNPL; ++i) row[i] = kFarAway;           // pads of the unrolled partner loop
        double half_l = ch.l / 2.0, rho = (double) ch.N / ch.l;
        Rng<kRngPhilox> rng;
        rng.k0 = (uint32_t) S.seed; rng.k1 = (uint32_t)(S.seed >> 32); rng.chain = (uint32_t)(S.chain_id0 + (own ? c : S.nchains - 1));
        auto until_event = [&]() -> uint32_t {
            uint64_t left = 0xffffffffull;
            auto upd = [&](uint64_t every) { if (every) left = min(left, every - sn % every); };
            upd(a.eci);
            if (a.adapt_device) { upd(a.mdai); upd(a.mvai); }
            if (relax_on && sn < 1000000ull) upd(10000);
            return (uint32_t) left;
        };
        uint32_t ev_left = until_event();

        // trial of step sn + 1
        rng.begin(sn + 1);
        uint32_t nm = rng.trial_type(ntt, scale), w1 = rng.b.w[1], w2 = rng.b.w[2];
        double rnm = 0.0, rT = 0.0;
        bool live = false;                                 // a displacement inside the walls: the loop warps' sums count
        auto publish = [&]() {                              // descriptor + sentinel of the current (nm, w1)
            const bool disp = nm < ch.N;
            rnm = row[disp ? nm : 0];
            rT = rnm + u01_shifted(w1, 1.5) * 2 * ch.maxStep;                          // (rn - 0.5) * 2 * maxStep, :1182
            live = disp && !(fabs(rT) > half_l);                                         // :1188
            T.rnm[lane] = rnm; T.rT[lane] = rT; T.nm[lane] = disp ? nm : 0u;
            if (disp) row[nm] = kFarAway;
            if (lane == 0) T.cmd = 1;
            __syncwarp();
            team_bar_arrive(go, kTeamWarps * 32);
        };
        publish();

        uint32_t n_acc = 0, n_rej = 0;
        for (uint32_t s = 0; s < count; ++s) {
            ++sn;
            // while the loop warps sum: the Philox block of the NEXT trial
            const bool more = s + 1 < count;
            uint32_t nm1 = 0, w11 = 0, w21 = 0;
            if (more) { rng.begin(sn + 1); nm1 = rng.trial_type(ntt, scale); w11 = rng.b.w[1]; w21 = rng.b.w[2]; }
            team_bar_sync(done, kTeamWarps * 32);           // the sums of this step are there; the loop warps are parked
            const bool disp = nm < ch.N;
            uint8_t flags = 0;
            bool acc = false;
            double s6 = 0.0, s12 = 0.0;
            if (disp) {                                     // qad2 :1160-1464
                s6 = T.s6[lane]; s12 = T.s12[lane];
                const double dE = 4 * s12 - 4 * s6;
                const double ran = u01_shifted(w2, 1.0);
                const double ea = (double) exp_neg_approx(dE * ch.invT);
                const bool down = dE <= 0;
                const bool acc_b = ran < ea - kMetropolisBand, rej_b = ran > ea + kMetropolisBand;
                acc = down | acc_b;
                if (live && !(down | acc_b | rej_b)) acc = metropolis_exact(dE, ch.T, ran);
                acc = acc && live;
                row[nm] = acc ? rT : rnm;                   // the sentinel goes, the particle is back (moved or not)
                if (LOG) flags = !live ? kLogWall : (acc ? kLogAccepted : 0);
            }
            // volume trials (qavLJ :1648-1730 / fav :2161-2293): in the lane; r *= s by the whole warp (as prod.cuh)
            double vscale = 0.0;
            if (own && !disp) {
                rng.b.w[2] = w2;                            // (the block of THIS step: rng now holds the next one)
                const double rn = u01(w1);
                if constexpr (POT == kPotLJ) {
                    flags = scaling_volume ? volume_trial_scaling<POT, false>(ch, rn, rng, &vscale) : volume_trial_full<POT, false>(ch, rn, rng, &vscale);
                } else flags = volume_trial_full<POT, false>(ch, rn, rng, &vscale);
            }
            {
                __syncwarp();
                unsigned pend = __ballot_sync(0xffffffffu, vscale != 0.0);
                while (pend) {
                    const int src = __ffs(pend) - 1;
                    pend &= pend - 1;
                    const double f = __shfl_sync(0xffffffffu, vscale, src);
                    double *col = T.row[src];
                    for (uint32_t i = lane; i < ch.N; i += 32) col[i] = col[i] * f;
                }
                __syncwarp();
                if (vscale != 0.0) { half_l = ch.l / 2.0; rho = (double) ch.N / ch.l; }
            }
            const bool event = --ev_left == 0;
            auto totals_and_thermo = [&]() {
                if (acc) {
                    const double dE12 = 4 * s12, dE6 = 4 * s6;
                    const double dV12 = 12 * dE12, dV6 = 6 * dE6, dH12 = 144 * dE12, dH6 = 36 * dE6;
                    ch.tot[0] += dE12 - dE6;  ch.tot[2] += dE12; ch.tot[4] += dE6;
                    ch.tot[1] += dV12 - dV6; ch.tot[3] += dV12; ch.tot[5] += dV6;
                    ch.tot[6] += dH12 - dH6; ch.tot[7] += dH12; ch.tot[8] += dH6;
                }
            };
            auto thermo = [&]() {                           // updateThermo :1941-1961 with the cached N/l
                const double E = ch.tot[0], Vir = ch.tot[1], HV = ch.tot[6];
                ch.acc[0] = ch.acc[0] + rho;        ch.acc[1] = ch.acc[1] + rho * rho;
                ch.acc[2] = ch.acc[2] + ch.l;       ch.acc[3] = ch.acc[3] + ch.l * ch.l;
                ch.acc[4] = ch.acc[4] + E;          ch.acc[5] = ch.acc[5] + E * E;
                ch.acc[6] = ch.acc[6] + ch.l * E;   ch.acc[7] = ch.acc[7] + Vir;
                ch.acc[8] = ch.acc[8] + Vir * Vir;  ch.acc[9] = ch.acc[9] + E * Vir;
                ch.acc[10] = ch.acc[10] + HV;       ch.acc[11] = ch.acc[11] + HV * HV;
            };
            if (event) {                                    // everything in the reference's order, before the next trial is set up
                totals_and_thermo();
                ch.cnt[0] += n_acc + (acc ? 1u : 0u); ch.cnt[1] += n_rej + ((disp && !acc) ? 1u : 0u); n_acc = n_rej = 0;
                if (own) {
                    if (a.eci && sn % a.eci == 0) energy_check<POT, false>(ch);            // Step :1800
                    thermo();                                                                // :1805
                    if (a.adapt_device) {                                                    // src/Main.cpp:145-176
                        if (a.mdai && sn % a.mdai == 0) adjust_max_step(ch, a.log_ideal);
                        if (a.mvai && sn % a.mvai == 0) adjust_max_dl(ch, a.log_ideal);
                        if (relax_on && sn % 10000 == 0 && sn < 1000000ull) relax_volume<POT, false>(ch);
                    }
                }
                half_l = ch.l / 2.0; rho = (double) ch.N / ch.l;
                ev_left = until_event();
            }
            // the next trial: descriptor, sentinel, release the loop warps
            const uint32_t nm_done = nm;
            nm = nm1; w1 = w11; w2 = w21;
            if (more) publish();
            // ... and, concurrently with their loop, what that loop does not need
            if (!event) {
                totals_and_thermo();
                if (own) thermo();
                n_acc += acc ? 1u : 0u;
                n_rej += (nm_done < ch.N && !acc) ? 1u : 0u;
            }
            if (LOG && own) a.accept_log[(uint64_t)(s0 + s) * S.nchains + c] = flags;
        }
        ch.cnt[0] += n_acc; ch.cnt[1] += n_rej;
        if (own) store_chain(ch, S, c, true);
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(progress + tile), "r"(k + 1) : "memory");
    }
    if (lane == 0) T.cmd = 0;                               // no more work: let the loop warps go
    __syncwarp();
    team_bar_arrive(go, kTeamWarps * 32);
}

// grid = co-resident CTAs (persistent); a CTA holds NT teams of nine warps; work items = (chunk, tile of 32 chains)
template <int POT, int NT, bool LOG>
__global__ void __launch_bounds__(NT * kTeamWarps * 32, 1) k_chains_step_team(ChainsDev S, StepArgs a, uint32_t chunk, uint32_t ntiles,
                                                                               uint32_t nchunks, unsigned int *work, unsigned int *progress) {
    extern __shared__ __align__(16) unsigned char team_smem[];
    const int warp = threadIdx.x >> 5;
    const int team = warp / kTeamWarps, w = warp % kTeamWarps;
    TeamShared &T = reinterpret_cast<TeamShared *>(team_smem)[team];
    const int go = 1 + 2 * team, done = 2 + 2 * team;     // named barriers 1 ... 2 NT (0 is __syncthreads)
    if (w < kTeamLoopWarps) team_loop_warp<POT>(T, w, go, done);
    else team_bookkeeper<POT, LOG>(T, S, a, chunk, ntiles, nchunks, work, progress, go, done);
}

}  // namespace jmm
