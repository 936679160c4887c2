// k_chains_step_solo — nearest-neighbour bond chains (HARMONIC, NBN 1: the INPUTstd shape, BASELINE config 2) with ONE
// CHAIN PER THREAD and ONE WARP PER SM.
//
// k_chains_step_bond (bond.cuh) gives a chain 16 lanes: 2048 warps for 4096 chains, but almost everything a step does
// is a scalar of the chain, so 346 warp instructions serve TWO chain-steps and a warp sharing its sub-partition with
// 3.3 others issues one of them every 5.3 cycles: 1830 cycles per step, whatever the lanes do.  What bounds a launch
// of 4096 serial Markov chains is the LATENCY of one step, and the cheapest step is the one in which the 32 lanes of a
// warp do 32 different chains' work with one instruction stream:
//   * one chain per lane, positions in a [N][32] shared-memory tile (lane = column: conflict-free at any nm);
//   * no divergence: displacement trial, volume trial (fav) and energy check are straight-line code with selects,
//     every lane evaluates all of it every step (a warp of 32 chains has a volume trial at 95 % of the steps anyway)
//     and commits under a predicate; only the 2e-5 events (a Metropolis or volume decision inside the approximation
//     band, an energy discrepancy) branch, warp-uniformly;
//   * a warp has a sub-partition (a whole SM) to itself, so the step time is the dependent-issue latency of the
//     stream divided by the instruction-level parallelism in it — four independent bond terms in the trial, nine in
//     fav, nine in the check, ten sums, and the Philox block of the NEXT step, which depends on nothing;
//   * every sum is taken by the thread in the reference's order (pair index order, left and right partner apart,
//     (acc - old) + new): there is no cross-lane arithmetic at all, so the result is bit-identical to the oracle by
//     construction, like chains.cuh's.
// Reference: Step :1758-1811, qad2 :1160-1464 (NBN 1), fav :2161-2293, ECheck :1965-2095, updateThermo :1941-1961,
// maxDisAdjust / maxDVAdjust :2100-2139 (src/jmmMCState.cpp), phiHarmoniccut src/pot.cpp:110-134.
#pragma once
#include "bond.cuh"

namespace jmm {

// NT > 0: N is a compile-time constant (loops unrolled, scaled positions in registers); NT == 0: any N
template <int NT, bool LOG, bool INF>
__global__ void __launch_bounds__(32) k_chains_step_solo(ChainsDev S, StepArgs a) {
    extern __shared__ double solo_smem[];                 // [N][32]
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x;
    const uint64_t C = S.nchains;
    const uint64_t c_raw = (uint64_t) blockIdx.x * 32 + lane;
    const bool own = c_raw < C;
    const uint64_t chain = own ? c_raw : C - 1;           // (lanes past the last chain shadow it, unsaved)
    const uint32_t N = NT ? (uint32_t) NT : (uint32_t) S.N;
    constexpr int NR = NT ? NT : 1;
    double *r = solo_smem + lane;                         // particle i: r[i * 32]

    double l = S.l[chain], maxStep = S.maxStep[chain], maxdl = S.maxdl[chain];
    const double P = S.P[chain], T = S.T[chain], invT = 1.0 / T, cutoff = S.cutoff;
    double half_l = l / 2.0, rho = (double) N / l, two_over_l = 2 / l;
    double E = S.tot[chain], Vir = S.tot[C + chain];
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = S.acc[k * C + chain];
    uint64_t cnt[kNCnt];
#pragma unroll
    for (int k = 0; k < kNCnt; ++k) cnt[k] = S.cnt[k * C + chain];
    uint64_t vAErr = S.vAErr[chain], echecks = S.echeck[chain], discrepancies = S.echeck[C + chain];
    uint32_t t_dacc = 0, t_drej = 0, t_vacc = 0, t_vrej = 0, t_checks = 0;   // 32-bit tallies, folded at adjustments and at the end
#pragma unroll
    for (uint32_t i = 0; i < N; ++i) r[i * 32] = S.r[(uint64_t) i * C + chain];

    const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32), cid = (uint32_t)(S.chain_id0 + chain);
    const uint32_t ntt = (uint32_t) S.numTrialTypes, scale = 0xffffffffu / ntt;
    uint64_t sn = a.sn0;
    auto until = [&](uint64_t every) -> uint32_t {
        if (!every) return 0xffffffffu;
        const uint64_t left = every - sn % every;
        return left > 0xfffffffeull ? 0xffffffffu : (uint32_t) left;
    };
    uint32_t eci_left = until(a.eci);
    uint32_t mdai_left = a.adapt_device ? until(a.mdai) : 0xffffffffu;
    uint32_t mvai_left = a.adapt_device ? until(a.mvai) : 0xffffffffu;
    const uint32_t eci32 = a.eci > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.eci;
    const uint32_t mdai32 = a.mdai > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.mdai;
    const uint32_t mvai32 = a.mvai > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.mvai;
    uint32_t adapt_span = min(mdai_left, mvai_left), adapt_left = adapt_span;

    auto draw = [&](uint64_t step, uint32_t &nm, uint32_t &w1, uint32_t &w2) {
        const Philox4 b = philox4x32_10((uint32_t) step, (uint32_t)(step >> 32), cid, kTagTrial, k0, k1);
        uint32_t k = b.w[0] / scale;                      // gsl_rng_uniform_int rule, see Rng<kRngPhilox>
        if (k >= ntt) k = b2_redraw(b.w[3], scale, ntt);  // probability ~ ntt / 2^32
        nm = k; w1 = b.w[1]; w2 = b.w[2];
    };
    auto fold = [&]() {
        cnt[0] += t_dacc; cnt[1] += t_drej; cnt[2] += t_vacc; cnt[3] += t_vrej; echecks += t_checks;
        t_dacc = t_drej = t_vacc = t_vrej = t_checks = 0;
    };
    // totals of the current positions in pair order (fad :907-946, the ECheck reset :2028-2071)
    auto recompute = [&]() {
        double e = 0, v = 0;
        for (uint32_t i = 0; i + 1 < N; ++i) {
            double pe, pv;
            b2_phi<INF>(r[(i + 1) * 32] - r[i * 32], cutoff, two_over_l, pe, pv);
            e += pe; v += pv;
        }
        E = e; Vir = v;
    };

    // ETest of ECheck :1965-2095: the energy again from the positions, in pair order
    auto etest = [&]() -> double {
        double et = 0;
        if constexpr (NT > 0) {
            double q[NR];
#pragma unroll
            for (int i = 0; i < NT; ++i) q[i] = r[i * 32];
#pragma unroll
            for (int i = 0; i + 1 < NT; ++i) et += b2_bond_energy<INF>(q[i + 1] - q[i], cutoff);
        } else {
            double qi = r[0];
            for (uint32_t i = 0; i + 1 < N; ++i) { const double qj = r[(i + 1) * 32]; et += b2_bond_energy<INF>(qj - qi, cutoff); qi = qj; }
        }
        return et;
    };
    // updateThermo :1941-1961 (HARMONIC defines no hypervirial: sums 10 and 11 stay +0)
    auto thermo = [&]() {
        acc[0] = acc[0] + rho;     acc[1] = acc[1] + rho * rho;
        acc[2] = acc[2] + l;       acc[3] = acc[3] + l * l;
        acc[4] = acc[4] + E;       acc[5] = acc[5] + E * E;
        acc[6] = acc[6] + l * E;   acc[7] = acc[7] + Vir;
        acc[8] = acc[8] + Vir * Vir; acc[9] = acc[9] + E * Vir;
    };

    // The loop is software-pipelined by one step: iteration t first evaluates, side by side and from the SAME state (the
    // one step t-1 left), the energy check of step t-1, the displacement trial of step t and the volume trial of step t —
    // three independent instruction streams in one basic block, plus the Philox block of step t+1 — then takes ONE vote on
    // everything rare (a discrepancy, a decision inside an approximation band), then adds step t-1's sample to the sums
    // and commits step t.  Order of effects per chain = the reference's: trial, ECheck, updateThermo, adjustments.
    uint32_t nm, w1, w2;
    draw(sn + 1, nm, w1, w2);
    const uint32_t nsteps = (uint32_t) a.nsteps;
    bool pending = false;                                 // step t-1 is committed; its sample (and `check`: its ECheck) is still due
    bool check = false;
    for (uint32_t s = 0; s < nsteps; ++s) {
        ++sn;                                                                         // incrementStep :1745
        uint32_t nm1 = 0, w11 = 0, w21 = 0;
        if (s + 1 < nsteps) draw(sn + 1, nm1, w11, w21);                              // (depends on nothing: fills the stalls below)
        const bool disp = nm < N;
        const double rnh = u01_shifted(w1, 1.5);                                      // rn - 0.5, exactly (rng.cuh)
        const double ran = u01_shifted(w2, 1.0);

        // ---- ECheck of step t-1
        const double et = check ? etest() : E;
        const bool bad = fabs(et - E) > 0.0001;

        // ---- displacement trial, qad2 :1160-1464 with NBN 1; a missing neighbour (chain end) contributes an exact 0
        const uint32_t i0 = disp ? nm : 0u;
        const bool hasL = i0 > 0, hasR = i0 + 1 < N;
        const double rnm = r[i0 * 32], rl = r[(hasL ? i0 - 1 : i0) * 32], rr = r[(hasR ? i0 + 1 : i0) * 32];
        const double rT = rnm + rnh * 2 * maxStep;                                    // :1182-1183
        const bool wall = fabs(rT) > half_l;                                          // :1188
        double po0, po1, pn0, pn1, qo0, qo1, qn0, qn1;
        b2_phi<INF>(rnm - rl, cutoff, two_over_l, po0, po1);
        b2_phi<INF>(rT - rl, cutoff, two_over_l, pn0, pn1);
        b2_phi<INF>(rr - rnm, cutoff, two_over_l, qo0, qo1);
        b2_phi<INF>(rr - rT, cutoff, two_over_l, qn0, qn1);
        const double l0 = hasL ? (0.0 - po0 + pn0) : 0.0, l1 = hasL ? (0.0 - po1 + pn1) : 0.0;   // :1244
        const double r0 = hasR ? (0.0 - qo0 + qn0) : 0.0, r1 = hasR ? (0.0 - qo1 + qn1) : 0.0;   // :1339
        const double dE = l0 + r0, dV = l1 + r1;                                      // :1354
        // Metropolis :1367-1377 by the approximation band of metropolis_accept()
        const double ea = (double) exp_neg_approx(dE * invT);
        const bool down = dE <= 0;
        const bool acc_b = ran < ea - kMetropolisBand, rej_b = ran > ea + kMetropolisBand;
        bool accept_d = down | acc_b;
        const bool undecided = disp && !wall && !(down | acc_b | rej_b);

        // ---- volume trial, fav :2161-2293: every pair term again on r * lRat1 (evaluated by every lane, committed by those it is for)
        const bool npt = ntt > N;
        const double dl = rnh * 2 * maxdl;
        const double lnew = l + dl;
        const double lRat1 = lnew / l;
        const double two_over_lnew = 2 / lnew;
        const double rho_new = (double) N / lnew;
        double rs[NR];
        double t0 = 0, t1 = 0;
        bool accept_v = false, v_open = false;
        if (npt) {
            if constexpr (NT > 0) {
#pragma unroll
                for (int i = 0; i < NT; ++i) rs[i] = r[i * 32] * lRat1;
#pragma unroll
                for (int i = 0; i + 1 < NT; ++i) {
                    double pe, pv;
                    b2_phi<INF>(rs[i + 1] - rs[i], cutoff, two_over_lnew, pe, pv);
                    t0 += pe; t1 += pv;
                }
            } else {
                double ri = r[0] * lRat1;
                for (uint32_t i = 0; i + 1 < N; ++i) {
                    const double rj = r[(i + 1) * 32] * lRat1;
                    double pe, pv;
                    b2_phi<INF>(rj - ri, cutoff, two_over_lnew, pe, pv);
                    t0 += pe; t1 += pv;
                    ri = rj;
                }
            }
            // volume_accept (pot.cuh) with its exact test under the common vote
            const double x = t0 - E + P * dl;                                         // :2249
            float lg;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float) lRat1));
            const double A = (double) N * ((double) lg * 0.6931471805599453) - x * invT;
            const double b = (double) exp_neg_approx(-A);
            const double band = volume_accept_band((double) N, (double) lg, A);
            const bool narrow = lRat1 > kVolumeBandLo && lRat1 < kVolumeBandHi;
            const bool v_yes = narrow && ran < b * (1.0 - band), v_no = narrow && ran > b * (1.0 + band);
            accept_v = v_yes;
            v_open = !disp && !(v_yes | v_no);
        }

        // ---- everything rare, one vote
        if (__any_sync(FULL, bad | undecided | v_open)) {
            if (bad) {                                                                // :2003-2071: the totals again from the positions
                discrepancies++;
                recompute();
                if (npt && !disp) accept_v = volume_accept_exact(t0 - E + P * dl, T, (double) N, lRat1, ran);
            } else if (v_open) accept_v = volume_accept_exact(t0 - E + P * dl, T, (double) N, lRat1, ran);
            if (undecided) accept_d = metropolis_exact(dE, T, ran);
        }
        // ---- the sample of step t-1 (updateThermo comes after ECheck, Step :1800-1805)
        if (pending) thermo();
        t_checks += check ? 1u : 0u;

        // ---- commit of step t
        const bool ok_d = disp && !wall && accept_d;                                  // :1384-1394
        const bool ok_v = npt && !disp && accept_v;                                   // :2257-2275
        if (ok_d) { r[nm * 32] = rT; E = E + dE; Vir = Vir + dV; }
        if (ok_v) {
            l = lnew; half_l = lnew / 2.0; two_over_l = two_over_lnew; rho = rho_new;
            E = t0; Vir = t1;
            if constexpr (NT > 0) {
#pragma unroll
                for (int i = 0; i < NT; ++i) r[i * 32] = rs[i];
            } else {
                for (uint32_t i = 0; i < N; ++i) r[i * 32] = r[i * 32] * lRat1;
            }
        }
        t_dacc += ok_d ? 1u : 0u;
        t_drej += (disp && !ok_d) ? 1u : 0u;
        t_vacc += ok_v ? 1u : 0u;
        t_vrej += (!disp && !ok_v) ? 1u : 0u;
        pending = true;
        check = eci32 == 1 || --eci_left == 0;
        if (check) eci_left = eci32;

        if (LOG && own) {
            const uint8_t f = disp ? (wall ? kLogWall : (ok_d ? kLogAccepted : 0)) : (uint8_t)(kLogVolume | (ok_v ? kLogAccepted : 0));
            a.accept_log[(uint64_t) s * C + chain] = f;
        }
        if (--adapt_left == 0) {                          // maxDisAdjust / maxDVAdjust steps (src/Main.cpp:145-165)
            mdai_left -= adapt_span; mvai_left -= adapt_span;
            const bool dis = mdai_left == 0, vol = mvai_left == 0;
            fold();
            if (dis) {                                                                // maxDisAdjust :2100-2115
                const double actualRatio = (double) cnt[0] / (double)(cnt[0] + cnt[1]);
                maxStep = maxStep * a.log_ideal / log(0.672924 * (actualRatio + 0.0644284));
                if (maxStep < 0.002) maxStep = 0.002;
                else if (maxStep > 0.5) maxStep = 0.5;
            }
            if (vol && (cnt[2] + cnt[3] - vAErr) > 0) {                               // maxDVAdjust :2120-2139
                vAErr = cnt[2] + cnt[3];
                const double actualRatio = (double) cnt[2] / (double)(cnt[2] + cnt[3]);
                maxdl = maxdl * a.log_ideal / log(0.672924 * (actualRatio + 0.0644284));
                if (maxdl < 0.002 * (double) N) maxdl = 0.002 * (double) N;
                else if (maxdl > 0.10 * (double) N) maxdl = 0.50 * (double) N;
            }
            if (dis) mdai_left = mdai32;
            if (vol) mvai_left = mvai32;
            adapt_span = adapt_left = min(mdai_left, mvai_left);
        }
        nm = nm1; w1 = w11; w2 = w21;
    }
    if (pending) {                                        // the last step's check and sample
        if (check) {
            ++t_checks;
            if (fabs(etest() - E) > 0.0001) { discrepancies++; recompute(); }
        }
        thermo();
    }
    fold();
    if (!own) return;
    for (uint32_t i = 0; i < N; ++i) S.r[(uint64_t) i * C + chain] = r[i * 32];
    S.l[chain] = l; S.maxStep[chain] = maxStep; S.maxdl[chain] = maxdl;
    S.tot[chain] = E; S.tot[C + chain] = Vir;
#pragma unroll
    for (int k = 2; k < kNTot; ++k) S.tot[k * C + chain] = 0.0;
#pragma unroll
    for (int k = 0; k < 10; ++k) S.acc[k * C + chain] = acc[k];
#pragma unroll
    for (int k = 0; k < kNCnt; ++k) S.cnt[k * C + chain] = cnt[k];
    S.vAErr[chain] = vAErr; S.echeck[chain] = echecks; S.echeck[C + chain] = discrepancies;
}


// ------------------------------------------------------------------------------------------------------------------
// k_chains_step_trio — the same chains, the step's instruction stream split over THREE warps of a CTA (one per
// sub-partition of the SM the CTA has to itself).  k_chains_step_solo issues ~560 instructions per step from one
// warp at 0.35 per cycle (profiles/r2w_c2solo_pipelined.txt: "wait" on dependent instructions); what shortens a step
// further is taking whatever the NEXT trial does not depend on out of that stream:
//   * warp P (producer) evaluates the Philox blocks and trial types of the 32 chains, 32 steps at a time, into a
//     double-buffered ring in shared memory: they depend on (seed, chain, step) only;
//   * warp T (trials) runs the Markov chains: displacement trial, volume trial, decisions, commit, E and l, counters,
//     step-size adjustments — ENERGIES ONLY: nothing T decides depends on a virial.  Per step it leaves a record
//     (what changed, the new value, E) in a second ring;
//   * warp V (verifier) replays the records on its own copy of the positions and does what only LOOKS at a step's
//     result: the virial (the four old/new bond terms of an accepted move in qad2's order :1244,1339,1354; all bonds
//     in pair order after an accepted volume change, fav :2212-2240), ECheck's energy from the positions and its
//     comparison with E, and updateThermo's sums.
// The rings are handed over 32 steps at a time through named barriers (st.shared, fence, barrier.arrive | barrier.sync,
// ld.shared: the producer/consumer use of barrier.arrive in the PTX ISA), so no warp polls.
// An energy discrepancy (ECheck :2003-2071: |ETest - E| > 1e-4, which the reference answers by recomputing the totals
// on the spot) would change what warp T has long passed.  It cannot be repaired in place, so it is repaired in
// time: the CTA stores NOTHING, raises its word in `redo`, and the launch is followed by k_chains_step_bond for the
// chains of the CTAs that raised it (jmm_gpu: launch_step_bond) — from the state the launch started with, through the
// kernel that takes the reset exactly where the reference takes it.  (E drifts by ~1e-16 per step against a
// threshold of 1e-4: the path exists for correctness and is exercised by JMM_SOLO_FORCE_REDO=1 in the tests.)
// Arithmetic = k_chains_step_solo's = the reference's, per chain in the reference's order: bit-identical.
constexpr int kTrioChunk = 32;                            // steps per ring buffer
// The warps of a CTA meet at named barriers from DIFFERENT places in the code (producer and consumer loops).  That is what
// barrier.sync / barrier.arrive WITHOUT .aligned are for (PTX ISA: .aligned = every thread of the warp executes the same barrier
// instruction, and tools such as compute-sanitizer synccheck hold the whole CTA to it); bar.sync is the .aligned form.
#ifndef JMM_BAR_SYNC
#define JMM_BAR_SYNC "barrier.sync"
#define JMM_BAR_ARRIVE "barrier.arrive"
#endif

__device__ __forceinline__ void trio_bar_sync(int id) { asm volatile(JMM_BAR_SYNC " %0, 64;" ::"r"(id) : "memory"); }
// the CTA-wide rendezvous at the end of the launch is reached from three places (one per role): a NAMED barrier with the thread
// count, which is defined per barrier resource, not __syncthreads(), which the programming model wants at one place
__device__ __forceinline__ void trio_bar_all() { asm volatile(JMM_BAR_SYNC " 9, 96;" ::: "memory"); }
__device__ __forceinline__ void trio_bar_arrive(int id) { __threadfence_block(); asm volatile(JMM_BAR_ARRIVE " %0, 64;" ::"r"(id) : "memory"); }

struct TrioRings {                                        // [buffer][step][lane]
    uint32_t nm[2][kTrioChunk][32], w1[2][kTrioChunk][32], w2[2][kTrioChunk][32];     // P -> T
    uint32_t code[2][kTrioChunk][32];                                                   // T -> V: 0 nothing, 1 | nm << 8 moved, 2 rescaled
    double val[2][kTrioChunk][32], e[2][kTrioChunk][32], lnew[2][kTrioChunk][32];
    int redo;                                                                           // V -> T at the end of the launch
};

template <int NT, bool LOG, bool INF>
__global__ void __launch_bounds__(96) k_chains_step_trio(ChainsDev S, StepArgs a, unsigned int *redo, int force_redo) {
    extern __shared__ __align__(16) unsigned char trio_smem[];
    constexpr uint32_t FULL = 0xffffffffu;
    // named barriers: a ring buffer b is FULL (producer arrives, consumer syncs) or FREE (the other way round)
    constexpr int P_FULL = 1, P_FREE = 3, D_FULL = 5, D_FREE = 7;                       // + buffer index 0/1
    const uint32_t lane = threadIdx.x & 31, role = threadIdx.x >> 5;                   // 0 = T, 1 = P, 2 = V
    const uint64_t C = S.nchains;
    const uint64_t c_raw = (uint64_t) blockIdx.x * 32 + lane;
    const bool own = c_raw < C;
    const uint64_t chain = own ? c_raw : C - 1;           // (lanes past the last chain shadow it, unsaved)
    const uint32_t N = NT ? (uint32_t) NT : (uint32_t) S.N;
    constexpr int NR = NT ? NT : 1;
    TrioRings &R = *reinterpret_cast<TrioRings *>(trio_smem);
    double *tiles = reinterpret_cast<double *>(trio_smem + sizeof(TrioRings));
    const uint32_t nsteps = (uint32_t) a.nsteps;
    const uint32_t nchunks = (nsteps + kTrioChunk - 1) / kTrioChunk;
    const double cutoff = S.cutoff;

    if (role == 1) {
        // ---------------------------------------------------------------- P: trial types and random words
        const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32), cid = (uint32_t)(S.chain_id0 + chain);
        const uint32_t ntt = (uint32_t) S.numTrialTypes, scale = 0xffffffffu / ntt;
        for (uint32_t k = 0; k < nchunks; ++k) {
            const int b = k & 1;
            if (k >= 2) trio_bar_sync(P_FREE + b);
            const uint32_t n = min((uint32_t) kTrioChunk, nsteps - k * kTrioChunk);
            const uint64_t step0 = a.sn0 + (uint64_t) k * kTrioChunk + 1;
            for (uint32_t j = 0; j < n; ++j) {
                const uint64_t step = step0 + j;
                const Philox4 blk = philox4x32_10((uint32_t) step, (uint32_t)(step >> 32), cid, kTagTrial, k0, k1);
                uint32_t t = blk.w[0] / scale;            // gsl_rng_uniform_int rule, see Rng<kRngPhilox>
                if (t >= ntt) t = b2_redraw(blk.w[3], scale, ntt);
                R.nm[b][j][lane] = t; R.w1[b][j][lane] = blk.w[1]; R.w2[b][j][lane] = blk.w[2];
            }
            trio_bar_arrive(P_FULL + b);
        }
    } else if (role == 2) {
        // ---------------------------------------------------------------- V: ECheck and updateThermo on a replayed copy
        double *r = tiles + (size_t) N * 32 + lane;
        double l = S.l[chain], rho = (double) N / l, two_over_l = 2 / l;
        double E = S.tot[chain], Vir = S.tot[C + chain];
        double acc[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[k] = S.acc[k * C + chain];
        uint64_t echecks = S.echeck[chain];
        uint32_t t_checks = 0;
        for (uint32_t i = 0; i < N; ++i) r[i * 32] = S.r[(uint64_t) i * C + chain];
        uint64_t sn = a.sn0;
        const uint32_t eci32 = a.eci > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.eci;
        uint32_t eci_left = 0xffffffffu;
        if (a.eci) { const uint64_t left = a.eci - sn % a.eci; eci_left = left > 0xfffffffeull ? 0xffffffffu : (uint32_t) left; }
        bool bad = force_redo != 0;
        for (uint32_t k = 0; k < nchunks; ++k) {
            const int b = k & 1;
            trio_bar_sync(D_FULL + b);
            const uint32_t n = min((uint32_t) kTrioChunk, nsteps - k * kTrioChunk);
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t code = R.code[b][j][lane];
                const double val = R.val[b][j][lane];
                double pe_unused;
                {   // an accepted move: its virial change, qad2 :1244,1339,1354.  (Keeping every bond's term in shared memory and reading
                    // the two OLD ones from there instead of evaluating them was measured: 1.00e10 against 1.13e10 — the loads and
                    // their write-backs sit on the step's dependency chain, profiles/r2zd_*, r2ze_c2_crew_ab.jsonl.)
                    const bool moved = code & 1u;
                    const uint32_t i0 = moved ? (code >> 8) : 0u;
                    const bool hasL = i0 > 0, hasR = i0 + 1 < N;
                    const double rnm = r[i0 * 32], rl = r[(hasL ? i0 - 1 : i0) * 32], rr = r[(hasR ? i0 + 1 : i0) * 32];
                    double pe, po1, pn1, qo1, qn1;
                    b2_phi<INF>(rnm - rl, cutoff, two_over_l, pe, po1);
                    b2_phi<INF>(val - rl, cutoff, two_over_l, pe, pn1);
                    b2_phi<INF>(rr - rnm, cutoff, two_over_l, pe, qo1);
                    b2_phi<INF>(rr - val, cutoff, two_over_l, pe, qn1);
                    const double l1 = hasL ? (0.0 - po1 + pn1) : 0.0, r1 = hasR ? (0.0 - qo1 + qn1) : 0.0;
                    if (moved) { Vir = Vir + (l1 + r1); r[i0 * 32] = val; }
                }
                if (__any_sync(FULL, code == 2u)) {                                   // an accepted volume trial: r *= lRat1 (:2264-2266), Vir from all bonds
                    if (code == 2u) {
                        l = R.lnew[b][j][lane];
                        rho = (double) N / l;
                        two_over_l = 2 / l;
                        double v = 0;
                        double ri = r[0] * val;
                        r[0] = ri;
                        for (uint32_t i = 0; i + 1 < N; ++i) {
                            const double rj = r[(i + 1) * 32] * val;
                            r[(i + 1) * 32] = rj;
                            double pv;
                            b2_phi<INF>(rj - ri, cutoff, two_over_l, pe_unused, pv);
                            v += pv;
                            ri = rj;
                        }
                        Vir = v;
                    }
                }
                E = R.e[b][j][lane];
                if (eci32 == 1 || --eci_left == 0) {                                  // ECheck :1965-2095
                    double et = 0;
                    if constexpr (NT > 0) {
                        double q[NR];
#pragma unroll
                        for (int i = 0; i < NT; ++i) q[i] = r[i * 32];
#pragma unroll
                        for (int i = 0; i + 1 < NT; ++i) et += b2_bond_energy<INF>(q[i + 1] - q[i], cutoff);
                    } else {
                        double qi = r[0];
                        for (uint32_t i = 0; i + 1 < N; ++i) { const double qj = r[(i + 1) * 32]; et += b2_bond_energy<INF>(qj - qi, cutoff); qi = qj; }
                    }
                    ++t_checks;
                    bad = bad || fabs(et - E) > 0.0001;
                    eci_left = eci32;
                }
                acc[0] = acc[0] + rho;     acc[1] = acc[1] + rho * rho;             // updateThermo :1941-1961
                acc[2] = acc[2] + l;       acc[3] = acc[3] + l * l;
                acc[4] = acc[4] + E;       acc[5] = acc[5] + E * E;
                acc[6] = acc[6] + l * E;   acc[7] = acc[7] + Vir;
                acc[8] = acc[8] + Vir * Vir; acc[9] = acc[9] + E * Vir;
            }
            if (k + 2 < nchunks) trio_bar_arrive(D_FREE + b);
        }
        const bool any_bad = __any_sync(FULL, bad);
        if (lane == 0) { R.redo = any_bad ? 1 : 0; if (any_bad) redo[blockIdx.x] = 1u; }
        trio_bar_all();
        if (!any_bad && own) {
#pragma unroll
            for (int k = 0; k < 10; ++k) S.acc[k * C + chain] = acc[k];
            S.echeck[chain] = echecks + t_checks;
            S.tot[C + chain] = Vir;
        }
        return;
    } else {
        // ---------------------------------------------------------------- T: the Markov chains
        double *r = tiles + lane;
        double l = S.l[chain], maxStep = S.maxStep[chain], maxdl = S.maxdl[chain];
        const double P = S.P[chain], T = S.T[chain], invT = 1.0 / T;
        double half_l = l / 2.0;
        double E = S.tot[chain];
        uint64_t cnt[kNCnt];
#pragma unroll
        for (int k = 0; k < kNCnt; ++k) cnt[k] = S.cnt[k * C + chain];
        uint64_t vAErr = S.vAErr[chain];
        uint32_t t_dacc = 0, t_drej = 0, t_vacc = 0, t_vrej = 0;
        for (uint32_t i = 0; i < N; ++i) r[i * 32] = S.r[(uint64_t) i * C + chain];
        const uint32_t ntt = (uint32_t) S.numTrialTypes;
        const bool npt = ntt > N;
        uint64_t sn = a.sn0;
        auto until = [&](uint64_t every) -> uint32_t {
            if (!every) return 0xffffffffu;
            const uint64_t left = every - sn % every;
            return left > 0xfffffffeull ? 0xffffffffu : (uint32_t) left;
        };
        uint32_t mdai_left = a.adapt_device ? until(a.mdai) : 0xffffffffu;
        uint32_t mvai_left = a.adapt_device ? until(a.mvai) : 0xffffffffu;
        const uint32_t mdai32 = a.mdai > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.mdai;
        const uint32_t mvai32 = a.mvai > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.mvai;
        uint32_t adapt_span = min(mdai_left, mvai_left), adapt_left = adapt_span;
        auto fold = [&]() {
            cnt[0] += t_dacc; cnt[1] += t_drej; cnt[2] += t_vacc; cnt[3] += t_vrej;
            t_dacc = t_drej = t_vacc = t_vrej = 0;
        };

        for (uint32_t k = 0; k < nchunks; ++k) {
            const int b = k & 1;
            trio_bar_sync(P_FULL + b);
            if (k >= 2) trio_bar_sync(D_FREE + b);
            const uint32_t n = min((uint32_t) kTrioChunk, nsteps - k * kTrioChunk);
            for (uint32_t j = 0; j < n; ++j) {
                ++sn;                                                                 // incrementStep :1745
                const uint32_t nm = R.nm[b][j][lane];
                const bool disp = nm < N;
                const double rnh = u01_shifted(R.w1[b][j][lane], 1.5);                // rn - 0.5, exactly (rng.cuh)
                const double ran = u01_shifted(R.w2[b][j][lane], 1.0);

                // ---- displacement trial, qad2 :1160-1464 with NBN 1; a missing neighbour contributes an exact 0
                const uint32_t i0 = disp ? nm : 0u;
                const bool hasL = i0 > 0, hasR = i0 + 1 < N;
                const double rnm = r[i0 * 32], rl = r[(hasL ? i0 - 1 : i0) * 32], rr = r[(hasR ? i0 + 1 : i0) * 32];
                const double rT = rnm + rnh * 2 * maxStep;                            // :1182-1183
                const bool wall = fabs(rT) > half_l;                                  // :1188
                const double po0 = b2_bond_energy<INF>(rnm - rl, cutoff), pn0 = b2_bond_energy<INF>(rT - rl, cutoff);
                const double qo0 = b2_bond_energy<INF>(rr - rnm, cutoff), qn0 = b2_bond_energy<INF>(rr - rT, cutoff);
                const double l0 = hasL ? (0.0 - po0 + pn0) : 0.0;                     // :1244
                const double r0 = hasR ? (0.0 - qo0 + qn0) : 0.0;                     // :1339
                const double dE = l0 + r0;                                            // :1354
                const double ea = (double) exp_neg_approx(dE * invT);                 // Metropolis :1367-1377, band of metropolis_accept()
                const bool down = dE <= 0;
                const bool acc_b = ran < ea - kMetropolisBand, rej_b = ran > ea + kMetropolisBand;
                bool accept_d = down | acc_b;
                const bool undecided = disp && !wall && !(down | acc_b | rej_b);

                // ---- volume trial, fav :2161-2293: every pair term again on r * lRat1
                const double dl = rnh * 2 * maxdl;
                const double lnew = l + dl;
                const double lRat1 = lnew / l;
                double rs[NR];
                double t0 = 0;
                bool accept_v = false, v_open = false;
                if (npt) {
                    if constexpr (NT > 0) {
#pragma unroll
                        for (int i = 0; i < NT; ++i) rs[i] = r[i * 32] * lRat1;
#pragma unroll
                        for (int i = 0; i + 1 < NT; ++i) t0 += b2_bond_energy<INF>(rs[i + 1] - rs[i], cutoff);
                    } else {
                        double ri = r[0] * lRat1;
                        for (uint32_t i = 0; i + 1 < N; ++i) {
                            const double rj = r[(i + 1) * 32] * lRat1;
                            t0 += b2_bond_energy<INF>(rj - ri, cutoff);
                            ri = rj;
                        }
                    }
                    const double x = t0 - E + P * dl;                                 // :2249, volume_accept (pot.cuh)
                    float lg;
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float) lRat1));
                    const double A = (double) N * ((double) lg * 0.6931471805599453) - x * invT;
                    const double bb = (double) exp_neg_approx(-A);
                    const double band = volume_accept_band((double) N, (double) lg, A);
                    const bool narrow = lRat1 > kVolumeBandLo && lRat1 < kVolumeBandHi;
                    const bool v_yes = narrow && ran < bb * (1.0 - band), v_no = narrow && ran > bb * (1.0 + band);
                    accept_v = v_yes;
                    v_open = !disp && !(v_yes | v_no);
                }
                if (__any_sync(FULL, undecided | v_open)) {                          // inside an approximation band: the exact expressions
                    if (v_open) accept_v = volume_accept_exact(t0 - E + P * dl, T, (double) N, lRat1, ran);
                    if (undecided) accept_d = metropolis_exact(dE, T, ran);
                }

                // ---- commit, and the record for warp V
                const bool ok_d = disp && !wall && accept_d;                          // :1384-1394
                const bool ok_v = npt && !disp && accept_v;                           // :2257-2275
                if (ok_d) { r[nm * 32] = rT; E = E + dE; }
                if (ok_v) {
                    l = lnew; half_l = lnew / 2.0;
                    E = t0;
                    if constexpr (NT > 0) {
#pragma unroll
                        for (int i = 0; i < NT; ++i) r[i * 32] = rs[i];
                    } else {
                        for (uint32_t i = 0; i < N; ++i) r[i * 32] = r[i * 32] * lRat1;
                    }
                    R.lnew[b][j][lane] = lnew;
                }
                R.code[b][j][lane] = ok_d ? (1u | (nm << 8)) : (ok_v ? 2u : 0u);
                R.val[b][j][lane] = ok_d ? rT : lRat1;
                R.e[b][j][lane] = E;
                t_dacc += ok_d ? 1u : 0u;
                t_drej += (disp && !ok_d) ? 1u : 0u;
                t_vacc += ok_v ? 1u : 0u;
                t_vrej += (!disp && !ok_v) ? 1u : 0u;
                if (LOG && own) {
                    const uint8_t f = disp ? (wall ? kLogWall : (ok_d ? kLogAccepted : 0)) : (uint8_t)(kLogVolume | (ok_v ? kLogAccepted : 0));
                    a.accept_log[(uint64_t)(k * kTrioChunk + j) * C + chain] = f;
                }
                if (--adapt_left == 0) {                  // maxDisAdjust / maxDVAdjust steps (src/Main.cpp:145-165)
                    mdai_left -= adapt_span; mvai_left -= adapt_span;
                    const bool dis = mdai_left == 0, vol = mvai_left == 0;
                    fold();
                    if (dis) {                                                        // maxDisAdjust :2100-2115
                        const double actualRatio = (double) cnt[0] / (double)(cnt[0] + cnt[1]);
                        maxStep = maxStep * a.log_ideal / log(0.672924 * (actualRatio + 0.0644284));
                        if (maxStep < 0.002) maxStep = 0.002;
                        else if (maxStep > 0.5) maxStep = 0.5;
                    }
                    if (vol && (cnt[2] + cnt[3] - vAErr) > 0) {                       // maxDVAdjust :2120-2139
                        vAErr = cnt[2] + cnt[3];
                        const double actualRatio = (double) cnt[2] / (double)(cnt[2] + cnt[3]);
                        maxdl = maxdl * a.log_ideal / log(0.672924 * (actualRatio + 0.0644284));
                        if (maxdl < 0.002 * (double) N) maxdl = 0.002 * (double) N;
                        else if (maxdl > 0.10 * (double) N) maxdl = 0.50 * (double) N;
                    }
                    if (dis) mdai_left = mdai32;
                    if (vol) mvai_left = mvai32;
                    adapt_span = adapt_left = min(mdai_left, mvai_left);
                }
            }
            trio_bar_arrive(D_FULL + b);
            if (k + 2 < nchunks) trio_bar_arrive(P_FREE + b);
        }
        fold();
        trio_bar_all();                                  // warp V's verdict
        if (R.redo || !own) return;
        for (uint32_t i = 0; i < N; ++i) S.r[(uint64_t) i * C + chain] = r[i * 32];
        S.l[chain] = l; S.maxStep[chain] = maxStep; S.maxdl[chain] = maxdl;
        S.tot[chain] = E;                                 // (the virial is warp V's)
#pragma unroll
        for (int k = 2; k < kNTot; ++k) S.tot[k * C + chain] = 0.0;
#pragma unroll
        for (int k = 0; k < kNCnt; ++k) S.cnt[k * C + chain] = cnt[k];
        S.vAErr[chain] = vAErr;
        return;
    }
    trio_bar_all();                                      // (warp P: the CTA's final barrier)
}


// ------------------------------------------------------------------------------------------------------------------
// k_chains_step_crew — five warps per 32 chains.  In k_chains_step_trio the trial warp evaluates the displacement trial
// AND the volume trial of every chain at every step (32 chains: some lane has a volume trial at 95 % of the steps),
// ~270 instructions of which a chain uses one half or the other; and warp V's replay became as long as that.  Here
//   * warp T evaluates displacement trials only and commits them for the chains whose step is one,
//   * warp F evaluates volume trials only (fav: every bond on r * lRat1) and commits them for the chains whose step
//     is one; it owns maxdl, the volume counters and maxDVAdjust, as T owns maxStep, the displacement counters and
//     maxDisAdjust,
//   * T and F work on ONE set of positions, E and l in shared memory (a chain's column is touched by exactly one of
//     them in a step — the other reads a column of zeros instead) and meet at one named barrier per step,
//   * warp V replays the records for the virial after a volume change and for updateThermo's sums (the virial change of a
//     move comes from T, which holds the four distances and waits for F anyway), warp W replays them for ECheck,
//   * warp P produces the Philox words, as before.
// Chunked rings, repair of an energy discrepancy by repeating the launch, arithmetic: as k_chains_step_trio.
#ifndef JMM_CREW_ZEROCOL
#define JMM_CREW_ZEROCOL 1        // 1: T / F read a column of zeros for the other's chains; 0: predicated loads
#endif
#ifndef JMM_CREW_DV_IN_T
#define JMM_CREW_DV_IN_T 1        // 1: warp T, which holds the four distances of a move, also evaluates its virial change; 0: warp V does
#endif
#ifndef JMM_CREW_BANDFMA
#define JMM_CREW_BANDFMA 1        // 1: the volume band with folded constants; 0: volume_accept_band()
#endif
struct CrewShared {                                       // [buffer][step][lane]
    uint32_t nm[2][kTrioChunk][32], w1[2][kTrioChunk][32], w2[2][kTrioChunk][32];     // P -> T, F
    uint32_t code[2][kTrioChunk][32];                                                   // T, F -> V, W: 0 nothing, 1 | nm << 8 moved, 2 rescaled
    double val[2][kTrioChunk][32], e[2][kTrioChunk][32], lnew[2][kTrioChunk][32];
    double dv[2][kTrioChunk][32];                                                       // T -> V: the virial change of a move
    double E[32], l[32];                                                                // the chains' live energy and length (T and F)
    int redo;
};

__device__ __forceinline__ void crew_bar_sync(int id, int n) { asm volatile(JMM_BAR_SYNC " %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void crew_bar_arrive(int id, int n) { __threadfence_block(); asm volatile(JMM_BAR_ARRIVE " %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int NT, bool LOG, bool INF>
__global__ void __launch_bounds__(160, 1) k_chains_step_crew(ChainsDev S, StepArgs a, unsigned int *redo, int force_redo) {
    extern __shared__ __align__(16) unsigned char crew_smem[];
    constexpr uint32_t FULL = 0xffffffffu;
    // named barriers (+ buffer index 0/1): a ring buffer is FULL (producers arrive, consumers sync) or FREE (the reverse)
    constexpr int P_FULL = 1, P_FREE = 3, D_FULL = 5, D_FREE = 7, STEP = 9, FINAL = 10;   // (FINAL: reached from five places, hence named)
    constexpr int kP = 96, kD = 128;                      // threads at a P barrier (P, T, F) and at a D barrier (T, F, V, W)
    // warp w runs on sub-partition w % 4: the light producer shares one with the energy checker, T and F have their own
    enum { ROLE_P = 0, ROLE_T = 1, ROLE_F = 2, ROLE_V = 3, ROLE_W = 4 };
    const uint32_t lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const uint64_t C = S.nchains;
    const uint64_t c_raw = (uint64_t) blockIdx.x * 32 + lane;
    const bool own = c_raw < C;
    const uint64_t chain = own ? c_raw : C - 1;           // (lanes past the last chain shadow it, unsaved)
    const uint32_t N = NT ? (uint32_t) NT : (uint32_t) S.N;
    constexpr int NR = NT ? NT : 1;
    CrewShared &R = *reinterpret_cast<CrewShared *>(crew_smem);
    double *tiles = reinterpret_cast<double *>(crew_smem + sizeof(CrewShared));
    const uint32_t nsteps = (uint32_t) a.nsteps;
    const uint32_t nchunks = (nsteps + kTrioChunk - 1) / kTrioChunk;
    const double cutoff = S.cutoff;

    const double *zeros = tiles + (size_t) 3 * N * 32 + lane;   // a column nobody writes: what T reads of F's chains and F of T's
    if (role == ROLE_T) {                                 // the live state, before anyone reads it
        double *r = tiles + lane;
        for (uint32_t i = 0; i < N; ++i) { r[i * 32] = S.r[(uint64_t) i * C + chain]; tiles[(size_t) 3 * N * 32 + i * 32 + lane] = 0.0; }
        R.E[lane] = S.tot[chain]; R.l[lane] = S.l[chain];
    }
    __syncthreads();

    if (role == ROLE_P) {
        // ---------------------------------------------------------------- P: trial types and random words
        const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32), cid = (uint32_t)(S.chain_id0 + chain);
        const uint32_t ntt = (uint32_t) S.numTrialTypes, scale = 0xffffffffu / ntt;
        for (uint32_t k = 0; k < nchunks; ++k) {
            const int b = k & 1;
            if (k >= 2) crew_bar_sync(P_FREE + b, kP);
            const uint32_t n = min((uint32_t) kTrioChunk, nsteps - k * kTrioChunk);
            const uint64_t step0 = a.sn0 + (uint64_t) k * kTrioChunk + 1;
            for (uint32_t j = 0; j < n; ++j) {
                const uint64_t step = step0 + j;
                const Philox4 blk = philox4x32_10((uint32_t) step, (uint32_t)(step >> 32), cid, kTagTrial, k0, k1);
                uint32_t t = blk.w[0] / scale;            // gsl_rng_uniform_int rule, see Rng<kRngPhilox>
                if (t >= ntt) t = b2_redraw(blk.w[3], scale, ntt);
                R.nm[b][j][lane] = t; R.w1[b][j][lane] = blk.w[1]; R.w2[b][j][lane] = blk.w[2];
            }
            crew_bar_arrive(P_FULL + b, kP);
        }
    } else if (role == ROLE_V) {
        // ---------------------------------------------------------------- V: virial and updateThermo on a replayed copy
        double *r = tiles + (size_t) N * 32 + lane;
        double l = S.l[chain], rho = (double) N / l, two_over_l = 2 / l;
        double E = S.tot[chain], Vir = S.tot[C + chain];
        double acc[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[k] = S.acc[k * C + chain];
        for (uint32_t i = 0; i < N; ++i) r[i * 32] = S.r[(uint64_t) i * C + chain];
        for (uint32_t k = 0; k < nchunks; ++k) {
            const int b = k & 1;
            crew_bar_sync(D_FULL + b, kD);
            const uint32_t n = min((uint32_t) kTrioChunk, nsteps - k * kTrioChunk);
            uint32_t code = R.code[b][0][lane];
            double val = R.val[b][0][lane], e_rec = R.e[b][0][lane];
#if JMM_CREW_DV_IN_T
            double dv = R.dv[b][0][lane];
#endif
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t code1 = j + 1 < n ? R.code[b][j + 1][lane] : 0u;       // (the next record: its latency hides behind this one)
                const double val1 = j + 1 < n ? R.val[b][j + 1][lane] : 0.0, e1 = j + 1 < n ? R.e[b][j + 1][lane] : 0.0;
#if JMM_CREW_DV_IN_T
                const double dv1 = j + 1 < n ? R.dv[b][j + 1][lane] : 0.0;
                if (code & 1u) { Vir = Vir + dv; r[(code >> 8) * 32] = val; }         // an accepted move: qad2 :1354,1390 (the change is warp T's)
#else
                {   // an accepted move: its virial change, qad2 :1244,1339,1354.  (Keeping every bond's term in shared memory and reading
                    // the two OLD ones from there instead of evaluating them was measured: 1.00e10 against 1.13e10 — the loads and
                    // their write-backs sit on the step's dependency chain, profiles/r2zd_*, r2ze_c2_crew_ab.jsonl.)
                    const bool moved = code & 1u;
                    const uint32_t i0 = moved ? (code >> 8) : 0u;
                    const bool hasL = i0 > 0, hasR = i0 + 1 < N;
                    const double rnm = r[i0 * 32], rl = r[(hasL ? i0 - 1 : i0) * 32], rr = r[(hasR ? i0 + 1 : i0) * 32];
                    double pe, po1, pn1, qo1, qn1;
                    b2_phi<INF>(rnm - rl, cutoff, two_over_l, pe, po1);
                    b2_phi<INF>(val - rl, cutoff, two_over_l, pe, pn1);
                    b2_phi<INF>(rr - rnm, cutoff, two_over_l, pe, qo1);
                    b2_phi<INF>(rr - val, cutoff, two_over_l, pe, qn1);
                    const double l1 = hasL ? (0.0 - po1 + pn1) : 0.0, r1 = hasR ? (0.0 - qo1 + qn1) : 0.0;
                    if (moved) { Vir = Vir + (l1 + r1); r[i0 * 32] = val; }
                }
#endif
                if (__any_sync(FULL, code == 2u)) {                                   // an accepted volume trial: r *= lRat1 (:2264-2266), Vir from all bonds
                    if (code == 2u) {
                        l = R.lnew[b][j][lane];
                        rho = (double) N / l;
                        two_over_l = 2 / l;
                        double v = 0;
                        double ri = r[0] * val;
                        r[0] = ri;
                        for (uint32_t i = 0; i + 1 < N; ++i) {
                            const double rj = r[(i + 1) * 32] * val;
                            r[(i + 1) * 32] = rj;
                            double pe, pv;
                            b2_phi<INF>(rj - ri, cutoff, two_over_l, pe, pv);
                            v += pv;
                            ri = rj;
                        }
                        Vir = v;
                    }
                }
                E = e_rec;
                acc[0] = acc[0] + rho;     acc[1] = acc[1] + rho * rho;             // updateThermo :1941-1961
                acc[2] = acc[2] + l;       acc[3] = acc[3] + l * l;
                acc[4] = acc[4] + E;       acc[5] = acc[5] + E * E;
                acc[6] = acc[6] + l * E;   acc[7] = acc[7] + Vir;
                acc[8] = acc[8] + Vir * Vir; acc[9] = acc[9] + E * Vir;
                code = code1; val = val1; e_rec = e1;
#if JMM_CREW_DV_IN_T
                dv = dv1;
#endif
            }
            if (k + 2 < nchunks) crew_bar_arrive(D_FREE + b, kD);
        }
        crew_bar_sync(FINAL, 160);                                  // warp W's verdict
        if (!R.redo && own) {
#pragma unroll
            for (int k = 0; k < 10; ++k) S.acc[k * C + chain] = acc[k];
            S.tot[C + chain] = Vir;
        }
        return;
    } else if (role == ROLE_W) {
        // ---------------------------------------------------------------- W: ECheck :1965-2095 on a replayed copy
        double *r = tiles + (size_t) 2 * N * 32 + lane;
        for (uint32_t i = 0; i < N; ++i) r[i * 32] = S.r[(uint64_t) i * C + chain];
        const uint64_t echecks = S.echeck[chain];
        uint32_t t_checks = 0;
        const uint32_t eci32 = a.eci > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.eci;
        uint32_t eci_left = 0xffffffffu;
        if (a.eci) { const uint64_t left = a.eci - a.sn0 % a.eci; eci_left = left > 0xfffffffeull ? 0xffffffffu : (uint32_t) left; }
        bool bad = force_redo != 0;
        for (uint32_t k = 0; k < nchunks; ++k) {
            const int b = k & 1;
            crew_bar_sync(D_FULL + b, kD);
            const uint32_t n = min((uint32_t) kTrioChunk, nsteps - k * kTrioChunk);
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t code = R.code[b][j][lane];
                const double val = R.val[b][j][lane], E = R.e[b][j][lane];
                if (__any_sync(FULL, code == 2u)) {
                    if (code == 2u) for (uint32_t i = 0; i < N; ++i) r[i * 32] = r[i * 32] * val;
                }
                if (code & 1u) r[(code >> 8) * 32] = val;
                if (eci32 == 1 || --eci_left == 0) {
                    double et = 0;
                    if constexpr (NT > 0) {
                        double q[NR];
#pragma unroll
                        for (int i = 0; i < NT; ++i) q[i] = r[i * 32];
#pragma unroll
                        for (int i = 0; i + 1 < NT; ++i) et += b2_bond_energy<INF>(q[i + 1] - q[i], cutoff);
                    } else {
                        double qi = r[0];
                        for (uint32_t i = 0; i + 1 < N; ++i) { const double qj = r[(i + 1) * 32]; et += b2_bond_energy<INF>(qj - qi, cutoff); qi = qj; }
                    }
                    ++t_checks;
                    bad = bad || fabs(et - E) > 0.0001;
                    eci_left = eci32;
                }
            }
            if (k + 2 < nchunks) crew_bar_arrive(D_FREE + b, kD);
        }
        const bool any_bad = __any_sync(FULL, bad);
        if (lane == 0) { R.redo = any_bad ? 1 : 0; if (any_bad) redo[blockIdx.x] = 1u; }
        crew_bar_sync(FINAL, 160);
        if (!any_bad && own) S.echeck[chain] = echecks + t_checks;
        return;
    } else if (role == ROLE_T) {
        // ---------------------------------------------------------------- T: displacement trials, qad2 :1160-1464 with NBN 1
        double *r = tiles + lane;
        double maxStep = S.maxStep[chain];
        const double T = S.T[chain], invT = 1.0 / T;
        uint64_t cnt0 = S.cnt[chain], cnt1 = S.cnt[C + chain];
        uint32_t t_acc = 0, t_rej = 0;
        const uint32_t mdai32 = a.mdai > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.mdai;
        uint32_t mdai_left = 0xffffffffu;
        if (a.adapt_device && a.mdai) { const uint64_t left = a.mdai - a.sn0 % a.mdai; mdai_left = left > 0xfffffffeull ? 0xffffffffu : (uint32_t) left; }
        for (uint32_t k = 0; k < nchunks; ++k) {
            const int b = k & 1;
            crew_bar_sync(P_FULL + b, kP);
            if (k >= 2) crew_bar_sync(D_FREE + b, kD);
            const uint32_t n = min((uint32_t) kTrioChunk, nsteps - k * kTrioChunk);
            uint32_t nm = R.nm[b][0][lane], w1 = R.w1[b][0][lane], w2 = R.w2[b][0][lane];
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t nm1 = j + 1 < n ? R.nm[b][j + 1][lane] : 0u, w11 = j + 1 < n ? R.w1[b][j + 1][lane] : 0u,
                               w21 = j + 1 < n ? R.w2[b][j + 1][lane] : 0u;
                const bool disp = nm < N;
                const double rnh = u01_shifted(w1, 1.5);                              // rn - 0.5, exactly (rng.cuh)
                const double ran = u01_shifted(w2, 1.0);
                const uint32_t i0 = disp ? nm : 0u;
                const bool hasL = i0 > 0, hasR = i0 + 1 < N;
                // (a chain on a volume trial belongs to warp F in this step: T reads the column of zeros instead of its positions)
                const double E = disp ? R.E[lane] : 0.0, lT = disp ? R.l[lane] : 1.0, half_l = lT / 2.0;
#if JMM_CREW_ZEROCOL
                const double *rc = disp ? r : zeros;
                const double rnm = rc[i0 * 32], rl = rc[(hasL ? i0 - 1 : i0) * 32], rr = rc[(hasR ? i0 + 1 : i0) * 32];
#else
                const double rnm = disp ? r[i0 * 32] : 0.0, rl = disp ? r[(hasL ? i0 - 1 : i0) * 32] : 0.0, rr = disp ? r[(hasR ? i0 + 1 : i0) * 32] : 0.0;
#endif
                const double rT = rnm + rnh * 2 * maxStep;                            // :1182-1183
                const bool wall = fabs(rT) > half_l;                                  // :1188
#if JMM_CREW_DV_IN_T
                // energy AND virial of the four bond terms: nothing T decides depends on the virial, but T holds the distances,
                // has the time (it waits for F), and warp V would have to load and rebuild them (190 instructions per step there)
                const double two_over_l = 2 / lT;
                double po0, po1, pn0, pn1, qo0, qo1, qn0, qn1;
                b2_phi<INF>(rnm - rl, cutoff, two_over_l, po0, po1);
                b2_phi<INF>(rT - rl, cutoff, two_over_l, pn0, pn1);
                b2_phi<INF>(rr - rnm, cutoff, two_over_l, qo0, qo1);
                b2_phi<INF>(rr - rT, cutoff, two_over_l, qn0, qn1);
                const double l1 = hasL ? (0.0 - po1 + pn1) : 0.0, r1 = hasR ? (0.0 - qo1 + qn1) : 0.0;
                const double dV = l1 + r1;
#else
                const double po0 = b2_bond_energy<INF>(rnm - rl, cutoff), pn0 = b2_bond_energy<INF>(rT - rl, cutoff);
                const double qo0 = b2_bond_energy<INF>(rr - rnm, cutoff), qn0 = b2_bond_energy<INF>(rr - rT, cutoff);
#endif
                const double l0 = hasL ? (0.0 - po0 + pn0) : 0.0;                     // :1244
                const double r0 = hasR ? (0.0 - qo0 + qn0) : 0.0;                     // :1339
                const double dE = l0 + r0;                                            // :1354
                const double ea = (double) exp_neg_approx(dE * invT);                 // Metropolis :1367-1377, band of metropolis_accept()
                const bool down = dE <= 0;
                const bool acc_b = ran < ea - kMetropolisBand, rej_b = ran > ea + kMetropolisBand;
                bool accept_d = down | acc_b;
                const bool undecided = disp && !wall && !(down | acc_b | rej_b);
                if (__any_sync(FULL, undecided)) {
                    if (undecided) accept_d = metropolis_exact(dE, T, ran);
                }
                const bool ok_d = disp && !wall && accept_d;                          // :1384-1394
                const double Enew = ok_d ? E + dE : E;
                if (ok_d) { r[nm * 32] = rT; R.E[lane] = Enew; }
                if (disp) {
                    R.code[b][j][lane] = ok_d ? (1u | (nm << 8)) : 0u;
                    R.val[b][j][lane] = rT;
                    R.e[b][j][lane] = Enew;
#if JMM_CREW_DV_IN_T
                    R.dv[b][j][lane] = dV;
#endif
                    if (LOG && own) a.accept_log[(uint64_t)(k * kTrioChunk + j) * C + chain] = wall ? kLogWall : (ok_d ? kLogAccepted : 0);
                }
                t_acc += ok_d ? 1u : 0u;
                t_rej += (disp && !ok_d) ? 1u : 0u;
                if (--mdai_left == 0) {                                               // maxDisAdjust :2100-2115 (src/Main.cpp:145-155)
                    cnt0 += t_acc; cnt1 += t_rej; t_acc = t_rej = 0;
                    const double actualRatio = (double) cnt0 / (double)(cnt0 + cnt1);
                    maxStep = maxStep * a.log_ideal / log(0.672924 * (actualRatio + 0.0644284));
                    if (maxStep < 0.002) maxStep = 0.002;
                    else if (maxStep > 0.5) maxStep = 0.5;
                    mdai_left = mdai32;
                }
                nm = nm1; w1 = w11; w2 = w21;
                crew_bar_sync(STEP, 64);                                              // warp F has committed its chains' step
            }
            crew_bar_arrive(D_FULL + b, kD);
            if (k + 2 < nchunks) crew_bar_arrive(P_FREE + b, kP);
        }
        cnt0 += t_acc; cnt1 += t_rej;
        crew_bar_sync(FINAL, 160);                                  // warp W's verdict
        if (R.redo || !own) return;
        for (uint32_t i = 0; i < N; ++i) S.r[(uint64_t) i * C + chain] = r[i * 32];
        S.l[chain] = R.l[lane]; S.maxStep[chain] = maxStep;
        S.tot[chain] = R.E[lane];                         // (the virial is warp V's)
#pragma unroll
        for (int k = 2; k < kNTot; ++k) S.tot[k * C + chain] = 0.0;
        S.cnt[chain] = cnt0; S.cnt[C + chain] = cnt1;
        return;
    } else {
        // ---------------------------------------------------------------- F: volume trials, fav :2161-2293
        double *r = tiles + lane;
        double maxdl = S.maxdl[chain];
        double l = S.l[chain];                                // (F is the only one to change a chain's length: a register, mirrored in R.l for T)
        const double P = S.P[chain], T = S.T[chain], invT = 1.0 / T;
        const double band_0 = 1.0000001 * 4.0 * (1.8e-7 * (double) N + 2.4e-7), band_lg = 1.0000001 * 4.0 * 4.2e-8 * (double) N;
        uint64_t cnt2 = S.cnt[2 * C + chain], cnt3 = S.cnt[3 * C + chain], vAErr = S.vAErr[chain];
        uint32_t t_acc = 0, t_rej = 0;
        const uint32_t mvai32 = a.mvai > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.mvai;
        uint32_t mvai_left = 0xffffffffu;
        if (a.adapt_device && a.mvai) { const uint64_t left = a.mvai - a.sn0 % a.mvai; mvai_left = left > 0xfffffffeull ? 0xffffffffu : (uint32_t) left; }
        for (uint32_t k = 0; k < nchunks; ++k) {
            const int b = k & 1;
            crew_bar_sync(P_FULL + b, kP);
            if (k >= 2) crew_bar_sync(D_FREE + b, kD);
            const uint32_t n = min((uint32_t) kTrioChunk, nsteps - k * kTrioChunk);
            uint32_t nm = R.nm[b][0][lane], w1 = R.w1[b][0][lane], w2 = R.w2[b][0][lane];
            // the trial length and its ratio (:2170-2172) depend on nothing warp T does: those of step j+1 are evaluated BEFORE
            // the barrier that ends step j, so that the division's latency is spent while waiting for T
            double dl = u01_shifted(w1, 1.5) * 2 * maxdl;                             // (rn - 0.5) * 2 * maxdl
            double lnew = l + dl;
            double lRat1 = lnew / l;
            for (uint32_t j = 0; j < n; ++j) {
                const uint32_t nm1 = j + 1 < n ? R.nm[b][j + 1][lane] : 0u, w11 = j + 1 < n ? R.w1[b][j + 1][lane] : 0u,
                               w21 = j + 1 < n ? R.w2[b][j + 1][lane] : 0u;
                const bool vol = !(nm < N);
                const double ran = u01_shifted(w2, 1.0);
                // (a chain on a displacement trial belongs to warp T in this step: F reads the column of zeros instead of its positions)
                const double E = vol ? R.E[lane] : 0.0;
#if JMM_CREW_ZEROCOL
                const double *rc = vol ? r : zeros;
#define JMM_CREW_LD(i) rc[(i) * 32]
#else
#define JMM_CREW_LD(i) (vol ? r[(i) * 32] : 0.0)
#endif
                double rs[NR];
                double t0 = 0;
                if constexpr (NT > 0) {
#pragma unroll
                    for (int i = 0; i < NT; ++i) rs[i] = JMM_CREW_LD(i) * lRat1;
#pragma unroll
                    for (int i = 0; i + 1 < NT; ++i) t0 += b2_bond_energy<INF>(rs[i + 1] - rs[i], cutoff);
                } else {
                    double ri = JMM_CREW_LD(0) * lRat1;
                    for (uint32_t i = 0; i + 1 < N; ++i) {
                        const double rj = JMM_CREW_LD(i + 1) * lRat1;
                        t0 += b2_bond_energy<INF>(rj - ri, cutoff);
                        ri = rj;
                    }
                }
                const double x = t0 - E + P * dl;                                     // :2249, volume_accept (pot.cuh)
                float lg;
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float) lRat1));
                const double A = (double) N * ((double) lg * 0.6931471805599453) - x * invT;
                const double bb = (double) exp_neg_approx(-A);
                // volume_accept_band() with its constants folded (two FMAs: the band only has to be no SMALLER than the bound)
#if JMM_CREW_BANDFMA
                const double band = __fma_rn(6.4e-7, fabs(A), __fma_rn(band_lg, (double) fabsf(lg), band_0));
#else
                const double band = volume_accept_band((double) N, (double) lg, A);
#endif
                const bool narrow = lRat1 > kVolumeBandLo && lRat1 < kVolumeBandHi;
                const bool v_yes = narrow && ran < bb * (1.0 - band), v_no = narrow && ran > bb * (1.0 + band);
                bool accept_v = v_yes;
                const bool v_open = vol && !(v_yes | v_no);
                if (__any_sync(FULL, v_open)) {                                       // inside the approximation band: the exact expression
                    if (v_open) accept_v = volume_accept_exact(x, T, (double) N, lRat1, ran);
                }
                const bool ok_v = vol && accept_v;                                    // :2257-2275
                if (ok_v) {
                    l = lnew;
                    R.l[lane] = lnew; R.E[lane] = t0;
                    if constexpr (NT > 0) {
#pragma unroll
                        for (int i = 0; i < NT; ++i) r[i * 32] = rs[i];
                    } else {
                        for (uint32_t i = 0; i < N; ++i) r[i * 32] = r[i * 32] * lRat1;
                    }
                    R.lnew[b][j][lane] = lnew;
                }
                if (vol) {
                    R.code[b][j][lane] = ok_v ? 2u : 0u;
                    R.val[b][j][lane] = lRat1;
                    R.e[b][j][lane] = ok_v ? t0 : E;
                    if (LOG && own) a.accept_log[(uint64_t)(k * kTrioChunk + j) * C + chain] = (uint8_t)(kLogVolume | (ok_v ? kLogAccepted : 0));
                }
                t_acc += ok_v ? 1u : 0u;
                t_rej += (vol && !ok_v) ? 1u : 0u;
                if (--mvai_left == 0) {                                               // maxDVAdjust :2120-2139 (src/Main.cpp:156-165)
                    cnt2 += t_acc; cnt3 += t_rej; t_acc = t_rej = 0;
                    if ((cnt2 + cnt3 - vAErr) > 0) {
                        vAErr = cnt2 + cnt3;
                        const double actualRatio = (double) cnt2 / (double)(cnt2 + cnt3);
                        maxdl = maxdl * a.log_ideal / log(0.672924 * (actualRatio + 0.0644284));
                        if (maxdl < 0.002 * (double) N) maxdl = 0.002 * (double) N;
                        else if (maxdl > 0.10 * (double) N) maxdl = 0.50 * (double) N;
                    }
                    mvai_left = mvai32;
                }
                nm = nm1; w1 = w11; w2 = w21;
                dl = u01_shifted(w1, 1.5) * 2 * maxdl;
                lnew = l + dl;
                lRat1 = lnew / l;
                crew_bar_sync(STEP, 64);                                              // warp T has committed its chains' step
            }
            crew_bar_arrive(D_FULL + b, kD);
            if (k + 2 < nchunks) crew_bar_arrive(P_FREE + b, kP);
        }
        cnt2 += t_acc; cnt3 += t_rej;
        crew_bar_sync(FINAL, 160);                                  // warp W's verdict
        if (R.redo || !own) return;
        S.maxdl[chain] = maxdl;
        S.cnt[2 * C + chain] = cnt2; S.cnt[3 * C + chain] = cnt3;
        S.vAErr[chain] = vAErr;
        return;
    }
    crew_bar_sync(FINAL, 160);                                      // (warp P: the CTA's final barrier)
}

}  // namespace jmm
