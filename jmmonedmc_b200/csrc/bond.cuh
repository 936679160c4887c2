// k_chains_step_bond — the C2 shape (nearest-neighbour bond chains: HARMONIC, NBN == 1, N-1 <= 16), one chain
// per 16-lane group like coop.cuh, but with the displacement trial and the energy check written WITHOUT
// control flow: a missing neighbour at a chain end contributes an exact 0 through a select, the potential's
// d <= 0 / d < cutoff cases are selects, a wall hit or a rejection is a predicate on the commit.
//
// Why: the two chains that share a warp (lanes 0-15 and 16-31) used to diverge on every one of those branches
// and the warp executed both sides; with straight-line code both groups run one instruction stream.
// Measured on C2 (4096 chains, trial moves/s): coop.cuh 3.00e9 -> this kernel 3.74e9.
// (Also tried and dropped: TWO chains interleaved in the same 16 lanes for instruction-level parallelism —
// 228 registers, half the warps, 2.4e9: the dependent chains did interleave but the lost warp-level
// parallelism cost more.)
//
// Arithmetic, order of operations and results are those of coop.cuh (bit-identical to the oracle); the rare
// paths (fav, ECheck fallback) ARE coop.cuh's functions on the chain's Coop object.
#pragma once
#include "coop.cuh"

namespace jmm {

constexpr int kB2G = 16;                                  // lanes per chain
using B2Chain = Coop<kPotHarmonic, kB2G>;

// the gsl_rng_uniform_int redraw (see Rng<kRngPhilox>): out of line so the common path carries no second division
__device__ __noinline__ uint32_t b2_redraw(uint32_t w3, uint32_t scale, uint32_t ntt) {
    const uint32_t k = w3 / scale;
    return k < ntt ? k : mulhi32(w3, ntt);
}

struct B2Trial {
    double rT, dE, dV;
    uint32_t nm;
    bool wall, decided, accept;
};

// phiHarmoniccut (src/pot.cpp:110-134) with selects instead of branches: same values, no control flow, so the
// two chains' evaluations can be interleaved by the instruction scheduler
// INF: the deck gave no cut-off (POT HARMONIC without a third token, src/readInput.cpp:125-130), so
// "d < cutOff" holds for every finite d and its select is dropped.
template <bool INF>
__device__ __forceinline__ void b2_phi(double d, double cutoff, double two_over_l, double &e, double &v) {
    const double rijm = d - 1.0;
    const double e_in = rijm * rijm, v_in = two_over_l * d * rijm;
    const bool pos = d > 0;
    if constexpr (INF) {
        e = pos ? e_in : 10E10;
        v = pos ? v_in : 10E10;
    } else {
        const bool in = d < cutoff;
        e = pos ? (in ? e_in : 0.0) : 10E10;
        v = pos ? (in ? v_in : 0.0) : 10E10;
    }
}

// qad2 :1160-1464 for NBN == 1, branch-free: a missing neighbour (chain end) contributes an exact 0
template <bool INF>
__device__ __forceinline__ B2Trial b2_displacement(const B2Chain &c, uint32_t nm, double rn, double ran) {
    B2Trial t;
    t.nm = nm;
    const double md = (rn - 0.5) * 2 * c.maxStep;
    const double rnm = c.r[nm];
    t.rT = rnm + md;
    t.wall = fabs(t.rT) > c.half_l;
    const bool hasL = nm > 0, hasR = nm + 1 < c.N;
    const double rl = c.r[hasL ? nm - 1 : nm], rr = c.r[hasR ? nm + 1 : nm];
    double po[2], pn[2], qo[2], qn[2];
    b2_phi<INF>(rnm - rl, c.cutoff, c.two_over_l, po[0], po[1]);
    b2_phi<INF>(t.rT - rl, c.cutoff, c.two_over_l, pn[0], pn[1]);
    b2_phi<INF>(rr - rnm, c.cutoff, c.two_over_l, qo[0], qo[1]);
    b2_phi<INF>(rr - t.rT, c.cutoff, c.two_over_l, qn[0], qn[1]);
    const double l0 = hasL ? (0.0 - po[0] + pn[0]) : 0.0, l1 = hasL ? (0.0 - po[1] + pn[1]) : 0.0;   // :1244
    const double r0 = hasR ? (0.0 - qo[0] + qn[0]) : 0.0, r1 = hasR ? (0.0 - qo[1] + qn[1]) : 0.0;   // :1339
    t.dE = l0 + r0;                                                                                  // :1354
    t.dV = l1 + r1;
    // Metropolis by the approximation band of metropolis_accept(); `decided` false = evaluate exp exactly
    const double ea = (double) exp_neg_approx(t.dE * c.invT);
    const bool down = t.dE <= 0;
    const bool acc_b = ran < ea - kMetropolisBand, rej_b = ran > ea + kMetropolisBand;
    t.decided = down | acc_b | rej_b;
    t.accept = down | acc_b;
    return t;
}

__device__ __forceinline__ void b2_resolve(const B2Chain &c, B2Trial &t, double ran) {
    if (!t.decided && !t.wall) t.accept = metropolis_accept(t.dE, c.T, c.invT, ran);
}

__device__ __forceinline__ uint8_t b2_commit(B2Chain &c, const B2Trial &t) {
    const bool ok = !t.wall && t.accept;
    c.cnt[0] += ok ? 1 : 0;
    c.cnt[1] += ok ? 0 : 1;
    c.tot[0] = ok ? c.tot[0] + t.dE : c.tot[0];
    c.tot[1] = ok ? c.tot[1] + t.dV : c.tot[1];
    if (ok && c.lane == 0) c.r[t.nm] = t.rT;
    return t.wall ? kLogWall : (ok ? kLogAccepted : 0);
}

template <bool INF>
__device__ __forceinline__ double b2_etest_partial(const B2Chain &c) {
    const bool has = c.lane + 1 < c.N;
    const uint32_t i = has ? c.lane : 0;                       // idle lanes read pair (0,1) and discard it
    const double d = c.r[i + 1] - c.r[i];
    const double rijm = d - 1.0;
    const double e = (d > 0) ? ((INF || d < c.cutoff) ? rijm * rijm : 0.0) : 10E10;
    return has ? e : 0.0;
}

__device__ __forceinline__ void b2_load(B2Chain &c, const ChainsDev &S, uint64_t chain, double *row, uint32_t lane, uint32_t gmask) {
    const uint64_t C = S.nchains;
    c.lane = lane; c.gmask = gmask; c.r = row; c.sc = nullptr; c.chain_of_bonds = true; c.lean = false; c.consistent_virial = false;
    c.N = (uint32_t) S.N; c.nbn = S.nbn; c.cutoff = S.cutoff;
    c.P = S.P[chain]; c.T = S.T[chain]; c.maxStep = S.maxStep[chain]; c.maxdl = S.maxdl[chain];
    c.invT = 1.0 / c.T;
    c.set_l(S.l[chain]);
#pragma unroll
    for (int k = 0; k < 2; ++k) c.tot[k] = S.tot[k * C + chain];
#pragma unroll
    for (int k = 0; k < kNAcc; ++k) c.acc[k] = S.acc[k * C + chain];
#pragma unroll
    for (int k = 0; k < kNCnt; ++k) c.cnt[k] = S.cnt[k * C + chain];
    c.vAErr = S.vAErr[chain]; c.echecks = S.echeck[chain]; c.discrepancies = S.echeck[C + chain];
    for (uint32_t i = lane; i < c.N; i += kB2G) row[i] = S.r[(uint64_t) i * C + chain];
}

__device__ __forceinline__ void b2_store(const B2Chain &c, const ChainsDev &S, uint64_t chain, bool with_acc = true) {
    const uint64_t C = S.nchains;
    for (uint32_t i = c.lane; i < c.N; i += kB2G) S.r[(uint64_t) i * C + chain] = c.r[i];
    if (c.lane == 0) {
        S.l[chain] = c.l; S.maxStep[chain] = c.maxStep; S.maxdl[chain] = c.maxdl;
#pragma unroll
        for (int k = 0; k < 2; ++k) S.tot[k * C + chain] = c.tot[k];
#pragma unroll
        for (int k = 2; k < kNTot; ++k) S.tot[k * C + chain] = 0.0;
        if (with_acc) {
#pragma unroll
            for (int k = 0; k < kNAcc; ++k) S.acc[k * C + chain] = c.acc[k];
        }
#pragma unroll
        for (int k = 0; k < kNCnt; ++k) S.cnt[k * C + chain] = c.cnt[k];
        S.vAErr[chain] = c.vAErr; S.echeck[chain] = c.echecks; S.echeck[C + chain] = c.discrepancies;
    }
}

__device__ __forceinline__ void b2_adapt(B2Chain &c, const StepArgs &a, bool dis, bool vol) {
    if (dis) {                                                                    // maxDisAdjust :2100-2115
        const double actualRatio = (double) c.cnt[0] / (double)(c.cnt[0] + c.cnt[1]);
        c.maxStep = c.maxStep * a.log_ideal / log(0.672924 * (actualRatio + 0.0644284));
        if (c.maxStep < 0.002) c.maxStep = 0.002;
        else if (c.maxStep > 0.5) c.maxStep = 0.5;
    }
    if (vol && (c.cnt[2] + c.cnt[3] - c.vAErr) > 0) {                             // maxDVAdjust :2120-2139
        c.vAErr = c.cnt[2] + c.cnt[3];
        const double actualRatio = (double) c.cnt[2] / (double)(c.cnt[2] + c.cnt[3]);
        c.maxdl = c.maxdl * a.log_ideal / log(0.672924 * (actualRatio + 0.0644284));
        if (c.maxdl < 0.002 * (double) c.N) c.maxdl = 0.002 * (double) c.N;
        else if (c.maxdl > 0.10 * (double) c.N) c.maxdl = 0.50 * (double) c.N;
    }
}

template <bool LOG, bool INF>
__global__ void __launch_bounds__(128) k_chains_step_bond(ChainsDev S, StepArgs a, int npad, const unsigned int *redo = nullptr) {
    extern __shared__ double smem[];
    const uint32_t gib = threadIdx.x / kB2G, lane = threadIdx.x % kB2G;
    const uint64_t chainA = (uint64_t) blockIdx.x * (blockDim.x / kB2G) + gib;
    const uint64_t C = S.nchains;
    if (chainA >= C) return;
    // after k_chains_step_trio (solo.cuh): only the chains of the 32-chain CTAs that met an energy discrepancy and stored nothing
    if (redo && !redo[chainA >> 5]) return;
    const uint32_t gmask = 0xffffu << ((threadIdx.x & 31) / kB2G * kB2G);
    B2Chain A;
    b2_load(A, S, chainA, smem + (size_t) gib * npad, lane, gmask);
    __syncwarp(gmask);

    const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32);
    const uint32_t cidA = (uint32_t)(S.chain_id0 + chainA);
    const uint32_t ntt = (uint32_t) S.numTrialTypes, scale = 0xffffffffu / ntt, N = (uint32_t) S.N;
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

    uint32_t adapt_span = min(mdai_left, mvai_left), adapt_left = adapt_span;   // steps to the next adjustment of either kind

    uint32_t nmA_l = 0, w1A_l = 0, w2A_l = 0;   // this lane's share of the Philox batch
    uint32_t batch_pos = kB2G;
    auto draw = [&](uint32_t cid, uint64_t step, uint32_t &nm, uint32_t &w1, uint32_t &w2) {
        const Philox4 b = philox4x32_10((uint32_t) step, (uint32_t)(step >> 32), cid, kTagTrial, k0, k1);
        uint32_t k = b.w[0] / scale;
        if (k >= ntt) k = b2_redraw(b.w[3], scale, ntt);          // probability ~ ntt / 2^32
        nm = k; w1 = b.w[1]; w2 = b.w[2];
    };

    for (uint32_t s = 0; s < (uint32_t) a.nsteps; ++s) {
        ++sn;
        if (batch_pos == kB2G) {                       // lane j draws step sn + j for both chains
            draw(cidA, sn + lane, nmA_l, w1A_l, w2A_l);
            batch_pos = 0;
        }
        const uint32_t nmA = __shfl_sync(gmask, nmA_l, batch_pos, kB2G);
        const double rnA = u01(__shfl_sync(gmask, w1A_l, batch_pos, kB2G));
        const double ranA = u01(__shfl_sync(gmask, w2A_l, batch_pos, kB2G));
        ++batch_pos;

        uint8_t fA;
        if (nmA < N) {
            B2Trial tA = b2_displacement<INF>(A, nmA, rnA, ranA);
            if (!(tA.decided | tA.wall)) b2_resolve(A, tA, ranA);
            __syncwarp(gmask);                         // every lane has read the positions
            fA = b2_commit(A, tA);
            __syncwarp(gmask);
        } else {
            fA = coop_volume_full(A, rnA, ranA);       // fav :2161-2293
        }
        if (eci32 == 1 || --eci_left == 0) {           // ECheck :1965-2095
            double eA = b2_etest_partial<INF>(A);
#pragma unroll
            for (int o = kB2G / 2; o > 0; o >>= 1) eA += __shfl_xor_sync(gmask, eA, o, kB2G);
            A.echecks++;
            if (fabs(eA - A.tot[0]) > 0.0001) { A.echecks--; coop_energy_check(A); }
            eci_left = eci32;
        }
        coop_update_thermo(A);
        if (LOG && lane == 0) a.accept_log[(uint64_t) s * C + chainA] = fA;
        if (--adapt_left == 0) {                       // maxDisAdjust / maxDVAdjust steps (src/Main.cpp:145-165)
            mdai_left -= adapt_span; mvai_left -= adapt_span;
            const bool dis = mdai_left == 0, vol = mvai_left == 0;
            b2_adapt(A, a, dis, vol);
            if (dis) mdai_left = mdai32;
            if (vol) mvai_left = mvai32;
            adapt_span = adapt_left = min(mdai_left, mvai_left);
        }
    }
    __syncwarp(gmask);
    b2_store(A, S, chainA);
}

// ------------------------------------------------------------------------------------------------------------------
// k_chains_step_bond2 — the same chains, but the step loop is built around the LATENCY of one step, which is what
// bounds a launch of 4096 serial Markov chains (k_chains_step_bond: 1825 cycles per step, 346 warp instructions, the
// fp64 pipe 37 % busy, profiles/r02p_c2_*):
//   * Positions live in REGISTERS: lane q of the chain's 16 lanes holds r[q] and copies of r[q-1], r[q+1].  Every lane
//     evaluates the trial "my particle moves by md" with the common md; the lane that owns particle nm has the real
//     one, and its decision reaches the others through one ballot.  No shared memory, no rendezvous, no position
//     shuffles on the critical path; the neighbours of nm update their copies with the same r + md (same operands,
//     same double).
//   * The energy check of step t (:1965-2095: fresh bond energies, a butterfly sum, a comparison that never fires) and
//     the updateThermo of step t (:1941-1961) are RESOLVED DURING STEP t+1, after the trial of t+1 has been issued and
//     before it is committed — so E, Vir, l and the positions are still those of step t, an "Energy discrepancy" reset
//     would be taken on exactly the state the reference takes it on, and the butterfly's latency hides behind the
//     trial's.  The twelve sums are spread over the lanes (ThermoLanes, coop.cuh): 4 instead of 19 fp64 instructions
//     per step, same products, same order of addition.
//   * The random numbers of step t+1 are shuffled out of the Philox batch during step t.
//   * Volume trials (fav, one step in eleven), adjustments and an ECheck that fires go through coop.cuh's functions:
//     the positions are parked in the chain's shared-memory row for the duration and read back afterwards.
// Arithmetic and order of operations are those of k_chains_step_bond / coop.cuh: bit-identical to the oracle.
template <bool INF>
__device__ __forceinline__ double b2_bond_energy(double d, double cutoff) {     // phi_energy<HARMONIC>, selects only
    const double rijm = d - 1.0;
    return (d > 0) ? ((INF || d < cutoff) ? rijm * rijm : 0.0) : 10E10;
}

template <bool LOG, bool INF, bool EVERY>          // EVERY: ENGCHECK 1 (the INPUTstd cadence): no check countdown at all
__global__ void __launch_bounds__(128, 4) k_chains_step_bond2(ChainsDev S, StepArgs a, int npad) {
    extern __shared__ double smem[];
    const uint32_t gib = threadIdx.x / kB2G, lane = threadIdx.x % kB2G;
    const uint64_t C = S.nchains;
    uint64_t chain = (uint64_t) blockIdx.x * (blockDim.x / kB2G) + gib;
    // The two chains of a warp run ONE converged instruction stream (full-mask votes and shuffles: a per-group mask
    // costs a MATCH + REDUX + VOTE each); a surplus half-warp shadows the last chain and stores nothing.
    const bool own = chain < C;
    if (!own) chain = C - 1;
    const uint32_t gbase = (threadIdx.x & 31) / kB2G * kB2G;
    const uint32_t gmask = 0xffffu << gbase;
    constexpr uint32_t FULL = 0xffffffffu;
    B2Chain c;
    double *row = smem + (size_t) gib * npad;
    b2_load(c, S, chain, row, lane, gmask);
    ThermoLanes<kB2G> th;
    th.init(row + ((S.N + 1) & ~1ull), lane, S.acc, C, chain);
    __syncwarp();
    const uint32_t N = (uint32_t) S.N;
    const bool hasL = lane > 0 && lane < N, hasR = lane + 1 < N;
    double r = 0, rprev = 0, rnext = 0;                       // my particle and copies of its neighbours
    auto fetch = [&]() {
        r = row[lane < N ? lane : 0];
        rprev = row[hasL ? lane - 1 : 0];
        rnext = row[hasR ? lane + 1 : 0];
    };
    auto park = [&]() {                                        // positions -> shared row (for coop.cuh's functions)
        if (lane < N) row[lane] = r;
        __syncwarp(gmask);
    };
    fetch();

    const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32);
    const uint32_t cid = (uint32_t)(S.chain_id0 + chain);
    const uint32_t ntt = (uint32_t) S.numTrialTypes, scale = 0xffffffffu / ntt;
    uint64_t sn = a.sn0;
    auto until_event = [&]() -> uint32_t {                    // steps until an adjustment is due (the ECheck has its own cadence)
        uint64_t left = 0xffffffffull;
        if (a.adapt_device) {
            if (a.mdai) left = min(left, a.mdai - sn % a.mdai);
            if (a.mvai) left = min(left, a.mvai - sn % a.mvai);
        }
        return (uint32_t) left;
    };
    uint32_t ev_left = until_event();
    const uint32_t eci32 = a.eci > 0xfffffffeull ? 0xffffffffu : (uint32_t) a.eci;
    uint32_t eci_left = a.eci ? (uint32_t) min((uint64_t) 0xffffffffull, a.eci - sn % a.eci) : 0xffffffffu;

    uint32_t my_nm = 0, my_w1 = 0, my_w2 = 0;                 // this lane's share of the Philox batch
    uint32_t batch_pos = kB2G;
    auto draw = [&](uint64_t step, uint32_t &nm_o, uint32_t &w1_o, uint32_t &w2_o) {
        if (batch_pos == kB2G) {                              // lane j draws the block of step `step` + j
            const uint64_t mine = step + lane;
            const Philox4 b = philox4x32_10((uint32_t) mine, (uint32_t)(mine >> 32), cid, kTagTrial, k0, k1);
            uint32_t k = b.w[0] / scale;
            if (k >= ntt) k = b2_redraw(b.w[3], scale, ntt);  // probability ~ ntt / 2^32
            my_nm = k; my_w1 = b.w[1]; my_w2 = b.w[2];
            batch_pos = 0;
        }
        nm_o = __shfl_sync(FULL, my_nm, batch_pos, kB2G);
        w1_o = __shfl_sync(FULL, my_w1, batch_pos, kB2G);
        w2_o = __shfl_sync(FULL, my_w2, batch_pos, kB2G);
        ++batch_pos;
    };

    // what the previous step left to be resolved: its energy check (etest = the butterfly sum issued then) and its
    // updateThermo.  Called with the warp converged; the state is still the previous step's.
    bool pend_check = false, pend_thermo = false;
    double etest = 0.0;
    auto resolve = [&]() {
        const bool fire = pend_check && fabs(etest - c.tot[0]) > 0.0001;
        c.echecks += pend_check ? 1 : 0;
        if (__any_sync(FULL, fire)) {                          // never in practice: decide again on the reference-order sum
            if (fire) { c.echecks--; park(); coop_energy_check(c); }
            __syncwarp();
        }
        if (pend_thermo) {                                     // (uniform: false only before the first step of a launch)
            th.push(c);
            if (th.fill == kThermoRing) th.flush(c);
        }
    };

    uint32_t nm, w1, w2;
    draw(sn + 1, nm, w1, w2);
    for (uint32_t s = 0; s < (uint32_t) a.nsteps; ++s) {
        ++sn;
        __syncwarp();
        const bool more = s + 1 < (uint32_t) a.nsteps;
        uint32_t nm1 = 0, w11 = 0, w21 = 0;
        if (more) draw(sn + 1, nm1, w11, w21);
        const bool disp = nm < N;

        // qad2 :1160-1464, NBN == 1, for "my particle" — evaluated by every lane of both chains, a volume step included
        // (its result is discarded): the warp stays converged
        const double md = u01_shifted(w1, 1.5) * 2 * c.maxStep;               // (rn - 0.5) * 2 * maxStep, :1182
        const double ran = u01_shifted(w2, 1.0);
        const double rT = r + md;
        const bool wall = fabs(rT) > c.half_l;                                 // :1188
        double po0, po1, pn0, pn1, qo0, qo1, qn0, qn1;
        b2_phi<INF>(r - rprev, c.cutoff, c.two_over_l, po0, po1);
        b2_phi<INF>(rT - rprev, c.cutoff, c.two_over_l, pn0, pn1);
        b2_phi<INF>(rnext - r, c.cutoff, c.two_over_l, qo0, qo1);
        b2_phi<INF>(rnext - rT, c.cutoff, c.two_over_l, qn0, qn1);
        const double l0 = hasL ? (0.0 - po0 + pn0) : 0.0, l1 = hasL ? (0.0 - po1 + pn1) : 0.0;   // :1244
        const double r0 = hasR ? (0.0 - qo0 + qn0) : 0.0, r1 = hasR ? (0.0 - qo1 + qn1) : 0.0;   // :1339
        const double dE_own = l0 + r0, dV_own = l1 + r1;                                         // :1354
        const double ea = (double) exp_neg_approx(dE_own * c.invT);
        const bool down = dE_own <= 0;
        const bool acc_b = ran < ea - kMetropolisBand, rej_b = ran > ea + kMetropolisBand;
        bool accept = down | acc_b;
        const bool mine = disp && lane == nm;                  // this lane owns the moved particle
        const bool undecided = mine && !(down | acc_b | rej_b) && !wall;
        if (__any_sync(FULL, undecided)) {                     // 2e-5 of the trials
            if (undecided) accept = metropolis_exact(dE_own, c.T, ran);
            __syncwarp();
        }
        const uint32_t votes = __ballot_sync(FULL, mine && accept && !wall);
        const bool ok = (votes & gmask) != 0;
        // the owner's deltas for the replicated totals (off the critical path of the next trial)
        const double dE = __shfl_sync(FULL, dE_own, disp ? nm : 0u, kB2G), dV = __shfl_sync(FULL, dV_own, disp ? nm : 0u, kB2G);
        resolve();                                             // step t-1's check and thermo, on step t-1's state
        c.cnt[0] += ok ? 1 : 0;
        c.cnt[1] += (disp && !ok) ? 1 : 0;
        c.tot[0] = ok ? c.tot[0] + dE : c.tot[0];
        c.tot[1] = ok ? c.tot[1] + dV : c.tot[1];
        r = (ok && lane == nm) ? rT : r;
        rnext = (ok && lane + 1 == nm) ? rnext + md : rnext;   // the same r[nm] + md
        rprev = (ok && lane == nm + 1) ? rprev + md : rprev;
        uint8_t flags = 0;
        if constexpr (LOG) {
            const bool wall_nm = (__ballot_sync(FULL, mine && wall) & gmask) != 0;
            flags = wall_nm ? kLogWall : (ok ? kLogAccepted : 0);
        }
        if (__any_sync(FULL, !disp)) {                         // a volume trial in at least one of the two chains
            if (!disp) {                                       // fav :2161-2293 through coop.cuh on the parked positions
                th.flush(c);
                park();
                flags = coop_volume_full(c, u01(w1), u01(w2));
                __syncwarp(gmask);
                fetch();
            }
            __syncwarp();
        }
        // this step's energy check: fresh bond energies + butterfly now, comparison during the next step
        if constexpr (EVERY) pend_check = true;
        else { pend_check = false; if (--eci_left == 0) { pend_check = true; eci_left = eci32; } }
        if (EVERY || __any_sync(FULL, pend_check)) {
            etest = hasR ? b2_bond_energy<INF>(rnext - r, c.cutoff) : 0.0;
#pragma unroll
            for (int o = kB2G / 2; o > 0; o >>= 1) etest += __shfl_xor_sync(FULL, etest, o, kB2G);
        }
        pend_thermo = true;
        if (LOG && own && lane == 0) a.accept_log[(uint64_t) s * C + chain] = flags;
        if (--ev_left == 0) {                                  // maxDisAdjust / maxDVAdjust (src/Main.cpp:145-165): after the thermo
            resolve();
            pend_check = false; pend_thermo = false;
            b2_adapt(c, a, a.mdai && sn % a.mdai == 0, a.mvai && sn % a.mvai == 0);
            ev_left = until_event();
        }
        nm = nm1; w1 = w11; w2 = w21;
    }
    __syncwarp();
    resolve();
    th.flush(c);
    if (!own) return;
    th.store(lane, S.acc, C, chain);
    park();
    b2_store(c, S, chain, false);
}

}  // namespace jmm
