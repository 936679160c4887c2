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

__device__ __forceinline__ void b2_store(const B2Chain &c, const ChainsDev &S, uint64_t chain) {
    const uint64_t C = S.nchains;
    for (uint32_t i = c.lane; i < c.N; i += kB2G) S.r[(uint64_t) i * C + chain] = c.r[i];
    if (c.lane == 0) {
        S.l[chain] = c.l; S.maxStep[chain] = c.maxStep; S.maxdl[chain] = c.maxdl;
#pragma unroll
        for (int k = 0; k < 2; ++k) S.tot[k * C + chain] = c.tot[k];
#pragma unroll
        for (int k = 2; k < kNTot; ++k) S.tot[k * C + chain] = 0.0;
#pragma unroll
        for (int k = 0; k < kNAcc; ++k) S.acc[k * C + chain] = c.acc[k];
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
__global__ void __launch_bounds__(128) k_chains_step_bond(ChainsDev S, StepArgs a, int npad) {
    extern __shared__ double smem[];
    const uint32_t gib = threadIdx.x / kB2G, lane = threadIdx.x % kB2G;
    const uint64_t chainA = (uint64_t) blockIdx.x * (blockDim.x / kB2G) + gib;
    const uint64_t C = S.nchains;
    if (chainA >= C) return;
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


}  // namespace jmm
