// Cooperative many-chain kernel: G lanes of a warp share ONE chain (G = 8, 16 or 32).
//
// Why: a Markov chain is strictly serial, so with few chains (config C2: 4096) the one-chain-per-thread
// kernel of chains.cuh runs one warp per SM sub-partition at ~0.18 IPC and pays for every branch any of
// its 32 chains takes (displacement + volume + edge cases: ~1600 warp instructions per step, ncu
// profiles/r01a).  Here control flow is uniform inside a group, the idle lanes do useful work, and the
// 4096 chains become 4096 warps (7 per sub-partition) that hide each other's latency:
//   * Philox: lane j of a group computes the block of step sn+1+j, so one round of Philox serves G steps;
//   * the pair loops of fav / ECheck / relaxVolume / moveVolume and the partner loop of qad2 are strided
//     over the lanes;
//   * sums that feed the state are still added in the reference's order: every lane writes its terms to a
//     per-group shared-memory scratch and all lanes then add them in index order (identical registers in
//     every lane, no broadcast needed).  ECheck's ETest only feeds a |dE| > 1e-4 test
//     (src/jmmMCState.cpp:1998), so it alone uses a butterfly reduction.
// Results are bit-identical to chains.cuh in RECOMPUTE mode and to the oracle's RECOMPUTE mode.
//
// Occupancy note (measured, C2, trial moves/s): __launch_bounds__(128) -> 128 registers, 4 CTAs/SM: 2.93e9;
// forcing 6 or 8 CTAs/SM (80 / 64 registers, 250-480 B of spills): 2.16e9 / 1.77e9; (128,1): 1.75e9.
//
// Reference lines restated: see chains.cuh (same functions, same order of operations).
#pragma once
#include "chains.cuh"

namespace jmm {

constexpr int kCoopChunk = 32;                   // scratch rows per group

template <int POT, int G>
struct Coop {
    static constexpr int NC = PotTraits<POT>::NC;
    // group-uniform (replicated in every lane of the group)
    double *r;                                    // shared memory row of this chain, r[0..N)
    double *sc;                                   // scratch [kCoopChunk][2*NC]
    uint32_t N; int nbn;
    double cutoff;
    double l, P, T, maxStep, maxdl;
    double half_l, rho, two_over_l;               // l/2.0, N/l, 2/l: recomputed whenever l changes
    double invT;                                  // 1/T, only for the acceptance bounds (metropolis_accept)
    double tot[NC];
    double acc[kNAcc];
    uint64_t cnt[kNCnt];
    uint64_t vAErr, echecks, discrepancies;
    // lane identity
    uint32_t lane;                                // 0..G-1 inside the group
    uint32_t gmask;                               // lanes of this group inside the warp
    bool chain_of_bonds;                          // NBN == 1 and N-1 <= G: one nearest-neighbour pair per lane
    bool consistent_virial;                       // JMM_FLAG_CONSISTENT_VIRIAL (include/jmm_gpu.h)
    bool lean;                                    // scratch is [G][NC] only: ordered sums spread over the lanes (lanes.cuh)

    __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
    __device__ __forceinline__ void set_l(double lnew) {
        l = lnew; half_l = lnew / 2.0; rho = (double) N / lnew; two_over_l = 2 / lnew;
    }
    __device__ __forceinline__ uint32_t rowlen(uint32_t i) const {
        const uint32_t rest = N - 1 - i;
        return (nbn < 0 || (uint32_t) nbn > rest) ? rest : (uint32_t) nbn;
    }
    __device__ __forceinline__ uint32_t npairs_included() const {
        if (nbn < 0 || (uint32_t) nbn >= N - 1) return N * (N - 1) / 2;
        const uint32_t k = (uint32_t) nbn;
        return (N - k) * k + k * (k - 1) / 2;     // N-k full rows of k, then k-1, ..., 1
    }
};

// pair term with the harmonic 2/l hoisted (same value: (2/l) is evaluated first in src/pot.cpp:126)
template <int POT, bool VIR>
__device__ __forceinline__ void phi_h(double d, double cutoff, double two_over_l, double (&o)[PotTraits<POT>::NC]) {
    if constexpr (POT == kPotHarmonic) {
        if (d <= 0) { o[0] = 10E10; o[1] = VIR ? 10E10 : 0.0; }
        else if (d < cutoff) {
            const double rijm = d - 1.0;
            o[0] = rijm * rijm;
            o[1] = VIR ? two_over_l * d * rijm : 0.0;
        } else { o[0] = 0; o[1] = 0; }
    } else {
        phi<POT, VIR>(d, cutoff, 0.0, o);
    }
}

// Ordered sum over the included pairs (pair-index order), terms computed lane-strided.
// term(i, j, out[NCX]) must be a pure function of shared positions; result identical in every lane.
template <int POT, int G, int NCX, class F>
__device__ __forceinline__ void coop_pair_sum(const Coop<POT, G> &c, F term, double (&out)[NCX]) {
#pragma unroll
    for (int k = 0; k < NCX; ++k) out[k] = 0;
    if (c.chain_of_bonds) {
        // NBN == 1 and N-1 <= G: pair q is (q, q+1) and lane q owns it.  The ordered sum is a walk over
        // the lanes with shuffles: no scratch, no barrier.
        double t[NCX];
#pragma unroll
        for (int k = 0; k < NCX; ++k) t[k] = 0;
        if (c.lane + 1 < c.N) term(c.lane, c.lane + 1, t);
        for (uint32_t q = 0; q + 1 < c.N; ++q) {
#pragma unroll
            for (int k = 0; k < NCX; ++k) out[k] += __shfl_sync(c.gmask, t[k], q, G);
        }
        return;
    }
    const uint32_t total = c.npairs_included();
    // per-lane cursor over pairs q = lane, lane+G, ... : (i, off) with j = i+1+off
    uint32_t i = 0, off = c.lane;
    if (c.lean) {
        // G pairs at a time through a [G][NCX] scratch; the NCX ordered sums are SPREAD over the lanes (lane k adds
        // component k, k+G, ... of the G terms in pair order), so a chunk costs G loads + G additions per lane instead
        // of G*NCX, and the scratch is G*NCX doubles instead of kCoopChunk*2*NC.  Same order of addition per component.
        constexpr int M = (NCX + G - 1) / G;
        double mine[M];
#pragma unroll
        for (int m = 0; m < M; ++m) mine[m] = 0;
        for (uint32_t base = 0; base < total; base += G) {
            const uint32_t n = min((uint32_t) G, total - base);
            if (c.lane < n) {
                while (off >= c.rowlen(i)) { off -= c.rowlen(i); ++i; }
                double t[NCX];
                term(i, i + 1 + off, t);
                double *dst = c.sc + c.lane * NCX;
#pragma unroll
                for (int k = 0; k < NCX; ++k) dst[k] = t[k];
                off += G;
            }
            c.sync();
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const uint32_t comp = c.lane + m * G;
                if (comp < NCX) {
                    if (n == G) {
#pragma unroll
                        for (int q = 0; q < G; ++q) mine[m] += c.sc[q * NCX + comp];
                    } else {
                        for (uint32_t q = 0; q < n; ++q) mine[m] += c.sc[q * NCX + comp];
                    }
                }
            }
            c.sync();
        }
#pragma unroll
        for (int k = 0; k < NCX; ++k) out[k] = __shfl_sync(c.gmask, mine[k / G], k % G, G);
        return;
    }
    for (uint32_t base = 0; base < total; base += kCoopChunk) {
        const uint32_t end = min(total, base + kCoopChunk);
        for (uint32_t q = base + c.lane; q < end; q += G) {
            while (off >= c.rowlen(i)) { off -= c.rowlen(i); ++i; }
            double t[NCX];
            term(i, i + 1 + off, t);
            double *dst = c.sc + (q - base) * NCX;
#pragma unroll
            for (int k = 0; k < NCX; ++k) dst[k] = t[k];
            off += G;
        }
        c.sync();
        for (uint32_t q = base; q < end; ++q) {
            const double *src = c.sc + (q - base) * NCX;
#pragma unroll
            for (int k = 0; k < NCX; ++k) out[k] += src[k];
        }
        c.sync();
    }
}

// totals from positions r*scale with virial (fav :2196-2235; scale = 1: fad/moveVolume/ECheck reset)
template <int POT, int G, bool SCALED>
__device__ __forceinline__ void coop_full_totals(const Coop<POT, G> &c, double scale, double two_over_l,
                                                 double (&out)[PotTraits<POT>::NC]) {
    constexpr int NC = PotTraits<POT>::NC;
    coop_pair_sum<POT, G, NC>(c, [&](uint32_t i, uint32_t j, double (&t)[NC]) {
        double ri = c.r[i], rj = c.r[j];
        if (SCALED) { ri = ri * scale; rj = rj * scale; }
        phi_h<POT, true>(rj - ri, c.cutoff, two_over_l, t);
    }, out);
}

// energy only, reference order (relaxVolume's EUp/EDown feed the new box length, so the order matters)
template <int POT, int G, bool SCALED>
__device__ __forceinline__ double coop_full_energy(const Coop<POT, G> &c, double scale) {
    double e[1];
    coop_pair_sum<POT, G, 1>(c, [&](uint32_t i, uint32_t j, double (&t)[1]) {
        double ri = c.r[i], rj = c.r[j];
        if (SCALED) { ri = ri * scale; rj = rj * scale; }
        t[0] = phi_energy<POT>(rj - ri, c.cutoff);
    }, e);
    return e[0];
}

// EUp and EDown of one relaxVolume iteration in ONE pass over the pairs (two independent ordered sums: the same
// doubles as two coop_full_energy calls, half the cursor arithmetic and two dependent addition chains in flight)
template <int POT, int G>
__device__ __forceinline__ void coop_full_energy2(const Coop<POT, G> &c, double scale_a, double scale_b, double &ea, double &eb) {
    double e[2];
    coop_pair_sum<POT, G, 2>(c, [&](uint32_t i, uint32_t j, double (&t)[2]) {
        const double ri = c.r[i], rj = c.r[j];
        t[0] = phi_energy<POT>(rj * scale_a - ri * scale_a, c.cutoff);
        t[1] = phi_energy<POT>(rj * scale_b - ri * scale_b, c.cutoff);
    }, e);
    ea = e[0]; eb = e[1];
}

// ECheck's ETest (:1974-1993): only compared against 1e-4, so lane partial sums + butterfly
template <int POT, int G>
__device__ __forceinline__ double coop_energy_unordered(const Coop<POT, G> &c) {
    double e = 0;
    if (c.chain_of_bonds) {
        if (c.lane + 1 < c.N) e = phi_energy<POT>(c.r[c.lane + 1] - c.r[c.lane], c.cutoff);
    } else {
        const uint32_t total = c.npairs_included();
        uint32_t i = 0, off = c.lane;
        for (uint32_t q = c.lane; q < total; q += G) {
            while (off >= c.rowlen(i)) { off -= c.rowlen(i); ++i; }
            e += phi_energy<POT>(c.r[i + 1 + off] - c.r[i], c.cutoff);
            off += G;
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) e += __shfl_xor_sync(c.gmask, e, o, G);
    return e;
}

template <int POT, int G>
__device__ __forceinline__ void coop_recompute_into_state(Coop<POT, G> &c) {
    double t[PotTraits<POT>::NC];
    coop_full_totals<POT, G, false>(c, 1.0, c.two_over_l, t);
#pragma unroll
    for (int k = 0; k < PotTraits<POT>::NC; ++k) c.tot[k] = t[k];
}

template <int POT, int G>
__device__ __forceinline__ void coop_scale_positions(Coop<POT, G> &c, double s, bool s_first) {
    c.sync();
    for (uint32_t i = c.lane; i < c.N; i += G) c.r[i] = s_first ? s * c.r[i] : c.r[i] * s;
    c.sync();
}

template <int POT, int G>
__device__ __forceinline__ void coop_move_volume(Coop<POT, G> &c, double lnew) {       // moveVolume :2831-2916
    const double lRat1 = lnew / c.l;
    coop_scale_positions(c, lRat1, false);
    c.set_l(lnew);
    coop_recompute_into_state(c);
}

template <int POT, int G>
__device__ __forceinline__ int coop_relax_volume(Coop<POT, G> &c) {                    // relaxVolume :2396-2679
    double lTryMin = 0, lTryMax = 1E10;
    for (int count = 0; count < 20; ++count) {
        const double h = 0.1;
        double EUp, EDown;
        coop_full_energy2<POT, G>(c, (c.l + h) / c.l, (c.l + (-h)) / c.l, EUp, EDown);
        const double first = (EUp - EDown) / (2 * h);
        const double second = (EUp - 2.0 * c.tot[0] + EDown) / (h * h);
        double dlEstimate = -(c.P - ((double) c.N / c.l) * c.T + first) / second;
        const double relaxMax = 0.10 * (double) c.N;
        if (fabs(dlEstimate) > relaxMax) dlEstimate = dlEstimate < 0 ? -relaxMax : relaxMax;
        if (c.l + dlEstimate > lTryMax) dlEstimate = 0.5 * (lTryMax - c.l);
        else if (c.l + dlEstimate < lTryMin) dlEstimate = 0.5 * (lTryMin - c.l);
        if (dlEstimate > 0.0) lTryMin = c.l; else lTryMax = c.l;
        const double relaxCrit = 0.0025 * (double) c.N;
        const bool converged = fabs(dlEstimate) < relaxCrit;
        coop_move_volume(c, c.l + dlEstimate);
        if (converged) return 0;
    }
    return 1;
}

template <int POT, int G>
__device__ __forceinline__ void coop_update_thermo(Coop<POT, G> &c) {                  // updateThermo :1941-1961
    const double E = c.tot[0], Vir = c.tot[1];
    c.acc[0] = c.acc[0] + c.rho;
    c.acc[1] = c.acc[1] + c.rho * c.rho;
    c.acc[2] = c.acc[2] + c.l;
    c.acc[3] = c.acc[3] + c.l * c.l;
    c.acc[4] = c.acc[4] + E;
    c.acc[5] = c.acc[5] + E * E;
    c.acc[6] = c.acc[6] + c.l * E;
    c.acc[7] = c.acc[7] + Vir;
    c.acc[8] = c.acc[8] + Vir * Vir;
    c.acc[9] = c.acc[9] + E * Vir;
    if constexpr (PotTraits<POT>::NC > 6) {          // HARMONIC: HV == 0, acc[10], acc[11] stay +0
        const double HV = c.tot[6];
        c.acc[10] = c.acc[10] + HV;
        c.acc[11] = c.acc[11] + HV * HV;
    }
}

// qad2 :1160-1464.  Few partners (2*NBN <= 8): every lane evaluates them itself (no communication).
// Otherwise the partners are strided over the lanes and (old, new) terms go through the scratch so that
// the (acc - old) + new chain runs in ascending partner index, left and right sums apart.
template <int POT, int G>
__device__ __forceinline__ uint8_t coop_displacement(Coop<POT, G> &c, uint32_t nm, double rn, double ran) {
    constexpr int NC = PotTraits<POT>::NC;
    const double md = (rn - 0.5) * 2 * c.maxStep;
    const double rnm = c.r[nm];
    const double rT = rnm + md;
    if (fabs(rT) > c.half_l) { c.cnt[1]++; return kLogWall; }
    const uint32_t N = c.N;
    const uint32_t lo = (c.nbn < 0 || (uint32_t) c.nbn > nm) ? 0u : nm - (uint32_t) c.nbn;
    const uint32_t hi = (c.nbn < 0 || nm + (uint32_t) c.nbn > N - 1) ? N - 1 : nm + (uint32_t) c.nbn;
    double dsum[NC], dleft[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { dsum[k] = 0; dleft[k] = 0; }
    if (c.nbn == 1) {
        // at most one partner on each side: straight-line, predicated
        double po[NC], pn[NC];
        if (nm > 0) {
            const double rp = c.r[nm - 1];
            phi_h<POT, true>(rnm - rp, c.cutoff, c.two_over_l, po);
            phi_h<POT, true>(rT - rp, c.cutoff, c.two_over_l, pn);
#pragma unroll
            for (int k = 0; k < NC; ++k) dleft[k] = 0.0 - po[k] + pn[k];
        }
        if (nm + 1 < N) {
            const double rp = c.r[nm + 1];
            phi_h<POT, true>(rp - rnm, c.cutoff, c.two_over_l, po);
            phi_h<POT, true>(rp - rT, c.cutoff, c.two_over_l, pn);
#pragma unroll
            for (int k = 0; k < NC; ++k) dsum[k] = 0.0 - po[k] + pn[k];
        }
    } else if (c.nbn >= 0 && c.nbn <= 4) {
        double po[NC], pn[NC];
        for (uint32_t p = lo; p <= hi; ++p) {
            if (p == nm) {
#pragma unroll
                for (int k = 0; k < NC; ++k) { dleft[k] = dsum[k]; dsum[k] = 0; }
                continue;
            }
            const bool left = p < nm;
            const double rp = c.r[p];
            phi_h<POT, true>(left ? rnm - rp : rp - rnm, c.cutoff, c.two_over_l, po);
            phi_h<POT, true>(left ? rT - rp : rp - rT, c.cutoff, c.two_over_l, pn);
#pragma unroll
            for (int k = 0; k < NC; ++k) dsum[k] = dsum[k] - po[k] + pn[k];
        }
    } else {
        const uint32_t count = hi - lo + 1;                   // includes nm itself (a no-op slot)
        for (uint32_t base = 0; base < count; base += kCoopChunk) {
            const uint32_t end = min(count, base + kCoopChunk);
            for (uint32_t q = base + c.lane; q < end; q += G) {
                const uint32_t p = lo + q;
                if (p == nm) continue;
                const bool left = p < nm;
                const double rp = c.r[p];
                double po[NC], pn[NC];
                phi_h<POT, true>(left ? rnm - rp : rp - rnm, c.cutoff, c.two_over_l, po);
                phi_h<POT, true>(left ? rT - rp : rp - rT, c.cutoff, c.two_over_l, pn);
                double *dst = c.sc + (q - base) * 2 * NC;
#pragma unroll
                for (int k = 0; k < NC; ++k) { dst[k] = po[k]; dst[NC + k] = pn[k]; }
            }
            c.sync();
            for (uint32_t q = base; q < end; ++q) {
                if (lo + q == nm) {
#pragma unroll
                    for (int k = 0; k < NC; ++k) { dleft[k] = dsum[k]; dsum[k] = 0; }
                    continue;
                }
                const double *src = c.sc + (q - base) * 2 * NC;
#pragma unroll
                for (int k = 0; k < NC; ++k) dsum[k] = dsum[k] - src[k] + src[NC + k];
            }
            c.sync();
        }
    }
    double d[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) d[k] = dleft[k] + dsum[k];
    if (!metropolis_accept(d[0], c.T, c.invT, ran)) { c.cnt[1]++; return 0; }
    c.cnt[0]++;
#pragma unroll
    for (int k = 0; k < NC; ++k) c.tot[k] += d[k];
    c.sync();                                   // every lane has read r[nm] and its partners
    if (c.lane == 0) c.r[nm] = rT;
    c.sync();
    return kLogAccepted;
}

template <int POT, int G>
__device__ __forceinline__ uint8_t coop_volume_scaling(Coop<POT, G> &c, double rn, double ran) {   // qavLJ :1648-1730
    static_assert(PotTraits<POT>::NC == 9, "LJ only");
    const double dl = (rn - 0.5) * 2 * c.maxdl;
    const double lRat1 = (c.l + dl) / c.l;
    const double lRat3 = lRat1 * lRat1 * lRat1;
    const double lRat6 = 1 / (lRat3 * lRat3);
    const double lRat12 = lRat6 * lRat6;
    const double E12Trial = lRat12 * c.tot[2];
    const double E6Trial = lRat6 * c.tot[4];
    const double dE = E12Trial - E6Trial - c.tot[0];
    if (!volume_accept(dE + c.P * dl, c.T, c.invT, (double) c.N, lRat1, ran)) { c.cnt[3]++; return kLogVolume; }   // :1666-1672
    c.cnt[2]++;
    c.tot[0] = c.tot[0] + dE;
    c.tot[2] = E12Trial;
    c.tot[4] = E6Trial;
    c.set_l(c.l + dl);
    if (c.consistent_virial) {
        c.tot[5] = lRat6 * c.tot[5];  c.tot[3] = lRat12 * c.tot[3];  c.tot[1] = c.tot[3] - c.tot[5];
        c.tot[8] = lRat6 * c.tot[8];  c.tot[7] = lRat12 * c.tot[7];  c.tot[6] = c.tot[7] - c.tot[8];
    } else {
        const double lRat7 = lRat6 / lRat1, lRat13 = lRat12 / lRat1;
        c.tot[5] = lRat7 * c.tot[5];
        c.tot[3] = lRat13 * c.tot[3];
        c.tot[1] = (double) c.N * c.T / c.l + c.tot[3] - c.tot[5];
    }
    coop_scale_positions(c, lRat1, true);
    return kLogVolume | kLogAccepted;
}

template <int POT, int G>
__device__ __forceinline__ uint8_t coop_volume_full(Coop<POT, G> &c, double rn, double ran) {      // fav :2161-2293
    constexpr int NC = PotTraits<POT>::NC;
    const double dl = (rn - 0.5) * 2 * c.maxdl;
    const double lnew = c.l + dl;
    const double lRat1 = lnew / c.l;
    double t[NC];
    coop_full_totals<POT, G, true>(c, lRat1, 2 / lnew, t);
    if (!volume_accept(t[0] - c.tot[0] + c.P * dl, c.T, c.invT, (double) c.N, lRat1, ran)) { c.cnt[3]++; return kLogVolume; }   // :2249-2255
    c.cnt[2]++;
    c.set_l(c.l + dl);
#pragma unroll
    for (int k = 0; k < NC; ++k) c.tot[k] = t[k];
    coop_scale_positions(c, lRat1, false);
    return kLogVolume | kLogAccepted;
}

template <int POT, int G>
__device__ __forceinline__ void coop_energy_check(Coop<POT, G> &c) {                  // ECheck :1965-2095
    const double ETest = coop_energy_unordered(c);
    c.echecks++;
    if (fabs(ETest - c.tot[0]) > 0.0001) {
        // decide on the reference-order sum so that the reset is taken exactly when the oracle takes it
        const double exact = coop_full_energy<POT, G, false>(c, 1.0);
        if (fabs(exact - c.tot[0]) > 0.0001) { c.discrepancies++; coop_recompute_into_state(c); }
    }
}

// maxDisAdjust :2100-2115, maxDVAdjust :2120-2139 (device variant)
template <int POT, int G>
__device__ __forceinline__ void coop_adjust_max_step(Coop<POT, G> &c, double log_ideal) {
    const double actualRatio = (double) c.cnt[0] / (double)(c.cnt[0] + c.cnt[1]);
    c.maxStep = c.maxStep * log_ideal / log(0.672924 * (actualRatio + 0.0644284));
    if (c.maxStep < 0.002) c.maxStep = 0.002;
    else if (c.maxStep > 0.5) c.maxStep = 0.5;
}
template <int POT, int G>
__device__ __forceinline__ void coop_adjust_max_dl(Coop<POT, G> &c, double log_ideal) {
    if ((c.cnt[2] + c.cnt[3] - c.vAErr) > 0) {
        c.vAErr = c.cnt[2] + c.cnt[3];
        const double actualRatio = (double) c.cnt[2] / (double)(c.cnt[2] + c.cnt[3]);
        c.maxdl = c.maxdl * log_ideal / log(0.672924 * (actualRatio + 0.0644284));
        if (c.maxdl < 0.002 * (double) c.N) c.maxdl = 0.002 * (double) c.N;
        else if (c.maxdl > 0.10 * (double) c.N) c.maxdl = 0.50 * (double) c.N;
    }
}

// nsteps x Step() :1758-1811, G lanes per chain, Philox stream, positions from shared memory
template <int POT, int G, bool LOG>
__global__ void __launch_bounds__(128) k_chains_step_coop(ChainsDev S, StepArgs a, int npad) {
    constexpr int NC = PotTraits<POT>::NC;
    extern __shared__ double smem[];
    const uint32_t groups_per_block = blockDim.x / G;
    const uint32_t gib = threadIdx.x / G;                     // group in block
    const uint64_t chain = (uint64_t) blockIdx.x * groups_per_block + gib;
    if (chain >= S.nchains) return;                           // whole groups leave together
    const uint64_t C = S.nchains;

    Coop<POT, G> c;
    c.lane = threadIdx.x % G;
    c.gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    c.chain_of_bonds = (S.nbn == 1) && (S.N - 1 <= (uint64_t) G);
    c.lean = false;
    c.consistent_virial = (S.flags & 1) != 0;
    const size_t per_group = (size_t) npad + (size_t) kCoopChunk * 2 * NC;
    c.r = smem + gib * per_group;
    c.sc = c.r + npad;
    c.N = (uint32_t) S.N; c.nbn = S.nbn; c.cutoff = S.cutoff;
    c.P = S.P[chain]; c.T = S.T[chain]; c.maxStep = S.maxStep[chain]; c.maxdl = S.maxdl[chain];
    c.invT = 1.0 / c.T;
    c.set_l(S.l[chain]);
#pragma unroll
    for (int k = 0; k < NC; ++k) c.tot[k] = S.tot[k * C + chain];
#pragma unroll
    for (int k = 0; k < kNAcc; ++k) c.acc[k] = S.acc[k * C + chain];
#pragma unroll
    for (int k = 0; k < kNCnt; ++k) c.cnt[k] = S.cnt[k * C + chain];
    c.vAErr = S.vAErr[chain]; c.echecks = S.echeck[chain]; c.discrepancies = S.echeck[C + chain];
    for (uint32_t i = c.lane; i < c.N; i += G) c.r[i] = S.r[(uint64_t) i * C + chain];
    c.sync();

    const uint32_t k0 = (uint32_t) S.seed, k1 = (uint32_t)(S.seed >> 32), cid = (uint32_t)(S.chain_id0 + chain);
    const uint32_t ntt = (uint32_t) S.numTrialTypes;
    const uint32_t scale = 0xffffffffu / ntt;
    const bool scaling_volume = (POT == kPotLJ) && S.nbn < 0;
    uint64_t sn = a.sn0;
    // countdowns to the next multiple of each interval; a launch is far shorter than 2^32 steps and the
    // host never asks for more than that per launch, so 32-bit counters saturated at 2^32-1 are exact
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

    for (uint32_t s = 0; s < (uint32_t) a.nsteps; ++s) {
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
        if (nm < c.N) flags = coop_displacement(c, nm, rn, ran);
        else {
            if constexpr (POT == kPotLJ) {
                flags = scaling_volume ? coop_volume_scaling(c, rn, ran) : coop_volume_full(c, rn, ran);
            } else flags = coop_volume_full(c, rn, ran);
        }
        if (--eci_left == 0) { coop_energy_check(c); eci_left = eci32; }
        coop_update_thermo(c);
        if (LOG && c.lane == 0) a.accept_log[(uint64_t) s * C + chain] = flags;
        if (a.adapt_device) {
            if (--mdai_left == 0) { coop_adjust_max_step(c, a.log_ideal); mdai_left = mdai32; }
            if (--mvai_left == 0) { coop_adjust_max_dl(c, a.log_ideal); mvai_left = mvai32; }
            if (--relax_left == 0) { if (sn < 1000000ull) coop_relax_volume(c); relax_left = 10000; }
        }
    }

    c.sync();
    for (uint32_t i = c.lane; i < c.N; i += G) S.r[(uint64_t) i * C + chain] = c.r[i];
    if (c.lane == 0) {
        S.l[chain] = c.l; S.maxStep[chain] = c.maxStep; S.maxdl[chain] = c.maxdl;
#pragma unroll
        for (int k = 0; k < NC; ++k) S.tot[k * C + chain] = c.tot[k];
#pragma unroll
        for (int k = NC; k < kNTot; ++k) S.tot[k * C + chain] = 0.0;
#pragma unroll
        for (int k = 0; k < kNAcc; ++k) S.acc[k * C + chain] = c.acc[k];
#pragma unroll
        for (int k = 0; k < kNCnt; ++k) S.cnt[k * C + chain] = c.cnt[k];
        S.vAErr[chain] = c.vAErr; S.echeck[chain] = c.echecks; S.echeck[C + chain] = c.discrepancies;
    }
}

// updateThermo :1941-1961 without twelve replicated accumulators.  Every lane of a group knows rho, l, E, Vir, HV of the
// step, and replicated sums would cost 19 fp64 instructions per step and 24 registers in EVERY lane.  Instead the
// five numbers of each step go into a small ring in shared memory (one lane writes), and every kThermoRing steps the
// twelve sums — SPREAD over the lanes: lane k owns acc[k], acc[k+G], ... — are brought up to date: lane k reads its
// operand pair (a, b) of every buffered step and does acc += a * b in step order (b = 1.0 for the linear terms: a * 1.0
// is exact).  Same products, same order of addition per sum: the twelve sums stay bit-identical to the reference's.
constexpr int kThermoRing = 8;                   // steps buffered
constexpr int kThermoSlots = 6;                  // rho, l, E, Vir, HV, 1.0
template <int G> struct ThermoLanes {
    static constexpr int M = (kNAcc + G - 1) / G;
    double acc[M];
    const double *pa[M], *pb[M];                  // this lane's operand slots in ring entry 0 (entry e: + e * kThermoSlots)
    double *ring;                                 // [kThermoRing][kThermoSlots] doubles of this group
    uint32_t fill;
    // sums in JMM_A_* order: rho, rho^2, l, l^2, E, E^2, l E, Vir, Vir^2, E Vir, HV, HV^2
    static __device__ __forceinline__ void operands(uint32_t k, uint32_t &a, uint32_t &b) {
        const uint32_t A[12] = {0, 0, 1, 1, 2, 2, 1, 3, 3, 2, 4, 4};
        const uint32_t B[12] = {5, 0, 5, 1, 5, 2, 2, 5, 3, 3, 5, 4};
        a = A[k < 12 ? k : 0]; b = B[k < 12 ? k : 0];
    }
    __device__ __forceinline__ void init(double *ring_, uint32_t lane, const double *acc_g, uint64_t C, uint64_t chain) {
        ring = ring_; fill = 0;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t k = lane + m * G;
            uint32_t ia, ib;
            operands(k, ia, ib);
            pa[m] = ring + ia; pb[m] = ring + ib;
            acc[m] = k < kNAcc ? __ldcg(acc_g + (uint64_t) k * C + chain) : 0.0;
        }
        if (lane < kThermoRing) ring[lane * kThermoSlots + 5] = 1.0;      // the constant operand of the linear sums, once
        if (G < kThermoRing && lane == 0)
            for (int e = G; e < kThermoRing; ++e) ring[e * kThermoSlots + 5] = 1.0;
    }
    template <int POT>
    __device__ __forceinline__ void push(const Coop<POT, G> &c) {                 // one step's numbers (any one lane writes)
        if (c.lane == 0) {
            double *e = ring + fill * kThermoSlots;
            e[0] = c.rho; e[1] = c.l; e[2] = c.tot[0]; e[3] = c.tot[1];
            e[4] = PotTraits<POT>::NC > 6 ? c.tot[6] : 0.0;
        }
        ++fill;
    }
    template <int POT>
    __device__ __forceinline__ void flush(const Coop<POT, G> &c) {                // bring the sums up to date
        c.sync();
        if (fill == kThermoRing) {
#pragma unroll
            for (int e = 0; e < kThermoRing; ++e) {
#pragma unroll
                for (int m = 0; m < M; ++m) acc[m] = acc[m] + pa[m][e * kThermoSlots] * pb[m][e * kThermoSlots];
            }
        } else {
            for (uint32_t e = 0; e < fill; ++e) {
#pragma unroll
                for (int m = 0; m < M; ++m) acc[m] = acc[m] + pa[m][e * kThermoSlots] * pb[m][e * kThermoSlots];
            }
        }
        fill = 0;
        c.sync();
    }
    __device__ __forceinline__ void store(uint32_t lane, double *acc_g, uint64_t C, uint64_t chain) const {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t k = lane + m * G;
            if (k < kNAcc) acc_g[(uint64_t) k * C + chain] = acc[m];
        }
    }
};


}  // namespace jmm
