// Many-chain Metropolis kernels: one chain per thread, chain index fastest in every device array.
//
// Device layout (SoA, "coalesced many-chain layout"): r[i][c], rij[pair][c], tot[k][c], acc[k][c],
// cnt[k][c] with c the chain index, so the 32 chains of a warp load/store 256 contiguous bytes per
// quantity.  A launch keeps a chain's scalars in registers and its positions in shared memory
// (column c of an [N][blockDim] tile: bank = f(thread) only, so dynamically indexed r[nm] is
// conflict-free) for all nsteps steps; HBM sees 2*(8N+256) bytes per chain per launch.
//
// Reference functions restated here (file:line in /root/reference/src/jmmMCState.cpp):
//   Step :1758-1811, qad2 :1160-1464, qavLJ :1648-1730, fav :2161-2293, fad :853-1003,
//   ECheck :1965-2095, updateThermo :1941-1961, maxDisAdjust :2100-2115, maxDVAdjust :2120-2139,
//   relaxVolume :2396-2679, calculateEnergyOfTrialVolumeChange :2783-2827, moveVolume :2831-2916.
#pragma once
#include <type_traits>
#include <math.h>
#include "pot.cuh"

namespace jmm {

constexpr int kRngTaus2 = 0, kRngPhilox = 1, kRngRecorded = 2;
constexpr int kEnsNPT = 0, kEnsNLT = 1;
constexpr int kNAcc = 12, kNCnt = 4, kNTot = 9;

struct ChainsDev {
    uint64_t nchains, N, npairs, numTrialTypes;
    int nbn, ensemble, relax, pot;
    double cutoff;
    double *r;        // [N][nchains]
    double *rij;      // [npairs][nchains]   TABLE mode only
    double *l, *P, *T, *maxStep, *maxdl;     // [nchains]
    double *tot;      // [9][nchains]
    double *acc;      // [12][nchains]
    uint64_t *cnt;    // [4][nchains]
    uint64_t *vAErr;  // [nchains]  static vAErrNtot of maxDVAdjust, :2121
    uint64_t *echeck; // [2][nchains] checks, discrepancies
    uint32_t *taus;   // [3][nchains]
    uint64_t seed, chain_id0;
    int flags;        // JMM_FLAG_* of the handle's jmm_config
};

struct StepArgs {
    uint64_t sn0, nsteps;
    uint64_t eci, mdai, mvai;
    int adapt_device;
    double log_ideal;             // log(0.672924*0.5 + 0.0644284), evaluated by the host's libm
    const uint32_t *stream;       // recorded words
    uint64_t n_words;
    uint64_t *cursor;             // device scalar
    int *err;                     // device scalar: 1 = stream exhausted
    uint8_t *accept_log;          // [nsteps][nchains] or nullptr
    int pos_in_smem;
};

__device__ __forceinline__ uint64_t pair_index(uint64_t N, uint64_t i, uint64_t j) {   // :142-148
    return i * (N - 1) - (i * (i + 1)) / 2 + j - 1;
}

// Density rho(x) and two-particle density g(x) histograms: fgrho :1069-1127, qagrho :2297-2384,
// ugrho :1131-1149.  The reference adds the WHOLE histogram to the accumulator on every step (54 % of its run
// time at N = 80).  Here a bin keeps its current count `val` and  W = sum over its changes of delta * u,  u = the
// number of ugrho calls made before the change.  A change of delta at u is seen by the ugrho calls u+1 .. U, so
//     accumulated count after U calls  =  sum_k delta_k (U - u_k)  =  U * val - W          (the same integers),
// and a change is two commutative additions with no result: two fire-and-forget RED instructions, no load, no
// dependent latency (the earlier read-modify-write form ran C4 at 2.7e7 trials/s).  A read-out
// returns U*val - W and rebases W := U*val.  All arithmetic is modulo 2^64, exact while the true sum fits int64.
// Layout: chain-major [chain][bin], one 16-byte record per bin (one 32-byte sector holds two bins);
// null pointers = histograms off.
struct __align__(16) HistBin {
    unsigned long long W;  // sum of delta * u over the changes since the last read-out (mod 2^64)
    int32_t val;           // current count                                   (rhol :180, gl :187)
    int32_t pad;
};

struct HistDev {
    uint64_t rhonb, gnb;
    int gns;
    double rbw, gsw, gbw;
    HistBin *rho;                // [nchains][rhonb]
    HistBin *g;                  // [nchains][gns][gnb]
    uint64_t *ucount;            // [nchains] ugrho calls so far
};

__device__ __forceinline__ void hist_bump(HistBin *bins, uint64_t idx, int delta, uint64_t u) {
    atomicAdd(&bins[idx].val, delta);                                                   // RED.ADD (result unused)
    atomicAdd(&bins[idx].W, (unsigned long long) ((long long) delta * (long long) u));  // RED.ADD.64
}

// Zero every bin of one chain's histogram at ugrho count u, by ALL the lanes of `wmask` together: lane `rank` of
// `nlanes` takes bins rank, rank+nlanes, ... (a warp-wide load covers 512 contiguous bytes), and only non-zero bins
// are touched.  One thread scanning the 11 000 bins of the RunJobs geometry alone is ~11 000 dependent L2 round
// trips (5 ms per accepted volume move, measured); 32 lanes with four loads in flight each take ~50 us.
__device__ __forceinline__ void hist_clear_coop(HistBin *bins, uint64_t n, uint64_t u, uint32_t rank, uint32_t nlanes) {
#pragma unroll 4
    for (uint64_t b = rank; b < n; b += nlanes) {
        const int v = __ldcg(&bins[b].val);              // L2: REDs never update L1
        if (v != 0) hist_bump(bins, b, -v, u);
    }
}

// The counting half of fgrho :1069-1127 for one chain (positions r, stride rs; rij table when TABLE), split over the
// `nlanes` lanes that call it together: rho over the particles, g over the partners j of every particle i.  The
// additions commute, so who counts which pair does not matter (one lane alone: 3160 pairs of ~150 instructions,
// ~1.9 ms per accepted volume move at N = 80, measured).
template <bool TABLE>
__device__ __forceinline__ void hist_recount(const HistDev &H, uint64_t c, const double *r, size_t rs, const double *rij, size_t ts,
                                             uint32_t N, uint64_t u, uint32_t rank, uint32_t nlanes) {
    HistBin *rb = H.rho + c * H.rhonb;
    for (uint32_t i = rank; i < N; i += nlanes) {
        const long long k = (long long) floor(r[i * rs] / H.rbw + (double) H.rhonb / 2.0);
        if (k >= 0 && k < (long long) H.rhonb) hist_bump(rb, (uint64_t) k, +1, u);
    }
    HistBin *gb_ = H.g + c * (uint64_t) H.gns * H.gnb;
    for (uint32_t i = 0; i + 1 < N; ++i) {
        const double ri = r[i * rs];
        const long long gs1 = (long long) floor(ri / H.gsw + H.gns / 2.0);
        for (uint32_t j = i + 1 + rank; j < N; j += nlanes) {
            const double rj = r[j * rs];
            const long long gs2 = (long long) floor(rj / H.gsw + H.gns / 2.0);
            const double d = TABLE ? rij[pair_index(N, i, j) * ts] : rj - ri;
            const long long gb = (long long) floor(d / H.gbw);
            if (gb >= 0 && gb < (long long) H.gnb) {
                if (gs1 >= 0 && gs1 < H.gns) hist_bump(gb_, (uint64_t) (gs1 * H.gnb + gb), +1, u);
                if (gs2 >= 0 && gs2 < H.gns) hist_bump(gb_, (uint64_t) (gs2 * H.gnb + gb), +1, u);
            }
        }
    }
}

// qagrho :2297-2384 after an accepted displacement of particle nm by md (old position re-derived as r[nm]-md).
// A decrement and an increment of the SAME bin cancel (the reference does both); they are skipped together.
__device__ __forceinline__ void hist_qagrho(const HistDev &H, uint64_t c, const double *r, size_t rs, uint32_t N, uint32_t nm,
                                            double md, uint64_t u) {
    const double rn = r[nm * rs];
    const long long rbn1 = (long long) floor((rn - md) / H.rbw + (double) H.rhonb / 2.0);
    const long long rbn2 = (long long) floor(rn / H.rbw + (double) H.rhonb / 2.0);
    const bool r1 = rbn1 >= 0 && rbn1 < (long long) H.rhonb, r2 = rbn2 >= 0 && rbn2 < (long long) H.rhonb;
    if (!(r1 && r2 && rbn1 == rbn2)) {
        if (r1) hist_bump(H.rho, c * H.rhonb + rbn1, -1, u);
        if (r2) hist_bump(H.rho, c * H.rhonb + rbn2, +1, u);
    }
    const long long gs11 = (long long) floor((rn - md) / H.gsw + H.gns / 2.0);
    const long long gs12 = (long long) floor(rn / H.gsw + H.gns / 2.0);
    const bool in11 = gs11 >= 0 && gs11 < H.gns, in12 = gs12 >= 0 && gs12 < H.gns;
    const uint64_t base = c * (uint64_t) H.gns * H.gnb;
    for (uint32_t i = 0; i < N; ++i) {
        if (i == nm) continue;
        const double ri = r[i * rs];
        const long long gs2 = (long long) floor(ri / H.gsw + H.gns / 2.0);
        const bool in2 = gs2 >= 0 && gs2 < H.gns;
        const long long gb1 = (long long) (unsigned long long) floor(fabs(rn - md - ri) / H.gbw);
        const long long gb2 = (long long) (unsigned long long) floor(fabs(rn - ri) / H.gbw);
        const bool b1 = gb1 >= 0 && gb1 < (long long) H.gnb, b2 = gb2 >= 0 && gb2 < (long long) H.gnb;
        // moved particle's segment(s)
        if (!(in11 && b1 && in12 && b2 && gs11 == gs12 && gb1 == gb2)) {
            if (in11 && b1) hist_bump(H.g, base + gs11 * H.gnb + gb1, -1, u);
            if (in12 && b2) hist_bump(H.g, base + gs12 * H.gnb + gb2, +1, u);
        }
        // the other particle's segment
        if (in2 && !(b1 && b2 && gb1 == gb2)) {
            if (b1) hist_bump(H.g, base + gs2 * H.gnb + gb1, -1, u);
            if (b2) hist_bump(H.g, base + gs2 * H.gnb + gb2, +1, u);
        }
    }
}

// ---------------------------------------------------------------- per-thread chain context

template <int POT>
struct Chain {
    static constexpr int NC = PotTraits<POT>::NC;
    double *r;   size_t rs;       // positions, stride between particles
    double *rij; size_t ts;       // pair table, stride between pairs
    uint32_t N; int nbn;
    double cutoff;
    double l, P, T, maxStep, maxdl;
    double invT;                  // 1/T for the acceptance bounds only (metropolis_accept)
    double tot[NC];
    double acc[kNAcc];
    uint64_t cnt[kNCnt];
    uint64_t vAErr, echecks, discrepancies;
    bool consistent_virial;       // JMM_FLAG_CONSISTENT_VIRIAL
};

// r *= f for the whole chain (qavLJ :1692, fav :2264-2266, moveVolume :2847-2849)
template <int POT>
__device__ __forceinline__ void scale_positions(Chain<POT> &ch, double f) {
    for (uint32_t i = 0; i < ch.N; ++i) ch.r[i * ch.rs] = ch.r[i * ch.rs] * f;
}

template <int POT>
__device__ __forceinline__ uint32_t row_end(const Chain<POT> &ch, uint32_t i) {       // last j of row i
    return (ch.nbn < 0 || i + (uint32_t) ch.nbn > ch.N - 1) ? ch.N - 1 : i + (uint32_t) ch.nbn;
}

// Sum of phi over pairs in pair-index order, honouring NBN.  SCALED: r*scale first,
// exactly rTrial[jj]-rTrial[ii] of fav :2179,2212.
template <int POT, bool VIR, bool SCALED>
__device__ __forceinline__ void full_totals(const Chain<POT> &ch, double scale, double lbox,
                                            double (&out)[PotTraits<POT>::NC]) {
    constexpr int NC = PotTraits<POT>::NC;
    double p[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) out[k] = 0;
    for (uint32_t i = 0; i + 1 < ch.N; ++i) {
        const uint32_t jmax = row_end(ch, i);
        double ri = ch.r[i * ch.rs];
        if (SCALED) ri = ri * scale;
        for (uint32_t j = i + 1; j <= jmax; ++j) {
            double rj = ch.r[j * ch.rs];
            if (SCALED) rj = rj * scale;
            phi<POT, VIR>(rj - ri, ch.cutoff, lbox, p);
#pragma unroll
            for (int k = 0; k < NC; ++k) out[k] += p[k];
        }
    }
}

template <int POT, bool SCALED>
__device__ __forceinline__ double full_energy(const Chain<POT> &ch, double scale) {
    double e = 0;
    for (uint32_t i = 0; i + 1 < ch.N; ++i) {
        const uint32_t jmax = row_end(ch, i);
        double ri = ch.r[i * ch.rs];
        if (SCALED) ri = ri * scale;
        for (uint32_t j = i + 1; j <= jmax; ++j) {
            double rj = ch.r[j * ch.rs];
            if (SCALED) rj = rj * scale;
            e += phi_energy<POT>(rj - ri, ch.cutoff);
        }
    }
    return e;
}

// fad :907-946 / moveVolume :2865-2904 / ECheck reset :2028-2071: totals from positions; in TABLE
// mode the included pairs also get rij = r[j]-r[i].
template <int POT, bool TABLE>
__device__ __forceinline__ void recompute_into_state(Chain<POT> &ch) {
    constexpr int NC = PotTraits<POT>::NC;
    double p[NC], out[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) out[k] = 0;
    for (uint32_t i = 0; i + 1 < ch.N; ++i) {
        const uint32_t jmax = row_end(ch, i);
        const double ri = ch.r[i * ch.rs];
        for (uint32_t j = i + 1; j <= jmax; ++j) {
            const double d = ch.r[j * ch.rs] - ri;
            if (TABLE) ch.rij[pair_index(ch.N, i, j) * ch.ts] = d;
            phi<POT, true>(d, ch.cutoff, ch.l, p);
#pragma unroll
            for (int k = 0; k < NC; ++k) out[k] += p[k];
        }
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) ch.tot[k] = out[k];
}

template <int POT, bool TABLE>
__device__ __forceinline__ void table_from_positions(Chain<POT> &ch) {                 // :768, :2273
    if (!TABLE) return;
    for (uint32_t i = 0; i + 1 < ch.N; ++i)
        for (uint32_t j = i + 1; j < ch.N; ++j)
            ch.rij[pair_index(ch.N, i, j) * ch.ts] = ch.r[j * ch.rs] - ch.r[i * ch.rs];
}

// updateThermo :1941-1961
template <int POT>
__device__ __forceinline__ void update_thermo(Chain<POT> &ch) {
    const double rhotmp = (double) ch.N / ch.l;
    const double E = ch.tot[0], Vir = ch.tot[1];
    double HV = 0.0;                       // HARMONIC defines no hypervirial (see pot.cuh)
    if constexpr (PotTraits<POT>::NC > 6) HV = ch.tot[6];
    ch.acc[0] = ch.acc[0] + rhotmp;
    ch.acc[1] = ch.acc[1] + rhotmp * rhotmp;
    ch.acc[2] = ch.acc[2] + ch.l;
    ch.acc[3] = ch.acc[3] + ch.l * ch.l;
    ch.acc[4] = ch.acc[4] + E;
    ch.acc[5] = ch.acc[5] + E * E;
    ch.acc[6] = ch.acc[6] + ch.l * E;
    ch.acc[7] = ch.acc[7] + Vir;
    ch.acc[8] = ch.acc[8] + Vir * Vir;
    ch.acc[9] = ch.acc[9] + E * Vir;
    ch.acc[10] = ch.acc[10] + HV;
    ch.acc[11] = ch.acc[11] + HV * HV;
}

// moveVolume :2831-2916
template <int POT, bool TABLE>
__device__ __forceinline__ void move_volume(Chain<POT> &ch, double lnew) {
    const double lRat1 = lnew / ch.l;
    scale_positions(ch, lRat1);
    ch.l = lnew;
    recompute_into_state<POT, TABLE>(ch);
}

// relaxVolume :2396-2679 — Newton step on dE/dL = -P with +-0.1 finite differences, <= 20 iterations
template <int POT, bool TABLE>
__device__ __forceinline__ int relax_volume(Chain<POT> &ch) {
    double lTryMin = 0, lTryMax = 1E10;
    for (int count = 0; count < 20; ++count) {
        const double h = 0.1;
        const double EUp = full_energy<POT, true>(ch, (ch.l + h) / ch.l);           // :2423, :2783-2827
        const double EDown = full_energy<POT, true>(ch, (ch.l + (-h)) / ch.l);      // :2488
        const double first = (EUp - EDown) / (2 * h);
        const double second = (EUp - 2.0 * ch.tot[0] + EDown) / (h * h);
        double dlEstimate = -(ch.P - ((double) ch.N / ch.l) * ch.T + first) / second;
        const double relaxMax = 0.10 * (double) ch.N;
        if (fabs(dlEstimate) > relaxMax) dlEstimate = dlEstimate < 0 ? -relaxMax : relaxMax;
        if (ch.l + dlEstimate > lTryMax) dlEstimate = 0.5 * (lTryMax - ch.l);
        else if (ch.l + dlEstimate < lTryMin) dlEstimate = 0.5 * (lTryMin - ch.l);
        if (dlEstimate > 0.0) lTryMin = ch.l; else lTryMax = ch.l;
        const double relaxCrit = 0.0025 * (double) ch.N;
        const bool converged = fabs(dlEstimate) < relaxCrit;
        move_volume<POT, TABLE>(ch, ch.l + dlEstimate);
        if (converged) return 0;
    }
    return 1;
}

// ---------------------------------------------------------------- random streams

template <int RNG> struct Rng;

template <> struct Rng<kRngPhilox> {
    uint32_t k0, k1, chain;
    Philox4 b;
    __device__ __forceinline__ void begin(uint64_t sn) {
        b = philox4x32_10((uint32_t) sn, (uint32_t)(sn >> 32), chain, kTagTrial, k0, k1);
    }
    __device__ __forceinline__ uint32_t trial_type(uint32_t n, uint32_t scale) {
        uint32_t k = b.w[0] / scale;
        if (k >= n) { k = b.w[3] / scale; if (k >= n) k = mulhi32(b.w[3], n); }
        return k;
    }
    __device__ __forceinline__ double rn() { return u01(b.w[1]); }
    __device__ __forceinline__ double ran() { return u01(b.w[2]); }
};

template <> struct Rng<kRngTaus2> {
    uint32_t s1, s2, s3;
    __device__ __forceinline__ void begin(uint64_t) {}
    __device__ __forceinline__ uint32_t trial_type(uint32_t n, uint32_t scale) {      // gsl_rng_uniform_int
        uint32_t k;
        do { k = taus2_next(s1, s2, s3) / scale; } while (k >= n);
        return k;
    }
    __device__ __forceinline__ double rn() { return u01(taus2_next(s1, s2, s3)); }
    __device__ __forceinline__ double ran() { return u01(taus2_next(s1, s2, s3)); }
};

template <> struct Rng<kRngRecorded> {
    const uint32_t *w;
    uint64_t n, cur;
    bool exhausted;
    __device__ __forceinline__ uint32_t next() {
        if (cur >= n) { exhausted = true; return 0u; }
        return w[cur++];
    }
    __device__ __forceinline__ void begin(uint64_t) {}
    __device__ __forceinline__ uint32_t trial_type(uint32_t nn, uint32_t scale) {
        uint32_t k;
        do { k = next() / scale; } while (k >= nn);
        return k;
    }
    __device__ __forceinline__ double rn() { return u01(next()); }
    __device__ __forceinline__ double ran() { return u01(next()); }
};

// Philox stream outside the lock-step (TABLE) mode: the acceptance word of a step exists whether or not the
// reference would have drawn it, so the banded decisions of pot.cuh apply.
template <class RNG, bool TABLE>
constexpr bool kPhiloxProduction = std::is_same<RNG, Rng<kRngPhilox>>::value && !TABLE;

// ---------------------------------------------------------------- trial moves

constexpr uint8_t kLogAccepted = 1, kLogVolume = 2, kLogWall = 4;

// qad2 :1160-1464.  Partners are visited in ascending index; the left (ii<nm) and right (ii>nm)
// partial sums are kept apart and added at the end (:1277-1285, :1354-1362), each updated as
// acc = (acc - old) + new (:1244, :1339).  In TABLE mode old/new distances come from the rij table
// (rij +- md, :1216, :1311); otherwise from positions.
template <int POT, bool TABLE, class RNG>
__device__ __forceinline__ uint8_t displacement_trial(Chain<POT> &ch, uint32_t nm, double rn, RNG &rng) {
    constexpr int NC = PotTraits<POT>::NC;
    const double md = (rn - 0.5) * 2 * ch.maxStep;                                   // :1182
    const double rnm = ch.r[nm * ch.rs];
    const double rT = rnm + md;                                                      // :1183
    if (fabs(rT) > ch.l / 2.0) { ch.cnt[1]++; return kLogWall; }                     // :1188-1193

    const uint32_t N = ch.N;
    const uint32_t lo = (ch.nbn < 0 || (uint32_t) ch.nbn > nm) ? 0u : nm - (uint32_t) ch.nbn;
    const uint32_t hi = (ch.nbn < 0 || nm + (uint32_t) ch.nbn > N - 1) ? N - 1 : nm + (uint32_t) ch.nbn;
    double dsum[NC], dleft[NC], po[NC], pn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) { dsum[k] = 0; dleft[k] = 0; }
    for (uint32_t p = lo; p <= hi; ++p) {
        if (p == nm) {            // end of the left partners: park their sum, restart for the right ones
#pragma unroll
            for (int k = 0; k < NC; ++k) { dleft[k] = dsum[k]; dsum[k] = 0; }
            continue;
        }
        const bool left = p < nm;
        double dold, dnew;
        if (TABLE) {
            dold = ch.rij[(left ? pair_index(N, p, nm) : pair_index(N, nm, p)) * ch.ts];
            dnew = left ? dold + md : dold - md;
        } else {
            const double rp = ch.r[p * ch.rs];
            dold = left ? rnm - rp : rp - rnm;
            dnew = left ? rT - rp : rp - rT;
        }
        phi<POT, true>(dold, ch.cutoff, ch.l, po);
        phi<POT, true>(dnew, ch.cutoff, ch.l, pn);
#pragma unroll
        for (int k = 0; k < NC; ++k) dsum[k] = dsum[k] - po[k] + pn[k];
    }
    double d[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) d[k] = dleft[k] + dsum[k];

    bool accept = d[0] <= 0;
    if (!accept) accept = metropolis_accept(d[0], ch.T, ch.invT, rng.ran());      // :1367-1377 (ran drawn only here)
    if (!accept) { ch.cnt[1]++; return 0; }                                          // :1447
    ch.cnt[0]++;                                                                     // :1384-1394
#pragma unroll
    for (int k = 0; k < NC; ++k) ch.tot[k] += d[k];
    ch.r[nm * ch.rs] = rT;
    if (TABLE) {                                                                     // :1399-1428 (every partner)
        for (uint32_t i = 0; i < nm; ++i) ch.rij[pair_index(N, i, nm) * ch.ts] += md;
        for (uint32_t j = nm + 1; j < N; ++j) ch.rij[pair_index(N, nm, j) * ch.ts] -= md;
    }
    return kLogAccepted;
}

// qavLJ :1648-1730 (POT LJ and NBN < 0 only: E12 ~ s^-12, E6 ~ s^-6)
template <int POT, bool TABLE, class RNG>
__device__ __forceinline__ uint8_t volume_trial_scaling(Chain<POT> &ch, double rn, RNG &rng, double *defer_scale = nullptr) {
    static_assert(PotTraits<POT>::NC == 9, "scaling shortcut is LJ only");
    const double dl = (rn - 0.5) * 2 * ch.maxdl;
    const double lRat1 = (ch.l + dl) / ch.l;
    const double lRat3 = lRat1 * lRat1 * lRat1;
    const double lRat6 = 1 / (lRat3 * lRat3);
    const double lRat12 = lRat6 * lRat6;
    const double E12Trial = lRat12 * ch.tot[2];
    const double E6Trial = lRat6 * ch.tot[4];
    const double dE = E12Trial - E6Trial - ch.tot[0];
    bool accepted;
    if constexpr (kPhiloxProduction<RNG, TABLE>) {                // banded decision (pot.cuh), same outcome
        accepted = volume_accept(dE + ch.P * dl, ch.T, ch.invT, (double) ch.N, lRat1, rng.ran());
    } else {
        const double bf = exp(-(dE + ch.P * dl) / ch.T + (double) ch.N * log(lRat1));
        double ran = 0;
        if (bf < 1.0) ran = rng.ran();                            // :1667: drawn only when needed
        accepted = bf >= 1.0 || bf > ran;
    }
    if (!accepted) { ch.cnt[3]++; return kLogVolume; }
    ch.cnt[2]++;
    ch.tot[0] = ch.tot[0] + dE;
    ch.tot[2] = E12Trial;
    ch.tot[4] = E6Trial;
    ch.l = ch.l + dl;
    if (ch.consistent_virial) {            // JMM_FLAG_CONSISTENT_VIRIAL: the definition fav / moveVolume / ECheck use
        ch.tot[5] = lRat6 * ch.tot[5];  ch.tot[3] = lRat12 * ch.tot[3];  ch.tot[1] = ch.tot[3] - ch.tot[5];
        ch.tot[8] = lRat6 * ch.tot[8];  ch.tot[7] = lRat12 * ch.tot[7];  ch.tot[6] = ch.tot[7] - ch.tot[8];
    } else {
        const double lRat7 = lRat6 / lRat1, lRat13 = lRat12 / lRat1;
        ch.tot[5] = lRat7 * ch.tot[5];
        ch.tot[3] = lRat13 * ch.tot[3];
        ch.tot[1] = (double) ch.N * ch.T / ch.l + ch.tot[3] - ch.tot[5];             // :1686 (HV, HV6, HV12 stay as they were)
    }
    if (defer_scale) *defer_scale = lRat1;                                           // the caller scales (prod.cuh: warp-cooperative)
    else scale_positions(ch, lRat1);                                                 // :1692
    if (TABLE) {
        const uint64_t np = (uint64_t) ch.N * (ch.N - 1) / 2;
        for (uint64_t q = 0; q < np; ++q) ch.rij[q * ch.ts] = lRat1 * ch.rij[q * ch.ts];   // :1699
    }
    return kLogVolume | kLogAccepted;
}

// fav :2161-2293
template <int POT, bool TABLE, class RNG>
__device__ __forceinline__ uint8_t volume_trial_full(Chain<POT> &ch, double rn, RNG &rng, double *defer_scale = nullptr) {
    constexpr int NC = PotTraits<POT>::NC;
    const double dl = (rn - 0.5) * 2 * ch.maxdl;
    const double lnew = ch.l + dl;
    const double lRat1 = lnew / ch.l;
    double t[NC];
    full_totals<POT, true, true>(ch, lRat1, lnew, t);
    bool accepted;
    if constexpr (kPhiloxProduction<RNG, TABLE>) {
        accepted = volume_accept(t[0] - ch.tot[0] + ch.P * dl, ch.T, ch.invT, (double) ch.N, lRat1, rng.ran());
    } else {
        const double bf = exp(-(t[0] - ch.tot[0] + ch.P * dl) / ch.T + (double) ch.N * log(lRat1));   // :2249
        double ran = 0;
        if (bf < 1.0) ran = rng.ran();                            // :2251: drawn only when needed
        accepted = bf >= 1.0 || bf > ran;
    }
    if (!accepted) { ch.cnt[3]++; return kLogVolume; }
    ch.cnt[2]++;
    ch.l = ch.l + dl;
#pragma unroll
    for (int k = 0; k < NC; ++k) ch.tot[k] = t[k];
    if (defer_scale && !TABLE) *defer_scale = lRat1;
    else scale_positions(ch, lRat1);
    table_from_positions<POT, TABLE>(ch);
    return kLogVolume | kLogAccepted;
}

// ECheck :1965-2095 (a discrepancy resets the totals from the full sums once; see oracle header)
template <int POT, bool TABLE>
__device__ __forceinline__ void energy_check(Chain<POT> &ch) {
    const double ETest = full_energy<POT, false>(ch, 1.0);
    ch.echecks++;
    if (fabs(ETest - ch.tot[0]) > 0.0001) {
        ch.discrepancies++;
        recompute_into_state<POT, TABLE>(ch);
    }
}

// maxDisAdjust :2100-2115, maxDVAdjust :2120-2139 (device variant; the host variant is in jmm_gpu.cu)
template <int POT>
__device__ __forceinline__ void adjust_max_step(Chain<POT> &ch, double log_ideal) {
    const double actualRatio = (double) ch.cnt[0] / (double)(ch.cnt[0] + ch.cnt[1]);
    ch.maxStep = ch.maxStep * log_ideal / log(0.672924 * (actualRatio + 0.0644284));
    if (ch.maxStep < 0.002) ch.maxStep = 0.002;
    else if (ch.maxStep > 0.5) ch.maxStep = 0.5;
}
template <int POT>
__device__ __forceinline__ void adjust_max_dl(Chain<POT> &ch, double log_ideal) {
    if ((ch.cnt[2] + ch.cnt[3] - ch.vAErr) > 0) {
        ch.vAErr = ch.cnt[2] + ch.cnt[3];
        const double actualRatio = (double) ch.cnt[2] / (double)(ch.cnt[2] + ch.cnt[3]);
        ch.maxdl = ch.maxdl * log_ideal / log(0.672924 * (actualRatio + 0.0644284));
        if (ch.maxdl < 0.002 * (double) ch.N) ch.maxdl = 0.002 * (double) ch.N;
        else if (ch.maxdl > 0.10 * (double) ch.N) ch.maxdl = 0.50 * (double) ch.N;
    }
}

// What qad2 / qavLJ do to the histograms after the trial (:1431, :1453, :1717, :1726); fav does nothing (:2161-2293).
// COLLECTIVE over the lanes of `wmask` (every one of them calls it once per step, chain owner or not): the
// displacement updates are the owner's alone, but an fgrho (accepted qavLJ move) — zero every bin, count every
// particle and pair again — is done by the whole warp for one chain at a time.  `own` = this lane owns chain c
// (helpers pass false and ignore the rest).
template <int POT, bool TABLE>
__device__ __forceinline__ void hist_after_trial(const HistDev &H, uint64_t c, const Chain<POT> &ch, uint32_t nm, double rn,
                                                 double maxStep_used, uint8_t flags, bool scaling_volume, uint64_t &u,
                                                 bool own, unsigned wmask) {
    const bool refill = own && (flags & kLogVolume) && scaling_volume && (flags & kLogAccepted);
    unsigned need = __ballot_sync(wmask, refill);
    if (need) {
        const uint32_t lane = threadIdx.x & 31;
        const uint32_t rank = __popc(wmask & ((1u << lane) - 1u)), nlanes = __popc(wmask);
        if (refill) __threadfence();                                   // the owner's earlier REDs are performed before the helpers read
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const uint64_t cc = __shfl_sync(wmask, c, src), uu = __shfl_sync(wmask, u, src);
            // the owner's positions (its shared-memory column or its global column) and pair table, by generic address
            const double *pr = (const double *) __shfl_sync(wmask, (unsigned long long) ch.r, src);
            const double *pt = (const double *) __shfl_sync(wmask, (unsigned long long) ch.rij, src);
            const size_t prs = (size_t) __shfl_sync(wmask, (unsigned long long) ch.rs, src);
            const size_t pts = (size_t) __shfl_sync(wmask, (unsigned long long) ch.ts, src);
            const uint32_t pn = __shfl_sync(wmask, ch.N, src);
            hist_clear_coop(H.rho + cc * H.rhonb, H.rhonb, uu, rank, nlanes);
            hist_clear_coop(H.g + cc * (uint64_t) H.gns * H.gnb, (uint64_t) H.gns * H.gnb, uu, rank, nlanes);
            __syncwarp(wmask);                                         // every lane has read its bins: counting may start
            hist_recount<TABLE>(H, cc, pr, prs, pt, pts, pn, uu, rank, nlanes);
        }
    }
    if (!own) return;
    if (flags & kLogVolume) {
        if (scaling_volume) ++u;                                       // fav: no fgrho, no ugrho
        return;
    }
    if (flags & kLogAccepted) hist_qagrho(H, c, ch.r, ch.rs, ch.N, nm, (rn - 0.5) * 2 * maxStep_used, u);
    ++u;                                                               // ugrho :1453 on every displacement trial
}

// ---------------------------------------------------------------- state load / store

// CG: read through L2 (ld.global.cg) — needed when another SM may have written the state earlier in the
// same launch (prod.cuh, time-sliced kernel); L1 is not coherent between SMs.
template <bool CG, class T>
__device__ __forceinline__ T ld_state(const T *p) {
    if constexpr (CG) return __ldcg(p); else return *p;
}

template <int POT, bool CG = false>
__device__ __forceinline__ void load_chain(Chain<POT> &ch, const ChainsDev &S, uint64_t c, double *smem_col,
                                           uint32_t smem_stride) {
    constexpr int NC = PotTraits<POT>::NC;
    ch.N = (uint32_t) S.N; ch.nbn = S.nbn; ch.cutoff = S.cutoff; ch.consistent_virial = (S.flags & 1) != 0;
    ch.l = ld_state<CG>(S.l + c); ch.P = ld_state<CG>(S.P + c); ch.T = ld_state<CG>(S.T + c);
    ch.maxStep = ld_state<CG>(S.maxStep + c); ch.maxdl = ld_state<CG>(S.maxdl + c);
    ch.invT = 1.0 / ch.T;
#pragma unroll
    for (int k = 0; k < NC; ++k) ch.tot[k] = ld_state<CG>(S.tot + k * S.nchains + c);
#pragma unroll
    for (int k = 0; k < kNAcc; ++k) ch.acc[k] = ld_state<CG>(S.acc + k * S.nchains + c);
#pragma unroll
    for (int k = 0; k < kNCnt; ++k) ch.cnt[k] = ld_state<CG>(S.cnt + k * S.nchains + c);
    ch.vAErr = ld_state<CG>(S.vAErr + c); ch.echecks = ld_state<CG>(S.echeck + c);
    ch.discrepancies = ld_state<CG>(S.echeck + S.nchains + c);
    ch.rij = S.rij ? S.rij + c : nullptr; ch.ts = S.nchains;
    if (smem_col) {
        for (uint32_t i = 0; i < ch.N; ++i) smem_col[i * smem_stride] = ld_state<CG>(S.r + i * S.nchains + c);
        ch.r = smem_col; ch.rs = smem_stride;
    } else { ch.r = S.r + c; ch.rs = S.nchains; }
}

template <int POT>
__device__ __forceinline__ void store_chain(const Chain<POT> &ch, const ChainsDev &S, uint64_t c, bool pos_in_smem) {
    constexpr int NC = PotTraits<POT>::NC;
    S.l[c] = ch.l; S.maxStep[c] = ch.maxStep; S.maxdl[c] = ch.maxdl;
#pragma unroll
    for (int k = 0; k < NC; ++k) S.tot[k * S.nchains + c] = ch.tot[k];
#pragma unroll
    for (int k = NC; k < kNTot; ++k) S.tot[k * S.nchains + c] = 0.0;    // HARMONIC: components 2..8 defined as 0
#pragma unroll
    for (int k = 0; k < kNAcc; ++k) S.acc[k * S.nchains + c] = ch.acc[k];
#pragma unroll
    for (int k = 0; k < kNCnt; ++k) S.cnt[k * S.nchains + c] = ch.cnt[k];
    S.vAErr[c] = ch.vAErr; S.echeck[c] = ch.echecks; S.echeck[S.nchains + c] = ch.discrepancies;
    if (pos_in_smem)
        for (uint32_t i = 0; i < ch.N; ++i) S.r[i * S.nchains + c] = ch.r[i * ch.rs];
}

// ---------------------------------------------------------------- kernels

// Prologue of main(): fad step 0 (src/Main.cpp:66-68), relaxVolume if RELAX (:71-73), updateThermo (:96)
constexpr int kStartFad = 1, kStartRelax = 2, kStartThermo = 4;

template <int POT, bool TABLE>
__global__ void k_chains_start(ChainsDev S, int pos_in_smem, int parts) {
    extern __shared__ double smem[];
    const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S.nchains) return;
    Chain<POT> ch;
    load_chain(ch, S, c, pos_in_smem ? smem + threadIdx.x : nullptr, blockDim.x);
    if (parts & kStartFad) {
        table_from_positions<POT, TABLE>(ch);      // fad sets rijTrial for every pair, :918
        recompute_into_state<POT, TABLE>(ch);
        ch.cnt[0]++;                               // :968
    }
    if ((parts & kStartRelax) && S.relax > 0 && S.ensemble == kEnsNPT) relax_volume<POT, TABLE>(ch);
    if (parts & kStartThermo) update_thermo(ch);
    store_chain(ch, S, c, pos_in_smem != 0);
}

template <int POT, bool TABLE>
__global__ void k_chains_relax(ChainsDev S, int pos_in_smem, int /*parts*/) {
    extern __shared__ double smem[];
    const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S.nchains) return;
    Chain<POT> ch;
    load_chain(ch, S, c, pos_in_smem ? smem + threadIdx.x : nullptr, blockDim.x);
    relax_volume<POT, TABLE>(ch);
    store_chain(ch, S, c, pos_in_smem != 0);
}

// configuration totals in the reference's summation order, one thread per chain (jmm_energy exact)
template <int POT>
__global__ void k_chains_totals_exact(ChainsDev S, double *out /*[9][nchains]*/) {
    const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S.nchains) return;
    Chain<POT> ch;
    ch.N = (uint32_t) S.N; ch.nbn = S.nbn; ch.cutoff = S.cutoff; ch.l = S.l[c];
    ch.r = S.r + c; ch.rs = S.nchains;
    constexpr int NC = PotTraits<POT>::NC;
    double t[NC];
    full_totals<POT, true, false>(ch, 1.0, ch.l, t);
#pragma unroll
    for (int k = 0; k < NC; ++k) out[k * S.nchains + c] = t[k];
#pragma unroll
    for (int k = NC; k < kNTot; ++k) out[k * S.nchains + c] = 0.0;
}

// nsteps x Step() for every chain
template <int POT, bool TABLE, int RNG>
__global__ void __launch_bounds__(128) k_chains_step(ChainsDev S, StepArgs a, HistDev H) {
    extern __shared__ double smem[];
    const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    Chain<POT> ch;
    if (c >= S.nchains) {
        // lanes without a chain: with histograms on they stay as helpers of the warp-wide refill (hist_after_trial)
        if (H.ucount) {
            uint64_t u = 0;
            for (uint64_t s = 0; s < a.nsteps; ++s) hist_after_trial<POT, TABLE>(H, 0, ch, 0, 0.0, 0.0, 0, false, u, false, 0xffffffffu);
        }
        return;
    }
    load_chain(ch, S, c, a.pos_in_smem ? smem + threadIdx.x : nullptr, blockDim.x);

    Rng<RNG> rng;
    if constexpr (RNG == kRngPhilox) {
        rng.k0 = (uint32_t) S.seed; rng.k1 = (uint32_t)(S.seed >> 32); rng.chain = (uint32_t)(S.chain_id0 + c);
    } else if constexpr (RNG == kRngTaus2) {
        rng.s1 = S.taus[c]; rng.s2 = S.taus[S.nchains + c]; rng.s3 = S.taus[2 * S.nchains + c];
    } else {
        rng.w = a.stream; rng.n = a.n_words; rng.cur = *a.cursor; rng.exhausted = false;
    }

    const uint32_t ntt = (uint32_t) S.numTrialTypes;
    const uint32_t scale = 0xffffffffu / ntt;
    const bool scaling_volume = (POT == kPotLJ) && S.nbn < 0;                         // dispatch :296-301
    uint64_t sn = a.sn0;
    // countdowns to the next multiple of each interval (sn % x == 0 tests of :1858, Main.cpp:145-176)
    uint64_t eci_left = a.eci ? a.eci - sn % a.eci : ~0ull;
    uint64_t mdai_left = (a.adapt_device && a.mdai) ? a.mdai - sn % a.mdai : ~0ull;
    uint64_t mvai_left = (a.adapt_device && a.mvai) ? a.mvai - sn % a.mvai : ~0ull;
    uint64_t relax_left = (a.adapt_device && S.relax > 0 && S.ensemble == kEnsNPT) ? 10000 - sn % 10000 : ~0ull;

    uint64_t hist_u = H.ucount ? H.ucount[c] : 0;
    for (uint64_t s = 0; s < a.nsteps; ++s) {
        ++sn;                                                                         // incrementStep :1745
        rng.begin(sn);
        const uint32_t nm = rng.trial_type(ntt, scale);                               // :1762
        const double rn = rng.rn();                                                   // :1763
        const double maxStep_used = ch.maxStep;
        uint8_t flags;
        if (nm < ch.N) flags = displacement_trial<POT, TABLE>(ch, nm, rn, rng);       // :1783-1785
        else {
            if constexpr (POT == kPotLJ) {
                flags = scaling_volume ? volume_trial_scaling<POT, TABLE>(ch, rn, rng)
                                       : volume_trial_full<POT, TABLE>(ch, rn, rng);
            } else flags = volume_trial_full<POT, TABLE>(ch, rn, rng);                // :1786-1788
        }
        if (H.ucount) hist_after_trial<POT, TABLE>(H, c, ch, nm, rn, maxStep_used, flags, scaling_volume, hist_u, true, 0xffffffffu);
        if (--eci_left == 0) { energy_check<POT, TABLE>(ch); eci_left = a.eci; }      // :1800-1802
        update_thermo(ch);                                                            // :1805
        if (a.accept_log) a.accept_log[s * S.nchains + c] = flags;
        if (a.adapt_device) {                                                         // src/Main.cpp:145-176
            if (--mdai_left == 0) { adjust_max_step(ch, a.log_ideal); mdai_left = a.mdai; }
            if (--mvai_left == 0) { adjust_max_dl(ch, a.log_ideal); mvai_left = a.mvai; }
            if (--relax_left == 0) { if (sn < 1000000ull) relax_volume<POT, TABLE>(ch); relax_left = 10000; }
        }
    }

    if constexpr (RNG == kRngTaus2) {
        S.taus[c] = rng.s1; S.taus[S.nchains + c] = rng.s2; S.taus[2 * S.nchains + c] = rng.s3;
    } else if constexpr (RNG == kRngRecorded) {
        *a.cursor = rng.cur;
        if (rng.exhausted) *a.err = 1;
    }
    if (H.ucount) H.ucount[c] = hist_u;
    store_chain(ch, S, c, a.pos_in_smem != 0);
}

// histogram set-up (setupMCS :773-776: fgrho + one ugrho on the initial configuration) and read-out
static __global__ void k_hist_init(ChainsDev S, HistDev H) {
    const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S.nchains) return;
    // distances from the positions: setupMCS fills rij = r[j]-r[i] (:768) just before its fgrho, the same doubles
    hist_recount<false>(H, c, S.r + c, S.nchains, nullptr, 0, (uint32_t) S.N, 0, 0, 1);    // the bins were allocated zeroed
    H.ucount[c] = 1;
}

// read-out: accumulated counts since the last read-out = U*val - W, then rebase W := U*val
// (printRho :1021-1038 / printG :1042-1064)
static __global__ void k_hist_take(HistBin *bins, const uint64_t *ucount, uint64_t bins_per_chain, uint64_t nchains, long long *out) {
    const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bins_per_chain * nchains) return;
    const unsigned long long U = ucount[t / bins_per_chain];
    HistBin b = bins[t];
    const unsigned long long uv = U * (unsigned long long) (long long) b.val;
    out[t] = (long long) (uv - b.W);
    b.W = uv;
    bins[t] = b;
}

// chain-major host layout <-> chain-fastest device layout
static __global__ void k_transpose_in(const double *__restrict__ src /*[nchains][n]*/, double *__restrict__ dst /*[n][nchains]*/,
                               uint64_t nchains, uint64_t n) {
    const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchains * n) return;
    const uint64_t i = t / nchains, c = t % nchains;
    dst[t] = src[c * n + i];
}
static __global__ void k_transpose_out(const double *__restrict__ src /*[n][nchains]*/, double *__restrict__ dst /*[nchains][n]*/,
                                uint64_t nchains, uint64_t n) {
    const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchains * n) return;
    const uint64_t c = t / n, i = t % n;
    dst[t] = src[i * nchains + c];
}
static __global__ void k_lattice_chain_major(double *r /*[nchains][N]*/, const double *l, uint64_t nchains, uint64_t N) {   // :561
    const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchains * N) return;
    const uint64_t c = t / N, i = t % N;
    r[t] = (((double) i + 0.5) / (double) N - 0.5) * l[c];
}
static __global__ void k_lattice(double *r /*[N][nchains]*/, const double *l, uint64_t nchains, uint64_t N) {   // :561
    const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchains * N) return;
    const uint64_t i = t / nchains, c = t % nchains;
    r[t] = (((double) i + 0.5) / (double) N - 0.5) * l[c];
}

}  // namespace jmm
