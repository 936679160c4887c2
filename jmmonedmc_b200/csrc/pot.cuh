// Pair potentials of src/pot.cpp, one pair term per call.
//   LJ / LJcut : src/pot.cpp:19-101 (phiLJcut), :104-108 (phiLJinfcutoff forces cutoff = inf)
//   HARMONIC   : src/pot.cpp:110-134 (phiHarmoniccut), :137-141
// The operation order below is the reference's, and the file is compiled with -fmad=false, so a
// term is the same double the x86-64 reference computes.
// Component order = the order phi[] is written, src/pot.cpp:90-100:
//   0 E, 1 Vir, 2 E12, 3 Vir12, 4 E6, 5 Vir6, 6 HV, 7 HV12, 8 HV6.
// HARMONIC only defines components 0 and 1 (the reference leaves 2..8 uninitialised): NC = 2.
#pragma once
#include "rng.cuh"

namespace jmm {

constexpr int kPotLJ = 0, kPotLJcut = 1, kPotHarmonic = 2;

template <int POT> struct PotTraits { static constexpr int NC = 9; };
template <> struct PotTraits<kPotHarmonic> { static constexpr int NC = 2; };

// VIR: params[0] of the reference (1 = also virial and hypervirial). lbox: params[1].
template <int POT, bool VIR>
__device__ __forceinline__ void phi(double d, double cutoff, double lbox, double (&o)[PotTraits<POT>::NC]) {
    if constexpr (POT == kPotHarmonic) {
        if (d <= 0) { o[0] = 10E10; o[1] = VIR ? 10E10 : 0.0; }
        else if (d < cutoff) {
            const double rijm = d - 1.0;
            o[0] = rijm * rijm;
            o[1] = VIR ? (2 / lbox) * d * rijm : 0.0;
        } else { o[0] = 0; o[1] = 0; }
    } else {
        const double rij3 = d * d * d;
        const double rij6 = 1 / (rij3 * rij3);
        const double rij12 = rij6 * rij6;
        const bool in = (POT == kPotLJ) ? true : (d <= cutoff);
        if (in) {
            const double phi6 = 4 * rij6, phi12 = 4 * rij12;
            o[0] = phi12 - phi6; o[2] = phi12; o[4] = phi6;
            if constexpr (VIR) {
                const double vir6 = 24 * rij6, vir12 = 48 * rij12, hv6 = 144 * rij6, hv12 = 576 * rij12;
                o[1] = vir12 - vir6; o[3] = vir12; o[5] = vir6;
                o[6] = hv12 - hv6;   o[7] = hv12;  o[8] = hv6;
            } else { o[1] = 0; o[3] = 0; o[5] = 0; o[6] = 0; o[7] = 0; o[8] = 0; }
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) o[k] = 0;
        }
    }
}

// energy only (ECheck's ETest, calculateEnergyOfTrialVolumeChange): component 0 of phi<POT,false>
template <int POT>
__device__ __forceinline__ double phi_energy(double d, double cutoff) {
    if constexpr (POT == kPotHarmonic) {
        if (d <= 0) return 10E10;
        if (d < cutoff) { const double rijm = d - 1.0; return rijm * rijm; }
        return 0.0;
    } else {
        const double rij3 = d * d * d;
        const double rij6 = 1 / (rij3 * rij3);
        const double rij12 = rij6 * rij6;
        const bool in = (POT == kPotLJ) ? true : (d <= cutoff);
        return in ? 4 * rij12 - 4 * rij6 : 0.0;
    }
}

// e^-x for x >= 0 in single precision: MUFU.EX2 of -x*log2(e).  Error budget: (float) x and the product round at
// 2^-24 each, so t = -x log2 e carries |t| 2^-22.6 and 2^t a relative |t| 2^-23.1; ex2.approx.ftz.f32 itself is
// good to 2^-22 (PTX ISA).  Absolute error <= e^-x (1.6e-7 x + 2.4e-7) <= 3e-7 for every x >= 0.
__device__ __forceinline__ float exp_neg_approx(double x) {
    const float t = (float) x * -1.4426950408889634f;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return e;
}
constexpr double kMetropolisBand = 1e-5;     // > 30 x the error bound of exp_neg_approx

// Exact Metropolis test, out of line: the band makes it a 2e-5 event, and inlined the compiler hoists its IEEE
// division and part of exp() above the branch (6 % of the instructions of the C2 kernel, profiles/r02b_c2_*).
static __device__ __noinline__ bool metropolis_exact(double dE, double T, double ran) { return dE <= 0 || exp(-dE / T) > ran; }

// Metropolis rule of qad2, src/jmmMCState.cpp:1367-1377: accept iff dE <= 0 || exp(-dE/T) > ran.
// exp() and the IEEE division are ~70 instructions; ran is uniform, so the decision is settled by a cheap
// approximation ea of exp(-dE/T) unless ran falls within 1e-5 of it (2e-5 of the trials):
//   ran > ea + 1e-5 -> reject,   ran < ea - 1e-5 -> accept,   otherwise evaluate exactly.
// The band is > 30 times the error of ea, so the result is always the one exp(-dE/T) > ran would give: decisions
// stay bit-identical to the reference-order code.  (Earlier version: Taylor bounds P3 <= e^-x <= P4, which left
// 4-20 % of the draws with x in (1, 1.5) undecided and a warp takes the slow path if ANY lane does: 7 % of the
// instructions of the C3 kernel, profiles/r01_c3_k_sweep_fast.txt.)  A NaN dE fails every comparison and is
// rejected by the exact test, like in the reference.
__device__ __forceinline__ bool metropolis_accept(double dE, double T, double invT, double ran) {
    if (dE <= 0) return true;
    const double ea = (double) exp_neg_approx(dE * invT);
    if (ran > ea + kMetropolisBand) return false;
    if (ran < ea - kMetropolisBand) return true;
    return metropolis_exact(dE, T, ran);
}

// Acceptance of a volume trial, qavLJ :1666-1672 / fav :2249-2255:  bf = exp(-(dE + P dl)/T + N log(lRat1)),
// accept iff bf >= 1.0 || bf > ran — which is bf > ran, because ran < 1.  (Philox streams only: ran is word 2 of the
// step's block whether or not the reference would have drawn it; the lock-step kernels draw it only if bf < 1.)
// exp and log in double precision are ~150 instructions that a warp executes whenever ANY of its chains makes a
// volume trial (17 % of the steps of C2), so the decision is taken on an approximation b of bf with a rigorous band:
//   log(s): lg2.approx.ftz.f32 of (float) s — input rounding 2^-24 relative, result within 2^-22.6 absolute (PTX ISA),
//           float result rounding <= 2^-24 |log2 s|: |error of ln s| <= 1.8e-7 for s in (0.5, 2);
//   exp(A): exp_neg_approx above, relative error <= 1.6e-7 |A| + 2.4e-7;
// so |b - bf| <= bf (1.8e-7 N + 1.6e-7 |A| + 2.4e-7).  Four times that is the band; inside it (<= ~1e-5 of the volume
// trials) the reference expression is evaluated.  NaN fails both comparisons and reaches the exact test, like every
// s outside the window below.
static __device__ __noinline__ bool volume_accept_exact(double x, double T, double n, double s, double ran) {
    const double bf = exp(-x / T + n * log(s));
    return bf >= 1.0 || bf > ran;
}

// The band is valid for 2^-6 < s < 2^6: lg2.approx is good to 2^-22.6 absolute on the mantissa part for any normal
// input, the float result rounds at 2^-24 |log2 s|, and the input rounding (float) s moves ln s by 2^-24, together
// <= 1.8e-7 + 4.2e-8 |log2 s| on ln s.  (The deck's own maxDVAdjust lets maxdl grow to N/2, :2136, so that a quarter of the
// volume trials of INPUTstd have s < 0.5: with the (0.5, 2) window of the first version they all paid for exp and log.)
constexpr double kVolumeBandLo = 0.015625, kVolumeBandHi = 64.0;
__device__ __forceinline__ double volume_accept_band(double n, double lg, double A) {
    return 4.0 * ((1.8e-7 + 4.2e-8 * fabs(lg)) * n + 1.6e-7 * fabs(A) + 2.4e-7);
}

__device__ __forceinline__ bool volume_accept(double x /* dE + P dl */, double T, double invT, double n, double s, double ran) {
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float) s));
    const double A = n * ((double) lg * 0.6931471805599453) - x * invT;
    const double b = (double) exp_neg_approx(-A);
    const double band = volume_accept_band(n, (double) lg, A);
    const bool narrow = s > kVolumeBandLo && s < kVolumeBandHi;
    if (narrow && ran < b * (1.0 - band)) return true;
    if (narrow && ran > b * (1.0 + band)) return false;
    return volume_accept_exact(x, T, n, s, ran);
}

}  // namespace jmm
