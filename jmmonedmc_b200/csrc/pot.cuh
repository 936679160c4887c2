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

// Metropolis rule of qad2, src/jmmMCState.cpp:1367-1377: accept iff dE <= 0 || exp(-dE/T) > ran.
// exp() and the division are ~70 instructions, and ran is uniform, so the decision is almost always
// settled by the Taylor bounds  P3(x) <= exp(-x) <= P4(x)  (valid for every x >= 0):
//   ran < P3(x) - 1e-9  -> accept,   ran > P4(x) + 1e-9 -> reject,   otherwise evaluate exactly
// (for x > 1.5 the reciprocal of the degree-4 partial sum of e^x bounds exp(-x) from above instead).
// The 1e-9 margins dwarf the rounding of x = dE*(1/T) and of the polynomials (~1e-15), so the result is
// always the one exp(-dE/T) > ran would give: decisions stay bit-identical to the reference-order code.
__device__ __forceinline__ bool metropolis_accept(double dE, double T, double invT, double ran) {
    if (dE <= 0) return true;
    const double x = dE * invT;
    if (x <= 1.5) {
        const double x2 = x * x;
        const double p3 = 1.0 - x + x2 * (0.5 - x * (1.0 / 6.0));
        if (ran < p3 - 1e-9) return true;
        const double p4 = p3 + x2 * x2 * (1.0 / 24.0);
        if (ran > p4 + 1e-9) return false;
    } else {
        // large x: e^x >= 1 + x + x^2/2 + x^3/6 + x^4/24, so exp(-x) <= 1/(that): almost every draw is rejected here
        const double q = 1.0 + x * (1.0 + x * (0.5 + x * ((1.0 / 6.0) + x * (1.0 / 24.0))));
        if (ran * q > 1.0 + 1e-9 * q) return false;
    }
    return exp(-dE / T) > ran;
}

}  // namespace jmm
