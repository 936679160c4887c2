// JMM_ARITH_FAST pair arithmetic for the LJ family (src/pot.cpp:19-108), shared by prod.cuh and sweep.cuh.
//
// A displacement trial needs, per partner, the OLD and the NEW pair term (qad2 :1231,1326) and nine
// (acc - old) + new sums (:1244,1339).  All nine follow from two numbers by the exact ratios of
// src/pot.cpp:56-66 (Vir12 = 12 E12, Vir6 = 6 E6, HV12 = 144 E12, HV6 = 36 E6):
//     s6  = sum_p ( b^-6  - a^-6  ),   s12 = sum_p ( b^-12 - a^-12 ),      a = old distance, b = new distance.
// With A = a^6, B = b^6 and ONE reciprocal  inv = 1/(A B):
//     b^-6 - a^-6   = (A - B) inv                      =: d6
//     b^-12 - a^-12 = (b^-6 - a^-6)(b^-6 + a^-6) = d6 * ((A + B) inv)
// = 2 DADD (a, b) + 6 DMUL (powers) + 1 DMUL (AB) + 3 DFMA (reciprocal) + 2 DADD + 2 DMUL + 1 DADD + 1 DFMA
// = 18 fp64-pipe instructions per partner for 33 algorithmic flop; the reference's own
// operation order costs ~76 (two IEEE divisions, 2 x 14 for phi, 18 accumulations).
// The cutoff of phiLJcut (`d <= cutOff` on the SIGNED distance, src/pot.cpp:53) masks A and B separately.
#pragma once
#include <stdint.h>

namespace jmm {

// 1/x for a positive normal x, far from overflow/underflow (x = a^6 b^6 with 1e-4 < a,b < 1e8).
// MUFU.RCP64H seed (relative error e0 <= 2^-23 by the PTX contract of rcp.approx.ftz.f64), then ONE cubic step
//     e = 1 - x y;   y' = y + y (e + e^2)      ->  relative error e0^3 = 2^-69, i.e. below the final rounding.
// x = 0 gives inf -> NaN, x = inf gives 0 -> NaN: a NaN dE is rejected by every comparison in
// metropolis_accept, which is what the reference does with the inf/NaN its own 1/(0) produces.
__device__ __forceinline__ double rcp_cubic(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = __fma_rn(-x, y, 1.0);
    const double e2 = __fma_rn(e, e, e);
    return __fma_rn(y, e2, y);
}

// Cutoff of phiLJcut: a pair term counts iff `d <= cutOff` on the SIGNED distance (src/pot.cpp:53).  The test is one
// DSETP (the fp64 pipe has slack wherever LJcut is used; a 64-bit integer compare is two issue slots), and a term is
// masked by clearing the HIGH word of its sixth power: what is left is a subnormal < 2^-1042, so that
// (A - B) inv and (A + B) inv round to exactly the values they have with a true zero whenever the other term is
// alive, and to (sub)normal noise below 1e-300 when both are masked — one SEL instead of two.
// In place: one DSETP and one predicated move of RZ into the high word (written with SEL on a copy, the compiler
// spends two more moves per term on rebuilding the register pair).
__device__ __forceinline__ void mask_beyond(double &x, double d, double cutoff) {
    asm("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi;\n\t"
        "setp.gtu.f64 p, %1, %2;\n\t"            // !(d <= cutoff): beyond the cutoff, or NaN
        "mov.b64 {lo, hi}, %0;\n\t"
        "@p mov.b32 hi, 0;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "+d"(x) : "d"(d), "d"(cutoff));
}

// One partner.  a, b = old and new SIGNED distance in the reference's orientation (r[j]-r[i], j > i by index).
template <bool CUT>
__device__ __forceinline__ void lj_partner(double a, double b, double cutoff, double &s6, double &s12) {
    const double a2 = a * a, b2 = b * b;
    double A = a2 * a2 * a2, B = b2 * b2 * b2;
    const double inv = rcp_cubic(A * B);
    if constexpr (CUT) {
        // new term b^-6 = A inv lives in A, old term a^-6 = B inv lives in B
        mask_beyond(A, b, cutoff);
        mask_beyond(B, a, cutoff);
    }
    const double d6 = (A - B) * inv;
    const double t6 = (A + B) * inv;
    s6 += d6;
    s12 = __fma_rn(d6, t6, s12);
}

// NP partners at once, written stage by stage (all first differences, all squares, all sixth powers, all reciprocals,
// ...) so that the NP independent dependency chains are issued interleaved: consecutive instructions of one chain
// are then NP issue slots (2 NP cycles) apart, which covers the fp64 latency without a second warp.  Written one
// partner after the other (lj_partner in an unrolled loop) ptxas keeps only ~2 chains in flight when registers are
// tight and the loop stalls on its own dependencies (profiles/r2b_c4lanes_*: "wait" 2.1 warps per issue).
// a[i], b[i] = old and new signed distance of partner i; the sums are accumulated in index order (as lj_partner).
template <bool CUT, int NP>
__device__ __forceinline__ void lj_partners(const double (&a)[NP], const double (&b)[NP], double cutoff, double &s6, double &s12) {
    double A[NP], B[NP], x[NP], y[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) { A[i] = a[i] * a[i]; B[i] = b[i] * b[i]; }
#pragma unroll
    for (int i = 0; i < NP; ++i) { x[i] = A[i] * A[i]; y[i] = B[i] * B[i]; }
#pragma unroll
    for (int i = 0; i < NP; ++i) { A[i] = x[i] * A[i]; B[i] = y[i] * B[i]; }          // a^6, b^6
#pragma unroll
    for (int i = 0; i < NP; ++i) x[i] = A[i] * B[i];
#pragma unroll
    for (int i = 0; i < NP; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(x[i]));
#pragma unroll
    for (int i = 0; i < NP; ++i) x[i] = __fma_rn(-x[i], y[i], 1.0);                    // e = 1 - x y
#pragma unroll
    for (int i = 0; i < NP; ++i) x[i] = __fma_rn(x[i], x[i], x[i]);                    // e + e^2
#pragma unroll
    for (int i = 0; i < NP; ++i) y[i] = __fma_rn(y[i], x[i], y[i]);                    // 1/(A B), see rcp_cubic
    if constexpr (CUT) {
#pragma unroll
        for (int i = 0; i < NP; ++i) { mask_beyond(A[i], b[i], cutoff); mask_beyond(B[i], a[i], cutoff); }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) { x[i] = (A[i] - B[i]) * y[i]; y[i] = (A[i] + B[i]) * y[i]; }   // d6, t6
#pragma unroll
    for (int i = 0; i < NP; ++i) { s6 += x[i]; s12 = __fma_rn(x[i], y[i], s12); }
}

// The nine deltas of qad2 from (s6, s12), component order of src/pot.cpp:90-100.
__device__ __forceinline__ void lj_nine(double s6, double s12, double (&d)[9]) {
    const double e12 = 4 * s12, e6 = 4 * s6;
    const double v12 = 12 * e12, v6 = 6 * e6, h12 = 144 * e12, h6 = 36 * e6;
    d[0] = e12 - e6; d[1] = v12 - v6; d[2] = e12; d[3] = v12; d[4] = e6; d[5] = v6; d[6] = h12 - h6; d[7] = h12; d[8] = h6;
}

}  // namespace jmm
