// Random streams of the hot path.
//   Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11): counter-based production stream.
//   taus2: the reference's generator (GNU GSL rng/taus.c, used at src/jmmMCState.cpp:779-781),
//          run on the device so a chain can replay the reference's own stream without a recording.
// Shared by host and device code (the host seeds taus2 state words).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define JMM_HD __host__ __device__ __forceinline__
#else
#define JMM_HD inline
#endif

namespace jmm {

struct Philox4 { uint32_t w[4]; };

JMM_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t) a * b) >> 32);
#endif
}

JMM_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 r; r.w[0] = c0; r.w[1] = c1; r.w[2] = c2; r.w[3] = c3;
    return r;
}

// The same block with the ten round keys precomputed (kernel parameters: the xor takes them as constant-bank
// operands, which removes the twenty key additions per block from the instruction stream).
struct PhiloxKeys { uint32_t k[20]; };

JMM_HD PhiloxKeys philox_keys(uint32_t k0, uint32_t k1) {
    PhiloxKeys K;
    for (int round = 0; round < 10; ++round) { K.k[2 * round] = k0; K.k[2 * round + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return K;
}

JMM_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &K) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ K.k[2 * round], n2 = hi0 ^ c3 ^ K.k[2 * round + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    Philox4 r; r.w[0] = c0; r.w[1] = c1; r.w[2] = c2; r.w[3] = c3;
    return r;
}

// stream tags in counter word 3 (keeps the three uses of one (seed) key disjoint)
constexpr uint32_t kTagTrial = 0u;            // (step, chain): one block per Step()
constexpr uint32_t kTagParticle = 0x80000000u; // | chain: (half-sweep, particle) in checkerboard mode
constexpr uint32_t kTagColour = 0x40000000u;   // | chain: colour of a half-sweep

JMM_HD uint32_t taus2_next(uint32_t &s1, uint32_t &s2, uint32_t &s3) {
    s1 = ((s1 & 4294967294u) << 12) ^ (((s1 << 13) ^ s1) >> 19);
    s2 = ((s2 & 4294967288u) << 4) ^ (((s2 << 2) ^ s2) >> 25);
    s3 = ((s3 & 4294967280u) << 17) ^ (((s3 << 3) ^ s3) >> 11);
    return s1 ^ s2 ^ s3;
}

JMM_HD void taus2_seed(uint64_t seed, uint32_t &s1, uint32_t &s2, uint32_t &s3) {
    if (seed == 0) seed = 1;
    const uint32_t s = (uint32_t) seed;
    s1 = 69069u * s;  if (s1 < 2) s1 += 2;
    s2 = 69069u * s1; if (s2 < 8) s2 += 8;
    s3 = 69069u * s2; if (s3 < 16) s3 += 16;
    for (int i = 0; i < 6; ++i) taus2_next(s1, s2, s3);
}

#if defined(__CUDACC__)
// (1 + w / 2^32) - c, exactly: the 32 random bits go straight into the mantissa of a double in [1, 2), so that
// u01(w) = u01_shifted(w, 1.0) and u01(w) - 0.5 = u01_shifted(w, 1.5) bit for bit (every intermediate is exact), for
// two integer instructions and one DADD instead of a conversion, a multiplication and the subtraction.
__device__ __forceinline__ double u01_shifted(uint32_t w, double c) {
    return __hiloint2double((int) (0x3FF00000u | (w >> 12)), (int) (w << 20)) - c;
}
#endif

// gsl_rng_uniform: x / 2^32 (exact in fp64)
JMM_HD double u01(uint32_t w) { return (double) w * (1.0 / 4294967296.0); }

}  // namespace jmm
