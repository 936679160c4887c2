/* TEST INFRASTRUCTURE — not product code.
 *
 * Minimal stand-in for GNU GSL's <gsl/gsl_rng.h>, so that the UNMODIFIED reference
 * (/root/reference/src/jmmMCState.cpp:779,781,862,868,963,1171,1177,1370,1668,1762,1763,2252)
 * can be compiled in a container that has no libgsl.  Only the five symbols the reference
 * uses exist.  Third-party module restated here: GNU GSL, rng/taus.c (gsl_rng_taus2) and
 * rng/rng.c (gsl_rng_uniform, gsl_rng_uniform_int); version unpinned by the reference
 * (src/Makefile:5 only says -lgsl).  The algorithm is L'Ecuyer's three-component combined
 * Tausworthe generator ("Maximally Equidistributed Combined Tausworthe Generators",
 * Math. Comp. 65 (1996) and the 1999 erratum for the seeding constraints).
 *
 * Known-answer pin: seed 1, the 10000th raw output is 2733957125 — the constant GSL's own
 * rng/test.c lists for gsl_rng_taus and gsl_rng_taus2 (checked by tests/test_oracle_rng.py).
 *
 * Recorder: when the environment variable JMM_RNG_LOG names a file, every raw 32-bit output
 * is appended to it (little-endian u32).  That is how the reference's random stream is
 * captured for the GPU lock-step mode without touching reference logic.
 */
#ifndef JMM_GSL_RNG_SHIM_H
#define JMM_GSL_RNG_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

typedef struct { int id; } gsl_rng_type;
typedef struct {
    uint32_t s1, s2, s3;
    FILE *log;
} gsl_rng;

static const gsl_rng_type jmm_shim_taus2_type = { 2 };
static const gsl_rng_type *gsl_rng_taus2 = &jmm_shim_taus2_type;

static inline unsigned long int jmm_shim_taus_step(gsl_rng *g) {
    /* one step of each component: (((s & c) << d) ^ (((s << a) ^ s) >> b)), 32-bit wrap */
    g->s1 = ((g->s1 & 4294967294u) << 12) ^ (((g->s1 << 13) ^ g->s1) >> 19);
    g->s2 = ((g->s2 & 4294967288u) <<  4) ^ (((g->s2 <<  2) ^ g->s2) >> 25);
    g->s3 = ((g->s3 & 4294967280u) << 17) ^ (((g->s3 <<  3) ^ g->s3) >> 11);
    return (unsigned long int)(g->s1 ^ g->s2 ^ g->s3);
}

static inline unsigned long int gsl_rng_get(gsl_rng *g) {
    uint32_t w = (uint32_t) jmm_shim_taus_step(g);
    if (g->log) fwrite(&w, sizeof w, 1, g->log);
    return w;
}

static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t) {
    (void) t;
    gsl_rng *g = (gsl_rng *) calloc(1, sizeof(gsl_rng));
    const char *p = getenv("JMM_RNG_LOG");
    g->log = (p && *p) ? fopen(p, "wb") : NULL;
    return g;
}

static inline void gsl_rng_set(gsl_rng *g, unsigned long int seed) {
    if (seed == 0) seed = 1;              /* GSL tests the full-width seed, then reduces mod 2^32 */
    uint32_t s = (uint32_t) seed;
    g->s1 = 69069u * s;     if (g->s1 < 2)  g->s1 += 2;
    g->s2 = 69069u * g->s1; if (g->s2 < 8)  g->s2 += 8;
    g->s3 = 69069u * g->s2; if (g->s3 < 16) g->s3 += 16;
    for (int i = 0; i < 6; i++) jmm_shim_taus_step(g);   /* warm-up outputs are discarded, not logged */
}

static inline double gsl_rng_uniform(gsl_rng *g) {
    return gsl_rng_get(g) / 4294967296.0;
}

static inline unsigned long int gsl_rng_uniform_int(gsl_rng *g, unsigned long int n) {
    unsigned long int scale = 0xffffffffUL / n, k;
    do { k = gsl_rng_get(g) / scale; } while (k >= n);
    return k;
}

static inline void gsl_rng_free(gsl_rng *g) {
    if (g && g->log) fclose(g->log);
    free(g);
}

#endif
