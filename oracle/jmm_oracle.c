/* TEST INFRASTRUCTURE — CPU restatement of the jmmOneDMC hot path (see jmm_oracle.h).
 *
 * Plain C, one thread, IEEE fp64 with no FMA contraction (-ffp-contract=off; the reference is
 * built without -march so it has none either).  Two arithmetic modes:
 *   TABLE      every pair distance lives in an incrementally updated rij table exactly as in the
 *              reference (rij +- md on a displacement, rij *= s on a qavLJ accept, r[j]-r[i] after
 *              fav/moveVolume).  The reference's 18 per-pair energy tables are NOT kept: each
 *              stored term is phi() of the stored rij under the same flags, so recomputing it from
 *              the rij table gives the identical double.  Bit-exact against the compiled reference.
 *   RECOMPUTE  distances are r[j]-r[i] from positions; O(N) memory.  This is what the GPU path
 *              computes, and the GPU parity tests compare against this mode bit for bit.
 *
 * Known, deliberate differences from the reference (all are undefined behaviour or dead paths there):
 *   - phiHarmoniccut writes only phi[0..1] (src/pot.cpp:116-131); the other seven components are
 *     uninitialised stack in the reference and are defined as 0 here.
 *   - ECheck's `resultFlag` is a file-static that is never cleared once a discrepancy was seen
 *     (src/jmmMCState.cpp:1998-2019,2086-2091); here a discrepancy resets the totals from a full
 *     recompute once, which is the evident intent.  No single-thread reference run reaches it.
 *   - relaxFlag is uninitialised when the deck has no RELAX line (src/readInput.cpp:197-199): 0 here.
 */
#include "jmm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct jmo_state {
    jmo_config cfg;
    uint64_t N, numTrialTypes, numPairs, sn;
    double l, maxStep, maxdl;
    double tot[JMO_NTOT];
    double acc[JMO_NACC];
    uint64_t dAcc[2], vAcc[2];
    uint64_t vAErrNtot;               /* static local of maxDVAdjust, src/jmmMCState.cpp:2121 */
    uint64_t echecks, discrepancies, relax_calls;
    double *r, *rTrial;
    double *rij;                      /* TABLE mode only */
    uint32_t taus[3];
    const uint32_t *rec; uint64_t nrec, irec;
    uint32_t phx[4]; int phx_valid;   /* Philox block of the current step */
    /* thermo print bookkeeping, src/jmmMCState.cpp:53-54,537 */
    uint64_t sltp;
    /* histograms, src/jmmMCState.cpp:58-63,128-132,180-192 */
    int hist;
    uint64_t rhonb, gnb, slrho, slg;
    int gns;
    double rbw, gsw, gbw;
    int64_t *rhol, *rhoA, *gl, *gA;          /* gl/gA are [gns][gnb] */
    double last_md; uint64_t last_nm;        /* qagrho reads mcs->md and mcs->nm */
};

/* ------------------------------------------------------------------ generators */

/* GNU GSL rng/taus.c (taus2 seeding); call sites src/jmmMCState.cpp:779-781 */
void jmo_taus2_seed(uint32_t st[3], uint64_t seed) {
    if (seed == 0) seed = 1;
    uint32_t s = (uint32_t) seed;
    st[0] = 69069u * s;     if (st[0] < 2)  st[0] += 2;
    st[1] = 69069u * st[0]; if (st[1] < 8)  st[1] += 8;
    st[2] = 69069u * st[1]; if (st[2] < 16) st[2] += 16;
    for (int i = 0; i < 6; i++) jmo_taus2_next(st);
}

uint32_t jmo_taus2_next(uint32_t st[3]) {
    st[0] = ((st[0] & 4294967294u) << 12) ^ (((st[0] << 13) ^ st[0]) >> 19);
    st[1] = ((st[1] & 4294967288u) <<  4) ^ (((st[1] <<  2) ^ st[1]) >> 25);
    st[2] = ((st[2] & 4294967280u) << 17) ^ (((st[2] <<  3) ^ st[2]) >> 11);
    return st[0] ^ st[1] ^ st[2];
}

/* Philox4x32-10 (Salmon et al., SC'11); not in the reference — the production stream */
void jmo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; round++) {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t raw_word(jmo_state *s) {
    if (s->cfg.rng_kind == JMO_RNG_RECORDED) {
        if (s->irec >= s->nrec) { fprintf(stderr, "jmm_oracle: recorded stream exhausted\n"); abort(); }
        return s->rec[s->irec++];
    }
    return jmo_taus2_next(s->taus);
}

static void philox_for_step(jmo_state *s) {
    uint32_t ctr[4] = { (uint32_t) s->sn, (uint32_t)(s->sn >> 32), (uint32_t) s->cfg.chain_id, 0u };
    uint32_t key[2] = { (uint32_t) s->cfg.seed, (uint32_t)(s->cfg.seed >> 32) };
    jmo_philox4x32_10(ctr, key, s->phx);
}

/* gsl_rng_uniform_int, GNU GSL rng/rng.c; call site src/jmmMCState.cpp:1762 */
static uint64_t draw_trial_type(jmo_state *s) {
    uint64_t n = s->numTrialTypes;
    if (s->cfg.rng_kind == JMO_RNG_PHILOX) {
        uint32_t scale = 0xffffffffu / (uint32_t) n;
        uint64_t k = s->phx[0] / scale;
        if (k >= n) { k = s->phx[3] / scale; if (k >= n) k = ((uint64_t) s->phx[3] * n) >> 32; }
        return k;
    }
    uint64_t scale = 0xffffffffUL / n, k;
    do { k = raw_word(s) / scale; } while (k >= n);
    return k;
}
/* gsl_rng_uniform; call sites src/jmmMCState.cpp:1763 (rn) and :1370,1668,2252 (ran) */
static double draw_rn(jmo_state *s)  { return (s->cfg.rng_kind == JMO_RNG_PHILOX ? s->phx[1] : raw_word(s)) / 4294967296.0; }
static double draw_ran(jmo_state *s) { return (s->cfg.rng_kind == JMO_RNG_PHILOX ? s->phx[2] : raw_word(s)) / 4294967296.0; }

/* ------------------------------------------------------------------ potentials */

/* src/pot.cpp:19-101 (phiLJcut), :104-108 (phiLJinfcutoff), :110-134 (phiHarmoniccut), :137-141 */
void jmo_phi(int pot, double d, double cutoff, int virflag, double l, double out[9]) {
    if (pot == JMO_POT_HARMONIC) {
        for (int k = 0; k < 9; k++) out[k] = 0;
        if (d <= 0) { out[0] = 10E10; out[1] = 10E10; }
        else if (d < cutoff) {
            double rijm = d - 1.0;
            out[0] = rijm * rijm;
            out[1] = (2 / l) * d * rijm;     /* l is only defined when the virial flag is 1 */
            if (!virflag) out[1] = 0;
        }
        return;
    }
    if (pot == JMO_POT_LJ) cutoff = INFINITY;
    double rij3 = d * d * d;
    double rij6 = 1 / (rij3 * rij3);
    double rij12 = rij6 * rij6;
    double phi6 = 0, phi12 = 0, phitot = 0, vir6 = 0, vir12 = 0, virtot = 0, hv6 = 0, hv12 = 0, hvtot = 0;
    if (d <= cutoff) {
        phi6 = 4 * rij6; phi12 = 4 * rij12; phitot = phi12 - phi6;
        if (virflag) {
            vir6 = 24 * rij6; vir12 = 48 * rij12; virtot = vir12 - vir6;
            hv6 = 144 * rij6; hv12 = 576 * rij12; hvtot = hv12 - hv6;
        }
    }
    out[0] = phitot; out[2] = phi12; out[4] = phi6;
    out[1] = virtot; out[3] = vir12; out[5] = vir6;
    out[6] = hvtot;  out[7] = hv12;  out[8] = hv6;
}

static inline uint64_t pair_index(uint64_t N, uint64_t i, uint64_t j) {   /* src/jmmMCState.cpp:142-148 */
    return i * (N - 1) - (i * (i + 1)) / 2 + j - 1;
}
static inline int within_nbn(int nbn, uint64_t i, uint64_t j) {           /* :919, :1217, :2200 */
    return nbn < 0 || (j - i) <= (uint64_t) nbn;
}

/* the §3.3 loop: for ind < numPairs in pair-index order, skip |j-i|>NBN, sum phi(r[j]-r[i]) */
void jmo_totals_of(const double *r, uint64_t N, int nbn, int pot, double cutoff,
                   double scale, int virflag, double l_for_vir, double out[9]) {
    double phi[9];
    for (int k = 0; k < 9; k++) out[k] = 0;
    for (uint64_t i = 0; i + 1 < N; i++) {
        uint64_t jmax = (nbn < 0 || i + (uint64_t) nbn > N - 1) ? N - 1 : i + (uint64_t) nbn;
        for (uint64_t j = i + 1; j <= jmax; j++) {
            double d = (scale == 1.0) ? r[j] - r[i] : r[j] * scale - r[i] * scale;
            jmo_phi(pot, d, cutoff, virflag, l_for_vir, phi);
            for (int k = 0; k < 9; k++) out[k] += phi[k];
        }
    }
}

void jmo_config_totals(const jmo_state *s, double scale, int virflag, double l_for_vir, double out[9]) {
    jmo_totals_of(s->r, s->N, s->cfg.nbn, s->cfg.pot, s->cfg.cutoff, scale, virflag, l_for_vir, out);
}

/* ------------------------------------------------------------------ set-up */

/* setupMCS, src/jmmMCState.cpp:261-790 (state, no files) */
jmo_state *jmo_create(const jmo_config *cfg) {
    jmo_state *s = (jmo_state *) calloc(1, sizeof *s);
    s->cfg = *cfg;
    s->N = cfg->N;
    s->numPairs = ((s->N - 1) * s->N) / 2;                                   /* :439 */
    if (cfg->ensemble == JMO_ENS_NPT) { s->numTrialTypes = s->N + 1; s->l = (double) s->N; }   /* :405,553 */
    else { s->numTrialTypes = s->N; s->l = cfg->L; s->cfg.relax = 0; }       /* :414-420 */
    s->maxStep = cfg->maxStep; s->maxdl = cfg->maxdl;
    s->tot[JMO_E] = s->tot[JMO_VIR] = s->tot[JMO_HV] = 10E10;                /* :542-544 */
    s->tot[JMO_E6] = s->tot[JMO_VIR6] = s->tot[JMO_HV6] = -5E10;             /* :302-307 */
    s->tot[JMO_E12] = s->tot[JMO_VIR12] = s->tot[JMO_HV12] = 5E10;
    s->r = (double *) malloc(s->N * sizeof(double));
    s->rTrial = (double *) malloc(s->N * sizeof(double));
    for (uint64_t i = 0; i < s->N; i++) s->r[i] = ((i + 0.5) / s->N - 0.5) * s->l;  /* :561 */
    if (cfg->mode == JMO_MODE_TABLE) {
        s->rij = (double *) malloc((s->numPairs ? s->numPairs : 1) * sizeof(double));
        for (uint64_t i = 0; i + 1 < s->N; i++)
            for (uint64_t j = i + 1; j < s->N; j++) s->rij[pair_index(s->N, i, j)] = s->r[j] - s->r[i];  /* :768 */
    }
    s->sltp = (uint64_t) -1;                                                 /* :537 */
    jmo_taus2_seed(s->taus, cfg->seed);                                      /* :779-781 */
    return s;
}

void jmo_destroy(jmo_state *s) {
    if (!s) return;
    free(s->r); free(s->rTrial); free(s->rij); free(s->rhol); free(s->rhoA); free(s->gl); free(s->gA); free(s);
}
void jmo_set_recorded(jmo_state *s, const uint32_t *w, uint64_t n) { s->rec = w; s->nrec = n; s->irec = 0; }
uint64_t jmo_recorded_cursor(const jmo_state *s) { return s->irec; }

/* full sums with side effects on the table: the bodies of fad :907-946, moveVolume :2865-2904 and
 * ECheck's reset :2028-2071 — included pairs get rij = r[j]-r[i] */
static void full_recompute_into_state(jmo_state *s) {
    double phi[9], out[9] = {0};
    for (uint64_t i = 0; i + 1 < s->N; i++)
        for (uint64_t j = i + 1; j < s->N; j++) {
            if (!within_nbn(s->cfg.nbn, i, j)) continue;
            double d = s->r[j] - s->r[i];
            if (s->rij) s->rij[pair_index(s->N, i, j)] = d;
            jmo_phi(s->cfg.pot, d, s->cfg.cutoff, 1, s->l, phi);
            for (int k = 0; k < 9; k++) out[k] += phi[k];
        }
    memcpy(s->tot, out, sizeof out);
}

/* fad(mcs,&0,&0.5): src/Main.cpp:66-68, src/jmmMCState.cpp:853-1003.  md = 0; ETrial < 1e11 so it is
 * always accepted (:967), counts as an accepted displacement (:968) and fills every table. */
void jmo_step0(jmo_state *s) {
    if (s->rij)
        for (uint64_t i = 0; i + 1 < s->N; i++)
            for (uint64_t j = i + 1; j < s->N; j++) s->rij[pair_index(s->N, i, j)] = s->r[j] - s->r[i];
    full_recompute_into_state(s);
    s->dAcc[0]++;
}

/* calculateEnergyOfTrialVolumeChange, src/jmmMCState.cpp:2783-2827 */
static double energy_of_trial_volume_change(jmo_state *s, double dl) {
    double lRat1 = (s->l + dl) / s->l, out[9];
    for (uint64_t i = 0; i < s->N; i++) s->rTrial[i] = s->r[i] * lRat1;
    jmo_totals_of(s->rTrial, s->N, s->cfg.nbn, s->cfg.pot, s->cfg.cutoff, 1.0, 0, 0.0, out);
    return out[JMO_E];
}

/* moveVolume, src/jmmMCState.cpp:2831-2916 */
static void move_volume(jmo_state *s, double lnew) {
    double lRat1 = lnew / s->l;
    for (uint64_t i = 0; i < s->N; i++) s->r[i] = s->r[i] * lRat1;
    s->l = lnew;
    full_recompute_into_state(s);
}

/* relaxVolume, src/jmmMCState.cpp:2396-2679 (live lines :2403-2423,2481-2488,2546-2592,2670-2677) */
int jmo_relax_volume(jmo_state *s) {
    double lTryMin = 0, lTryMax = 1E10, h, EUp, EDown, dlEstimate;
    s->relax_calls++;
    for (int count = 0; count < 20; count++) {
        h = 0.1;
        EUp = energy_of_trial_volume_change(s, h);
        EDown = energy_of_trial_volume_change(s, -h);
        double first = (EUp - EDown) / (2 * h);
        double second = (EUp - 2.0 * s->tot[JMO_E] + EDown) / (h * h);
        dlEstimate = -(s->cfg.P - ((double) s->N / s->l) * s->cfg.T + first) / second;
        double relaxMax = (double) 0.10 * s->N;
        if (fabs(dlEstimate) > relaxMax) dlEstimate = dlEstimate < 0 ? -relaxMax : relaxMax;
        if (s->l + dlEstimate > lTryMax) dlEstimate = 0.5 * (lTryMax - s->l);
        else if (s->l + dlEstimate < lTryMin) dlEstimate = 0.5 * (lTryMin - s->l);
        if (dlEstimate > 0.0) lTryMin = s->l; else lTryMax = s->l;
        double relaxCrit = 0.0025 * s->N;
        int converged = fabs(dlEstimate) < relaxCrit;
        move_volume(s, s->l + dlEstimate);
        if (converged) return 0;
    }
    return 1;
}

/* updateThermo, src/jmmMCState.cpp:1941-1961 */
void jmo_update_thermo(jmo_state *s) {
    double rhotmp = s->N / s->l, E = s->tot[JMO_E], Vir = s->tot[JMO_VIR], HV = s->tot[JMO_HV];
    s->acc[JMO_A_RHO]  = s->acc[JMO_A_RHO]  + rhotmp;
    s->acc[JMO_A_RHO2] = s->acc[JMO_A_RHO2] + rhotmp * rhotmp;
    s->acc[JMO_A_L]    = s->acc[JMO_A_L]    + s->l;
    s->acc[JMO_A_L2]   = s->acc[JMO_A_L2]   + s->l * s->l;
    s->acc[JMO_A_E]    = s->acc[JMO_A_E]    + E;
    s->acc[JMO_A_E2]   = s->acc[JMO_A_E2]   + E * E;
    s->acc[JMO_A_LE]   = s->acc[JMO_A_LE]   + s->l * E;
    s->acc[JMO_A_VIR]  = s->acc[JMO_A_VIR]  + Vir;
    s->acc[JMO_A_VIR2] = s->acc[JMO_A_VIR2] + Vir * Vir;
    s->acc[JMO_A_EVIR] = s->acc[JMO_A_EVIR] + E * Vir;
    s->acc[JMO_A_HV]   = s->acc[JMO_A_HV]   + HV;
    s->acc[JMO_A_HV2]  = s->acc[JMO_A_HV2]  + HV * HV;
}

/* ------------------------------------------------------------------ histograms */

/* fgrho, src/jmmMCState.cpp:1069-1127.  Pair distances: the rij table when there is one (the reference reads
 * mcs->rij[ind], :1103), else r[jj]-r[ii]. */
static void hist_fgrho(jmo_state *s) {
    for (uint64_t b = 0; b < s->rhonb; b++) s->rhol[b] = 0;
    for (uint64_t i = 0; i < s->N; i++) {
        long rb = (long) floor(s->r[i] / s->rbw + s->rhonb / 2.0);
        if (rb >= 0 && rb < (long) s->rhonb) s->rhol[rb]++;
    }
    for (uint64_t q = 0; q < (uint64_t) s->gns * s->gnb; q++) s->gl[q] = 0;
    for (uint64_t i = 0; i + 1 < s->N; i++)
        for (uint64_t j = i + 1; j < s->N; j++) {
            long gs1 = (long) floor(s->r[i] / s->gsw + s->gns / 2.0);
            long gs2 = (long) floor(s->r[j] / s->gsw + s->gns / 2.0);
            double d = s->rij ? s->rij[pair_index(s->N, i, j)] : s->r[j] - s->r[i];
            long gb = (long) floor(d / s->gbw);
            if (gb < (long) s->gnb && gb >= 0) {       /* the reference does not test gb >= 0 (UB for crossed particles) */
                if (gs1 >= 0 && gs1 < (long) s->gns) s->gl[gs1 * s->gnb + gb]++;
                if (gs2 >= 0 && gs2 < (long) s->gns) s->gl[gs2 * s->gnb + gb]++;
            }
        }
}

/* ugrho, :1131-1149 */
static void hist_ugrho(jmo_state *s) {
    for (uint64_t b = 0; b < s->rhonb; b++) s->rhoA[b] += s->rhol[b];
    for (uint64_t q = 0; q < (uint64_t) s->gns * s->gnb; q++) s->gA[q] += s->gl[q];
}

/* qagrho, :2297-2384: incremental update after an accepted displacement of particle nm by md.
 * The old position is re-derived as r[nm] - md (not the stored old value), as in the reference. */
static void hist_qagrho(jmo_state *s, uint64_t nm, double md) {
    const double rn = s->r[nm];
    long rbn1 = (long) floor((rn - md) / s->rbw + s->rhonb / 2.0);
    long rbn2 = (long) floor(rn / s->rbw + s->rhonb / 2.0);
    if (rbn1 >= 0 && rbn1 < (long) s->rhonb) s->rhol[rbn1]--;
    if (rbn2 >= 0 && rbn2 < (long) s->rhonb) s->rhol[rbn2]++;
    long gs11 = (long) floor((rn - md) / s->gsw + s->gns / 2.0);
    long gs12 = (long) floor(rn / s->gsw + s->gns / 2.0);
    int in11 = gs11 >= 0 && gs11 < s->gns, in12 = gs12 >= 0 && gs12 < s->gns;
    for (uint64_t i = 0; i < s->N; i++) {
        if (i == nm) continue;
        long gs2 = (long) floor(s->r[i] / s->gsw + s->gns / 2.0);
        int in2 = gs2 >= 0 && gs2 < s->gns;
        long gb1 = (long) (unsigned long) floor(fabs(rn - md - s->r[i]) / s->gbw);
        long gb2 = (long) (unsigned long) floor(fabs(rn - s->r[i]) / s->gbw);
        int b1 = gb1 >= 0 && gb1 < (long) s->gnb, b2 = gb2 >= 0 && gb2 < (long) s->gnb;
        if (in11 && b1) s->gl[gs11 * s->gnb + gb1]--;
        if (in12 && b2) s->gl[gs12 * s->gnb + gb2]++;
        if (in2 && b1) s->gl[gs2 * s->gnb + gb1]--;
        if (in2 && b2) s->gl[gs2 * s->gnb + gb2]++;
    }
}

void jmo_enable_histograms(jmo_state *s, uint64_t rhonb, double rbw, int gns, uint64_t gnb, double gsw, double gbw) {
    s->hist = 1; s->rhonb = rhonb; s->rbw = rbw; s->gns = gns; s->gnb = gnb; s->gsw = gsw; s->gbw = gbw;
    s->slrho = (uint64_t) -1; s->slg = (uint64_t) -1;                              /* :538-539 */
    s->rhol = (int64_t *) calloc(rhonb ? rhonb : 1, sizeof(int64_t));
    s->rhoA = (int64_t *) calloc(rhonb ? rhonb : 1, sizeof(int64_t));
    const uint64_t ng = (uint64_t) gns * gnb;
    s->gl = (int64_t *) calloc(ng > 0 ? ng : 1, sizeof(int64_t));
    s->gA = (int64_t *) calloc(ng > 0 ? ng : 1, sizeof(int64_t));
    hist_fgrho(s);                                                                 /* setupMCS :773-776 */
    hist_ugrho(s);
}

void jmo_take_histograms(jmo_state *s, int64_t *rhoA, int64_t *gA) {
    if (rhoA) { memcpy(rhoA, s->rhoA, s->rhonb * sizeof(int64_t)); memset(s->rhoA, 0, s->rhonb * sizeof(int64_t)); }
    if (gA) { memcpy(gA, s->gA, (uint64_t) s->gns * s->gnb * sizeof(int64_t)); memset(s->gA, 0, (uint64_t) s->gns * s->gnb * sizeof(int64_t)); }
}

/* ------------------------------------------------------------------ trial moves */

/* qad2, src/jmmMCState.cpp:1160-1464.  Left partners (ii<nm, :1213-1268) and right partners
 * (ii>nm, :1308-1348) are summed in ascending index into SEPARATE accumulators, each as
 * acc = (acc - old) + new, then added (:1354-1362). */
static int displacement_trial(jmo_state *s, uint64_t nm, double rn) {
    const uint64_t N = s->N;
    const int nbn = s->cfg.nbn, pot = s->cfg.pot;
    const double cut = s->cfg.cutoff, l = s->l;
    double md = (rn - 0.5) * 2 * s->maxStep;                         /* :1182 */
    double rT = s->r[nm] + md;                                       /* :1183 */
    if (fabs(rT) > l / 2.0) { s->dAcc[1]++; if (s->hist) hist_ugrho(s); return 0; }   /* :1188-1193, ugrho :1453 */

    double dL[9] = {0}, dR[9] = {0}, po[9], pn[9];
    uint64_t lo = (nbn < 0 || (uint64_t) nbn > nm) ? 0 : nm - (uint64_t) nbn;
    uint64_t hi = (nbn < 0 || nm + (uint64_t) nbn > N - 1) ? N - 1 : nm + (uint64_t) nbn;
    for (uint64_t i = lo; i < nm; i++) {
        double dold, dnew;
        if (s->rij) { dold = s->rij[pair_index(N, i, nm)]; dnew = dold + md; }   /* :1216 */
        else        { dold = s->r[nm] - s->r[i];           dnew = rT - s->r[i]; }
        jmo_phi(pot, dold, cut, 1, l, po);
        jmo_phi(pot, dnew, cut, 1, l, pn);
        for (int k = 0; k < 9; k++) dL[k] = dL[k] - po[k] + pn[k];               /* :1244-1267 */
    }
    for (uint64_t j = nm + 1; j <= hi && j < N; j++) {
        double dold, dnew;
        if (s->rij) { dold = s->rij[pair_index(N, nm, j)]; dnew = dold - md; }   /* :1311 */
        else        { dold = s->r[j] - s->r[nm];           dnew = s->r[j] - rT; }
        jmo_phi(pot, dold, cut, 1, l, po);
        jmo_phi(pot, dnew, cut, 1, l, pn);
        for (int k = 0; k < 9; k++) dR[k] = dR[k] - po[k] + pn[k];               /* :1339-1347 */
    }
    double d[9];
    for (int k = 0; k < 9; k++) d[k] = dL[k] + dR[k];                            /* :1354-1362 */

    int accept = d[JMO_E] <= 0;
    if (!accept) {                                                               /* :1367-1377 */
        double ran = draw_ran(s);
        double bf = exp(-d[JMO_E] / s->cfg.T);
        accept = bf > ran;
    }
    if (accept) {                                                                /* :1384-1428 */
        s->dAcc[0]++;
        for (int k = 0; k < 9; k++) s->tot[k] += d[k];
        s->r[nm] = rT;
        if (s->rij) {
            /* the reference copies rijTrial for every partner, interacting or not */
            for (uint64_t i = 0; i < nm; i++)     s->rij[pair_index(N, i, nm)] += md;
            for (uint64_t j = nm + 1; j < N; j++) s->rij[pair_index(N, nm, j)] -= md;
        }
        if (s->hist) hist_qagrho(s, nm, md);                                     /* :1431 */
    } else s->dAcc[1]++;                                                         /* :1447 */
    if (s->hist) hist_ugrho(s);                                                  /* :1453 */
    return accept;
}

/* shared acceptance rule of qavLJ :1665-1672 and fav :2249-2255 */
static int volume_accept(jmo_state *s, double dE, double dl, double lRat1) {
    double bf = exp(-(dE + s->cfg.P * dl) / s->cfg.T + s->N * log(lRat1));
    double ran = 0;
    if (bf < 1.0) ran = draw_ran(s);
    return bf >= 1.0 || bf > ran;
}

/* qavLJ, src/jmmMCState.cpp:1648-1730: un-truncated LJ scales as s^-12, s^-6 */
static int volume_trial_LJ(jmo_state *s, double rn) {
    double dl = (rn - 0.5) * 2 * s->maxdl;
    double lRat1 = (s->l + dl) / s->l;
    double lRat3 = lRat1 * lRat1 * lRat1;
    double lRat6 = 1 / (lRat3 * lRat3);
    double lRat12 = lRat6 * lRat6;
    double E12Trial = lRat12 * s->tot[JMO_E12];
    double E6Trial = lRat6 * s->tot[JMO_E6];
    double dE = E12Trial - E6Trial - s->tot[JMO_E];
    double bf = exp(-(dE + s->cfg.P * dl) / s->cfg.T + s->N * log(lRat1));
    double ran = 0;
    if (bf < 1.0) ran = draw_ran(s);
    if (bf >= 1.0 || bf > ran) {
        s->vAcc[0]++;
        s->tot[JMO_E] = s->tot[JMO_E] + dE;
        s->tot[JMO_E12] = E12Trial;
        s->tot[JMO_E6] = E6Trial;
        s->l = s->l + dl;
        double lRat7 = lRat6 / lRat1, lRat13 = lRat12 / lRat1;
        s->tot[JMO_VIR6] = lRat7 * s->tot[JMO_VIR6];
        s->tot[JMO_VIR12] = lRat13 * s->tot[JMO_VIR12];
        s->tot[JMO_VIR] = s->N * s->cfg.T / s->l + s->tot[JMO_VIR12] - s->tot[JMO_VIR6];   /* :1686 */
        /* HV, HV12, HV6 are left untouched by the reference (:1675-1686) */
        for (uint64_t i = 0; i < s->N; i++) s->r[i] = lRat1 * s->r[i];                     /* :1692 */
        if (s->rij) for (uint64_t p = 0; p < s->numPairs; p++) s->rij[p] = lRat1 * s->rij[p];  /* :1699 */
        if (s->hist) { hist_fgrho(s); hist_ugrho(s); }                                     /* :1717, :1726 */
        return 1;
    }
    s->vAcc[1]++;
    if (s->hist) hist_ugrho(s);                                                            /* :1726 */
    return 0;
}

/* fav, src/jmmMCState.cpp:2161-2293: volume trial by full recompute */
static int volume_trial_full(jmo_state *s, double rn) {
    double dl = (rn - 0.5) * 2 * s->maxdl;
    double lnew = s->l + dl;
    double lRat1 = lnew / s->l;
    double t[9];
    for (uint64_t i = 0; i < s->N; i++) s->rTrial[i] = s->r[i] * lRat1;
    jmo_totals_of(s->rTrial, s->N, s->cfg.nbn, s->cfg.pot, s->cfg.cutoff, 1.0, 1, lnew, t);
    if (volume_accept(s, t[JMO_E] - s->tot[JMO_E], dl, lRat1)) {
        s->vAcc[0]++;
        s->l = s->l + dl;
        memcpy(s->tot, t, sizeof t);
        memcpy(s->r, s->rTrial, s->N * sizeof(double));
        if (s->rij)
            for (uint64_t i = 0; i + 1 < s->N; i++)
                for (uint64_t j = i + 1; j < s->N; j++) s->rij[pair_index(s->N, i, j)] = s->r[j] - s->r[i];  /* :2273 */
        return 1;
    }
    s->vAcc[1]++;
    return 0;
}

/* ECheck, src/jmmMCState.cpp:1965-2095 */
static void energy_check(jmo_state *s) {
    double t[9];
    jmo_config_totals(s, 1.0, 0, 0.0, t);
    s->echecks++;
    if (fabs(t[JMO_E] - s->tot[JMO_E]) > 0.0001) {
        s->discrepancies++;
        full_recompute_into_state(s);
    }
}

/* incrementStep :1734-1754 + Step :1758-1811 */
int jmo_step(jmo_state *s) {
    s->sn++;
    if (s->cfg.rng_kind == JMO_RNG_PHILOX) philox_for_step(s);
    uint64_t nm = draw_trial_type(s);                                  /* :1762 */
    double rn = draw_rn(s);                                            /* :1763 */
    int acc;
    if (nm < s->N) acc = displacement_trial(s, nm, rn);                /* :1783-1785 */
    else if (s->cfg.pot == JMO_POT_LJ && s->cfg.nbn < 0) acc = volume_trial_LJ(s, rn);   /* :296-301 */
    else acc = volume_trial_full(s, rn);
    if (s->cfg.eci && s->sn % s->cfg.eci == 0) energy_check(s);        /* :1800-1802,1858 */
    jmo_update_thermo(s);                                              /* :1805 */
    return acc;
}

static void jmo_cadence_adjust_only(jmo_state *s);
/* maxDisAdjust :2100-2115, maxDVAdjust :2120-2139, periodic relax src/Main.cpp:173-176 */
void jmo_cadence(jmo_state *s) {
    jmo_cadence_adjust_only(s);
    if (s->sn % 10000 == 0 && s->sn < 1E6 && s->cfg.relax > 0) jmo_relax_volume(s);
}
static void jmo_cadence_adjust_only(jmo_state *s) {
    if (s->cfg.mdai && s->sn % s->cfg.mdai == 0) {
        double idealRatio = 0.5;
        double actualRatio = (double) s->dAcc[0] / (s->dAcc[0] + s->dAcc[1]);
        s->maxStep = s->maxStep * log(0.672924 * idealRatio + 0.0644284) / log(0.672924 * (actualRatio + 0.0644284));
        if (s->maxStep < 0.002) s->maxStep = 0.002;
        else if (s->maxStep > 0.5) s->maxStep = 0.5;
    }
    if (s->cfg.mvai && s->sn % s->cfg.mvai == 0) {
        if ((s->vAcc[0] + s->vAcc[1] - s->vAErrNtot) > 0) {
            double idealRatio = 0.5;
            s->vAErrNtot = s->vAcc[0] + s->vAcc[1];
            double actualRatio = (double) s->vAcc[0] / (s->vAcc[0] + s->vAcc[1]);
            s->maxdl = s->maxdl * log(0.672924 * idealRatio + 0.0644284) / log(0.672924 * (actualRatio + 0.0644284));
            if (s->maxdl < 0.002 * s->N) s->maxdl = 0.002 * s->N;
            else if (s->maxdl > 0.10 * s->N) s->maxdl = 0.50 * s->N;
        }
    }
}

void jmo_run(jmo_state *s, uint64_t nsteps) {
    for (uint64_t k = 0; k < nsteps; k++) { jmo_step(s); jmo_cadence(s); }
}

/* printThermo :1896-1937 and printCoords :1007-1017, byte-compatible */
static void print_thermo(jmo_state *s, FILE *tf, FILE *log) {
    uint64_t ss = s->sn - s->sltp;
    const double *a = s->acc;
    if (tf) fprintf(tf, "%lu\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\n",
            (unsigned long) s->sn, a[JMO_A_E] / ss, a[JMO_A_E2] / ss, a[JMO_A_L] / ss, a[JMO_A_L2] / ss,
            a[JMO_A_LE] / ss, a[JMO_A_RHO] / ss, a[JMO_A_RHO2] / ss, a[JMO_A_VIR] / ss, a[JMO_A_VIR2] / ss,
            a[JMO_A_EVIR] / ss, a[JMO_A_HV] / ss, a[JMO_A_HV2] / ss);
    if (log) fprintf(log, "%lu  %.8G  %.8G  %.8G  %.8G\n", (unsigned long) s->sn, s->tot[JMO_E], s->l,
                     s->tot[JMO_VIR], s->tot[JMO_HV]);
    memset(s->acc, 0, sizeof s->acc);
    s->sltp = s->sn;
}
static void print_coords(jmo_state *s, FILE *cf) {
    if (!cf) return;
    fprintf(cf, "%lu\nStep no.: %lu  Box length: %.5f\n", (unsigned long) s->N, (unsigned long) s->sn, s->l);
    for (uint64_t i = 0; i < s->N; i++) fprintf(cf, "%lu  0.0  0.0  %.8G\n", (unsigned long)(i + 1), s->r[i]);
}

/* src/Main.cpp:66-180 without the histogram files */
void jmo_run_deck(jmo_state *s, uint64_t numsteps, uint64_t tpi, uint64_t cpi,
                  FILE *thermo, FILE *config, FILE *log) {
    if (thermo) fprintf(thermo, "Step    Econf           Econf2          L       L2  "
                                "    LEconf          rho             rho2            Virial      "
                                "   Virial2         EconfVir        HV              HV2 \n");   /* src/jmmMCState.cpp:566-568 */
    jmo_step0(s);
    if (s->cfg.relax > 0) jmo_relax_volume(s);
    print_coords(s, config);
    jmo_update_thermo(s);
    print_thermo(s, thermo, log);
    while (s->sn != numsteps) {
        jmo_step(s);
        if (cpi && s->sn % cpi == 0) print_coords(s, config);
        if (tpi && s->sn % tpi == 0) print_thermo(s, thermo, log);
        jmo_cadence(s);
    }
    if (log) {
        fprintf(log, "\nE = %.8G\n", s->tot[JMO_E]);
        fprintf(log, "Accepted/Rejected: %lu/%lu %lu/%lu\n", (unsigned long) s->dAcc[0], (unsigned long) s->dAcc[1],
                (unsigned long) s->vAcc[0], (unsigned long) s->vAcc[1]);
    }
}

/* printRho :1021-1038 and printG :1042-1064, byte-compatible */
static void print_rho(jmo_state *s, FILE *f) {
    uint64_t ns = s->sn - s->slrho;
    if (f) fprintf(f, "%lu", (unsigned long) s->sn);
    for (uint64_t b = 0; b < s->rhonb; b++) {
        double m = (double) (int) s->rhoA[b] / ns / s->rbw;
        if (f) fprintf(f, " %.8G", m);
        s->rhoA[b] = 0;
    }
    if (f) fprintf(f, "\n");
    s->slrho = s->sn;
}
static void print_g(jmo_state *s, FILE **gf) {
    uint64_t ns = s->sn - s->slg;
    for (int k = 0; k < s->gns; k++) {
        FILE *f = gf ? gf[k] : NULL;
        if (f) fprintf(f, "%lu", (unsigned long) s->sn);
        for (uint64_t b = 0; b < s->gnb; b++) {
            double m = (double) (int) s->gA[k * s->gnb + b] / ns / s->gsw / s->gbw;
            if (f) fprintf(f, " %.8G", m);
            s->gA[k * s->gnb + b] = 0;
        }
        if (f) fprintf(f, "\n");
    }
    s->slg = s->sn;
}

/* src/Main.cpp:66-180 with all four kinds of output */
void jmo_run_deck_hist(jmo_state *s, uint64_t numsteps, uint64_t tpi, uint64_t cpi, uint64_t rhopi, uint64_t gpi,
                       FILE *thermo, FILE *config, FILE *rho, FILE **gfiles, FILE *log) {
    if (thermo) fprintf(thermo, "Step    Econf           Econf2          L       L2  "
                                "    LEconf          rho             rho2            Virial      "
                                "   Virial2         EconfVir        HV              HV2 \n");
    jmo_step0(s);
    if (s->cfg.relax > 0) jmo_relax_volume(s);
    print_coords(s, config);
    print_rho(s, rho);
    jmo_update_thermo(s);
    print_thermo(s, thermo, log);
    print_g(s, gfiles);
    while (s->sn != numsteps) {
        jmo_step(s);
        if (cpi && s->sn % cpi == 0) print_coords(s, config);
        if (tpi && s->sn % tpi == 0) print_thermo(s, thermo, log);
        if (rhopi && s->sn % rhopi == 0) print_rho(s, rho);
        jmo_cadence_adjust_only(s);
        if (gpi && s->sn % gpi == 0) print_g(s, gfiles);
        if (s->sn % 10000 == 0 && s->sn < 1E6 && s->cfg.relax > 0) jmo_relax_volume(s);
    }
}

/* ------------------------------------------------------------------ accessors */
uint64_t jmo_N(const jmo_state *s) { return s->N; }
uint64_t jmo_sn(const jmo_state *s) { return s->sn; }
double jmo_l(const jmo_state *s) { return s->l; }
void jmo_get_r(const jmo_state *s, double *r) { memcpy(r, s->r, s->N * sizeof(double)); }
void jmo_set_r(jmo_state *s, const double *r, double l) {
    memcpy(s->r, r, s->N * sizeof(double));
    s->l = l;
    if (s->rij)
        for (uint64_t i = 0; i + 1 < s->N; i++)
            for (uint64_t j = i + 1; j < s->N; j++) s->rij[pair_index(s->N, i, j)] = s->r[j] - s->r[i];
    full_recompute_into_state(s);
}
/* continue at a given step number: what the restart branch of setupMCS does with the step count of the last frame
 * (src/jmmMCState.cpp:641); the print/histogram marks follow as at :692,711,739 */
void jmo_set_sn(jmo_state *s, uint64_t sn) { s->sn = sn; s->sltp = sn; s->slrho = sn; s->slg = sn; }
void jmo_get_totals(const jmo_state *s, double out[9]) { memcpy(out, s->tot, sizeof s->tot); }
void jmo_get_accum(const jmo_state *s, double out[12]) { memcpy(out, s->acc, sizeof s->acc); }
void jmo_zero_accum(jmo_state *s) { memset(s->acc, 0, sizeof s->acc); }
void jmo_get_counters(const jmo_state *s, uint64_t out[4]) {
    out[0] = s->dAcc[0]; out[1] = s->dAcc[1]; out[2] = s->vAcc[0]; out[3] = s->vAcc[1];
}
void jmo_get_step_sizes(const jmo_state *s, double *a, double *b) { *a = s->maxStep; *b = s->maxdl; }
void jmo_set_step_sizes(jmo_state *s, double a, double b) { s->maxStep = a; s->maxdl = b; }
uint64_t jmo_echeck_count(const jmo_state *s, uint64_t *disc) { if (disc) *disc = s->discrepancies; return s->echecks; }
uint64_t jmo_relax_calls(const jmo_state *s) { return s->relax_calls; }

/* ------------------------------------------------------------------ large-chain checkerboard
 * Not in the reference (it moves one particle per step and cannot allocate N > ~1e4).  With NBN = k,
 * particles i and j interact iff |i-j| <= k (src/jmmMCState.cpp:1217,1312), so all particles of one
 * colour i mod (k+1) are mutually independent and their qad2 trials commute.  Each trial is exactly
 * displacement_trial() in RECOMPUTE mode; the Philox block is keyed per (sweep step, pair of consecutive same-colour
 * particles): all four words of a block are used. */
int jmo_colour_of_step(uint64_t seed, uint64_t chain_id, uint64_t sweep_step, int ncolours) {
    uint32_t ctr[4] = { (uint32_t) sweep_step, (uint32_t)(sweep_step >> 32), 0xFFFFFFFFu, 0x40000000u | (uint32_t) chain_id };
    uint32_t key[2] = { (uint32_t) seed, (uint32_t)(seed >> 32) }, w[4];
    jmo_philox4x32_10(ctr, key, w);
    return (int)(((uint64_t) w[0] * (uint64_t) ncolours) >> 32);
}

uint64_t jmo_colour_halfsweep(double *r, uint64_t N, double l, int nbn, int pot, double cutoff,
                              double T, double maxStep, uint64_t seed, uint64_t chain_id,
                              uint64_t sweep_step, int ncolours, int colour, double dtot[9]) {
    uint64_t accepted = 0;
    uint32_t key[2] = { (uint32_t) seed, (uint32_t)(seed >> 32) };
    for (int k = 0; k < 9; k++) dtot[k] = 0;
    for (uint64_t nm = (uint64_t) colour; nm < N; nm += (uint64_t) ncolours) {
        /* one Philox block serves two trials: trial j = nm / ncolours of this half-sweep (nm = colour + j ncolours)
         * takes words 0,1 (j even) or 2,3 (j odd) of the block keyed by (sweep step, j / 2) */
        uint64_t j = nm / (uint64_t) ncolours;
        uint32_t ctr[4] = { (uint32_t) sweep_step, (uint32_t)(sweep_step >> 32), (uint32_t)(j >> 1), 0x80000000u | (uint32_t) chain_id }, w[4];
        jmo_philox4x32_10(ctr, key, w);
        double rn = w[2 * (j & 1)] / 4294967296.0, ran = w[2 * (j & 1) + 1] / 4294967296.0;
        double md = (rn - 0.5) * 2 * maxStep;
        double rT = r[nm] + md;
        if (fabs(rT) > l / 2.0) continue;
        double dL[9] = {0}, dR[9] = {0}, po[9], pn[9];
        uint64_t lo = (nbn < 0 || (uint64_t) nbn > nm) ? 0 : nm - (uint64_t) nbn;
        uint64_t hi = (nbn < 0 || nm + (uint64_t) nbn > N - 1) ? N - 1 : nm + (uint64_t) nbn;
        for (uint64_t i = lo; i < nm; i++) {
            jmo_phi(pot, r[nm] - r[i], cutoff, 1, l, po);
            jmo_phi(pot, rT - r[i], cutoff, 1, l, pn);
            for (int k = 0; k < 9; k++) dL[k] = dL[k] - po[k] + pn[k];
        }
        for (uint64_t j = nm + 1; j <= hi; j++) {
            jmo_phi(pot, r[j] - r[nm], cutoff, 1, l, po);
            jmo_phi(pot, r[j] - rT, cutoff, 1, l, pn);
            for (int k = 0; k < 9; k++) dR[k] = dR[k] - po[k] + pn[k];
        }
        double dE = dL[0] + dR[0];
        if (dE <= 0 || exp(-dE / T) > ran) {
            r[nm] = rT;
            accepted++;
            for (int k = 0; k < 9; k++) dtot[k] += dL[k] + dR[k];
        }
    }
    return accepted;
}
