/* TEST INFRASTRUCTURE — CPU restatement of the jmmOneDMC hot path.  Not product code.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  Every function cites the reference lines it restates (paths under /root/reference).
 *
 * Pinning: the reference ships no golden vectors, so this restatement is pinned
 * against outputs of the reference itself, compiled here by oracle/Makefile (`make ref`) and run
 * with OMP_NUM_THREADS=1: tests/golden/ holds those outputs and tests/golden/make_golden.py is
 * the script that produced them.  tests/test_oracle_vs_reference.py is the pin.
 */
#ifndef JMM_ORACLE_H
#define JMM_ORACLE_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* potentials: selection by POT string, src/jmmMCState.cpp:292-399 */
enum { JMO_POT_LJ = 0, JMO_POT_LJCUT = 1, JMO_POT_HARMONIC = 2 };
/* ensembles: src/jmmMCState.cpp:402-425 */
enum { JMO_ENS_NPT = 0, JMO_ENS_NLT = 1 };
/* random streams */
enum { JMO_RNG_TAUS2 = 0,      /* gsl_rng_taus2 restated, same draws as the reference            */
       JMO_RNG_PHILOX = 1,     /* Philox4x32-10, one block per step: production stream            */
       JMO_RNG_RECORDED = 2 }; /* raw u32 words recorded from the reference (JMM_RNG_LOG)          */
/* how pair distances are obtained */
enum { JMO_MODE_TABLE = 0,     /* incremental rij table, exactly the reference's arithmetic      */
       JMO_MODE_RECOMPUTE = 1  /* r[j]-r[i] from positions, O(N) memory: what the GPU does        */ };

/* order of the nine totals everywhere: the order phi() writes them, src/pot.cpp:90-100 */
enum { JMO_E = 0, JMO_VIR, JMO_E12, JMO_VIR12, JMO_E6, JMO_VIR6, JMO_HV, JMO_HV12, JMO_HV6, JMO_NTOT };
/* order of the twelve running sums: updateThermo, src/jmmMCState.cpp:1941-1961 */
enum { JMO_A_RHO = 0, JMO_A_RHO2, JMO_A_L, JMO_A_L2, JMO_A_E, JMO_A_E2, JMO_A_LE,
       JMO_A_VIR, JMO_A_VIR2, JMO_A_EVIR, JMO_A_HV, JMO_A_HV2, JMO_NACC };

typedef struct jmo_config {
    uint64_t N;
    int32_t  nbn;        /* NBN: index-distance neighbour limit, <0 = none                       */
    int32_t  pot;
    double   cutoff;     /* INFINITY for LJ / no third POT token                                 */
    int32_t  ensemble;
    int32_t  relax;      /* RELAX flag (ignored under NLT, src/jmmMCState.cpp:416-420)           */
    double   P, T, L;
    double   maxStep, maxdl;
    uint64_t eci, mdai, mvai;  /* ENGCHECK, DADJ, VADJ intervals; 0 = never                      */
    uint64_t seed;
    uint64_t chain_id;   /* Philox subsequence                                                   */
    int32_t  rng_kind;
    int32_t  mode;
} jmo_config;

typedef struct jmo_state jmo_state;

jmo_state *jmo_create(const jmo_config *cfg);
void       jmo_destroy(jmo_state *s);
void       jmo_set_recorded(jmo_state *s, const uint32_t *words, uint64_t nwords);
uint64_t   jmo_recorded_cursor(const jmo_state *s);

/* pair potential, src/pot.cpp:19-141 */
void jmo_phi(int pot, double d, double cutoff, int virflag, double l, double out[9]);

/* the pieces of the reference's run, in the order Main.cpp calls them */
void jmo_step0(jmo_state *s);                 /* fad(0, 0.5), src/Main.cpp:66-68                 */
int  jmo_relax_volume(jmo_state *s);          /* src/jmmMCState.cpp:2396-2679                    */
void jmo_update_thermo(jmo_state *s);         /* src/jmmMCState.cpp:1941-1961                    */
int  jmo_step(jmo_state *s);                  /* incrementStep + Step; returns accept flag       */
void jmo_cadence(jmo_state *s);               /* maxDisAdjust, maxDVAdjust, periodic relax       */
void jmo_run(jmo_state *s, uint64_t nsteps);  /* nsteps x (jmo_step + jmo_cadence)               */

/* full run with the reference's file output (thermo.dat.mcs, config.dat.mcs; either may be NULL) */
void jmo_run_deck(jmo_state *s, uint64_t numsteps, uint64_t tpi, uint64_t cpi,
                  FILE *thermo, FILE *config, FILE *log);

/* configuration energy from positions: the loop shared by fad/fav/ECheck/moveVolume, SURVEY §3.3 */
void jmo_config_totals(const jmo_state *s, double scale, int virflag, double l_for_vir, double out[9]);
/* stand-alone version for arbitrary positions */
void jmo_totals_of(const double *r, uint64_t N, int nbn, int pot, double cutoff,
                   double scale, int virflag, double l_for_vir, double out[9]);

/* getters / setters */
uint64_t jmo_N(const jmo_state *s);
uint64_t jmo_sn(const jmo_state *s);
double   jmo_l(const jmo_state *s);
void     jmo_get_r(const jmo_state *s, double *r);
void     jmo_set_r(jmo_state *s, const double *r, double l);   /* rebuilds rij table and totals  */
void     jmo_get_totals(const jmo_state *s, double out[9]);
void     jmo_get_accum(const jmo_state *s, double out[12]);
void     jmo_zero_accum(jmo_state *s);
void     jmo_get_counters(const jmo_state *s, uint64_t out[4]); /* dAcc0,dAcc1,vAcc0,vAcc1      */
void     jmo_get_step_sizes(const jmo_state *s, double *maxStep, double *maxdl);
void     jmo_set_step_sizes(jmo_state *s, double maxStep, double maxdl);
uint64_t jmo_echeck_count(const jmo_state *s, uint64_t *discrepancies);
uint64_t jmo_relax_calls(const jmo_state *s);

/* raw generators, for known-answer tests */
void     jmo_taus2_seed(uint32_t st[3], uint64_t seed);
uint32_t jmo_taus2_next(uint32_t st[3]);
void     jmo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* density / two-particle-density histograms: fgrho :1069-1127, qagrho :2297-2384,
 * ugrho :1131-1149, printRho :1021-1038, printG :1042-1064.  Enabled per chain; the restatement keeps the
 * reference's add-the-whole-histogram-every-step accumulation. */
void jmo_enable_histograms(jmo_state *s, uint64_t rhonb, double rbw, int gns, uint64_t gnb, double gsw, double gbw);
/* accumulated counts since the last call (then zeroed, like printRho/printG); either pointer may be NULL */
void jmo_take_histograms(jmo_state *s, int64_t *rhoA /*[rhonb]*/, int64_t *gA /*[gns][gnb]*/);
/* full run with all four kinds of output files; gfiles[k] may be NULL */
void jmo_run_deck_hist(jmo_state *s, uint64_t numsteps, uint64_t tpi, uint64_t cpi, uint64_t rhopi, uint64_t gpi,
                       FILE *thermo, FILE *config, FILE *rho, FILE **gfiles, FILE *log);

/* large-chain checkerboard sweep (configs C3/C5): one colour half-sweep over every particle of
 * colour `colour` (index mod ncolours), recompute mode, Philox keyed by (seed, chain, sweep step,
 * particle).  Returns the number of accepted moves; dtot[9] receives the summed deltas. */
uint64_t jmo_colour_halfsweep(double *r, uint64_t N, double l, int nbn, int pot, double cutoff,
                              double T, double maxStep, uint64_t seed, uint64_t chain_id,
                              uint64_t sweep_step, int ncolours, int colour, double dtot[9]);
void     jmo_set_sn(jmo_state *s, uint64_t sn);
int      jmo_colour_of_step(uint64_t seed, uint64_t chain_id, uint64_t sweep_step, int ncolours);

#ifdef __cplusplus
}
#endif
#endif
