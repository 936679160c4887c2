/* jmm_gpu.h — C ABI of the B200-native jmmOneDMC hot path (libjmmgpu.so).
 *
 * The reference (mmansell7/jmmOneDMC) has no plugin/FFI interface.  Its hot path is reached through
 *   (1) three function-pointer slots inside `struct MCState`   (src/jmmMCState.cpp:198-206), filled
 *       from the POT string by setupMCS                         (src/jmmMCState.cpp:292-399), and
 *   (2) the opaque-handle free functions of src/jmmMCState.h:7-48 that src/Main.cpp:66-180 calls.
 * Every entry point below names the reference function(s) it replaces.  Plain pointers and sizes
 * only; no C++ or torch types cross this boundary; no exception leaves the library.
 *
 * Ownership: host buffers belong to the caller, device buffers to the handle.  All per-chain host
 * arrays are chain-major: r is [nchains][N], totals [nchains][9], accumulators [nchains][12],
 * counters [nchains][4].  One host thread per handle; calls are ordered on the handle's stream and
 * every call that returns data to the host synchronises that stream before returning.
 *
 * Status: 0 = JMM_OK, negative = error (jmm_last_error() has the text).  A missing/failed CUDA
 * device is an error, never a CPU fallback.
 */
#ifndef JMM_GPU_H
#define JMM_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t jmm_status;
enum {
    JMM_OK = 0,
    JMM_ERR_INVALID = -1,      /* bad argument / unsupported combination                        */
    JMM_ERR_CUDA = -2,         /* CUDA runtime error (no device, launch failure, out of memory)  */
    JMM_ERR_IO = -3,           /* INPUT file missing or unreadable                               */
    JMM_ERR_STREAM = -4,       /* recorded random stream exhausted                               */
    JMM_ERR_UNKNOWN_POT = -5,  /* "FATAL ERROR: Unknown potential."  src/jmmMCState.cpp:396-399  */
    JMM_ERR_UNKNOWN_ENS = -6,  /* "FATAL ERROR: Unknown ensemble."   src/jmmMCState.cpp:422-425  */
    JMM_ERR_NCCL = -7          /* NCCL missing or a collective failed                            */
};

/* POT keyword -> phi / qav dispatch, src/jmmMCState.cpp:292-367 */
enum { JMM_POT_LJ = 0,         /* phiLJinfcutoff; volume trial = qavLJ iff NBN < 0, else fav     */
       JMM_POT_LJCUT = 1,      /* phiLJcut + cut-off; volume trial = fav                         */
       JMM_POT_HARMONIC = 2 }; /* phiHarmoniccut + cut-off; volume trial = fav                   */
/* ENSEMBLE keyword, src/jmmMCState.cpp:402-425 */
enum { JMM_ENS_NPT = 0,        /* numTrialTypes = N+1, l0 = N                                    */
       JMM_ENS_NLT = 1 };      /* numTrialTypes = N,   l0 = L, RELAX ignored                     */
/* where the random numbers of Step() come from (src/jmmMCState.cpp:1762-1763,1370,1668,2252) */
enum { JMM_RNG_TAUS2 = 0,      /* gsl_rng_taus2 run on the device: the reference's own stream    */
       JMM_RNG_PHILOX = 1,     /* Philox4x32-10, counter = (step, chain): production             */
       JMM_RNG_RECORDED = 2 }; /* raw u32 words recorded from the reference (lock-step mode)     */
/* pair-distance arithmetic */
enum { JMM_MODE_TABLE = 0,     /* incremental rij table = the reference's arithmetic, bit-exact  */
       JMM_MODE_RECOMPUTE = 1, /* r[j]-r[i] from positions, O(N) memory: production              */
       JMM_MODE_CHECKERBOARD = 2 }; /* one long chain, colour-decomposed sweeps (jmm_sweep)      */
/* who runs maxDisAdjust / maxDVAdjust (src/jmmMCState.cpp:2100-2139) */
enum { JMM_ADAPT_HOST = 0,     /* host libm log(): bit-identical step sizes to the reference      */
       JMM_ADAPT_DEVICE = 1,   /* inside the kernel: no launch boundary every DADJ/VADJ steps     */
       JMM_ADAPT_CALLER = 2 }; /* jmm_step never adjusts or relaxes: the caller drives the cadence
                                  of src/Main.cpp:145-176 itself (jmm_adjust_step_sizes,
                                  jmm_relax_volume) — what the jmmMCState.h-compatible shim does   */

/* arithmetic of the production displacement trial (JMM_MODE_RECOMPUTE, Philox, one chain per thread) */
enum { JMM_ARITH_REFERENCE = 0, /* every pair term and sum as in src/pot.cpp / qad2: bit-identical     */
       JMM_ARITH_FAST = 1 };    /* LJ/LJcut: one division per partner, r^-6/r^-12 differences only;
                                   totals equal to <= 1e-12 relative, same decisions (see prod.cuh)   */

/* jmm_config.flags.  The reference's running virial is path-dependent: its pair virial carries no 1/l
 * (src/pot.cpp:60-63), fav / moveVolume / ECheck set Vir = sum of pair virials (src/jmmMCState.cpp:2261, 2908, 2077),
 * but an accepted qavLJ move rescales Vir6, Vir12 by (l'/l)^-7, ^-13, adds the ideal term N T / l, and leaves HV, HV6,
 * HV12 untouched (:1675-1686).  Lock-step and the default production mode reproduce that bit for bit.
 * JMM_FLAG_CONSISTENT_VIRIAL makes an accepted qavLJ move keep the definition the other paths use — Vir6, HV6 scale
 * by (l'/l)^-6, Vir12, HV12 by ^-12, Vir = Vir12 - Vir6, HV = HV12 - HV6, no ideal term — so that the running Vir and HV
 * always equal the configuration sums (jmm_energy).  Positions, box length, E and every accept/reject decision are the
 * same with and without the flag; only the Virial / HV columns of thermo.dat.mcs and Summary.dat change. */
enum { JMM_FLAG_CONSISTENT_VIRIAL = 1 };

/* index of each total in a 9-vector: the order phi() writes them, src/pot.cpp:90-100 */
enum { JMM_E = 0, JMM_VIR, JMM_E12, JMM_VIR12, JMM_E6, JMM_VIR6, JMM_HV, JMM_HV12, JMM_HV6, JMM_NTOT };
/* index of each running sum in a 12-vector: updateThermo, src/jmmMCState.cpp:1941-1961 */
enum { JMM_A_RHO = 0, JMM_A_RHO2, JMM_A_L, JMM_A_L2, JMM_A_E, JMM_A_E2, JMM_A_LE,
       JMM_A_VIR, JMM_A_VIR2, JMM_A_EVIR, JMM_A_HV, JMM_A_HV2, JMM_NACC };
/* counters: dAcc[0], dAcc[1], vAcc[0], vAcc[1], src/jmmMCState.cpp:66-69 */
enum { JMM_C_DACC = 0, JMM_C_DREJ, JMM_C_VACC, JMM_C_VREJ, JMM_NCNT };

/* Mirror of `struct MCInput` (src/jmmMCState.h:50-85) for the fields the hot path uses, plus the
 * batch/device fields the reference has no notion of. */
typedef struct jmm_config {
    uint64_t N;            /* N                                                                  */
    int32_t  nbn;          /* NBN (index-distance neighbour limit, <0 = none)                    */
    int32_t  pot;          /* JMM_POT_*                                                          */
    double   cutoff;       /* POT <name> <cutoff>; INFINITY when absent (readInput.cpp:125-130)  */
    int32_t  ensemble;     /* JMM_ENS_*                                                          */
    int32_t  relax;        /* RELAX                                                              */
    double   P, T, L;      /* P, T, L — defaults for every chain; see jmm_set_state              */
    double   maxStep;      /* MAXSTEP                                                            */
    double   maxdl;        /* MAXDV                                                              */
    uint64_t eci;          /* ENGCHECK (0 = never)                                               */
    uint64_t mdai;         /* DADJ     (0 = never)                                               */
    uint64_t mvai;         /* VADJ     (0 = never)                                               */
    uint64_t seed;         /* SEED                                                               */
    uint64_t nchains;      /* independent chains held by this handle                             */
    uint64_t chain_id0;    /* global id of chain 0 (Philox subsequence = chain_id0 + index;
                              TAUS2 seed = seed + chain_id0 + index)                             */
    int32_t  rng_kind;     /* JMM_RNG_*                                                          */
    int32_t  mode;         /* JMM_MODE_*                                                         */
    int32_t  adapt;        /* JMM_ADAPT_*                                                        */
    int32_t  device;       /* CUDA device ordinal                                                */
    int32_t  arith;        /* JMM_ARITH_*                                                        */
    int32_t  flags;        /* JMM_FLAG_* bits; 0 = everything as the reference does it           */
} jmm_config;

/* The print/cadence keywords of the INPUT deck that the hot path itself does not consume
 * (src/readInput.cpp:85-191); filled by jmm_read_input for the host driver. */
typedef struct jmm_deck {
    uint64_t numsteps, cpi, tpi, gpi, rhopi, gnb, rhonb;
    double   rbw, gsw, gbw;
    int32_t  gns;
    int32_t  is_restart;
    int32_t  n_unknown;        /* lines answered with "Property command %s not understood."      */
    char     pot_str[80], ensemble_str[80];
} jmm_deck;

typedef struct jmm_handle jmm_handle;

/* readInput(char*), src/readInput.cpp:8-262.  Same keywords, same delimiters, same
 * "not understood" message on stdout; a missing file is JMM_ERR_IO instead of a crash. */
jmm_status jmm_read_input(const char *path, jmm_config *cfg, jmm_deck *deck);

/* setupMCS(), src/jmmMCState.cpp:261-790 (state only, no files): allocates nchains chains on the
 * device, l = N (NPT) or L (NLT), r[i] = ((i+.5)/N-.5)*l, sentinel totals, zero sums/counters. */
jmm_status jmm_create(const jmm_config *cfg, jmm_handle **out);
/* freeMCS(), src/jmmMCState.cpp:792-842 */
jmm_status jmm_destroy(jmm_handle *h);

/* Overwrite per-chain state before jmm_start; any pointer may be NULL (= keep).  r [nchains][N],
 * l/P/T [nchains].  (The reference can only start from its lattice; sweeps need per-chain P,T:
 * scripts/RunJobs.bash:16-27 launches one process per state point.) */
jmm_status jmm_set_state(jmm_handle *h, const double *r, const double *l, const double *P, const double *T);
/* maxStep / maxdl per chain: the values maxDisAdjust/maxDVAdjust maintain, :2100-2139 */
jmm_status jmm_set_step_sizes(jmm_handle *h, const double *maxStep, const double *maxdl);
jmm_status jmm_get_step_sizes(jmm_handle *h, double *maxStep, double *maxdl);

/* The prologue of main(): fad(mcs,&0,&0.5) "step 0" (src/Main.cpp:66-68, jmmMCState.cpp:853-1003),
 * relaxVolume if RELAX (src/Main.cpp:71-73) and the first updateThermo (src/Main.cpp:96). */
jmm_status jmm_start(jmm_handle *h);
/* the same three pieces individually, for callers that keep main()'s own call sequence:
 * fad (src/Main.cpp:66-68), relaxVolume (:71-73), updateThermo (:96) */
jmm_status jmm_start_parts(jmm_handle *h, int32_t do_fad, int32_t do_relax, int32_t do_thermo);

/* Configuration totals from the current positions: the pair loop shared by fad :907-946,
 * fav :2196-2235, ECheck :1974-1993, moveVolume :2865-2904.  totals [nchains][9].
 * exact_order != 0 sums in the reference's pair-index order (bit-exact, one thread per chain);
 * 0 uses the parallel reduction (fast; differs by summation rounding only). */
jmm_status jmm_energy(jmm_handle *h, double *totals, int32_t exact_order);

/* nsteps x { incrementStep :1734 ; Step :1758-1811 } for every chain: draw, qad2 :1160-1464 or
 * qavLJ :1648-1730 / fav :2161-2293, ECheck every eci steps :1965-2095, updateThermo :1941-1961.
 * With JMM_ADAPT_DEVICE also maxDisAdjust/maxDVAdjust and the periodic relaxVolume of
 * src/Main.cpp:173-176; with JMM_ADAPT_HOST the library splits the launch at those steps itself.
 * rng_stream/n_words: the recorded u32 words (JMM_RNG_RECORDED, nchains == 1), else NULL/0.
 * accept_log: optional [nsteps][nchains] bytes; bit0 accepted, bit1 volume trial, bit2 wall reject. */
jmm_status jmm_step(jmm_handle *h, uint64_t nsteps, const uint32_t *rng_stream, uint64_t n_words,
                    uint8_t *accept_log);

/* relaxVolume(), src/jmmMCState.cpp:2396-2679 (+ calculateEnergyOfTrialVolumeChange :2783-2827,
 * moveVolume :2831-2916) on every chain */
jmm_status jmm_relax_volume(jmm_handle *h);
/* maxDisAdjust :2100-2115 / maxDVAdjust :2120-2139 on the host (glibc log), for every chain */
jmm_status jmm_adjust_step_sizes(jmm_handle *h, int32_t do_dis, int32_t do_vol);

/* Read back; any pointer may be NULL.  r [nchains][N], l [nchains], totals [nchains][9],
 * accum [nchains][12], counters [nchains][4]. */
jmm_status jmm_get_state(jmm_handle *h, double *r, double *l, double *totals, double *accum, uint64_t *counters);
/* what printThermo does to the sums after printing, :1922-1933 */
jmm_status jmm_zero_accum(jmm_handle *h);
/* overwrite the twelve running sums (accum [nchains][12]) and the number of updateThermo calls they hold: the inverse
 * of jmm_get_state + jmm_zero_accum, for a driver that prints-and-zeroes like printThermo but wants the device-side
 * summary records (jmm_summaries) to carry its whole-run sums */
jmm_status jmm_set_accum(jmm_handle *h, const double *accum, uint64_t samples);
/* getStepNum(), :2777 */
uint64_t   jmm_step_number(const jmm_handle *h);
/* Continue at a given step number: what a restart does with the step count of the last frame of config.dat.mcs
 * (setupMCS restart branch, src/jmmMCState.cpp:572-765, `sscanf(... &(mcs->sn) ...)` at :641).  The step number selects the Philox
 * blocks and decides whether the relaxVolume cadence (every 10 000 steps below 1 000 000, src/Main.cpp:173) is
 * still active.  Many-chain handles only. */
jmm_status jmm_set_step_number(jmm_handle *h, uint64_t sn);
/* ECheck statistics summed over chains: number of checks and of "Energy discrepancy" resets */
jmm_status jmm_echeck_stats(jmm_handle *h, uint64_t *checks, uint64_t *discrepancies);
/* words of the recorded stream consumed so far */
uint64_t   jmm_stream_cursor(const jmm_handle *h);

/* Density rho(x) and two-particle density g(x) histograms for many-chain handles:
 * fgrho :1069-1127, qagrho :2297-2384, ugrho :1131-1149 with the reference's bins —
 *   rho bin  = floor(r/RBW + RHONB/2);  g segment = floor(r/GSW + GNS/2);  g bin = floor(|rij|/GBW).
 * Call after jmm_create/jmm_set_state and before jmm_start: it performs the fgrho + ugrho of setupMCS
 * (:773-776) on the current configuration.  The reference's quirks are kept: fav and relaxVolume never
 * touch the histograms, a fav step is not accumulated.  Accumulation is lazy (per changed bin), the
 * integers are the reference's. */
jmm_status jmm_enable_histograms(jmm_handle *h, uint64_t rhonb, double rbw, int32_t gns, uint64_t gnb, double gsw, double gbw);
/* What printRho :1021-1038 / printG :1042-1064 read and reset: the counts accumulated since the last call.
 * rhoA [nchains][rhonb], gA [nchains][gns][gnb]; a NULL pointer leaves that histogram untouched. */
jmm_status jmm_take_histograms(jmm_handle *h, int64_t *rhoA, int64_t *gA);

/* Exact restart.  The reference's RESTART (src/jmmMCState.cpp:572-765, src/readInput.cpp:69-78)
 * re-reads the last complete frame of config.dat.mcs and re-seeds the generator with the SAME seed: maxStep, maxdl,
 * the acceptance counters and the generator state are lost, so a resumed run is only statistically a continuation.
 * jmm_checkpoint_save writes everything the next step depends on — step number, positions, box, the nine totals,
 * twelve sums, four counters, step sizes, vAErrNtot, ECheck statistics, taus2 states / recorded-stream cursor, the
 * pair table in table mode, the histogram bins when enabled (a Philox stream needs nothing: its counter IS the step
 * number) — to one binary file; jmm_checkpoint_load restores it into a handle created from the same jmm_config
 * (checked: N, nchains, mode, pot, NBN, ensemble, generator, seed, chain_id0), after which jmm_step / jmm_sweep
 * continue bit for bit as if the run had never stopped.  JMM_ERR_IO on file errors, JMM_ERR_INVALID on a mismatch. */
jmm_status jmm_checkpoint_save(jmm_handle *h, const char *path);
jmm_status jmm_checkpoint_load(jmm_handle *h, const char *path);

/* JMM_MODE_CHECKERBOARD (no reference counterpart — the reference moves
 * one particle per Step and cannot allocate its O(N^2) tables beyond N ~ 1e4).  One call performs
 * n_halfsweeps colour half-sweeps: in each, a colour c in [0, NBN+1) is drawn (Philox) and every
 * particle i with i mod (NBN+1) == c makes one qad2 trial; same-colour particles do not interact
 * (|i-j| <= NBN rule, :1217,1312), so the trials commute.  Totals and counters are maintained;
 * the twelve sums are updated once per half-sweep.  trials_out = number of trial moves made. */
jmm_status jmm_sweep(jmm_handle *h, uint64_t n_halfsweeps, uint64_t *trials_out);

/* ---- Multi-GPU.  The reference runs one process per state point (scripts/RunJobs.bash:99-101, one
 * LSF job each) and joins the results through files (scripts/Analyze_Mean.py reads every thermo.dat.mcs).  Here chains
 * are sharded over ranks (one process per GPU, cfg.chain_id0 = first global chain id of the rank, so the Philox streams
 * do not depend on the sharding) with no data-path traffic, and the join is ONE ncclAllGather of per-chain records.
 * Record = JMM_SUMMARY_DOUBLES doubles:
 *   0 global chain id   1 P   2 T   3 N   4 samples = updateThermo calls in the sums (since jmm_zero_accum)
 *   5..16 the twelve running sums (JMM_A_* order)   17..20 dAcc[0], dAcc[1], vAcc[0], vAcc[1]   21 l   22 E   23 Vir
 * NCCL is loaded at run time (libnccl.so.2); without it these calls return JMM_ERR_NCCL and everything else works. */
enum { JMM_SUMMARY_DOUBLES = 24, JMM_COMM_ID_BYTES = 128 };
typedef struct jmm_comm jmm_comm;
/* ncclGetUniqueId: rank 0 calls it and hands the bytes to the other ranks (file, environment, launcher) */
jmm_status jmm_comm_unique_id(uint8_t id[JMM_COMM_ID_BYTES]);
/* ncclCommInitRank on `device`; collective over the `world` ranks */
jmm_status jmm_comm_create(const uint8_t id[JMM_COMM_ID_BYTES], int32_t rank, int32_t world, int32_t device, jmm_comm **out);
jmm_status jmm_comm_destroy(jmm_comm *c);
/* this handle's records, packed on the device: records [nchains][JMM_SUMMARY_DOUBLES] (host) */
jmm_status jmm_summaries(jmm_handle *h, double *records);
/* Collective: every rank packs its records on the device, one ncclAllGather (NVLink) on the handle's stream, and the
 * gathered table comes back to every rank's host buffer out [total_chains][JMM_SUMMARY_DOUBLES] in global chain order.
 * Each rank may hold at most ceil(total_chains / world) chains; every chain id in [0, total_chains) must be held by
 * exactly one rank. */
jmm_status jmm_allgather_summaries(jmm_handle *h, jmm_comm *c, uint64_t total_chains, double *out);
/* NCCL version code (e.g. 22809) of the library bound at run time, 0 if none */
int32_t    jmm_nccl_version(void);

/* instrumentation for bench.py: kernels launched so far, device time of the last jmm_step /
 * jmm_sweep / jmm_energy kernel(s) in ms (CUDA events on the handle's stream) */
uint64_t   jmm_kernel_launches(const jmm_handle *h);
double     jmm_last_kernel_ms(const jmm_handle *h);
/* name of the kernel family jmm_step (jmm_sweep for a checkerboard handle) launches for this handle, e.g.
 * "k_chains_step_crew", "k_chains_step_lanes<G=8>", "k_chains_step_prod"; a static string */
const char *jmm_engine(const jmm_handle *h);
/* run on this CUDA stream (cudaStream_t as void*) instead of the handle's own */
jmm_status jmm_set_stream(jmm_handle *h, void *cuda_stream);
/* pinned host memory for the caller's staging buffers (cudaHostAlloc / cudaFreeHost) */
void      *jmm_host_alloc(uint64_t bytes);
void       jmm_host_free(void *p);

/* fp64 pipe microbenchmark for the roofline denominator (MEASURED_PEAKS.json has no fp64 entry):
 * independent DFMA chains on every SM; returns TFLOP/s counting one FMA as 2 flop, <0 on error */
double     jmm_fp64_peak_tflops(int32_t device);

const char *jmm_last_error(void);
const char *jmm_version(void);
/* device-side self-test of the generators: fills out[0..3] with Philox4x32-10(ctr,key) and
 * taus_out[0..n-1] with the first n taus2 words for `seed`, computed ON THE GPU */
jmm_status jmm_rng_selftest(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4],
                            uint64_t seed, uint32_t *taus_out, uint32_t n, int32_t device);
/* device-side self-test of the banded acceptance rules (pot.cuh): n random cases each of the Metropolis rule
 * (src/jmmMCState.cpp:1367-1377) and of the volume rule (:1666-1672, :2249-2255), half of them with the random
 * number placed within a few ulp .. 1e-4 of the exact acceptance probability, decided both by the banded fast path
 * and by the reference expression.  counts[0] = Metropolis mismatches, [1] = volume mismatches, [2] / [3] = cases
 * that reached the exact expression.  A correct build returns counts[0] = counts[1] = 0. */
jmm_status jmm_accept_selftest(uint64_t n, uint64_t seed, uint64_t counts[4], int32_t device);

#ifdef __cplusplus
}
#endif
#endif /* JMM_GPU_H */
