// jmm_mcstate_compat.cpp — the reference's own `jmmMCState.h` API, implemented on the GPU C ABI.
//
// Purpose: the drop-in proof.  The reference's UNMODIFIED src/Main.cpp and src/readInput.cpp are compiled
// from /root/reference and linked against this file + libjmmgpu.so instead of src/jmmMCState.cpp,
// src/pot.cpp and src/compute.cpp (recipe: oracle/Makefile, target `compat`).  The resulting program
// reads the same INPUT, runs every Step() on the B200 and writes thermo.dat.mcs / config.dat.mcs that
// are byte-identical to the reference's (tests/test_gpu_compat.py).
//
// This file includes the REFERENCE's jmmMCState.h (for `struct MCInput`, src/jmmMCState.h:50-85, and
// the prototypes :7-48) at build time; nothing of the reference is copied into this repository.
// Each function cites the reference function it replaces.  `struct MCState` is opaque in the reference
// (defined in the .cpp, src/jmmMCState.cpp:33-211), so its layout here is ours.
//
// Default mode = lock-step: rij-table arithmetic + gsl_rng_taus2 on the device + the caller (Main.cpp)
// driving maxDisAdjust/maxDVAdjust/relaxVolume exactly as in the reference.  JMM_COMPAT_PRODUCTION=1
// switches to the production arithmetic (positions only, Philox).
//
// rho(x) / g(x) histograms: jmm_enable_histograms / jmm_take_histograms; printRho and printG write
// rho.dat.mcs and g<k>.dat.mcs in the reference's format.
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "jmmMCState.h"                 // the reference's header, found via -I/root/reference/src
#include "../../../include/jmm_gpu.h"

double lRat1, lRat3, lRat6, lRat7, lRat12, lRat13;      // declared extern in jmmMCState.h:4

struct MCState {
    jmm_handle *h;
    jmm_config cfg;
    bool isRestart;
    unsigned long int N, sn, numSteps, cpi, tpi, mdai, mvai, gpi, rhopi, sltp;
    int relaxFlag;
    char potStr[80], ensembleStr[80];
    FILE *cf, *tf, *rhof;
    std::vector<FILE *> gf;
    unsigned long int rhonb, gnb, slrho, slg;
    int gns;
    double rbw, gsw, gbw;
    std::vector<int64_t> hist;
    std::vector<double> r;
    double l, tot[9], acc[12];
    uint64_t cnt[4];
};

// every reference function is entered by all threads of one OpenMP team (src/Main.cpp:50); only the
// master talks to the GPU, the others wait
#define MASTER_ONLY(body)            \
    do {                             \
        _Pragma("omp barrier")       \
        _Pragma("omp master")        \
        { body }                     \
        _Pragma("omp barrier")       \
    } while (0)

// Functions the reference calls from `omp sections` / `omp single` (src/Main.cpp:77-106,128-165) can run on different
// threads AT THE SAME TIME (printCoords, printRho, updateThermo + printThermo at step 0).  A jmm_handle is a
// one-host-thread object (include/jmm_gpu.h) and `m->r/tot/acc` are shared, so every such function runs its body
// under one named critical section.
#define GPU_LOCKED(...)                     \
    do {                                    \
        _Pragma("omp critical(jmm_gpu)")    \
        { __VA_ARGS__ }                     \
    } while (0)

static void die(const char *what) {
    fprintf(stderr, "jmm compat: %s failed: %s\n", what, jmm_last_error());
    exit(2);
}
#define JCK(call) do { if ((call) != JMM_OK) die(#call); } while (0)

static void pull(struct MCState *m) {
    JCK(jmm_get_state(m->h, m->r.data(), &m->l, m->tot, m->acc, m->cnt));
}

// printMCP, src/jmmMCState.cpp:216-257 (same lines; stdout is informational)
int printMCP(struct MCState *m) {
    printf("Printing Monte Carlo parameters...\nEnsemble: %s\n", m->ensembleStr);
    if (m->cfg.ensemble == JMM_ENS_NPT) printf("N: %lu\nP: %.5G\nT: %.5G\n", m->N, m->cfg.P, m->cfg.T);
    else printf("N: %lu\nL: %.5G\nT: %.5G\n", m->N, m->cfg.L, m->cfg.T);
    printf("Potential: %s\nPotential cut-off: %.5G\n", m->potStr, m->cfg.cutoff);
    if (m->cfg.nbn > 0) printf("Number of neighbors with which each particle can interact: %u\n", m->cfg.nbn);
    else printf("No neighbor number limit.\n");
    printf("numSteps: %lu\nmaxStep: %.5G\nmax vol change: %.5G\n", m->numSteps, m->cfg.maxStep, m->cfg.maxdl);
    printf("Adjust max. displacement every %lu steps.\nAdjust max. volume change every %lu steps.\n", m->mdai, m->mvai);
    printf("Engine: jmmonedmc_b200 (%s)\n", jmm_version());
    fflush(stdout);
    return 0;
}

// setupMCS, src/jmmMCState.cpp:261-790
struct MCState *setupMCS(struct MCInput inp) {
    struct MCState *m = new MCState();
    jmm_config &c = m->cfg;
    memset(&c, 0, sizeof c);
    c.N = inp.N; c.nbn = inp.nbn; c.P = inp.P; c.T = inp.T; c.L = inp.L;
    c.maxStep = inp.maxStep; c.maxdl = inp.maxdl; c.eci = inp.eci; c.mdai = inp.mdai; c.mvai = inp.mvai;
    c.seed = inp.seed; c.nchains = 1; c.chain_id0 = 0; c.device = 0;
    strncpy(m->potStr, inp.potStr, 79);
    strncpy(m->ensembleStr, inp.ensembleStr, 79);
    if (!strncmp(inp.potStr, "LJ", 10)) { c.pot = JMM_POT_LJ; c.cutoff = INFINITY; }                    // :292-329
    else if (!strncmp(inp.potStr, "LJcut", 10)) { c.pot = JMM_POT_LJCUT; c.cutoff = inp.potCutOff; }     // :330-362
    else if (!strncmp(inp.potStr, "HARMONIC", 10)) { c.pot = JMM_POT_HARMONIC; c.cutoff = inp.potCutOff; }  // :363-395
    else { printf("FATAL ERROR: Unknown potential.\nABORTING SIMULATION\n\n"); delete m; return NULL; }  // :396-399
    if (!strncmp(inp.ensembleStr, "NPT", 80)) {                                                          // :402-411
        printf("ENSEMBLE = NPT\n"); c.ensemble = JMM_ENS_NPT; c.relax = inp.relaxFlag > 0 && inp.relaxFlag < 2 ? 1 : 0;
    } else if (!strncmp(inp.ensembleStr, "NLT", 80)) {                                                   // :412-421
        printf("ENSEMBLE = NLT\n"); c.ensemble = JMM_ENS_NLT;
        if (inp.relaxFlag == 1) printf("This is an NLT ensemble simulation. Ignoring requested volume relaxation.\n");
        c.relax = 0;
    } else { printf("FATAL ERROR: Unknown ensemble.\nABORTING SIMULATION\n\n"); delete m; return NULL; } // :422-425
    const bool production = getenv("JMM_COMPAT_PRODUCTION") && atoi(getenv("JMM_COMPAT_PRODUCTION")) > 0;
    c.mode = production ? JMM_MODE_RECOMPUTE : JMM_MODE_TABLE;
    c.rng_kind = production ? JMM_RNG_PHILOX : JMM_RNG_TAUS2;
    c.adapt = JMM_ADAPT_CALLER;
    m->isRestart = inp.isRestart; m->N = inp.N; m->numSteps = inp.ns; m->relaxFlag = c.relax;
    m->cpi = inp.cpi; m->tpi = inp.tpi; m->mdai = inp.mdai; m->mvai = inp.mvai; m->gpi = inp.gpi; m->rhopi = inp.rhopi;
    m->sn = 0; m->sltp = (unsigned long int) -1;                                                         // :535-537
    m->r.resize(inp.N);
    printMCP(m);
    if (m->isRestart) { printf("jmm compat: RESTART is not supported (SURVEY §2: out of scope)\n"); exit(1); }
    if (jmm_create(&c, &m->h) != JMM_OK) die("jmm_create");
    m->rhonb = inp.rhonb; m->gnb = inp.gnb; m->gns = inp.gns; m->rbw = inp.rbw; m->gsw = inp.gsw; m->gbw = inp.gbw;
    m->slrho = (unsigned long int) -1; m->slg = (unsigned long int) -1;                                  // :538-539
    JCK(jmm_enable_histograms(m->h, m->rhonb, m->rbw, m->gns, m->gnb, m->gsw, m->gbw));                  // fgrho + ugrho, :773-776
    m->cf = fopen("config.dat.mcs", "w");                                                                // :526-528
    m->tf = fopen("thermo.dat.mcs", "w");
    m->rhof = fopen("rho.dat.mcs", "w");
    for (int k = 0; k < m->gns; k++) {                                                                   // :530-533
        char name[32];
        snprintf(name, sizeof name, "g%d.dat.mcs", k);
        m->gf.push_back(fopen(name, "w"));
    }
    fprintf(m->tf, "Step    Econf           Econf2          L       L2  "                                // :566-568
                   "    LEconf          rho             rho2            Virial      "
                   "   Virial2         EconfVir        HV              HV2 \n");
    fflush(m->tf);
    fflush(stdout);
    return m;
}

void freeMCS(struct MCState *m) {                                              // :792-842
    if (!m) return;
    jmm_destroy(m->h);
    if (m->cf) fclose(m->cf);
    if (m->tf) fclose(m->tf);
    if (m->rhof) fclose(m->rhof);
    for (FILE *f : m->gf) if (f) fclose(f);
    delete m;
}

void printStep(struct MCState *m) { printf("Step: %lu...\n", m->sn); fflush(stdout); }   // :845-848

// fad(mcs,&0,&0.5) — only ever called as "step 0" (src/Main.cpp:66-68)
int fad(struct MCState *m, unsigned long int *, double *) {
    MASTER_ONLY(JCK(jmm_start_parts(m->h, 1, 0, 0)););
    return 0;
}
int getRelaxFlag(struct MCState *m) { return m->relaxFlag; }                   // :2386
bool getRestartFlag(struct MCState *m) { return m->isRestart; }                // :2390
unsigned long int getStepNum(struct MCState *m) { return m->sn; }              // :2777

int relaxVolume(struct MCState *m) {                                           // :2396-2679
    MASTER_ONLY(
        pull(m);
        printf("Relaxing the Volume...Starting Length,Energy: %.5G,%.5G\n", m->l, m->tot[JMM_E]);
        JCK(jmm_relax_volume(m->h));
        pull(m);
        printf("Relaxation converged. Length,energy: %.10G,%.10G\n", m->l, m->tot[JMM_E]);
    );
    return 0;
}

int updateThermo(struct MCState *m) {                                          // :1941-1961 (step-0 call, Main.cpp:96)
    GPU_LOCKED(JCK(jmm_start_parts(m->h, 0, 0, 1)););
    return 0;
}

int printCoords(struct MCState *m) {                                           // :1007-1017
    GPU_LOCKED(
        pull(m);
        fprintf(m->cf, "%lu\nStep no.: %lu  Box length: %.5f\n", m->N, m->sn, m->l);
        for (unsigned long int i = 0; i < m->N; i++) fprintf(m->cf, "%lu  0.0  0.0  %.8G\n", i + 1, m->r[i]);
        fflush(m->cf);
    );
    return 0;
}

int printThermo(struct MCState *m) {                                           // :1896-1937
    GPU_LOCKED(
    pull(m);
    const unsigned long int ss = m->sn - m->sltp;
    const double *a = m->acc;
    fprintf(m->tf, "%lu\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\n", m->sn,
            a[JMM_A_E] / ss, a[JMM_A_E2] / ss, a[JMM_A_L] / ss, a[JMM_A_L2] / ss, a[JMM_A_LE] / ss, a[JMM_A_RHO] / ss,
            a[JMM_A_RHO2] / ss, a[JMM_A_VIR] / ss, a[JMM_A_VIR2] / ss, a[JMM_A_EVIR] / ss, a[JMM_A_HV] / ss, a[JMM_A_HV2] / ss);
    fflush(m->tf);
    printf("%lu  %.8G  %.8G  %.8G  %.8G\n", m->sn, m->tot[JMM_E], m->l, m->tot[JMM_VIR], m->tot[JMM_HV]);
    fflush(stdout);
    JCK(jmm_zero_accum(m->h));
    m->sltp = m->sn;
    );
    return 0;
}

int printRho(struct MCState *m) {                                              // :1021-1038
    GPU_LOCKED(
    m->hist.resize(m->rhonb);
    JCK(jmm_take_histograms(m->h, m->hist.data(), NULL));
    const unsigned long int ns = m->sn - m->slrho;
    fprintf(m->rhof, "%lu", m->sn);
    for (unsigned long int b = 0; b < m->rhonb; b++) fprintf(m->rhof, " %.8G", (double) (int) m->hist[b] / ns / m->rbw);
    fprintf(m->rhof, "\n");
    fflush(m->rhof);
    m->slrho = m->sn;
    );
    return 0;
}

int printG(struct MCState *m) {                                                // :1042-1064 (entered by every thread)
    MASTER_ONLY(
        m->hist.resize((size_t) m->gns * m->gnb);
        JCK(jmm_take_histograms(m->h, NULL, m->hist.data()));
        const unsigned long int ns = m->sn - m->slg;
        for (int k = 0; k < m->gns; k++) {
            fprintf(m->gf[k], "%lu", m->sn);
            for (unsigned long int b = 0; b < m->gnb; b++)
                fprintf(m->gf[k], " %.8G", (double) (int) m->hist[(size_t) k * m->gnb + b] / ns / m->gsw / m->gbw);
            fprintf(m->gf[k], "\n");
            fflush(m->gf[k]);
        }
        m->slg = m->sn;
    );
    return 0;
}

unsigned long int incrementStep(struct MCState *m) {                           // :1734-1754
    if (m->sn == m->numSteps) return 0;
    MASTER_ONLY(
        m->sn++;
        if (m->sn % 10000 == 0) printStep(m);
    );
    return m->sn;
}

int Step(struct MCState *m) {                                                  // :1758-1811
    MASTER_ONLY(JCK(jmm_step(m->h, 1, NULL, 0, NULL)););
    return 0;
}

int isCoordPrint(struct MCState *m) { return m->sn % m->cpi == 0; }            // :1815
int isThermoPrint(struct MCState *m) { return m->sn % m->tpi == 0; }           // :1825
int isRhoPrint(struct MCState *m) { return m->sn % m->rhopi == 0; }            // :1836
int isGPrint(struct MCState *m) { return m->sn % m->gpi == 0; }                // :1847
int isMaxDisAdjust(struct MCState *m) { return m->sn % m->mdai == 0; }         // :1870
int isMaxDVAdjust(struct MCState *m) { return m->sn % m->mvai == 0; }          // :1883

int maxDisAdjust(struct MCState *m) {                                          // :2100-2115
    GPU_LOCKED(
        JCK(jmm_adjust_step_sizes(m->h, 1, 0));
        double ms, mv;
        JCK(jmm_get_step_sizes(m->h, &ms, &mv));
        pull(m);
        printf("Step: %lu  Updating max Step...dAcc: %lu,%lu   new maxStep: %.5G\n", m->sn, (unsigned long) m->cnt[0],
               (unsigned long) m->cnt[1], ms);
    );
    return 0;
}

int maxDVAdjust(struct MCState *m) {                                           // :2120-2139
    GPU_LOCKED(JCK(jmm_adjust_step_sizes(m->h, 0, 1)););
    return 0;
}

int printE(struct MCState *m) { GPU_LOCKED(pull(m); printf("\nE = %.8G\n", m->tot[JMM_E]);); return 0; }   // :2143
int printAcc(struct MCState *m) {                                              // :2151
    GPU_LOCKED(
        pull(m);
        printf("Accepted/Rejected: Displacements VolumeChanges\n              \
          %lu/%lu          %lu/%lu\n", (unsigned long) m->cnt[0], (unsigned long) m->cnt[1], (unsigned long) m->cnt[2],
               (unsigned long) m->cnt[3]);
    );
    return 0;
}
