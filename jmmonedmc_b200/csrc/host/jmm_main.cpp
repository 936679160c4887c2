// jmm_run — batch driver of the GPU engine: the role of src/Main.cpp (one chain) and of
// scripts/RunJobs.bash + 100 LSF jobs (a P x T grid of state points) in one process per GPU.
//
//   jmm_run [INPUT] [--chains C] [--sweep-p lo hi n] [--sweep-t lo hi n] [--outdir DIR] [--lockstep]
//           [--rank R --world W]        (chain range [R*C/W, (R+1)*C/W), Philox subsequence = global chain id)
//
// Reads the reference's INPUT format (jmm_read_input == readInput, src/readInput.cpp:8) and keeps
// Main.cpp's cadence (src/Main.cpp:114-176): it launches min(next CPI/TPI boundary) - sn steps at a
// time, then prints.  Outputs:
//   one chain      thermo.dat.mcs / config.dat.mcs in DIR, byte-compatible with the reference
//                  (header :566-568, rows :1916-1918, frames :1009-1013), stdout lines of :1920, :2143-2156
//   many chains    thermo_chains.dat.mcs (chain id + the same 13 columns per row) and Summary.dat:
//                  per chain P, T, N, samples, the 12 run means, acceptance ratios — the table
//                  scripts/Analyze_Mean.py builds from 100 thermo files
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../../include/jmm_gpu.h"

#define JCK(call)                                                                         \
    do {                                                                                  \
        if ((call) != JMM_OK) { fprintf(stderr, "jmm_run: %s: %s\n", #call, jmm_last_error()); return 2; } \
    } while (0)

static const char *kThermoHeader =                                       // src/jmmMCState.cpp:566-568
    "Step    Econf           Econf2          L       L2  "
    "    LEconf          rho             rho2            Virial      "
    "   Virial2         EconfVir        HV              HV2 \n";

int main(int argc, char **argv) {
    std::string input = "INPUT", outdir = ".";
    uint64_t chains = 0, rank = 0, world = 1;
    bool lockstep = false;
    double sp[3] = {0, 0, 0}, st[3] = {0, 0, 0};
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](int n) { if (i + n >= argc) { fprintf(stderr, "jmm_run: %s needs %d value(s)\n", a.c_str(), n); exit(1); } };
        if (a == "--chains") { need(1); chains = strtoull(argv[++i], nullptr, 10); }
        else if (a == "--outdir") { need(1); outdir = argv[++i]; }
        else if (a == "--lockstep") lockstep = true;
        else if (a == "--rank") { need(1); rank = strtoull(argv[++i], nullptr, 10); }
        else if (a == "--world") { need(1); world = strtoull(argv[++i], nullptr, 10); }
        else if (a == "--sweep-p") { need(3); for (int k = 0; k < 3; ++k) sp[k] = atof(argv[++i]); }
        else if (a == "--sweep-t") { need(3); for (int k = 0; k < 3; ++k) st[k] = atof(argv[++i]); }
        else input = a;
    }
    printf("#####################################################\n#        jmmOneDMC hot path on B200: %s\n"
           "#####################################################\n\n", jmm_version());
    jmm_config cfg;
    jmm_deck deck;
    JCK(jmm_read_input(input.c_str(), &cfg, &deck));
    const uint64_t np = sp[2] > 0 ? (uint64_t) sp[2] : 1, nt = st[2] > 0 ? (uint64_t) st[2] : 1;
    uint64_t total = chains ? chains : cfg.nchains;
    if (np * nt > 1) total = np * nt * std::max<uint64_t>(1, total / (np * nt) ? total / (np * nt) : 1);
    const uint64_t c0 = rank * total / world, c1 = (rank + 1) * total / world, C = c1 - c0;
    if (C == 0) { fprintf(stderr, "jmm_run: rank %lu has no chains\n", (unsigned long) rank); return 1; }
    cfg.nchains = C;
    cfg.chain_id0 = c0;
    // lock-step: the host drives maxDisAdjust/maxDVAdjust/relaxVolume itself, in Main.cpp's order between the prints
    if (lockstep) { cfg.rng_kind = JMM_RNG_TAUS2; cfg.mode = JMM_MODE_TABLE; cfg.adapt = JMM_ADAPT_CALLER; }
    if (cfg.ensemble == JMM_ENS_NPT) printf("ENSEMBLE = NPT\n"); else printf("ENSEMBLE = NLT\n");
    printf("N: %lu  chains: %lu (global %lu..%lu of %lu)  numSteps: %lu  POT %s  NBN %d\n", (unsigned long) cfg.N,
           (unsigned long) C, (unsigned long) c0, (unsigned long) c1 - 1, (unsigned long) total,
           (unsigned long) deck.numsteps, deck.pot_str, cfg.nbn);

    jmm_handle *h = nullptr;
    JCK(jmm_create(&cfg, &h));
    // state points: chain g -> (P, T) on the grid, replicas of a point are consecutive chains
    std::vector<double> P(C, cfg.P), T(C, cfg.T);
    if (np * nt > 1) {
        const uint64_t reps = total / (np * nt);
        for (uint64_t c = 0; c < C; ++c) {
            const uint64_t point = (c0 + c) / reps, ip = point / nt, it = point % nt;
            if (np > 1) P[c] = sp[0] + (sp[1] - sp[0]) * (double) ip / (double) (np - 1);
            if (nt > 1) T[c] = st[0] + (st[1] - st[0]) * (double) it / (double) (nt - 1);
        }
        JCK(jmm_set_state(h, nullptr, nullptr, P.data(), T.data()));
    }
    printf("Setup completed\n");

    const bool single = (total == 1);
    FILE *tf = fopen((outdir + (single ? "/thermo.dat.mcs" : "/thermo_chains.dat.mcs")).c_str(), "w");
    FILE *cf = single ? fopen((outdir + "/config.dat.mcs").c_str(), "w") : nullptr;
    if (!tf || (single && !cf)) { fprintf(stderr, "jmm_run: cannot open output files in %s\n", outdir.c_str()); return 1; }
    if (single) fputs(kThermoHeader, tf); else fprintf(tf, "chain\t%s", kThermoHeader);

    // histograms: a single chain writes rho.dat.mcs and g<k>.dat.mcs like the reference (src/jmmMCState.cpp:526-533)
    const bool hist = single && deck.rhonb > 0 && deck.rbw > 0 && deck.gsw > 0 && deck.gbw > 0 && !getenv("JMM_RUN_NO_HIST");
    FILE *rhof = nullptr;
    std::vector<FILE *> gf;
    std::vector<int64_t> hbuf;
    uint64_t slrho = (uint64_t) -1, slg = (uint64_t) -1;
    if (hist) {
        JCK(jmm_enable_histograms(h, deck.rhonb, deck.rbw, deck.gns, deck.gnb, deck.gsw, deck.gbw));
        rhof = fopen((outdir + "/rho.dat.mcs").c_str(), "w");
        for (int k = 0; k < deck.gns; ++k) gf.push_back(fopen((outdir + "/g" + std::to_string(k) + ".dat.mcs").c_str(), "w"));
    }
    std::vector<double> r(single ? cfg.N : 0), l(C), tot(C * 9), acc(C * 12), run(C * 12, 0.0);
    std::vector<uint64_t> cnt(C * 4);
    uint64_t sn = 0, sltp = (uint64_t) -1, samples = 0;
    auto print_thermo = [&]() -> int {                                   // printThermo, :1896-1937
        JCK(jmm_get_state(h, nullptr, l.data(), tot.data(), acc.data(), cnt.data()));
        const uint64_t ss = sn - sltp;
        for (uint64_t c = 0; c < C; ++c) {
            const double *a = &acc[c * 12];
            if (!single) fprintf(tf, "%lu\t", (unsigned long) (c0 + c));
            fprintf(tf, "%lu\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\n", (unsigned long) sn,
                    a[JMM_A_E] / ss, a[JMM_A_E2] / ss, a[JMM_A_L] / ss, a[JMM_A_L2] / ss, a[JMM_A_LE] / ss, a[JMM_A_RHO] / ss,
                    a[JMM_A_RHO2] / ss, a[JMM_A_VIR] / ss, a[JMM_A_VIR2] / ss, a[JMM_A_EVIR] / ss, a[JMM_A_HV] / ss,
                    a[JMM_A_HV2] / ss);
            for (int k = 0; k < 12; ++k) run[c * 12 + k] += a[k];
        }
        samples += ss;
        fflush(tf);
        if (single) printf("%lu  %.8G  %.8G  %.8G  %.8G\n", (unsigned long) sn, tot[JMM_E], l[0], tot[JMM_VIR], tot[JMM_HV]);
        JCK(jmm_zero_accum(h));
        sltp = sn;
        return 0;
    };
    auto print_coords = [&]() -> int {                                   // printCoords, :1007-1017
        if (!single) return 0;
        JCK(jmm_get_state(h, r.data(), l.data(), nullptr, nullptr, nullptr));
        fprintf(cf, "%lu\nStep no.: %lu  Box length: %.5f\n", (unsigned long) cfg.N, (unsigned long) sn, l[0]);
        for (uint64_t i = 0; i < cfg.N; ++i) fprintf(cf, "%lu  0.0  0.0  %.8G\n", (unsigned long) (i + 1), r[i]);
        fflush(cf);
        return 0;
    };

    auto print_rho = [&]() -> int {                                      // printRho, :1021-1038
        if (!hist) return 0;
        hbuf.resize(deck.rhonb);
        JCK(jmm_take_histograms(h, hbuf.data(), nullptr));
        const uint64_t ns = sn - slrho;
        fprintf(rhof, "%lu", (unsigned long) sn);
        for (uint64_t b = 0; b < deck.rhonb; ++b) fprintf(rhof, " %.8G", (double) (int) hbuf[b] / ns / deck.rbw);
        fprintf(rhof, "\n");
        fflush(rhof);
        slrho = sn;
        return 0;
    };
    auto print_g = [&]() -> int {                                        // printG, :1042-1064
        if (!hist) return 0;
        hbuf.resize((size_t) deck.gns * deck.gnb);
        JCK(jmm_take_histograms(h, nullptr, hbuf.data()));
        const uint64_t ns = sn - slg;
        for (int k = 0; k < deck.gns; ++k) {
            fprintf(gf[k], "%lu", (unsigned long) sn);
            for (uint64_t b = 0; b < deck.gnb; ++b) fprintf(gf[k], " %.8G", (double) (int) hbuf[(size_t) k * deck.gnb + b] / ns / deck.gsw / deck.gbw);
            fprintf(gf[k], "\n");
            fflush(gf[k]);
        }
        slg = sn;
        return 0;
    };

    printf("Step: 0...\n");
    JCK(jmm_start(h));                                                   // src/Main.cpp:66-96
    if (print_coords()) return 2;
    if (print_rho()) return 2;
    if (print_thermo()) return 2;
    if (print_g()) return 2;
    const uint64_t tpi = deck.tpi ? deck.tpi : deck.numsteps, cpi = deck.cpi ? deck.cpi : deck.numsteps;
    while (sn < deck.numsteps) {                                         // src/Main.cpp:114-180, batched
        uint64_t n = deck.numsteps - sn;
        if (tpi) n = std::min(n, tpi - sn % tpi);
        if (single && cpi) n = std::min(n, cpi - sn % cpi);
        if (hist && deck.rhopi) n = std::min(n, deck.rhopi - sn % deck.rhopi);
        if (hist && deck.gpi) n = std::min(n, deck.gpi - sn % deck.gpi);
        const bool relax_on = lockstep && cfg.relax > 0 && cfg.ensemble == JMM_ENS_NPT;
        if (lockstep && cfg.mdai) n = std::min(n, cfg.mdai - sn % cfg.mdai);
        if (lockstep && cfg.mvai) n = std::min(n, cfg.mvai - sn % cfg.mvai);
        if (relax_on && sn < 1000000) n = std::min<uint64_t>(n, 10000 - sn % 10000);
        JCK(jmm_step(h, n, nullptr, 0, nullptr));
        const uint64_t before = sn;
        sn += n;
        if (sn / 10000 != before / 10000) printf("Step: %lu...\n", (unsigned long) (sn / 10000 * 10000));
        if (single && cpi && sn % cpi == 0 && print_coords()) return 2;
        if (tpi && sn % tpi == 0 && print_thermo()) return 2;
        if (hist && deck.rhopi && sn % deck.rhopi == 0 && print_rho()) return 2;
        if (lockstep) {                                                  // src/Main.cpp:145-165
            const bool dis = cfg.mdai && sn % cfg.mdai == 0, vol = cfg.mvai && sn % cfg.mvai == 0;
            if (dis || vol) JCK(jmm_adjust_step_sizes(h, dis, vol));
        }
        if (hist && deck.gpi && sn % deck.gpi == 0 && print_g()) return 2;
        if (relax_on && sn % 10000 == 0 && sn < 1000000) JCK(jmm_relax_volume(h));   // src/Main.cpp:173-176
        fflush(stdout);
    }
    JCK(jmm_get_state(h, nullptr, l.data(), tot.data(), acc.data(), cnt.data()));
    uint64_t checks = 0, disc = 0;
    JCK(jmm_echeck_stats(h, &checks, &disc));
    printf("\nPROGRAM COMPLETED SUCCESSFULLY!\n");
    if (single) {
        printf("\nE = %.8G\n", tot[JMM_E]);                              // printE :2143, printAcc :2151
        printf("Accepted/Rejected: Displacements VolumeChanges\n                        %lu/%lu          %lu/%lu\n",
               (unsigned long) cnt[0], (unsigned long) cnt[1], (unsigned long) cnt[2], (unsigned long) cnt[3]);
    } else {
        FILE *sf = fopen((outdir + (world > 1 ? "/Summary.rank" + std::to_string(rank) + ".dat" : "/Summary.dat")).c_str(), "w");
        fprintf(sf, "chain\tP\tT\tN\tsamples\tEconf\tEconf2\tL\tL2\tLEconf\trho\trho2\tVirial\tVirial2\tEconfVir\tHV\tHV2\t"
                    "dAccRatio\tvAccRatio\tEfinal\tLfinal\n");
        const int order[12] = {JMM_A_E, JMM_A_E2, JMM_A_L, JMM_A_L2, JMM_A_LE, JMM_A_RHO, JMM_A_RHO2, JMM_A_VIR, JMM_A_VIR2,
                               JMM_A_EVIR, JMM_A_HV, JMM_A_HV2};
        for (uint64_t c = 0; c < C; ++c) {
            fprintf(sf, "%lu\t%.8G\t%.8G\t%lu\t%lu", (unsigned long) (c0 + c), P[c], T[c], (unsigned long) cfg.N, (unsigned long) samples);
            for (int k = 0; k < 12; ++k) fprintf(sf, "\t%.8G", samples ? run[c * 12 + order[k]] / (double) samples : 0.0);
            const double d = (double) (cnt[c * 4] + cnt[c * 4 + 1]), v = (double) (cnt[c * 4 + 2] + cnt[c * 4 + 3]);
            fprintf(sf, "\t%.6G\t%.6G\t%.8G\t%.8G\n", d > 0 ? cnt[c * 4] / d : 0.0, v > 0 ? cnt[c * 4 + 2] / v : 0.0, tot[c * 9], l[c]);
        }
        fclose(sf);
    }
    printf("ECheck: %lu checks, %lu discrepancies; kernel launches: %lu\n", (unsigned long) checks, (unsigned long) disc,
           (unsigned long) jmm_kernel_launches(h));
    fclose(tf);
    if (cf) fclose(cf);
    if (rhof) fclose(rhof);
    for (FILE *f : gf) if (f) fclose(f);
    jmm_destroy(h);
    return 0;
}
