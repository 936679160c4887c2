// jmm_run — batch driver of the GPU engine: the role of src/Main.cpp (one chain) and of
// scripts/RunJobs.bash + 100 LSF jobs (a P x T grid of state points) in one process per GPU.
//
//   jmm_run [INPUT] [--chains C] [--sweep-p lo hi n] [--sweep-t lo hi n] [--outdir DIR] [--lockstep] [--fast]
//           [--rank R --world W [--id-file F]]   one process per GPU: chain range [R*C/W, (R+1)*C/W), Philox
//                                       subsequence = global chain id, no traffic while sampling; at the end ONE
//                                       ncclAllGather of the per-chain summary records (jmm_allgather_summaries) and
//                                       rank 0 writes the merged Summary.dat.  The NCCL id travels through file F
//                                       (default DIR/.jmm_nccl_id, written by rank 0).
//           [--layout runjobs]          also write data/<POT>/m<NBN>/N<N>/P<P>_T<T>/{INPUT,thermo.dat.mcs} per state
//                                       point under DIR: the tree scripts/RunJobs.bash:27 creates with one LSF job each
//           [--consistent-virial]       JMM_FLAG_CONSISTENT_VIRIAL: the running Virial/HV columns equal the configuration sums
//           [--device D]                CUDA device (default: the deck's GPU keyword; the rank when --world > 1)
//           [--checkpoint F] [--resume F]   exact restart (jmm_checkpoint_save / jmm_checkpoint_load): write F at the
//                                       end of the run / continue from F (outputs are appended, never truncated)
//
// A deck with a RESTART line is refused: the reference's RESTART re-reads config.dat.mcs and re-seeds
// (src/jmmMCState.cpp:572-765); the exact equivalent here is --resume.
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

#include <sys/stat.h>
#include <unistd.h>

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

static void mkdirs(const std::string &path) {
    for (size_t i = 1; i <= path.size(); ++i)
        if (i == path.size() || path[i] == '/') mkdir(path.substr(0, i).c_str(), 0777);
}

// "%g"-style number for directory names, the way the shell variables of RunJobs.bash read: 0.1, 0.55, 1
static std::string short_num(double x) {
    char b[64];
    snprintf(b, sizeof b, "%.10g", x);
    return b;
}

int main(int argc, char **argv) {
    std::string input = "INPUT", outdir = ".", ckpt_out, ckpt_in, id_file, layout;
    uint64_t chains = 0, rank = 0, world = 1;
    bool lockstep = false, fast = false, consistent = false;
    int device = -1;                     // --device D; default: the deck's GPU keyword, or the rank when --world > 1
    double sp[3] = {0, 0, 0}, st[3] = {0, 0, 0};
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](int n) { if (i + n >= argc) { fprintf(stderr, "jmm_run: %s needs %d value(s)\n", a.c_str(), n); exit(1); } };
        if (a == "--chains") { need(1); chains = strtoull(argv[++i], nullptr, 10); }
        else if (a == "--outdir") { need(1); outdir = argv[++i]; }
        else if (a == "--lockstep") lockstep = true;
        else if (a == "--fast") fast = true;
        else if (a == "--consistent-virial") consistent = true;
        else if (a == "--checkpoint") { need(1); ckpt_out = argv[++i]; }
        else if (a == "--resume") { need(1); ckpt_in = argv[++i]; }
        else if (a == "--id-file") { need(1); id_file = argv[++i]; }
        else if (a == "--layout") { need(1); layout = argv[++i]; }
        else if (a == "--device") { need(1); device = atoi(argv[++i]); }
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
    if (deck.is_restart) {
        fprintf(stderr, "jmm_run: this deck has a RESTART line.  The reference's restart re-reads config.dat.mcs and re-seeds "
                        "(src/jmmMCState.cpp:572-765); jmm_run does not do that and will not truncate the outputs of the "
                        "previous run.  Remove the line and continue exactly with --resume <checkpoint> (written by --checkpoint).\n");
        return 1;
    }
    if (layout.size() && layout != "runjobs") { fprintf(stderr, "jmm_run: unknown --layout %s\n", layout.c_str()); return 1; }
    if (rank >= world) { fprintf(stderr, "jmm_run: --rank must be < --world\n"); return 1; }
    if (fast) cfg.arith = JMM_ARITH_FAST;
    if (consistent) cfg.flags |= JMM_FLAG_CONSISTENT_VIRIAL;
    if (device >= 0) cfg.device = device;
    else if (world > 1 && cfg.device == 0) cfg.device = (int32_t) rank;
    if (id_file.empty()) id_file = outdir + "/.jmm_nccl_id";
    if (world > 1 && rank == 0) unlink(id_file.c_str());            // a stale id of an earlier run
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
    const bool resume = !ckpt_in.empty();
    const char *fmode = resume ? "a" : "w";                            // a resumed run appends: nothing is truncated
    const std::string rank_tag = world > 1 ? ".rank" + std::to_string(rank) : "";
    FILE *tf = fopen((outdir + (single ? "/thermo.dat.mcs" : "/thermo_chains" + rank_tag + ".dat.mcs")).c_str(), fmode);
    FILE *cf = single ? fopen((outdir + "/config.dat.mcs").c_str(), fmode) : nullptr;
    if (!tf || (single && !cf)) { fprintf(stderr, "jmm_run: cannot open output files in %s\n", outdir.c_str()); return 1; }
    if (!resume) { if (single) fputs(kThermoHeader, tf); else fprintf(tf, "chain\t%s", kThermoHeader); }

    // --layout runjobs: one directory per state point (scripts/RunJobs.bash:27), replicas beyond the first get _r<k>
    std::vector<std::string> point_dir;
    if (!layout.empty()) {
        const uint64_t reps = std::max<uint64_t>(1, total / (np * nt));
        point_dir.resize(C);
        for (uint64_t c = 0; c < C; ++c) {
            const uint64_t rep = (c0 + c) % reps;
            point_dir[c] = outdir + "/data/" + deck.pot_str + "/m" + std::to_string(cfg.nbn) + "/N" + std::to_string(cfg.N) + "/P" +
                           short_num(P[c]) + "_T" + short_num(T[c]) + (rep ? "_r" + std::to_string(rep) : "");
            if (resume) continue;
            mkdirs(point_dir[c]);
            FILE *in = fopen((point_dir[c] + "/INPUT").c_str(), "w"), *src = fopen(input.c_str(), "r");
            if (!in || !src) { fprintf(stderr, "jmm_run: cannot write %s/INPUT\n", point_dir[c].c_str()); return 1; }
            char line[512];
            while (fgets(line, sizeof line, src)) {                  // the deck with this point's P and T
                char key[64] = "";
                sscanf(line, " %63s", key);
                if (!strcmp(key, "P")) fprintf(in, "P          %s\n", short_num(P[c]).c_str());
                else if (!strcmp(key, "T")) fprintf(in, "T          %s\n", short_num(T[c]).c_str());
                else fputs(line, in);
            }
            fclose(in); fclose(src);
            FILE *t = fopen((point_dir[c] + "/thermo.dat.mcs").c_str(), "w");
            if (t) { fputs(kThermoHeader, t); fclose(t); }
        }
    }

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
            if (!point_dir.empty()) {                                    // the reference's own 13-column row, one file per point
                FILE *t = fopen((point_dir[c] + "/thermo.dat.mcs").c_str(), "a");
                if (t) {
                    fprintf(t, "%lu\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\t%.8G\n", (unsigned long) sn,
                            a[JMM_A_E] / ss, a[JMM_A_E2] / ss, a[JMM_A_L] / ss, a[JMM_A_L2] / ss, a[JMM_A_LE] / ss, a[JMM_A_RHO] / ss,
                            a[JMM_A_RHO2] / ss, a[JMM_A_VIR] / ss, a[JMM_A_VIR2] / ss, a[JMM_A_EVIR] / ss, a[JMM_A_HV] / ss,
                            a[JMM_A_HV2] / ss);
                    fclose(t);
                }
            }
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

    if (resume) {
        // exact continuation: everything the next step depends on comes from the file (include/jmm_gpu.h, N4)
        JCK(jmm_checkpoint_load(h, ckpt_in.c_str()));
        sn = jmm_step_number(h);
        // the sums in the file run since the last thermo row before the checkpoint: sltp = the last TPI multiple
        sltp = deck.tpi ? sn - sn % deck.tpi : 0;
        if (deck.rhopi) slrho = sn - sn % deck.rhopi;
        if (deck.gpi) slg = sn - sn % deck.gpi;
        printf("Resumed from %s at step %lu\n", ckpt_in.c_str(), (unsigned long) sn);
    } else {
        printf("Step: 0...\n");
        JCK(jmm_start(h));                                               // src/Main.cpp:66-96
        if (print_coords()) return 2;
        if (print_rho()) return 2;
        if (print_thermo()) return 2;
        if (print_g()) return 2;
    }
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
    if (!ckpt_out.empty()) {
        JCK(jmm_checkpoint_save(h, (ckpt_out + rank_tag).c_str()));
        printf("Checkpoint written: %s%s (step %lu)\n", ckpt_out.c_str(), rank_tag.c_str(), (unsigned long) sn);
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
        const int order[12] = {JMM_A_E, JMM_A_E2, JMM_A_L, JMM_A_L2, JMM_A_LE, JMM_A_RHO, JMM_A_RHO2, JMM_A_VIR, JMM_A_VIR2,
                               JMM_A_EVIR, JMM_A_HV, JMM_A_HV2};
        // one row per chain; the run means are sums over ALL thermo rows (kept on the host in `run`) / samples
        std::vector<double> rows;                                        // [chains][21] as printed below
        auto add_row = [&](double id, double p, double t, const double *means, const uint64_t *k4, double e, double lf) {
            rows.push_back(id); rows.push_back(p); rows.push_back(t); rows.push_back((double) cfg.N); rows.push_back((double) samples);
            for (int k = 0; k < 12; ++k) rows.push_back(means[order[k]]);
            const double d = (double) (k4[0] + k4[1]), v = (double) (k4[2] + k4[3]);
            rows.push_back(d > 0 ? k4[0] / d : 0.0); rows.push_back(v > 0 ? k4[2] / v : 0.0); rows.push_back(e); rows.push_back(lf);
        };
        std::vector<double> means(12);
        for (uint64_t c = 0; c < C; ++c) {
            for (int k = 0; k < 12; ++k) means[k] = samples ? run[c * 12 + k] / (double) samples : 0.0;
            add_row((double) (c0 + c), P[c], T[c], means.data(), &cnt[c * 4], tot[c * 9], l[c]);
        }
        if (world > 1) {
            // The only exchange of the job: the run sums go back into the handle's accumulators so that
            // the device-side records carry the whole run, then ONE ncclAllGather; rank 0 writes the merged table.
            uint8_t id[JMM_COMM_ID_BYTES];
            if (rank == 0) {
                JCK(jmm_comm_unique_id(id));
                FILE *f = fopen((id_file + ".tmp").c_str(), "wb");
                if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "jmm_run: cannot write %s\n", id_file.c_str()); return 1; }
                fclose(f);
                rename((id_file + ".tmp").c_str(), id_file.c_str());
            } else {
                bool ok = false;
                for (int tries = 0; tries < 6000 && !ok; ++tries) {      // up to 10 minutes: ranks finish at different times
                    FILE *f = fopen(id_file.c_str(), "rb");
                    if (f) { ok = fread(id, 1, sizeof id, f) == sizeof id; fclose(f); }
                    if (!ok) usleep(100000);
                }
                if (!ok) { fprintf(stderr, "jmm_run: rank %lu never saw %s\n", (unsigned long) rank, id_file.c_str()); return 1; }
            }
            jmm_comm *comm = nullptr;
            JCK(jmm_comm_create(id, (int32_t) rank, (int32_t) world, cfg.device, &comm));
            std::vector<double> table(total * JMM_SUMMARY_DOUBLES);
            JCK(jmm_set_accum(h, run.data(), samples));
            JCK(jmm_allgather_summaries(h, comm, total, table.data()));
            JCK(jmm_comm_destroy(comm));
            rows.clear();
            for (uint64_t g = 0; g < total; ++g) {
                const double *r = &table[g * JMM_SUMMARY_DOUBLES];
                for (int k = 0; k < 12; ++k) means[k] = r[4] > 0 ? r[5 + k] / r[4] : 0.0;
                const uint64_t k4[4] = {(uint64_t) r[17], (uint64_t) r[18], (uint64_t) r[19], (uint64_t) r[20]};
                add_row(r[0], r[1], r[2], means.data(), k4, r[22], r[21]);
            }
            if (rank == 0) { printf("Summary: %lu chains gathered from %lu ranks with one ncclAllGather (NCCL %d)\n", (unsigned long) total,
                                    (unsigned long) world, (int) jmm_nccl_version()); unlink(id_file.c_str()); }
        }
        if (rank == 0) {
            FILE *sf = fopen((outdir + "/Summary.dat").c_str(), "w");
            if (!sf) { fprintf(stderr, "jmm_run: cannot write %s/Summary.dat\n", outdir.c_str()); return 1; }
            fprintf(sf, "chain\tP\tT\tN\tsamples\tEconf\tEconf2\tL\tL2\tLEconf\trho\trho2\tVirial\tVirial2\tEconfVir\tHV\tHV2\t"
                        "dAccRatio\tvAccRatio\tEfinal\tLfinal\n");
            for (size_t i = 0; i + 21 <= rows.size(); i += 21) {
                const double *r = &rows[i];
                fprintf(sf, "%lu\t%.8G\t%.8G\t%lu\t%lu", (unsigned long) r[0], r[1], r[2], (unsigned long) r[3], (unsigned long) r[4]);
                for (int k = 0; k < 12; ++k) fprintf(sf, "\t%.8G", r[5 + k]);
                fprintf(sf, "\t%.6G\t%.6G\t%.8G\t%.8G\n", r[17], r[18], r[19], r[20]);
            }
            fclose(sf);
        }
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
