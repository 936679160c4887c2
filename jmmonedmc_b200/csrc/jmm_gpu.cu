// libjmmgpu.so — C ABI of the B200-native jmmOneDMC hot path (see include/jmm_gpu.h).
// Host-side orchestration only; the arithmetic is in chains.cuh / sweep.cuh / pot.cuh / rng.cuh.
// There is no CPU fallback anywhere in this file: without a usable CUDA device every entry point
// that computes returns JMM_ERR_CUDA.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "handle.h"

using namespace jmm;

static thread_local std::string g_err;

static jmm_status fail(jmm_status code, const std::string &msg) {
    g_err = msg;
    return code;
}
jmm_status jmm_fail(jmm_status code, const std::string &msg) { return fail(code, msg); }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(JMM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

template <class T>
static cudaError_t dalloc(jmm_handle *h, T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **) p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) {
        h->allocs.push_back(*p);
        e = cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream);
    }
    return e;
}

static jmm_status ensure_stage(jmm_handle *h, size_t bytes) {
    if (h->stage_bytes >= bytes) return JMM_OK;
    if (h->d_stage) CK(cudaFree(h->d_stage));
    h->d_stage = nullptr; h->stage_bytes = 0;
    CK(cudaMalloc((void **) &h->d_stage, bytes));
    h->stage_bytes = bytes;
    return JMM_OK;
}

static bool is_cb(const jmm_handle *h) { return h->cfg.mode == JMM_MODE_CHECKERBOARD; }

static void tick(jmm_handle *h) { if (!h->timed) { cudaEventRecord(h->ev0, h->stream); h->timed = true; } }
static void tock(jmm_handle *h) { cudaEventRecord(h->ev1, h->stream); }

// ------------------------------------------------------------------------------------------------
// readInput(), src/readInput.cpp:8-262
// ------------------------------------------------------------------------------------------------
extern "C" jmm_status jmm_read_input(const char *path, jmm_config *cfg, jmm_deck *deck) {
    if (!path || !cfg) return fail(JMM_ERR_INVALID, "jmm_read_input: null argument");
    FILE *f = fopen(path, "r");
    if (!f) return fail(JMM_ERR_IO, std::string("FATAL ERROR: INPUT not found: ") + path);   // :51-54
    jmm_deck dk;
    memset(&dk, 0, sizeof dk);
    memset(cfg, 0, sizeof *cfg);
    cfg->cutoff = INFINITY;
    cfg->nbn = -1;                                     // no NBN line = no neighbour limit (the reference leaves nbn
                                                       // uninitialised, src/readInput.cpp:103; 0 would exclude every pair)
    cfg->ensemble = JMM_ENS_NPT;                       // default "NPT", :57
    cfg->nchains = 1;
    cfg->rng_kind = JMM_RNG_PHILOX;
    cfg->mode = JMM_MODE_RECOMPUTE;
    cfg->adapt = JMM_ADAPT_DEVICE;
    strncpy(dk.ensemble_str, "NPT", sizeof dk.ensemble_str - 1);
    char line[512];
    const char *delim = "[ \t\n]";                     // :43 (the brackets are delimiters too)
    while (fgets(line, sizeof line, f)) {
        char *tok[32];
        int n = 0;
        for (char *t = strtok(line, delim); t && n < 31; t = strtok(nullptr, delim)) tok[n++] = t;
        tok[n] = nullptr;
        if (n == 0) continue;                          // the reference dereferences NULL here (:69)
        const std::string k = tok[0];
        auto num = [&](int i) { return (i < n) ? strtod(tok[i], nullptr) : 0.0; };
        if (k == "RESTART") {
            dk.is_restart = 1;
            if (n > 1 && (!strcmp(tok[1], "FALSE") || !strcmp(tok[1], "False") || !strcmp(tok[1], "false"))) dk.is_restart = 0;
        } else if (k == "N") cfg->N = (uint64_t) num(1);
        else if (k == "P") cfg->P = num(1);
        else if (k == "L") cfg->L = num(1);
        else if (k == "T") cfg->T = num(1);
        else if (k == "NBN") cfg->nbn = (int32_t) num(1);
        else if (k == "NUMSTEPS") dk.numsteps = (uint64_t) num(1);
        else if (k == "POT") {                         // :108-131
            const char *name = n > 1 ? tok[1] : "";
            if (!strcmp(name, "LJcut")) cfg->pot = JMM_POT_LJCUT;
            else if (!strcmp(name, "LJ")) cfg->pot = JMM_POT_LJ;
            else if (!strcmp(name, "HARMONIC")) cfg->pot = JMM_POT_HARMONIC;
            else {
                printf("Unknown potential type requested (%s). Using Lennard-Jones potential with no cut-off.\n", name);
                cfg->pot = JMM_POT_LJ; name = "LJ";
            }
            strncpy(dk.pot_str, name, sizeof dk.pot_str - 1);
            cfg->cutoff = (n > 2) ? strtod(tok[2], nullptr) : INFINITY;
        } else if (k == "MAXSTEP") cfg->maxStep = num(1);
        else if (k == "MAXDV") cfg->maxdl = num(1);
        else if (k == "CPI") dk.cpi = (uint64_t) num(1);
        else if (k == "TPI") dk.tpi = (uint64_t) num(1);
        else if (k == "GPI") dk.gpi = (uint64_t) num(1);
        else if (k == "RHOPI") dk.rhopi = (uint64_t) num(1);
        else if (k == "RBW") dk.rbw = num(1);
        else if (k == "RHONB") dk.rhonb = (uint64_t) num(1);
        else if (k == "GBW") dk.gbw = num(1);
        else if (k == "GNB") dk.gnb = (uint64_t) num(1);
        else if (k == "GSW") dk.gsw = num(1);
        else if (k == "GNS") dk.gns = (int32_t) num(1);
        else if (k == "SEED") cfg->seed = (uint64_t) num(1);
        else if (k == "ENGCHECK") cfg->eci = (uint64_t) num(1);
        else if (k == "DADJ") cfg->mdai = (uint64_t) num(1);
        else if (k == "VADJ") cfg->mvai = (uint64_t) num(1);
        else if (k == "RELAX") cfg->relax = 1;
        else if (k == "ENSEMBLE") {                    // :201-226
            const char *name = n > 1 ? tok[1] : "";
            if (!strcmp(name, "NPT")) cfg->ensemble = JMM_ENS_NPT;
            else if (!strcmp(name, "NLT")) cfg->ensemble = JMM_ENS_NLT;
            else {
                printf("Unknown ensemble type requested (%s). Using NPT ensemble.\n", name);
                cfg->ensemble = JMM_ENS_NPT; name = "NPT";
            }
            strncpy(dk.ensemble_str, name, sizeof dk.ensemble_str - 1);
        } else if (k == "COMPUTE" || k == "compute") {
            // `Compute` objects are constructed and never invoked by the reference: ignored
        } else if (k == "NCHAINS") cfg->nchains = (uint64_t) num(1);   // extension keywords: the reference
        else if (k == "GPU") cfg->device = (int32_t) num(1);           // answers these with "not understood"
        else if (k == "RNG") {
            if (n > 1 && !strcmp(tok[1], "TAUS2")) cfg->rng_kind = JMM_RNG_TAUS2;
            else cfg->rng_kind = JMM_RNG_PHILOX;
        } else {
            printf("Property command %s not understood.\n", tok[0]);   // :254-256
            dk.n_unknown++;
        }
    }
    fclose(f);
    if (cfg->pot == JMM_POT_LJ) cfg->cutoff = INFINITY;                // src/jmmMCState.cpp:294
    if (deck) *deck = dk;
    return JMM_OK;
}

// ------------------------------------------------------------------------------------------------
// setupMCS / freeMCS
// ------------------------------------------------------------------------------------------------
static jmm_status validate(const jmm_config *c) {
    if (c->N < 2) return fail(JMM_ERR_INVALID, "N must be >= 2");
    if (c->N > 0x7fffffffull) return fail(JMM_ERR_INVALID, "N too large");
    if (c->nchains < 1) return fail(JMM_ERR_INVALID, "nchains must be >= 1");
    if (c->nbn == 0) return fail(JMM_ERR_INVALID, "NBN 0 excludes every pair (use NBN -1 for no neighbour limit)");
    if (!(c->T > 0)) return fail(JMM_ERR_INVALID, "T must be > 0");
    if (c->pot < JMM_POT_LJ || c->pot > JMM_POT_HARMONIC)
        return fail(JMM_ERR_UNKNOWN_POT, "FATAL ERROR: Unknown potential.");
    if (c->ensemble != JMM_ENS_NPT && c->ensemble != JMM_ENS_NLT)
        return fail(JMM_ERR_UNKNOWN_ENS, "FATAL ERROR: Unknown ensemble.");
    if (c->rng_kind < JMM_RNG_TAUS2 || c->rng_kind > JMM_RNG_RECORDED) return fail(JMM_ERR_INVALID, "bad rng_kind");
    if (c->mode < JMM_MODE_TABLE || c->mode > JMM_MODE_CHECKERBOARD) return fail(JMM_ERR_INVALID, "bad mode");
    if (c->adapt < JMM_ADAPT_HOST || c->adapt > JMM_ADAPT_CALLER) return fail(JMM_ERR_INVALID, "bad adapt");
    if (c->arith != JMM_ARITH_REFERENCE && c->arith != JMM_ARITH_FAST) return fail(JMM_ERR_INVALID, "bad arith");
    if (c->flags & ~JMM_FLAG_CONSISTENT_VIRIAL) return fail(JMM_ERR_INVALID, "unknown bits in jmm_config.flags");
    if ((c->flags & JMM_FLAG_CONSISTENT_VIRIAL) && (c->mode == JMM_MODE_TABLE || c->rng_kind != JMM_RNG_PHILOX))
        return fail(JMM_ERR_INVALID, "JMM_FLAG_CONSISTENT_VIRIAL is a production option (Philox stream, RECOMPUTE or CHECKERBOARD mode); "
                                     "lock-step reproduces the reference's running virial as it is");
    if (c->arith == JMM_ARITH_FAST && (c->pot == JMM_POT_HARMONIC || c->mode == JMM_MODE_TABLE || c->rng_kind != JMM_RNG_PHILOX))
        return fail(JMM_ERR_INVALID, "JMM_ARITH_FAST is for LJ/LJcut with the Philox stream (RECOMPUTE or CHECKERBOARD mode)");
    if (c->rng_kind == JMM_RNG_RECORDED && c->nchains != 1)
        return fail(JMM_ERR_INVALID, "a recorded stream drives exactly one chain");
    if (c->mode == JMM_MODE_CHECKERBOARD) {
        if (c->nbn < 1) return fail(JMM_ERR_INVALID, "checkerboard sweeps need NBN >= 1 (colours = NBN+1)");
        if (c->ensemble != JMM_ENS_NLT) return fail(JMM_ERR_INVALID, "checkerboard sweeps are NLT (fixed L) only");
        if (c->rng_kind != JMM_RNG_PHILOX) return fail(JMM_ERR_INVALID, "checkerboard sweeps need the Philox stream");
        if (c->N >= 0xffffffffull) return fail(JMM_ERR_INVALID, "N too large");
    }
    if (c->mode == JMM_MODE_TABLE) {
        const double bytes = 8.0 * (double) c->nchains * (double) c->N * (double) (c->N - 1) / 2.0;
        if (bytes > 64e9) return fail(JMM_ERR_INVALID, "rij table would exceed 64 GB; use JMM_MODE_RECOMPUTE");
    }
    return JMM_OK;
}

extern "C" jmm_status jmm_destroy(jmm_handle *h) {
    if (!h) return JMM_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void *p : h->allocs) cudaFree(p);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->d_log) cudaFree(h->d_log);
    if (h->d_partial) cudaFree(h->d_partial);
    if (h->d_stream) cudaFree(h->d_stream);
    if (h->d_work) cudaFree(h->d_work);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return JMM_OK;
}

extern "C" jmm_status jmm_create(const jmm_config *cfg, jmm_handle **out) {
    if (!cfg || !out) return fail(JMM_ERR_INVALID, "jmm_create: null argument");
    *out = nullptr;
    jmm_status st = validate(cfg);
    if (st != JMM_OK) return st;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(JMM_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(JMM_ERR_INVALID, "device ordinal out of range");
    CK(cudaSetDevice(cfg->device));

    jmm_handle *h = new jmm_handle;
    h->cfg = *cfg;
    if (h->cfg.pot == JMM_POT_LJ) h->cfg.cutoff = INFINITY;
    if (h->cfg.ensemble == JMM_ENS_NLT) h->cfg.relax = 0;           // src/jmmMCState.cpp:416-420
    auto bail = [&](jmm_status s) { jmm_destroy(h); return s; };
#define CKH(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) { fail(JMM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); return bail(JMM_ERR_CUDA); } \
    } while (0)
    CKH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    CKH(cudaEventCreate(&h->ev0));
    CKH(cudaEventCreate(&h->ev1));

    const uint64_t C = cfg->nchains, N = cfg->N;
    ChainsDev &S = h->S;
    S.nchains = C; S.N = N; S.npairs = N * (N - 1) / 2;
    S.numTrialTypes = (cfg->ensemble == JMM_ENS_NPT) ? N + 1 : N;   // :405,415
    S.nbn = cfg->nbn; S.ensemble = cfg->ensemble; S.relax = h->cfg.relax; S.pot = cfg->pot;
    S.cutoff = h->cfg.cutoff; S.seed = cfg->seed; S.chain_id0 = cfg->chain_id0; S.flags = cfg->flags;
    CKH(dalloc(h, &S.l, C)); CKH(dalloc(h, &S.P, C)); CKH(dalloc(h, &S.T, C));
    CKH(dalloc(h, &S.maxStep, C)); CKH(dalloc(h, &S.maxdl, C));

    const double l0 = (cfg->ensemble == JMM_ENS_NPT) ? (double) N : cfg->L;   // :553, :414
    {   // five per-chain constants: one staging buffer each, one synchronisation
        std::vector<double> v5(5 * C);
        const double vals[5] = {l0, cfg->P, cfg->T, cfg->maxStep, cfg->maxdl};
        double *dst[5] = {S.l, S.P, S.T, S.maxStep, S.maxdl};
        for (int k = 0; k < 5; ++k) {
            std::fill(v5.begin() + k * C, v5.begin() + (k + 1) * C, vals[k]);
            CKH(cudaMemcpyAsync(dst[k], v5.data() + k * C, C * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        }
        CKH(cudaStreamSynchronize(h->stream));
    }

    if (is_cb(h)) {
        CKH(dalloc(h, &h->cb_r[0], C * N)); CKH(dalloc(h, &h->cb_r[1], C * N));
        CKH(dalloc(h, &h->cb_tot, C * 9)); CKH(dalloc(h, &h->cb_acc, C * 12));
        CKH(dalloc(h, &h->cb_counts, C * 2));
        CKH(dalloc(h, &h->cb_tile_done, C));
        // lattice, chain-major: same formula as :561, every chain in one launch
        k_lattice_chain_major<<<nblk(C * N, 256), 256, 0, h->stream>>>(h->cb_r[0], S.l, C, N);
        h->launches++;
        CKH(cudaGetLastError());
    } else {
        CKH(dalloc(h, &S.r, C * N));
        CKH(dalloc(h, &S.tot, C * 9)); CKH(dalloc(h, &S.acc, C * 12));
        CKH(dalloc(h, &S.cnt, C * 4)); CKH(dalloc(h, &S.vAErr, C)); CKH(dalloc(h, &S.echeck, C * 2));
        CKH(dalloc(h, &S.taus, C * 3));
        if (cfg->mode == JMM_MODE_TABLE) CKH(dalloc(h, &S.rij, C * S.npairs));
        CKH(dalloc(h, &h->d_cursor, 1)); CKH(dalloc(h, &h->d_err, 1));
        // sentinels of :302-307, :542-544
        std::vector<double> t(C * 9);
        const double init[9] = {10E10, 10E10, 5E10, 5E10, -5E10, -5E10, 10E10, 5E10, -5E10};
        for (int k = 0; k < 9; ++k) std::fill(t.begin() + k * C, t.begin() + (k + 1) * C, init[k]);
        CKH(cudaMemcpyAsync(S.tot, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CKH(cudaStreamSynchronize(h->stream));
        // taus2 state per chain, gsl_rng_set(seed) :781; chain c uses seed + chain_id0 + c
        std::vector<uint32_t> ts(C * 3);
        for (uint64_t c = 0; c < C; ++c) {
            uint32_t a, b, d;
            taus2_seed(cfg->seed + cfg->chain_id0 + c, a, b, d);
            ts[c] = a; ts[C + c] = b; ts[2 * C + c] = d;
        }
        CKH(cudaMemcpyAsync(S.taus, ts.data(), ts.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        CKH(cudaStreamSynchronize(h->stream));
        k_lattice<<<nblk(C * N, 256), 256, 0, h->stream>>>(S.r, S.l, C, N);
        h->launches++;
        CKH(cudaGetLastError());
        // one warp per block spreads few chains over as many SMs as possible; positions live in
        // shared memory when an [N][32] tile fits
        h->block = 32;
        h->smem = (size_t) N * h->block * sizeof(double);
        h->pos_in_smem = h->smem <= 200 * 1024;
        if (!h->pos_in_smem) h->smem = 0;
        // Few chains: let G lanes share a chain (coop.cuh) so that ~4096+ warps are in flight.
        if (cfg->mode == JMM_MODE_RECOMPUTE && cfg->rng_kind == JMM_RNG_PHILOX) {
            int g = 0;
            // measured on the C2 workload (4096 chains): G=16 2.84e9, G=32 1.95e9, G=8 1.75e9, thread/chain 0.89e9 trials/s
            if (C <= 8192) g = 16; else if (C <= 16384) g = 8;
            if (const char *e = getenv("JMM_COOP_G")) g = atoi(e);
            if (g == 8 || g == 16 || g == 32) {
                const int nc = (cfg->pot == JMM_POT_HARMONIC) ? 2 : 9;
                const int scratch = kCoopScratchRows * 2 * nc;
                int npad = (int) N;
                npad += (17 - ((npad + scratch) % 16)) % 16;          // rows of different groups hit different banks
                const size_t bytes = (size_t) (128 / g) * (npad + scratch) * sizeof(double);
                if (bytes <= 200 * 1024) { h->coop_g = g; h->coop_npad = npad; h->coop_smem = bytes; }
            }
            // JMM_ARITH_FAST with too few chains for one chain per thread (prod.cuh wants >= ~2048 warps = 65 536
            // chains): G lanes per chain with the same arithmetic (lanes.cuh), G = the power of two that brings the
            // launch to ~2048 warps, at most one lane per ~partner.  (A sweep of 65 536 state points sharded over 8
            // GPUs is 8192 chains per GPU: G = 8.)
            if (cfg->arith == JMM_ARITH_FAST && cfg->pot != JMM_POT_HARMONIC) {
                const uint64_t partners = cfg->nbn < 0 ? N - 1 : std::min<uint64_t>(2 * (uint64_t) cfg->nbn, N - 1);
                int lg = 1;
                while (lg < 32 && (uint64_t) (2 * lg) * C <= 65536 && (uint64_t) (2 * lg) <= std::max<uint64_t>(2, partners)) lg *= 2;
                if (const char *e = getenv("JMM_LANES_G")) lg = atoi(e);
                h->coop_g = 0;                                       // never the reference-arithmetic kernel for a FAST handle
                if (lg == 2 || lg == 4 || lg == 8 || lg == 16 || lg == 32) jmm_lanes_shape(h, lg);
            }
            // nearest-neighbour bond chains (the INPUTstd shape): branch-free kernel, one chain per 16 lanes
            const char *eb = getenv("JMM_BOND");
            if (h->coop_g && cfg->pot == JMM_POT_HARMONIC && cfg->nbn == 1 && N - 1 <= 16 && !(h->cfg.relax > 0) &&
                C <= 16384 && !(eb && atoi(eb) == 0))
            {
                // 5 = k_chains_step_crew (solo.cuh: five warps per 32 chains — displacement | volume | Philox | virial + sums |
                //     ECheck; 1.13e10 trial moves/s on C2) — the default while every 32-chain CTA is resident at once (two per
                //     SM); a deck without volume trials (NLT) runs 4 instead,
                // 4 = k_chains_step_trio (three warps per 32 chains: trials | Philox | virial + ECheck + sums; 9.15e9),
                // 3 = k_chains_step_solo (one warp per 32 chains doing all of it: 4.97e9),
                // 2 = k_chains_step_bond2 (registers-only, deferred ECheck: 3.99e9, profiles/r2e_c2_*),
                // 1 = k_chains_step_bond (16 lanes per chain: 4.47e9; also what repeats a launch of 4 / 5 after an energy
                //     discrepancy, and what more than 2 x SMs x 32 chains get).  All parity-tested.
                int nsm = 0;
                cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, cfg->device);
                const int fallback = (C + 31) / 32 <= (uint64_t) 2 * (uint64_t) std::max(nsm, 1) ? 5 : 1;
                h->bond = (eb && atoi(eb) >= 1 && atoi(eb) <= 5) ? atoi(eb) : fallback;
            }
        }
    }
    CKH(cudaStreamSynchronize(h->stream));
#undef CKH
    *out = h;
    return JMM_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel dispatch over (POT, TABLE, RNG)
// ------------------------------------------------------------------------------------------------
template <int POT, bool TABLE>
static cudaError_t launch_start(jmm_handle *h, bool relax_only, int parts = kStartFad | kStartRelax | kStartThermo) {
    auto kern = relax_only ? k_chains_relax<POT, TABLE> : k_chains_start<POT, TABLE>;
    if (h->smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<nblk(h->S.nchains, h->block), h->block, h->smem, h->stream>>>(h->S, h->pos_in_smem, parts);
    h->launches++;
    return cudaGetLastError();
}

template <int POT, bool TABLE, int RNG>
static cudaError_t launch_step_rng(jmm_handle *h, const StepArgs &a) {
    auto kern = k_chains_step<POT, TABLE, RNG>;
    if (h->smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<nblk(h->S.nchains, h->block), h->block, h->smem, h->stream>>>(h->S, a, h->H);
    h->launches++;
    return cudaGetLastError();
}

template <int POT, bool TABLE>
static cudaError_t launch_step_table(jmm_handle *h, const StepArgs &a) {
    if constexpr (!TABLE) {
        if (h->lanes_g) return jmm_launch_lanes(h, a);
        if (h->coop_g) return jmm_launch_coop(h, a);
        if (h->cfg.rng_kind == JMM_RNG_PHILOX && h->pos_in_smem && h->block == 32 && !getenv("JMM_NO_PROD"))
            return jmm_launch_prod(h, a);
    }
    switch (h->cfg.rng_kind) {
        case JMM_RNG_TAUS2: return launch_step_rng<POT, TABLE, kRngTaus2>(h, a);
        case JMM_RNG_PHILOX: return launch_step_rng<POT, TABLE, kRngPhilox>(h, a);
        default: return launch_step_rng<POT, TABLE, kRngRecorded>(h, a);
    }
}

#define DISPATCH_POT_TABLE(h, FN, ...)                                                              \
    ([&]() -> cudaError_t {                                                                         \
        const bool tbl = (h)->cfg.mode == JMM_MODE_TABLE;                                           \
        switch ((h)->cfg.pot) {                                                                     \
            case JMM_POT_LJ: return tbl ? FN<kPotLJ, true>(__VA_ARGS__) : FN<kPotLJ, false>(__VA_ARGS__);               \
            case JMM_POT_LJCUT: return tbl ? FN<kPotLJcut, true>(__VA_ARGS__) : FN<kPotLJcut, false>(__VA_ARGS__);      \
            default: return tbl ? FN<kPotHarmonic, true>(__VA_ARGS__) : FN<kPotHarmonic, false>(__VA_ARGS__);           \
        }                                                                                           \
    })()

// ------------------------------------------------------------------------------------------------
// state in / out
// ------------------------------------------------------------------------------------------------
static jmm_status upload_transposed(jmm_handle *h, const double *src, double *dst, uint64_t n) {
    const uint64_t C = h->S.nchains;
    jmm_status st = ensure_stage(h, C * n * sizeof(double));
    if (st != JMM_OK) return st;
    CK(cudaMemcpyAsync(h->d_stage, src, C * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_transpose_in<<<nblk(C * n, 256), 256, 0, h->stream>>>(h->d_stage, dst, C, n);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));          // the caller may reuse src right away
    return JMM_OK;
}

static jmm_status download_transposed(jmm_handle *h, const double *src, double *dst, uint64_t n) {
    const uint64_t C = h->S.nchains;
    jmm_status st = ensure_stage(h, C * n * sizeof(double));
    if (st != JMM_OK) return st;
    k_transpose_out<<<nblk(C * n, 256), 256, 0, h->stream>>>(src, h->d_stage, C, n);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, h->d_stage, C * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return JMM_OK;
}

extern "C" jmm_status jmm_set_state(jmm_handle *h, const double *r, const double *l, const double *P, const double *T) {
    if (!h) return fail(JMM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains, N = h->S.N;
    if (r) {
        if (is_cb(h)) {
            CK(cudaMemcpyAsync(h->cb_r[h->cb_cur], r, C * N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        } else {
            jmm_status st = upload_transposed(h, r, h->S.r, N);
            if (st != JMM_OK) return st;
        }
    }
    if (l) CK(cudaMemcpyAsync(h->S.l, l, C * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (P) CK(cudaMemcpyAsync(h->S.P, P, C * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (T) CK(cudaMemcpyAsync(h->S.T, T, C * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return JMM_OK;
}

extern "C" jmm_status jmm_set_step_sizes(jmm_handle *h, const double *maxStep, const double *maxdl) {
    if (!h) return fail(JMM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains;
    if (maxStep) CK(cudaMemcpyAsync(h->S.maxStep, maxStep, C * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (maxdl) CK(cudaMemcpyAsync(h->S.maxdl, maxdl, C * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return JMM_OK;
}

extern "C" jmm_status jmm_get_step_sizes(jmm_handle *h, double *maxStep, double *maxdl) {
    if (!h) return fail(JMM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains;
    if (maxStep) CK(cudaMemcpyAsync(maxStep, h->S.maxStep, C * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (maxdl) CK(cudaMemcpyAsync(maxdl, h->S.maxdl, C * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return JMM_OK;
}

extern "C" jmm_status jmm_get_state(jmm_handle *h, double *r, double *l, double *totals, double *accum, uint64_t *counters) {
    if (!h) return fail(JMM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains, N = h->S.N;
    jmm_status st;
    if (is_cb(h)) {
        if (r) CK(cudaMemcpyAsync(r, h->cb_r[h->cb_cur], C * N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (totals) CK(cudaMemcpyAsync(totals, h->cb_tot, C * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (accum) CK(cudaMemcpyAsync(accum, h->cb_acc, C * 12 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (counters) {
            std::vector<unsigned long long> c2(C * 2);
            CK(cudaMemcpyAsync(c2.data(), h->cb_counts, C * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            for (uint64_t c = 0; c < C; ++c) {
                counters[4 * c + 0] = c2[2 * c]; counters[4 * c + 1] = c2[2 * c + 1] - c2[2 * c];
                counters[4 * c + 2] = 0; counters[4 * c + 3] = 0;
            }
        }
    } else {
        if (r && (st = download_transposed(h, h->S.r, r, N)) != JMM_OK) return st;
        if (totals && (st = download_transposed(h, h->S.tot, totals, 9)) != JMM_OK) return st;
        if (accum && (st = download_transposed(h, h->S.acc, accum, 12)) != JMM_OK) return st;
        if (counters) {
            std::vector<uint64_t> t(C * 4);
            CK(cudaMemcpyAsync(t.data(), h->S.cnt, C * 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            for (uint64_t c = 0; c < C; ++c)
                for (int k = 0; k < 4; ++k) counters[4 * c + k] = t[k * C + c];
        }
    }
    if (l) CK(cudaMemcpyAsync(l, h->S.l, C * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return JMM_OK;
}

extern "C" jmm_status jmm_zero_accum(jmm_handle *h) {
    if (!h) return fail(JMM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    double *a = is_cb(h) ? h->cb_acc : h->S.acc;
    CK(cudaMemsetAsync(a, 0, h->S.nchains * 12 * sizeof(double), h->stream));
    h->samples = 0;
    return JMM_OK;
}

extern "C" jmm_status jmm_set_accum(jmm_handle *h, const double *accum, uint64_t samples) {
    if (!h || !accum) return fail(JMM_ERR_INVALID, "jmm_set_accum: null argument");
    CK(cudaSetDevice(h->cfg.device));
    if (is_cb(h)) {
        CK(cudaMemcpyAsync(h->cb_acc, accum, h->S.nchains * 12 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    } else {
        jmm_status st = upload_transposed(h, accum, h->S.acc, 12);
        if (st != JMM_OK) return st;
    }
    h->samples = samples;
    return JMM_OK;
}

extern "C" uint64_t jmm_step_number(const jmm_handle *h) { return h ? (is_cb(h) ? h->halfsweeps : h->sn) : 0; }

extern "C" jmm_status jmm_set_step_number(jmm_handle *h, uint64_t sn) {
    if (!h) return fail(JMM_ERR_INVALID, "null handle");
    if (is_cb(h)) return fail(JMM_ERR_INVALID, "jmm_set_step_number: many-chain handles only");
    h->sn = sn;
    return JMM_OK;
}
extern "C" uint64_t jmm_stream_cursor(const jmm_handle *h) { return h ? h->cursor : 0; }
extern "C" uint64_t jmm_kernel_launches(const jmm_handle *h) { return h ? h->launches : 0; }

extern "C" const char *jmm_engine(const jmm_handle *h) {          // mirrors launch_step_table / jmm_launch_coop / jmm_launch_lanes
    if (!h) return "";
    if (h->cfg.mode == JMM_MODE_CHECKERBOARD) return h->cfg.arith == JMM_ARITH_FAST ? "k_sweep_fast" : "k_sweep";
    if (h->cfg.mode == JMM_MODE_TABLE) return "k_chains_step<TABLE>";
    if (h->lanes_g) {
        const char *t = getenv("JMM_TEAM");
        if (t && atoi(t) != 0 && h->lanes_g == 8 && h->lanes_npl == 10) return "k_chains_step_team";
        switch (h->lanes_g) {
            case 2: return "k_chains_step_lanes<G=2>";
            case 4: return "k_chains_step_lanes<G=4>";
            case 8: return "k_chains_step_lanes<G=8>";
            case 16: return "k_chains_step_lanes<G=16>";
            default: return "k_chains_step_lanes<G=32>";
        }
    }
    if (h->coop_g) {
        if (h->cfg.pot == JMM_POT_HARMONIC && h->bond) {
            switch (h->bond) {
                case 5: return (uint64_t) h->S.numTrialTypes > h->S.N ? "k_chains_step_crew" : "k_chains_step_trio";
                case 4: return "k_chains_step_trio";
                case 3: return "k_chains_step_solo";
                case 2: return "k_chains_step_bond2";
                default: return "k_chains_step_bond";
            }
        }
        return "k_chains_step_coop";
    }
    if (h->cfg.rng_kind == JMM_RNG_PHILOX && h->pos_in_smem && h->block == 32 && !getenv("JMM_NO_PROD")) return "k_chains_step_prod";
    return "k_chains_step";
}

extern "C" double jmm_last_kernel_ms(const jmm_handle *hc) {
    jmm_handle *h = const_cast<jmm_handle *>(hc);
    if (!h || !h->ev0 || !h->ev1) return -1.0;
    cudaSetDevice(h->cfg.device);
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
    return (double) ms;
}

extern "C" jmm_status jmm_set_stream(jmm_handle *h, void *cuda_stream) {
    if (!h) return fail(JMM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->own_stream) CK(cudaStreamDestroy(h->stream));
    h->stream = (cudaStream_t) cuda_stream;
    h->own_stream = false;
    return JMM_OK;
}

extern "C" void *jmm_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void jmm_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" jmm_status jmm_echeck_stats(jmm_handle *h, uint64_t *checks, uint64_t *discrepancies) {
    if (!h || is_cb(h)) return fail(JMM_ERR_INVALID, "jmm_echeck_stats: many-chain handles only");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains;
    std::vector<uint64_t> t(2 * C);
    CK(cudaMemcpyAsync(t.data(), h->S.echeck, 2 * C * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    uint64_t a = 0, b = 0;
    for (uint64_t c = 0; c < C; ++c) { a += t[c]; b += t[C + c]; }
    if (checks) *checks = a;
    if (discrepancies) *discrepancies = b;
    return JMM_OK;
}

// ------------------------------------------------------------------------------------------------
// histograms
// ------------------------------------------------------------------------------------------------
extern "C" jmm_status jmm_enable_histograms(jmm_handle *h, uint64_t rhonb, double rbw, int32_t gns, uint64_t gnb, double gsw, double gbw) {
    if (!h || is_cb(h)) return fail(JMM_ERR_INVALID, "jmm_enable_histograms: many-chain handles only");
    if (h->H.ucount) return fail(JMM_ERR_INVALID, "histograms are already enabled");
    if (h->sn != 0) return fail(JMM_ERR_INVALID, "enable histograms before the first step");
    if (!(rbw > 0) || !(gsw > 0) || !(gbw > 0) || gns < 0) return fail(JMM_ERR_INVALID, "bad histogram geometry");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains, ng = (uint64_t) gns * gnb;
    HistDev H{};
    H.rhonb = rhonb; H.gnb = gnb; H.gns = gns; H.rbw = rbw; H.gsw = gsw; H.gbw = gbw;
    CK(dalloc(h, &H.rho, C * rhonb)); CK(dalloc(h, &H.g, C * ng));
    CK(dalloc(h, &H.ucount, C));
    h->H = H;
    h->coop_g = 0; h->bond = 0; h->lanes_g = 0;    // the per-thread kernels carry the histogram hooks
    k_hist_init<<<nblk(C, 32), 32, 0, h->stream>>>(h->S, h->H);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return JMM_OK;
}

extern "C" jmm_status jmm_take_histograms(jmm_handle *h, int64_t *rhoA, int64_t *gA) {
    if (!h || !h->H.ucount) return fail(JMM_ERR_INVALID, "histograms are not enabled on this handle");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains;
    auto take = [&](HistBin *b, uint64_t bins, int64_t *out) -> jmm_status {
        if (!out || bins == 0) return JMM_OK;
        jmm_status st = ensure_stage(h, C * bins * sizeof(long long));
        if (st != JMM_OK) return st;
        k_hist_take<<<nblk(C * bins, 256), 256, 0, h->stream>>>(b, h->H.ucount, bins, C, (long long *) h->d_stage);
        h->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out, h->d_stage, C * bins * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return JMM_OK;
    };
    jmm_status st = take(h->H.rho, h->H.rhonb, rhoA);
    if (st != JMM_OK) return st;
    return take(h->H.g, (uint64_t) h->H.gns * h->H.gnb, gA);
}

// ------------------------------------------------------------------------------------------------
// start / relax / adjust
// ------------------------------------------------------------------------------------------------
static jmm_status totals_parallel(jmm_handle *h, const double *r, uint64_t ps, uint64_t cs, double *out, uint64_t ks, uint64_t ocs);

// ------------------------------------------------------------------------------------------------
// exact restart (N4): one binary file with everything the next step depends on
// ------------------------------------------------------------------------------------------------
namespace {
struct CkptHeader {
    char magic[8];                 // "JMMCKPT2"
    uint64_t N, nchains, seed, chain_id0, sn, halfsweeps, cursor;
    int32_t mode, pot, nbn, ensemble, rng_kind, cb_cur, hist, gns;
    uint64_t rhonb, gnb;
    double rbw, gsw, gbw;
    // (format 2) what decides the cadence and the arithmetic of the continuation
    double cutoff;
    uint64_t eci, mdai, mvai;
    int32_t adapt, arith, relax, flags;
    uint64_t samples;
};

// the device arrays of a handle, in file order
template <class F>
jmm_status ckpt_arrays(jmm_handle *h, F &&io) {
    const uint64_t C = h->S.nchains, N = h->S.N;
    jmm_status st;
#define IO(ptr, count) if ((st = io((void *) (ptr), (size_t) (count) * sizeof(*(ptr)))) != JMM_OK) return st
    IO(h->S.l, C); IO(h->S.P, C); IO(h->S.T, C); IO(h->S.maxStep, C); IO(h->S.maxdl, C);
    if (is_cb(h)) {
        IO(h->cb_r[h->cb_cur], C * N); IO(h->cb_tot, C * 9); IO(h->cb_acc, C * 12); IO(h->cb_counts, C * 2);
    } else {
        IO(h->S.r, C * N); IO(h->S.tot, C * 9); IO(h->S.acc, C * 12); IO(h->S.cnt, C * 4); IO(h->S.vAErr, C);
        IO(h->S.echeck, C * 2); IO(h->S.taus, C * 3);
        if (h->S.rij) IO(h->S.rij, C * h->S.npairs);
    }
    if (h->H.ucount) {
        IO(h->H.rho, C * h->H.rhonb); IO(h->H.g, C * (uint64_t) h->H.gns * h->H.gnb); IO(h->H.ucount, C);
    }
#undef IO
    return JMM_OK;
}

CkptHeader ckpt_header(const jmm_handle *h) {
    CkptHeader k{};
    memcpy(k.magic, "JMMCKPT2", 8);
    k.N = h->S.N; k.nchains = h->S.nchains; k.seed = h->cfg.seed; k.chain_id0 = h->cfg.chain_id0;
    k.sn = h->sn; k.halfsweeps = h->halfsweeps; k.cursor = h->cursor;
    k.mode = h->cfg.mode; k.pot = h->cfg.pot; k.nbn = h->cfg.nbn; k.ensemble = h->cfg.ensemble; k.rng_kind = h->cfg.rng_kind;
    k.cb_cur = 0; k.hist = h->H.ucount ? 1 : 0; k.gns = h->H.gns; k.rhonb = h->H.rhonb; k.gnb = h->H.gnb;
    k.rbw = h->H.rbw; k.gsw = h->H.gsw; k.gbw = h->H.gbw;
    k.cutoff = h->cfg.cutoff; k.eci = h->cfg.eci; k.mdai = h->cfg.mdai; k.mvai = h->cfg.mvai;
    k.adapt = h->cfg.adapt; k.arith = h->cfg.arith; k.relax = h->cfg.relax; k.flags = h->cfg.flags; k.samples = h->samples;
    return k;
}
}  // namespace

extern "C" jmm_status jmm_checkpoint_save(jmm_handle *h, const char *path) {
    if (!h || !path) return fail(JMM_ERR_INVALID, "jmm_checkpoint_save: null argument");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    FILE *f = fopen(path, "wb");
    if (!f) return fail(JMM_ERR_IO, std::string("cannot write ") + path);
    const CkptHeader k = ckpt_header(h);
    bool ok = fwrite(&k, sizeof(k), 1, f) == 1;
    std::vector<char> buf;
    jmm_status st = ckpt_arrays(h, [&](void *d, size_t bytes) -> jmm_status {
        buf.resize(bytes);
        CK(cudaMemcpy(buf.data(), d, bytes, cudaMemcpyDeviceToHost));
        ok = ok && fwrite(buf.data(), 1, bytes, f) == bytes;
        return JMM_OK;
    });
    ok = (fclose(f) == 0) && ok;
    if (st != JMM_OK) return st;
    return ok ? JMM_OK : fail(JMM_ERR_IO, std::string("short write to ") + path);
}

extern "C" jmm_status jmm_checkpoint_load(jmm_handle *h, const char *path) {
    if (!h || !path) return fail(JMM_ERR_INVALID, "jmm_checkpoint_load: null argument");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaStreamSynchronize(h->stream));
    FILE *f = fopen(path, "rb");
    if (!f) return fail(JMM_ERR_IO, std::string("cannot read ") + path);
    CkptHeader k{};
    if (fread(&k, sizeof(k), 1, f) != 1 || memcmp(k.magic, "JMMCKPT2", 8) != 0) {
        fclose(f);
        return fail(JMM_ERR_IO, std::string(path) + " is not a jmm checkpoint");
    }
    const CkptHeader me = ckpt_header(h);
    if (k.N != me.N || k.nchains != me.nchains || k.seed != me.seed || k.chain_id0 != me.chain_id0 || k.mode != me.mode ||
        k.pot != me.pot || k.nbn != me.nbn || k.ensemble != me.ensemble || k.rng_kind != me.rng_kind) {
        fclose(f);
        return fail(JMM_ERR_INVALID, "checkpoint was written by a handle with a different configuration "
                                     "(N, nchains, mode, pot, NBN, ensemble, generator, seed or chain_id0)");
    }
    if (memcmp(&k.cutoff, &me.cutoff, sizeof(double)) != 0 || k.eci != me.eci || k.mdai != me.mdai || k.mvai != me.mvai ||
        k.adapt != me.adapt || k.arith != me.arith || k.relax != me.relax || k.flags != me.flags) {
        fclose(f);
        return fail(JMM_ERR_INVALID, "checkpoint was written by a handle with a different cadence or arithmetic "
                                     "(cutoff, ENGCHECK, DADJ, VADJ, adapt, arith, flags or RELAX): the continuation would not be exact");
    }
    if (k.hist != me.hist || (k.hist && (k.rhonb != me.rhonb || k.gnb != me.gnb || k.gns != me.gns || k.rbw != me.rbw ||
                                         k.gsw != me.gsw || k.gbw != me.gbw))) {
        fclose(f);
        return fail(JMM_ERR_INVALID, "checkpoint and handle disagree about the histograms (call jmm_enable_histograms "
                                     "with the same geometry before jmm_checkpoint_load, or not at all)");
    }
    // read and size-check the whole file before any device state is touched: a truncated file leaves the handle as it was
    size_t need = 0;
    ckpt_arrays(h, [&](void *, size_t bytes) -> jmm_status { need += bytes; return JMM_OK; });
    std::vector<char> buf(need);
    const size_t got = need ? fread(buf.data(), 1, need, f) : 0;
    const bool trailing = fgetc(f) != EOF;
    fclose(f);
    if (got != need || trailing) return fail(JMM_ERR_IO, std::string(got != need ? "truncated checkpoint " : "checkpoint longer than this handle's state: ") + path);
    size_t off = 0;
    jmm_status st = ckpt_arrays(h, [&](void *d, size_t bytes) -> jmm_status {
        // on the handle's own stream (non-blocking: the legacy stream does not order against it), then one synchronise
        CK(cudaMemcpyAsync(d, buf.data() + off, bytes, cudaMemcpyHostToDevice, h->stream));
        off += bytes;
        return JMM_OK;
    });
    if (st != JMM_OK) return st;
    h->sn = k.sn; h->halfsweeps = k.halfsweeps; h->cursor = k.cursor; h->samples = k.samples;
    if (h->d_cursor) CK(cudaMemcpyAsync(h->d_cursor, &h->cursor, sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->keep_cursor = h->cfg.rng_kind == JMM_RNG_RECORDED;
    return JMM_OK;
}

extern "C" jmm_status jmm_start(jmm_handle *h) {
    if (!h) return fail(JMM_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    if (is_cb(h)) {
        // totals of the initial configuration (what the step-0 fad establishes, :907-946), then the
        // first updateThermo (src/Main.cpp:96) as a zero-length "finish"
        jmm_status st = totals_parallel(h, h->cb_r[h->cb_cur], 1, h->S.N, h->cb_tot, 1, 9);
        if (st != JMM_OK) return st;
        k_sweep_finish<<<(unsigned) h->S.nchains, 32, 0, h->stream>>>(nullptr, 0, h->S.N, h->S.l, h->cb_tot, h->cb_acc, 1);
        h->launches++;
        CK(cudaGetLastError());
        h->samples += 1;
        return JMM_OK;
    }
    CK(DISPATCH_POT_TABLE(h, launch_start, h, false));
    h->samples += 1;
    return JMM_OK;
}

extern "C" jmm_status jmm_start_parts(jmm_handle *h, int32_t do_fad, int32_t do_relax, int32_t do_thermo) {
    if (!h || is_cb(h)) return fail(JMM_ERR_INVALID, "jmm_start_parts: many-chain handles only");
    CK(cudaSetDevice(h->cfg.device));
    const int parts = (do_fad ? kStartFad : 0) | (do_relax ? kStartRelax : 0) | (do_thermo ? kStartThermo : 0);
    if (!parts) return JMM_OK;
    CK(DISPATCH_POT_TABLE(h, launch_start, h, false, parts));
    if (do_thermo) h->samples += 1;
    return JMM_OK;
}

extern "C" jmm_status jmm_relax_volume(jmm_handle *h) {
    if (!h || is_cb(h)) return fail(JMM_ERR_INVALID, "jmm_relax_volume: many-chain handles only");
    CK(cudaSetDevice(h->cfg.device));
    CK(DISPATCH_POT_TABLE(h, launch_start, h, true));
    return JMM_OK;
}

// maxDisAdjust :2100-2115 / maxDVAdjust :2120-2139 with the host's libm, like the reference
extern "C" jmm_status jmm_adjust_step_sizes(jmm_handle *h, int32_t do_dis, int32_t do_vol) {
    if (!h || is_cb(h)) return fail(JMM_ERR_INVALID, "jmm_adjust_step_sizes: many-chain handles only");
    if (!do_dis && !do_vol) return JMM_OK;
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains;
    const double N = (double) h->S.N;
    std::vector<uint64_t> cnt(4 * C), vA(C);
    std::vector<double> ms(C), mv(C);
    CK(cudaMemcpyAsync(cnt.data(), h->S.cnt, 4 * C * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(vA.data(), h->S.vAErr, C * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(ms.data(), h->S.maxStep, C * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(mv.data(), h->S.maxdl, C * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    volatile double idealRatio = 0.5;
    for (uint64_t c = 0; c < C; ++c) {
        const uint64_t d0 = cnt[c], d1 = cnt[C + c], v0 = cnt[2 * C + c], v1 = cnt[3 * C + c];
        if (do_dis) {
            const double actualRatio = (double) d0 / (d0 + d1);
            ms[c] = ms[c] * log(0.672924 * idealRatio + 0.0644284) / log(0.672924 * (actualRatio + 0.0644284));
            if (ms[c] < 0.002) ms[c] = 0.002;
            else if (ms[c] > 0.5) ms[c] = 0.5;
        }
        if (do_vol && (v0 + v1 - vA[c]) > 0) {
            vA[c] = v0 + v1;
            const double actualRatio = (double) v0 / (v0 + v1);
            mv[c] = mv[c] * log(0.672924 * idealRatio + 0.0644284) / log(0.672924 * (actualRatio + 0.0644284));
            if (mv[c] < 0.002 * N) mv[c] = 0.002 * N;
            else if (mv[c] > 0.10 * N) mv[c] = 0.50 * N;
        }
    }
    CK(cudaMemcpyAsync(h->S.maxStep, ms.data(), C * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->S.maxdl, mv.data(), C * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->S.vAErr, vA.data(), C * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return JMM_OK;
}

// ------------------------------------------------------------------------------------------------
// energy
// ------------------------------------------------------------------------------------------------
template <int POT>
static cudaError_t launch_totals_partial(jmm_handle *h, const double *r, uint64_t ps, uint64_t cs, int nblocks, int threads) {
    k_totals_partial<POT><<<(unsigned) (h->S.nchains * nblocks), threads, 0, h->stream>>>(
        r, ps, cs, h->S.N, h->S.nbn, h->S.cutoff, h->S.l, nblocks, h->d_partial);
    h->launches++;
    return cudaGetLastError();
}

static jmm_status ensure_partial(jmm_handle *h, size_t bytes) {
    if (h->partial_bytes >= bytes) return JMM_OK;
    if (h->d_partial) CK(cudaFree(h->d_partial));
    h->d_partial = nullptr; h->partial_bytes = 0;
    CK(cudaMalloc((void **) &h->d_partial, bytes));
    h->partial_bytes = bytes;
    return JMM_OK;
}

static jmm_status totals_parallel(jmm_handle *h, const double *r, uint64_t ps, uint64_t cs, double *out, uint64_t ks, uint64_t ocs) {
    const uint64_t C = h->S.nchains, N = h->S.N;
    const int threads = N >= 256 ? 256 : (int) std::max<uint64_t>(32, ((N + 31) / 32) * 32);
    uint64_t want = (N + threads - 1) / threads;                     // blocks per chain
    const uint64_t cap = std::max<uint64_t>(1, (148ull * 8) / C);    // keep the grid near 8 CTAs per SM
    const int nblocks = (int) std::max<uint64_t>(1, std::min(want, cap));
    jmm_status st = ensure_partial(h, C * nblocks * 9 * sizeof(double));
    if (st != JMM_OK) return st;
    cudaError_t e;
    switch (h->cfg.pot) {
        case JMM_POT_LJ: e = launch_totals_partial<kPotLJ>(h, r, ps, cs, nblocks, threads); break;
        case JMM_POT_LJCUT: e = launch_totals_partial<kPotLJcut>(h, r, ps, cs, nblocks, threads); break;
        default: e = launch_totals_partial<kPotHarmonic>(h, r, ps, cs, nblocks, threads); break;
    }
    CK(e);
    k_totals_finish<<<nblk(C * 9, 128), 128, 0, h->stream>>>(h->d_partial, nblocks, C, out, ks, ocs);
    h->launches++;
    CK(cudaGetLastError());
    return JMM_OK;
}

extern "C" jmm_status jmm_energy(jmm_handle *h, double *totals, int32_t exact_order) {
    if (!h || !totals) return fail(JMM_ERR_INVALID, "jmm_energy: null argument");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains;
    jmm_status st = ensure_stage(h, 2 * C * 9 * sizeof(double));
    if (st != JMM_OK) return st;
    double *d_out = h->d_stage + C * 9;          // [9][C] (many-chain) or [C][9] (checkerboard)
    h->timed = false; tick(h);
    if (is_cb(h)) {
        if (exact_order) return fail(JMM_ERR_INVALID, "exact_order is a many-chain (one thread per chain) path");
        st = totals_parallel(h, h->cb_r[h->cb_cur], 1, h->S.N, d_out, 1, 9);
        if (st != JMM_OK) return st;
        tock(h);
        CK(cudaMemcpyAsync(totals, d_out, C * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return JMM_OK;
    }
    if (exact_order) {
        switch (h->cfg.pot) {
            case JMM_POT_LJ: k_chains_totals_exact<kPotLJ><<<nblk(C, 32), 32, 0, h->stream>>>(h->S, d_out); break;
            case JMM_POT_LJCUT: k_chains_totals_exact<kPotLJcut><<<nblk(C, 32), 32, 0, h->stream>>>(h->S, d_out); break;
            default: k_chains_totals_exact<kPotHarmonic><<<nblk(C, 32), 32, 0, h->stream>>>(h->S, d_out); break;
        }
        h->launches++;
        CK(cudaGetLastError());
    } else {
        st = totals_parallel(h, h->S.r, C, 1, d_out, C, 1);
        if (st != JMM_OK) return st;
    }
    tock(h);
    k_transpose_out<<<nblk(C * 9, 256), 256, 0, h->stream>>>(d_out, h->d_stage, C, 9);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(totals, h->d_stage, C * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return JMM_OK;
}

// ------------------------------------------------------------------------------------------------
// step
// ------------------------------------------------------------------------------------------------
extern "C" jmm_status jmm_step(jmm_handle *h, uint64_t nsteps, const uint32_t *rng_stream, uint64_t n_words, uint8_t *accept_log) {
    if (!h || is_cb(h)) return fail(JMM_ERR_INVALID, "jmm_step: many-chain handles only (use jmm_sweep)");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains;
    if (h->cfg.rng_kind == JMM_RNG_RECORDED) {
        if (rng_stream) {                        // (re)load a stream; NULL keeps consuming the loaded one
            if (h->stream_cap < n_words) {
                if (h->d_stream) CK(cudaFree(h->d_stream));
                h->d_stream = nullptr; h->stream_cap = 0;
                CK(cudaMalloc((void **) &h->d_stream, std::max<uint64_t>(n_words, 1) * sizeof(uint32_t)));
                h->stream_cap = n_words;
            }
            CK(cudaMemcpyAsync(h->d_stream, rng_stream, n_words * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
            h->stream_len = n_words;
            if (h->keep_cursor) {
                // first (re)load after jmm_checkpoint_load: the caller passes the SAME recording again and the run
                // continues at the restored cursor
                h->keep_cursor = false;
                if (h->cursor > n_words) return fail(JMM_ERR_STREAM, "the restored cursor lies beyond the recorded stream passed to jmm_step");
            } else {
                CK(cudaMemsetAsync(h->d_cursor, 0, sizeof(uint64_t), h->stream));
                h->cursor = 0;
            }
        } else n_words = h->stream_len;
        if (!h->d_stream) return fail(JMM_ERR_INVALID, "JMM_RNG_RECORDED needs rng_stream");
    }
    if (accept_log && nsteps) {
        const size_t need = (size_t) nsteps * C;
        if (h->log_bytes < need) {
            if (h->d_log) CK(cudaFree(h->d_log));
            h->d_log = nullptr; h->log_bytes = 0;
            CK(cudaMalloc((void **) &h->d_log, need));
            h->log_bytes = need;
        }
    }
    StepArgs a{};
    a.eci = h->cfg.eci; a.mdai = h->cfg.mdai; a.mvai = h->cfg.mvai;
    a.adapt_device = h->cfg.adapt == JMM_ADAPT_DEVICE;
    const bool host_cadence = h->cfg.adapt == JMM_ADAPT_HOST;
    volatile double idealRatio = 0.5;
    a.log_ideal = log(0.672924 * idealRatio + 0.0644284);
    a.stream = h->d_stream; a.n_words = n_words; a.cursor = h->d_cursor; a.err = h->d_err;
    a.pos_in_smem = h->pos_in_smem;
    const bool relax_on = h->S.relax > 0 && h->S.ensemble == kEnsNPT;

    h->timed = false;
    uint64_t remaining = nsteps, done = 0;
    while (remaining) {
        uint64_t n = std::min<uint64_t>(remaining, 1ull << 30);      // kernels count steps in 32 bits
        if (host_cadence) {                      // split at the reference's host-side events, src/Main.cpp:145-176
            if (a.mdai) n = std::min(n, a.mdai - h->sn % a.mdai);
            if (a.mvai) n = std::min(n, a.mvai - h->sn % a.mvai);
            if (relax_on && h->sn < 1000000ull) n = std::min(n, 10000 - h->sn % 10000);
        }
        a.sn0 = h->sn; a.nsteps = n;
        a.accept_log = accept_log ? h->d_log + done * C : nullptr;
        tick(h);
        CK(DISPATCH_POT_TABLE(h, launch_step_table, h, a));
        tock(h);
        h->sn += n; remaining -= n; done += n; h->samples += n;
        if (host_cadence) {
            const bool dis = a.mdai && h->sn % a.mdai == 0, vol = a.mvai && h->sn % a.mvai == 0;
            if (dis || vol) {
                jmm_status st = jmm_adjust_step_sizes(h, dis, vol);
                if (st != JMM_OK) return st;
            }
            if (relax_on && h->sn % 10000 == 0 && h->sn < 1000000ull) {
                jmm_status st = jmm_relax_volume(h);
                if (st != JMM_OK) return st;
            }
        }
    }
    if (accept_log && nsteps) {
        CK(cudaMemcpyAsync(accept_log, h->d_log, (size_t) nsteps * C, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (h->cfg.rng_kind == JMM_RNG_RECORDED) {
        int err = 0;
        CK(cudaMemcpyAsync(&h->cursor, h->d_cursor, sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(&err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (err) return fail(JMM_ERR_STREAM, "recorded random stream exhausted");
    }
    return JMM_OK;
}

// ------------------------------------------------------------------------------------------------
// checkerboard sweeps
// ------------------------------------------------------------------------------------------------

static int env_int(const char *name, int fallback) {
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : fallback;
}

// Tiling of one launch for a given number k of CTAs per SM: k CTAs per SM and ONE wave (152 CTAs on 148 SMs would
// cost a whole second wave): tiles per chain = m * floor(148 k / nchains), the smallest m whose tile plus halos fits
// in `budget` doubles of shared memory; nsub is halved until the redundantly recomputed halo is below ~25 % of the tile.
static void sweep_tiling(uint64_t N, uint64_t C, int nbn, int ncol, uint64_t ctas, int budget, uint64_t want_sub,
                         int *tile_out, int *halo_out, int *nsub_out) {
    const uint64_t base = std::max<uint64_t>(1, ctas / C);
    int nsub = (int) std::min<uint64_t>(want_sub, 64);
    int tile = 0, halo = 0;
    for (;;) {
        halo = nsub * nbn;
        halo += halo & 1;
        uint64_t tiles = base;
        for (;; tiles += base) {
            uint64_t t = (N + tiles - 1) / tiles;
            t += t & 1;
            if ((int64_t) t + 2 * halo <= budget || t <= (uint64_t) 4 * ncol) { tile = (int) t; break; }
        }
        if (tile >= 8 * halo || nsub == 1) break;
        nsub = std::max(1, nsub / 2);
    }
    if (tile < 2) tile = 2;
    *tile_out = tile; *halo_out = halo; *nsub_out = nsub;
}

static SweepShape sweep_shape(const jmm_handle *h, uint64_t want_sub) {
    SweepShape s{};
    const int nbn = h->cfg.nbn, ncol = nbn + 1;
    const uint64_t N = h->S.N, C = h->S.nchains;
    const bool fast = h->cfg.arith == JMM_ARITH_FAST && h->cfg.pot != JMM_POT_HARMONIC;
    s.fast = fast ? 1 : 0;
    if (!fast) {
        s.G = nbn >= 16 ? 32 : 1;
        sweep_tiling(N, C, nbn, ncol, 148, 200 * 1024 / 8 - 1024, want_sub, &s.tile, &s.halo, &s.nsub);
        const int per_sub = (s.tile + 2 * std::max(0, s.halo - nbn) + ncol - 1) / ncol;
        int threads = s.G == 1 ? per_sub : per_sub * 32;
        threads = std::min(512, std::max(64, ((threads + 31) / 32) * 32));
        s.threads = threads;
        s.smem = (size_t) (s.tile + 2 * s.halo) * 8 + (size_t) 2 * (threads / 32) * 9 * 8 + (size_t) s.nsub * 8 + 16;
        return s;
    }
    // ---- k_sweep_fast (<= 64 registers, so up to 32 warps per SM whatever the CTA size)
    // lanes per trial: the largest power of two <= min(NBN, 32) that keeps a half-sweep at 2-3 rounds of the
    // resident threads (148 SMs x JMM_SWEEP_MAXT).  Measured on C5 (NBN 64): G = 8 3.87e9, G = 4 3.54e9, G = 2
    // 2.9e9 trials/s; on C3 (NBN 4): G = 1 5.3e10, G = 2 3.5e10.
    const uint64_t trials = std::max<uint64_t>(1, C * ((N + ncol - 1) / ncol));
    int g = 1;
    while (g * 2 <= std::min(nbn, 32) && trials * (uint64_t) (g * 2) * 10 <= 148ull * JMM_SWEEP_MAXT * 24) g *= 2;
    {
        const int v = env_int("JMM_SWEEP_G", 0);
        if (v >= 1 && v <= 32 && (v & (v - 1)) == 0) g = v;
    }
    s.G = g;
    const int gpw = 32 / g;
    // Candidates: k = 1, 2, 4 CTAs per SM x 2, 3, ... warps; score = modelled trials per unit time x owned
    // (non-halo) share x the share of a half-sweep that is not the neighbour hand-shake.
    const int k_env = env_int("JMM_SWEEP_K", 0), w_env = env_int("JMM_SWEEP_WARPS", 0);
    const uint64_t sub_cap = (uint64_t) std::max(1, std::min(64, env_int("JMM_SWEEP_NSUB", 64)));
    double best = -1.0;
    for (int k = 1; k <= 4; k *= 2) {
        if (k_env && k != k_env) continue;
        int tile, halo, nsub;
        // shared memory: 200 KB / k, less the per-warp sums [nsub][nwarps][2] and the other small arrays (<= 40 KB / k)
        sweep_tiling(N, C, nbn, ncol, 148ull * k, (160 / k) * 1024 / 8, std::min(want_sub, sub_cap), &tile, &halo, &nsub);
        // first half-sweep: halos still tried; + 1: the base of a stretch moves down to an even trial index (sweep.cuh)
        const int per_sub = (tile + 2 * std::max(0, halo - nbn) + ncol - 1) / ncol + 1;
        const double owned = (double) tile / (double) (tile + std::max(0, halo - nbn));   // mean over the half-sweeps
        const double x = (double) (((N + tile - 1) / tile) * C) / (148.0 * k);            // CTAs / resident CTA slots
        const double fill = x <= 1.0 ? x : x / ceil(x);
        for (int nw = 2; nw * k * 32 <= JMM_SWEEP_MAXT; ++nw) {          // registers: 65536 / MAXT per thread
            if (w_env && nw != w_env) continue;
            const int rounds = (per_sub + nw * gpw - 1) / (nw * gpw);
            // trials per SM and half-sweep over the time of its rounds; a round of w warps per sub-partition takes
            // ~ w^0.7 (measured, profiles/r01t_sweep_shapes_*: more warps keep the fp64 pipe busier, so idle lanes in
            // the last round cost less than a warp count that divides the trials evenly but is smaller)
            const double wps = (double) (nw * k) / 4.0;                               // warps per sub-partition
            const double work = rounds * (120.0 + 40.0 * ((2 * nbn + g - 1) / g));       // warp instructions per half-sweep
            // launch ramp, window load and the closing reduction cost about one half-sweep per launch
            const double score = fill * owned * ((double) per_sub * k) / (rounds * pow(wps, 0.7)) * work / (work + 80.0) * nsub / (nsub + 1.0);
            if (score > best) {
                best = score;
                s.tile = tile; s.halo = halo; s.nsub = nsub; s.rounds = rounds; s.threads = nw * 32;
            }
        }
    }
    // how far (in warps) a stretch can collide: 1 whenever a stretch holds >= 6 trials (sweep.cuh: "Interior first");
    // for shorter stretches the conservative bound on the drift of the colour offsets over the whole launch
    s.rad = (s.rounds * gpw >= 6) ? 1 : 1 + (s.nsub * nbn + ncol + 2 * nbn) / (s.rounds * gpw * ncol);
    const size_t nw = (size_t) s.threads / 32;
    s.smem = (size_t) (s.tile + 2 * s.halo) * 8 + ((3 * (size_t) s.nsub + nw + 3) / 4) * 16 + (size_t) s.nsub * nw * 16 + (size_t) s.nsub * 72 + 32;
    return s;
}

extern "C" jmm_status jmm_sweep(jmm_handle *h, uint64_t n_halfsweeps, uint64_t *trials_out) {
    if (!h || !is_cb(h)) return fail(JMM_ERR_INVALID, "jmm_sweep: JMM_MODE_CHECKERBOARD handles only");
    CK(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains, N = h->S.N;
    uint64_t trials = 0;
    h->timed = false;
    uint64_t remaining = n_halfsweeps;
    tick(h);
    while (remaining) {
        const SweepShape s = sweep_shape(h, remaining);
        const int nsub = (int) std::min<uint64_t>(remaining, (uint64_t) s.nsub);
        const unsigned ntiles = (unsigned) ((N + s.tile - 1) / s.tile);
        // k_sweep: per-tile deltas [C][nsub][ntiles][9], then their sums over the tiles [C][nsub][9] (k_sweep_reduce);
        // k_sweep_fast: per-tile (s12, s6) [C][nsub][ntiles][2], reduced inside the kernel
        const size_t n_partial = s.fast ? C * (size_t) nsub * ntiles * 2 : C * (size_t) nsub * ntiles * 9;
        jmm_status st = ensure_partial(h, (n_partial + C * (size_t) nsub * 9) * sizeof(double));
        if (st != JMM_OK) return st;
        double *tsum = h->d_partial + n_partial;
        SweepDev W{};
        W.nchains = C; W.N = N; W.nbn = h->cfg.nbn; W.ncol = h->cfg.nbn + 1; W.cutoff = h->S.cutoff;
        W.r_in = h->cb_r[h->cb_cur]; W.r_out = h->cb_r[h->cb_cur ^ 1];
        W.l = h->S.l; W.T = h->S.T; W.maxStep = h->S.maxStep; W.seed = h->cfg.seed; W.chain_id0 = h->cfg.chain_id0;
        // trials of this launch, counted on the host (the colour of a half-sweep is a pure function of
        // (seed, chain, half-sweep)): no device round trip, launches stay asynchronous
        for (uint64_t c = 0; c < C; ++c)
            for (int t = 0; t < nsub; ++t) {
                const uint64_t step = h->halfsweeps + t;
                const Philox4 b = philox4x32_10((uint32_t) step, (uint32_t)(step >> 32), 0xFFFFFFFFu,
                                                kTagColour | (uint32_t)(h->cfg.chain_id0 + c), (uint32_t) h->cfg.seed,
                                                (uint32_t)(h->cfg.seed >> 32));
                const uint64_t col = ((uint64_t) b.w[0] * (uint64_t) W.ncol) >> 32;
                if (col < N) trials += (N - col + W.ncol - 1) / W.ncol;
            }
        CK(jmm_launch_sweep(h, s, W, h->halfsweeps, nsub, ntiles));
        if (!s.fast) {
            k_sweep_reduce<<<dim3((unsigned) nsub, (unsigned) C), 288, 0, h->stream>>>(h->d_partial, nsub, (int) ntiles, tsum);
            k_sweep_finish<<<(unsigned) C, 32, (size_t) nsub * 9 * sizeof(double), h->stream>>>(tsum, nsub, N, h->S.l, h->cb_tot, h->cb_acc, 0);
            h->launches += 2;
            CK(cudaGetLastError());
        }
        h->cb_cur ^= 1;
        h->halfsweeps += nsub; h->samples += nsub;
        remaining -= nsub;
    }
    tock(h);
    if (trials_out) *trials_out = trials;
    return JMM_OK;
}

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
extern "C" const char *jmm_last_error(void) { return g_err.c_str(); }
extern "C" const char *jmm_version(void) { return "jmmonedmc_b200 0.1 (sm_100a)"; }

__global__ void k_rng_selftest(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                               uint64_t seed, uint32_t *out, uint32_t *taus_out, uint32_t n) {
    const Philox4 b = philox4x32_10(c0, c1, c2, c3, k0, k1);
    for (int i = 0; i < 4; ++i) out[i] = b.w[i];
    uint32_t s1, s2, s3;
    taus2_seed(seed, s1, s2, s3);
    for (uint32_t i = 0; i < n; ++i) taus_out[i] = taus2_next(s1, s2, s3);
}

// 16 independent DFMA chains per thread, 2 x 148 x 4 CTAs of 256 threads: saturates the fp64 pipe
__global__ void k_fp64_peak(double *out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __fma_rn(x[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;      // never true; keeps the chains alive
}

extern "C" double jmm_fp64_peak_tflops(int32_t device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { fail(JMM_ERR_CUDA, "no CUDA device"); return -1.0; }
    cudaSetDevice(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    double *d = nullptr;
    if (cudaMalloc((void **) &d, 8) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_fp64_peak<<<blocks, threads>>>(d, iters, 1.0000001, 0.9999999);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 16.0 * (double) iters * (double) blocks * threads;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return best;
}

// ---- self-test of the banded acceptance rules ---------------------------------------------------------------
__global__ void k_accept_selftest(uint64_t n, uint64_t seed, unsigned long long *counts) {
    const uint64_t i0 = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bad_m = 0, bad_v = 0, ex_m = 0, ex_v = 0;
    for (uint64_t i = i0; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const Philox4 a = philox4x32_10((uint32_t) i, (uint32_t)(i >> 32), 0x5e1f7e57u, 1u, (uint32_t) seed, (uint32_t)(seed >> 32));
        const Philox4 b = philox4x32_10((uint32_t) i, (uint32_t)(i >> 32), 0x5e1f7e57u, 2u, (uint32_t) seed, (uint32_t)(seed >> 32));
        const double T = 0.1 + 1.9 * u01(a.w[0]);
        // energy changes from -5 T to +30 T, log-uniform magnitudes included through the square
        const double u = u01(a.w[1]) * 2 - 0.3;
        const double dE = T * 18.0 * u * fabs(u);
        double ran = u01(a.w[2]);
        if (i & 1) {                                   // adversarial: ran next to the exact probability
            const double p = exp(-dE / T);
            const double eps = ldexp(1.0, -(int) (a.w[3] % 40) - 13) * ((a.w[3] >> 8 & 1) ? 1.0 : -1.0);    // 1e-4 .. 1e-16
            ran = p * (1.0 + eps);
            if (!(ran >= 0.0 && ran < 1.0)) ran = u01(a.w[2]);
        }
        const bool want = dE <= 0 || exp(-dE / T) > ran;                              // :1377
        const double ea = (double) exp_neg_approx(dE * (1.0 / T));
        if (!(dE <= 0) && !(ran > ea + kMetropolisBand) && !(ran < ea - kMetropolisBand)) ++ex_m;
        if (metropolis_accept(dE, T, 1.0 / T, ran) != want) ++bad_m;

        // volume rule: N from 2 to 2000, box ratio within +-20 %, x = dE + P dl of either sign
        const double N = 2.0 + floor(1998.0 * u01(b.w[0]) * u01(b.w[0]));
        const double s = 0.8 + 0.4 * u01(b.w[1]);
        const double v = u01(b.w[2]) * 2 - 1;
        const double x = T * (N * log(s) + 12.0 * v * fabs(v));                        // keeps bf in a useful range
        double ran2 = u01(b.w[3]);
        const double bf = exp(-x / T + N * log(s));
        if (i & 2) {
            const double eps = ldexp(1.0, -(int) (a.w[3] >> 16 & 31) - 13) * ((a.w[3] >> 9 & 1) ? 1.0 : -1.0);
            ran2 = bf * (1.0 + eps);
            if (!(ran2 >= 0.0 && ran2 < 1.0)) ran2 = u01(b.w[3]);
        }
        const bool want_v = bf >= 1.0 || bf > ran2;                                   // :1672, :2255
        if (volume_accept(x, T, 1.0 / T, N, s, ran2) != want_v) ++bad_v;
        {   // the same band arithmetic as volume_accept, to count how often the exact expression is reached
            float lg;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float) s));
            const double A = N * ((double) lg * 0.6931471805599453) - x * (1.0 / T);
            const double bb = (double) exp_neg_approx(-A);
            const double band = 4.0 * (1.8e-7 * N + 1.6e-7 * fabs(A) + 2.4e-7);
            if (!(ran2 < bb * (1.0 - band)) && !(ran2 > bb * (1.0 + band))) ++ex_v;
        }
    }
    atomicAdd(&counts[0], bad_m); atomicAdd(&counts[1], bad_v); atomicAdd(&counts[2], ex_m); atomicAdd(&counts[3], ex_v);
}

extern "C" jmm_status jmm_accept_selftest(uint64_t n, uint64_t seed, uint64_t counts[4], int32_t device) {
    if (!counts) return fail(JMM_ERR_INVALID, "jmm_accept_selftest: null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(JMM_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    CK(cudaSetDevice(device));
    unsigned long long *d = nullptr;
    CK(cudaMalloc((void **) &d, 4 * sizeof(unsigned long long)));
    CK(cudaMemset(d, 0, 4 * sizeof(unsigned long long)));
    k_accept_selftest<<<148 * 4, 256>>>(n, seed, d);
    cudaError_t le = cudaGetLastError();
    if (le == cudaSuccess) le = cudaMemcpy(counts, d, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (le != cudaSuccess) return fail(JMM_ERR_CUDA, cudaGetErrorString(le));
    return JMM_OK;
}

extern "C" jmm_status jmm_rng_selftest(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], uint64_t seed,
                                       uint32_t *taus_out, uint32_t n, int32_t device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(JMM_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    CK(cudaSetDevice(device));
    uint32_t *d = nullptr;
    CK(cudaMalloc((void **) &d, (4 + (size_t) n) * sizeof(uint32_t)));
    k_rng_selftest<<<1, 1>>>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], seed, d, d + 4, n);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) { cudaFree(d); return fail(JMM_ERR_CUDA, cudaGetErrorString(le)); }
    std::vector<uint32_t> hbuf(4 + (size_t) n);
    e = cudaMemcpy(hbuf.data(), d, hbuf.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(JMM_ERR_CUDA, cudaGetErrorString(e));
    memcpy(out, hbuf.data(), 4 * sizeof(uint32_t));
    if (n) memcpy(taus_out, hbuf.data() + 4, (size_t) n * sizeof(uint32_t));
    return JMM_OK;
}
