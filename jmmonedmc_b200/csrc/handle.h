// The state behind the opaque jmm_handle of include/jmm_gpu.h, and the launch entry points of the kernel
// families.  libjmmgpu.so is built from four translation units so that the template instantiations compile in
// parallel: jmm_gpu.cu (ABI, chains.cuh), launch_prod.cu (prod.cuh), launch_coop.cu (coop.cuh, bond.cuh),
// launch_sweep.cu (sweep.cuh: k_sweep).  Nothing declared here is exported from the library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../../include/jmm_gpu.h"
#include "chains.cuh"
#include "sweep.cuh"

struct jmm_handle {
    jmm_config cfg{};
    jmm::ChainsDev S{};
    jmm::HistDev H{};                    // null pointers = histograms off
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    uint64_t sn = 0, launches = 0;
    uint64_t samples = 0;                // updateThermo calls accumulated in the running sums since jmm_zero_accum
    // many-chain launch shape
    int block = 32, pos_in_smem = 1;
    size_t smem = 0;
    int bond = 0;                   // bond.cuh serves this handle (HARMONIC, NBN 1, N <= 17, no RELAX): 1 = k_chains_step_bond, 2 = ..._bond2
    int coop_g = 0;                 // lanes per chain of the cooperative kernel (0 = one chain per thread)
    int coop_npad = 0;
    size_t coop_smem = 0;
    // lanes.cuh (JMM_ARITH_FAST, G lanes per chain): lanes per chain (0 = not this kernel), unrolled slots per lane
    // (0 = run-time loop), row length incl. pads, doubles per group in shared memory
    int lanes_g = 0, lanes_npl = 0, lanes_npad = 0, lanes_stride = 0;
    // recorded stream
    uint32_t *d_stream = nullptr;
    uint64_t stream_cap = 0;               // allocated words
    uint64_t stream_len = 0;               // words of the stream loaded last (<= stream_cap)
    bool keep_cursor = false;              // set by jmm_checkpoint_load: the next stream load keeps the restored cursor
    uint64_t *d_cursor = nullptr;
    int *d_err = nullptr;
    uint64_t cursor = 0;
    // time-sliced production launches: work counter + per-tile progress words
    unsigned int *d_work = nullptr;
    size_t work_words = 0;
    // scratch
    double *d_stage = nullptr;
    size_t stage_bytes = 0;
    uint8_t *d_log = nullptr;
    size_t log_bytes = 0;
    double *d_partial = nullptr;
    size_t partial_bytes = 0;
    // checkerboard mode: chain-major positions, double-buffered
    double *cb_r[2] = {nullptr, nullptr};
    int cb_cur = 0;
    double *cb_tot = nullptr, *cb_acc = nullptr;
    unsigned long long *cb_counts = nullptr;
    uint64_t halfsweeps = 0;
    unsigned int *cb_tile_done = nullptr;   // [nchains] tiles of the running k_sweep_fast launch that have finished
    std::vector<void *> allocs;
};


static inline unsigned nblk(uint64_t n, unsigned b) { return (unsigned) ((n + b - 1) / b); }

constexpr int kCoopScratchRows = 32;        // = kCoopChunk of coop.cuh (checked there)

struct SweepShape {
    int tile, halo, nsub, threads, G;
    size_t smem;
    int fast;          // k_sweep_fast serves this launch (JMM_ARITH_FAST, LJ family)
    int rounds, rad;   // k_sweep_fast: trials per group per half-sweep; how far (in warps) a warp's stretch can collide
};

#define JMM_INTERNAL __attribute__((visibility("hidden")))
#include <string>
// records the message for jmm_last_error() and returns `code` (jmm_gpu.cu)
JMM_INTERNAL jmm_status jmm_fail(jmm_status code, const std::string &msg);
// prod.cuh: many chains, one chain per thread or per G lanes (POT / arithmetic / G dispatch inside)
JMM_INTERNAL cudaError_t jmm_launch_prod(jmm_handle *h, const jmm::StepArgs &a);
// coop.cuh / bond.cuh: few chains, G lanes per chain
JMM_INTERNAL cudaError_t jmm_launch_coop(jmm_handle *h, const jmm::StepArgs &a);
// lanes.cuh: a moderate number of chains, G lanes per chain, fast arithmetic
JMM_INTERNAL cudaError_t jmm_launch_lanes(jmm_handle *h, const jmm::StepArgs &a);
JMM_INTERNAL void jmm_lanes_shape(jmm_handle *h, int g);
// sweep.cuh: nsub colour half-sweeps of every chain of a checkerboard handle
JMM_INTERNAL cudaError_t jmm_launch_sweep(jmm_handle *h, const SweepShape &s, const jmm::SweepDev &W, uint64_t step0, int nsub,
                                          unsigned ntiles);
