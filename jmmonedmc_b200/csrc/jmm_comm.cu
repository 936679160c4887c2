// Multi-GPU side of the C ABI: chains and state points are sharded over ranks with NO data-path
// traffic; the one exchange of a job is a single ncclAllGather of fixed-size per-chain summary records at the end
// (what scripts/RunJobs.bash + scripts/Analyze_Mean.py do through 100 directories of thermo files).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded by the host process — e.g. the one
// PyTorch bundles — or the system's), so libjmmgpu.so loads and runs on one GPU without it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "handle.h"

using namespace jmm;

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string error;
};

void nccl_load(NcclApi &api);

NcclApi *nccl_api() {                       // bound once per process, whichever thread asks first
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] { nccl_load(api); });
    return &api;
}

void nccl_load(NcclApi &api) {
    const char *names[] = {getenv("JMM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) { api.error = std::string("cannot load NCCL (libnccl.so.2): ") + dlerror(); return; }
    auto sym = [&](const char *name) { void *p = dlsym(api.lib, name); if (!p) api.error = std::string("NCCL symbol missing: ") + name; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId)) sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank)) sym("ncclCommInitRank");
    api.AllGather = (decltype(api.AllGather)) sym("ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy)) sym("ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString)) sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion)) sym("ncclGetVersion");
    if (!api.error.empty()) api.lib = nullptr;
}

}  // namespace

struct jmm_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    double *d_send = nullptr, *d_recv = nullptr;
    size_t slot = 0;                       // records per rank the buffers are sized for
};

static_assert(JMM_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "jmm_gpu.h mirrors ncclUniqueId");

#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return jmm_fail(JMM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define CKN(call)                                                                                  \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) return jmm_fail(JMM_ERR_NCCL, std::string(#call) + ": " + api->GetErrorString(r_)); \
    } while (0)

extern "C" jmm_status jmm_comm_unique_id(uint8_t id[JMM_COMM_ID_BYTES]) {
    if (!id) return jmm_fail(JMM_ERR_INVALID, "jmm_comm_unique_id: null argument");
    NcclApi *api = nccl_api();
    if (!api->lib) return jmm_fail(JMM_ERR_NCCL, api->error);
    ncclUniqueId u;
    CKN(api->GetUniqueId(&u));
    memcpy(id, u.internal, JMM_COMM_ID_BYTES);
    return JMM_OK;
}

extern "C" jmm_status jmm_comm_create(const uint8_t id[JMM_COMM_ID_BYTES], int32_t rank, int32_t world, int32_t device, jmm_comm **out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return jmm_fail(JMM_ERR_INVALID, "jmm_comm_create: bad argument");
    *out = nullptr;
    NcclApi *api = nccl_api();
    if (!api->lib) return jmm_fail(JMM_ERR_NCCL, api->error);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return jmm_fail(JMM_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return jmm_fail(JMM_ERR_INVALID, "device ordinal out of range");
    CKC(cudaSetDevice(device));
    jmm_comm *c = new jmm_comm;
    c->rank = rank; c->world = world; c->device = device;
    ncclUniqueId u;
    memcpy(u.internal, id, JMM_COMM_ID_BYTES);
    ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) { delete c; return jmm_fail(JMM_ERR_NCCL, std::string("ncclCommInitRank: ") + api->GetErrorString(r)); }
    *out = c;
    return JMM_OK;
}

extern "C" jmm_status jmm_comm_destroy(jmm_comm *c) {
    if (!c) return JMM_OK;
    NcclApi *api = nccl_api();
    cudaSetDevice(c->device);
    if (c->d_send) cudaFree(c->d_send);
    if (c->d_recv) cudaFree(c->d_recv);
    if (c->comm && api->lib) api->CommDestroy(c->comm);
    delete c;
    return JMM_OK;
}

extern "C" int32_t jmm_nccl_version(void) {
    NcclApi *api = nccl_api();
    int v = 0;
    if (!api->lib || !api->GetVersion || api->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

// ---- summary records --------------------------------------------------------------------------------------------
// One record per chain, JMM_SUMMARY_DOUBLES doubles (layout in jmm_gpu.h).  Slots beyond nchains (a ragged last
// rank) carry chain id -1.
static __global__ void k_pack_summaries(ChainsDev S, const double *cb_tot, const double *cb_acc, const unsigned long long *cb_counts,
                                        int checkerboard, double samples, uint64_t slot, double *out) {
    const uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= slot) return;
    double *o = out + c * JMM_SUMMARY_DOUBLES;
    if (c >= S.nchains) {
        o[0] = -1.0;
        for (int k = 1; k < JMM_SUMMARY_DOUBLES; ++k) o[k] = 0.0;
        return;
    }
    const uint64_t C = S.nchains;
    o[0] = (double) (S.chain_id0 + c); o[1] = S.P[c]; o[2] = S.T[c]; o[3] = (double) S.N; o[4] = samples;
    if (checkerboard) {
        for (int k = 0; k < 12; ++k) o[5 + k] = cb_acc[c * 12 + k];
        o[17] = (double) cb_counts[2 * c]; o[18] = (double) (cb_counts[2 * c + 1] - cb_counts[2 * c]); o[19] = 0.0; o[20] = 0.0;
        o[21] = S.l[c]; o[22] = cb_tot[c * 9 + 0]; o[23] = cb_tot[c * 9 + 1];
    } else {
        for (int k = 0; k < 12; ++k) o[5 + k] = S.acc[k * C + c];
        for (int k = 0; k < 4; ++k) o[17 + k] = (double) S.cnt[k * C + c];
        o[21] = S.l[c]; o[22] = S.tot[0 * C + c]; o[23] = S.tot[1 * C + c];
    }
}

static cudaError_t pack(jmm_handle *h, uint64_t slot, double *d_out) {
    const int cb = h->cfg.mode == JMM_MODE_CHECKERBOARD;
    k_pack_summaries<<<nblk(slot, 128), 128, 0, h->stream>>>(h->S, h->cb_tot, h->cb_acc, h->cb_counts, cb, (double) h->samples, slot, d_out);
    h->launches++;
    return cudaGetLastError();
}

extern "C" jmm_status jmm_summaries(jmm_handle *h, double *records) {
    if (!h || !records) return jmm_fail(JMM_ERR_INVALID, "jmm_summaries: null argument");
    CKC(cudaSetDevice(h->cfg.device));
    const uint64_t C = h->S.nchains;
    double *d = nullptr;
    CKC(cudaMalloc((void **) &d, C * JMM_SUMMARY_DOUBLES * sizeof(double)));
    cudaError_t e = pack(h, C, d);
    if (e == cudaSuccess) e = cudaMemcpyAsync(records, d, C * JMM_SUMMARY_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess) return jmm_fail(JMM_ERR_CUDA, cudaGetErrorString(e));
    return JMM_OK;
}

extern "C" jmm_status jmm_allgather_summaries(jmm_handle *h, jmm_comm *c, uint64_t total_chains, double *out) {
    if (!h || !c || !out) return jmm_fail(JMM_ERR_INVALID, "jmm_allgather_summaries: null argument");
    NcclApi *api = nccl_api();
    if (!api->lib) return jmm_fail(JMM_ERR_NCCL, api->error);
    if (h->cfg.device != c->device) return jmm_fail(JMM_ERR_INVALID, "handle and communicator live on different devices");
    CKC(cudaSetDevice(h->cfg.device));
    // every rank sends the same number of records: ceil(total / world) slots, the unused ones marked with id -1
    const uint64_t slot = (total_chains + (uint64_t) c->world - 1) / (uint64_t) c->world;
    if (h->S.nchains > slot) return jmm_fail(JMM_ERR_INVALID, "this rank holds more chains than ceil(total_chains / world)");
    if (c->slot < slot) {
        if (c->d_send) cudaFree(c->d_send);
        if (c->d_recv) cudaFree(c->d_recv);
        c->d_send = c->d_recv = nullptr; c->slot = 0;
        CKC(cudaMalloc((void **) &c->d_send, slot * JMM_SUMMARY_DOUBLES * sizeof(double)));
        CKC(cudaMalloc((void **) &c->d_recv, slot * JMM_SUMMARY_DOUBLES * sizeof(double) * c->world));
        c->slot = slot;
    }
    CKC(pack(h, slot, c->d_send));
    CKN(api->AllGather(c->d_send, c->d_recv, slot * JMM_SUMMARY_DOUBLES, ncclDouble, c->comm, h->stream));   // the job's only collective
    std::vector<double> host(slot * JMM_SUMMARY_DOUBLES * c->world);
    CKC(cudaMemcpyAsync(host.data(), c->d_recv, host.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CKC(cudaStreamSynchronize(h->stream));
    // records in global chain order; every chain must arrive exactly once
    std::vector<uint8_t> seen(total_chains, 0);
    uint64_t n = 0;
    for (size_t s = 0; s < slot * (size_t) c->world; ++s) {
        const double *rec = host.data() + s * JMM_SUMMARY_DOUBLES;
        if (rec[0] < 0) continue;
        const uint64_t id = (uint64_t) rec[0];
        if (id >= total_chains || seen[id]) return jmm_fail(JMM_ERR_INVALID, "allgather: chain ids of the ranks overlap or exceed total_chains");
        seen[id] = 1; ++n;
        memcpy(out + id * JMM_SUMMARY_DOUBLES, rec, JMM_SUMMARY_DOUBLES * sizeof(double));
    }
    if (n != total_chains) return jmm_fail(JMM_ERR_INVALID, "allgather: the ranks together hold fewer chains than total_chains");
    return JMM_OK;
}
