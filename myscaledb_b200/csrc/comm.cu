// comm.cu -- multi-GPU as a product feature: a communicator below the C ABI so that a C++ host (one process per GPU, or
// one ClickHouse worker per device) runs "shard search -> all-gather of the per-shard top-k -> merge" without Python.
//
// Replaces, across GPUs, what MergeTreeBaseSearchManager::getTotalTopSearchResultImpl does across parts (reference:
// src/VectorIndex/Storages/MergeTreeBaseSearchManager.cpp:207-299) and the table-wide statistics sum of
// ReadWithHybridSearch::getStatisticForTextSearch (src/VectorIndex/Processors/ReadWithHybridSearch.cpp:89-209).
// The only data-path collective of the vector side is ONE ncclAllGather of the packed record
// {float dis[nq * k]; int64 id[nq * k]} per rank and batch (122 KB at nq = 1024, k = 10: latency-bound over NVSwitch),
// followed by the merge kernel (b200_topk_merge_device_ex); BM25 adds one ncclAllReduce(sum) of a few uint64 counters.
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already carries, e.g. PyTorch's): the library has no
// link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>

#include <cstring>
#include <string>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace {

typedef struct ncclComm *ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { ncclUint8 = 1, ncclUint64 = 5 };
enum { ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};

NcclApi &nccl_api(const char *path_hint) {
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (api.handle) return api;
    const char *env = getenv("B200_NCCL_LIB");
    for (const char *cand : {path_hint, env, "libnccl.so.2", "libnccl.so"}) {
        if (!cand || !*cand) continue;
        api.handle = dlopen(cand, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.error = std::string("libnccl.so.2 not found (pass its path or set B200_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
        return api;
    }
    auto sym = [&](const char *n) { return dlsym(api.handle, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.AllReduce) {
        api.error = "libnccl lacks an expected symbol";
        dlclose(api.handle);
        api.handle = nullptr;
    }
    return api;
}

}  // namespace

using namespace b200;

struct b200_comm {
    ncclComm_t comm = nullptr;
    NcclApi *api = nullptr;
    int rank = 0, world = 1, device = 0;
    void *send = nullptr, *recv = nullptr;   // packed records: one / world of them
    size_t send_cap = 0, recv_cap = 0;
    uint64_t *d_counters = nullptr;          // all-reduce scratch
    size_t counters_cap = 0;
    void *d_host_out = nullptr;              // merged result of the host-buffer gather
    size_t host_out_cap = 0;
    std::mutex mu;
    // CUDA graphs of whole sharded search steps, keyed by everything that is baked into the nodes
    std::map<std::tuple<const void *, const void *, int64_t, int, const void *, int64_t, void *, void *, void *>, cudaGraphExec_t> graphs;
    std::map<cudaGraphExec_t, int64_t> graph_launches;   // kernels / collectives one replay stands for (launch accounting)
    void *host_stage = nullptr;                          // device staging of the host-buffer entry point
    size_t host_stage_cap = 0;
};

#define B200_NCCL_OK(c, expr)                                                                                              \
    do {                                                                                                                   \
        int _r = (expr);                                                                                                   \
        if (_r != 0)                                                                                                       \
            return fail(B200_ERR_CUDA, std::string(#expr) + ": " + ((c)->api->GetErrorString ? (c)->api->GetErrorString(_r) : "NCCL error")); \
    } while (0)

extern "C" int b200_comm_unique_id(const char *nccl_lib_path, void *out_id_128_bytes) {
    if (!out_id_128_bytes) return fail(B200_ERR_INVALID, "null output");
    NcclApi &api = nccl_api(nccl_lib_path);
    if (!api.handle) return fail(B200_ERR_UNSUPPORTED, api.error);
    ncclUniqueId id;
    const int r = api.GetUniqueId(&id);
    if (r != 0) return fail(B200_ERR_CUDA, std::string("ncclGetUniqueId: ") + (api.GetErrorString ? api.GetErrorString(r) : "error"));
    memcpy(out_id_128_bytes, id.internal, 128);
    return B200_OK;
}

extern "C" int b200_comm_create(const char *nccl_lib_path, const void *unique_id_128_bytes, int rank, int world, b200_comm **out) {
    if (!out || !unique_id_128_bytes || world < 1 || rank < 0 || rank >= world) return fail(B200_ERR_INVALID, "bad arguments");
    *out = nullptr;
    NcclApi &api = nccl_api(nccl_lib_path);
    if (!api.handle) return fail(B200_ERR_UNSUPPORTED, api.error);
    b200_comm *c = new b200_comm();
    c->api = &api;
    c->rank = rank;
    c->world = world;
    B200_CUDA_OK(cudaGetDevice(&c->device));
    ncclUniqueId id;
    memcpy(id.internal, unique_id_128_bytes, 128);
    const int r = api.CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        delete c;
        return fail(B200_ERR_CUDA, std::string("ncclCommInitRank: ") + (api.GetErrorString ? api.GetErrorString(r) : "error"));
    }
    *out = c;
    return B200_OK;
}

extern "C" int b200_comm_free(b200_comm *c) {
    if (!c) return B200_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (auto &kv : c->graphs) cudaGraphExecDestroy(kv.second);
    if (c->comm) c->api->CommDestroy(c->comm);
    for (void *p : {c->send, c->recv, (void *)c->d_counters, c->host_stage, c->d_host_out})
        if (p) cudaFree(p);
    delete c;
    return B200_OK;
}

extern "C" int b200_comm_info(const b200_comm *c, int *rank, int *world) {
    if (!c) return fail(B200_ERR_INVALID, "null communicator");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return B200_OK;
}

static int comm_reserve(b200_comm *c, int64_t nq, int k) {
    const size_t rec = (size_t)nq * k * 12;
    if (rec > c->send_cap) {
        if (c->send) cudaFree(c->send);
        if (c->recv) cudaFree(c->recv);
        c->send = c->recv = nullptr;
        c->send_cap = c->recv_cap = 0;
        for (auto &kv : c->graphs) cudaGraphExecDestroy(kv.second);   // captured pointers are gone
        c->graphs.clear();
        const size_t want = rec + rec / 4 + 256;
        if (cudaMalloc(&c->send, want) != cudaSuccess || cudaMalloc(&c->recv, want * c->world) != cudaSuccess) {
            cudaGetLastError();
            return fail(B200_ERR_NOMEM, "cudaMalloc of the all-gather buffers failed");
        }
        c->send_cap = want;
        c->recv_cap = want * c->world;
    }
    return B200_OK;
}

// where a shard search should write its [nq][k] result so that no pack step is needed before the all-gather
extern "C" int b200_comm_local_buffers(b200_comm *c, int64_t nq, int k, float **d_dis, int64_t **d_ids) {
    if (!c || !d_dis || !d_ids || nq < 0 || k <= 0) return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA_OK(cudaSetDevice(c->device));
    B200_TRY(comm_reserve(c, std::max<int64_t>(nq, 1), k));
    *d_dis = reinterpret_cast<float *>(c->send);
    *d_ids = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(c->send) + (size_t)nq * k * 4);
    return B200_OK;
}

// all-gather of the records written at b200_comm_local_buffers(nq, k) + merge -> the global top-k on every rank
extern "C" int b200_comm_gather_merge(b200_comm *c, int64_t nq, int k, int descending, float *d_out_dis, int64_t *d_out_ids, void *stream) {
    if (!c || !d_out_dis || !d_out_ids || nq < 0 || k <= 0) return fail(B200_ERR_INVALID, "bad arguments");
    if (nq == 0) return B200_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t rec = (size_t)nq * k * 12;
    if (rec > c->send_cap) return fail(B200_ERR_INVALID, "call b200_comm_local_buffers(nq, k) first");
    if (c->world == 1) {
        B200_CUDA_OK(cudaMemcpyAsync(d_out_dis, c->send, (size_t)nq * k * 4, cudaMemcpyDeviceToDevice, s));
        B200_CUDA_OK(cudaMemcpyAsync(d_out_ids, reinterpret_cast<char *>(c->send) + (size_t)nq * k * 4, (size_t)nq * k * 8, cudaMemcpyDeviceToDevice, s));
        return B200_OK;
    }
    B200_NCCL_OK(c, c->api->AllGather(c->send, c->recv, rec, ncclUint8, c->comm, s));
    g_launches++;
    // list l of the gathered buffer: dis at recv + l * rec, ids 4 nq k bytes further (rec is a multiple of 8 when nq * k is even;
    // the strides are passed in elements of the respective type, so rec must be divisible by 8)
    if (rec % 8) return fail(B200_ERR_UNSUPPORTED, "nq * k must be even for the packed all-gather record");
    return b200_topk_merge_device_ex(reinterpret_cast<const float *>(c->recv),
                                     reinterpret_cast<const int64_t *>(reinterpret_cast<const char *>(c->recv) + (size_t)nq * k * 4), c->world,
                                     (int64_t)(rec / 4), (int64_t)(rec / 8), nq, k, k, descending, 0, d_out_dis, d_out_ids, nullptr,
                                     stream ? stream : nullptr);
}

// Host-buffer form for lists that are produced on the host (the per-shard BM25 top-k of b200_bm25_search_batch): uploads
// this rank's [nq][k] scores / ids (unused slots: score -inf (descending) or +inf, id -1), all-gathers + merges, and
// returns the table-wide top-k to the host.  Synchronous.
extern "C" int b200_comm_gather_merge_host(b200_comm *c, const float *h_dis, const int64_t *h_ids, int64_t nq, int k, int descending,
                                           float *h_out_dis, int64_t *h_out_ids) {
    if (!c || !h_dis || !h_ids || !h_out_dis || !h_out_ids || nq < 0 || k <= 0) return fail(B200_ERR_INVALID, "bad arguments");
    if (nq == 0) return B200_OK;
    if (c->world == 1) {
        if (h_out_dis != h_dis) memcpy(h_out_dis, h_dis, (size_t)nq * k * 4);
        if (h_out_ids != h_ids) memcpy(h_out_ids, h_ids, (size_t)nq * k * 8);
        return B200_OK;
    }
    float *d_dis = nullptr;
    int64_t *d_ids = nullptr;
    B200_TRY(b200_comm_local_buffers(c, nq, k, &d_dis, &d_ids));
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t nd = (size_t)nq * k * 4, ni = (size_t)nq * k * 8;
    if (nd + ni > c->host_out_cap) {
        if (c->d_host_out) cudaFree(c->d_host_out);
        c->d_host_out = nullptr;
        c->host_out_cap = 0;
        B200_CUDA_OK(cudaMalloc(&c->d_host_out, nd + ni + 256));
        c->host_out_cap = nd + ni + 256;
    }
    float *o_dis = reinterpret_cast<float *>(c->d_host_out);
    int64_t *o_ids = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(c->d_host_out) + ((nd + 7) & ~(size_t)7));
    B200_CUDA_OK(cudaMemcpyAsync(d_dis, h_dis, nd, cudaMemcpyHostToDevice, nullptr));
    B200_CUDA_OK(cudaMemcpyAsync(d_ids, h_ids, ni, cudaMemcpyHostToDevice, nullptr));
    B200_TRY(b200_comm_gather_merge(c, nq, k, descending, o_dis, o_ids, nullptr));
    B200_CUDA_OK(cudaMemcpyAsync(h_out_dis, o_dis, nd, cudaMemcpyDeviceToHost, nullptr));
    B200_CUDA_OK(cudaMemcpyAsync(h_out_ids, o_ids, ni, cudaMemcpyDeviceToHost, nullptr));
    B200_CUDA_OK(cudaStreamSynchronize(nullptr));
    return B200_OK;
}

// table-wide statistics: in-place sum over the ranks of n uint64 counters held on the host
// (total_docs, total_tokens[field], doc_freq[(field, term)]; getStatisticForTextSearch)
extern "C" int b200_comm_allreduce_sum_u64(b200_comm *c, uint64_t *host_counters, int64_t n) {
    if (!c || (!host_counters && n > 0) || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    if (n == 0 || c->world == 1) return B200_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA_OK(cudaSetDevice(c->device));
    if ((size_t)n * 8 > c->counters_cap) {
        if (c->d_counters) cudaFree(c->d_counters);
        c->d_counters = nullptr;
        B200_CUDA_OK(cudaMalloc(&c->d_counters, (size_t)n * 8 + 256));
        c->counters_cap = (size_t)n * 8 + 256;
    }
    B200_CUDA_OK(cudaMemcpy(c->d_counters, host_counters, (size_t)n * 8, cudaMemcpyHostToDevice));
    B200_NCCL_OK(c, c->api->AllReduce(c->d_counters, c->d_counters, (size_t)n, ncclUint64, ncclSum, c->comm, nullptr));
    g_launches++;
    B200_CUDA_OK(cudaStreamSynchronize(nullptr));
    B200_CUDA_OK(cudaMemcpy(host_counters, c->d_counters, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return B200_OK;
}

// ------------------------------------------------------------------------------------
// whole sharded steps
// ------------------------------------------------------------------------------------
extern "C" int b200_corpus_search_device(b200_corpus *c, const float *d_queries, int64_t nq, int k, const uint8_t *d_alive_bits,
                                         int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, void *stream);
extern "C" int b200_index_search_device(b200_index *ix, const float *d_queries, int64_t nq, int k, const char *params, int first_stage_only,
                                        const uint8_t *d_alive_bits, int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, void *stream);
namespace b200 {
int corpus_metric(const b200_corpus *c);
bool corpus_timing_enabled(const b200_corpus *c);
}

static int sharded_corpus_step(b200_comm *cm, b200_corpus *corpus, const float *d_queries, int64_t nq, int k, const uint8_t *d_alive,
                               int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, cudaStream_t s) {
    float *l_dis = reinterpret_cast<float *>(cm->send);
    int64_t *l_ids = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(cm->send) + (size_t)nq * k * 4);
    B200_TRY(b200_corpus_search_device(corpus, d_queries, nq, k, d_alive, id_offset, l_dis, l_ids, s));
    return b200_comm_gather_merge(cm, nq, k, corpus_metric(corpus) == B200_METRIC_IP ? 1 : 0, d_out_dis, d_out_ids, s);
}

// FLAT corpus sharded by rows over the communicator's GPUs: this rank scans its shard (ids + id_offset), the per-shard
// top-k lists are all-gathered and merged; every rank ends with the global answer in d_out_*.  Asynchronous on `stream`
// (must be a real stream, not NULL).  use_graph: replay the whole step (query conversion, scan, all-gather, merge) as
// ONE CUDA graph after the first call with the same arguments -- the step is launch-latency-bound at 8 GPUs.
extern "C" int b200_sharded_corpus_search(b200_comm *cm, b200_corpus *corpus, const float *d_queries, int64_t nq, int k,
                                          const uint8_t *d_alive_bits, int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, void *stream,
                                          int use_graph) {
    if (!cm || !corpus || (!d_queries && nq > 0) || !d_out_dis || !d_out_ids || nq < 0 || k <= 0 || !stream)
        return fail(B200_ERR_INVALID, "bad arguments (a non-NULL stream is required)");
    if (nq == 0) return B200_OK;
    std::lock_guard<std::mutex> lk(cm->mu);
    B200_CUDA_OK(cudaSetDevice(cm->device));
    B200_TRY(comm_reserve(cm, nq, k));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (!use_graph || corpus_timing_enabled(corpus))
        return sharded_corpus_step(cm, corpus, d_queries, nq, k, d_alive_bits, id_offset, d_out_dis, d_out_ids, s);
    const auto key = std::make_tuple((const void *)corpus, (const void *)d_queries, nq, k, (const void *)d_alive_bits, id_offset, (void *)d_out_dis,
                                     (void *)d_out_ids, (void *)s);
    auto it = cm->graphs.find(key);
    if (it == cm->graphs.end()) {
        // first call: run eagerly once (sizes every workspace), then capture the identical sequence
        B200_TRY(sharded_corpus_step(cm, corpus, d_queries, nq, k, d_alive_bits, id_offset, d_out_dis, d_out_ids, s));
        B200_CUDA_OK(cudaStreamSynchronize(s));
        cudaGraph_t g = nullptr;
        B200_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const int64_t launches_before = g_launches;
        int rc = sharded_corpus_step(cm, corpus, d_queries, nq, k, d_alive_bits, id_offset, d_out_dis, d_out_ids, s);
        const int64_t per_step = g_launches - launches_before;
        cudaError_t e = cudaStreamEndCapture(s, &g);
        if (rc != B200_OK || e != cudaSuccess || !g) {
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            if (rc != B200_OK) return rc;
            return fail(B200_ERR_CUDA, std::string("stream capture of the sharded step failed: ") + cudaGetErrorString(e));
        }
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(B200_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        it = cm->graphs.emplace(key, exec).first;
        cm->graph_launches[exec] = per_step;
        g_launches = launches_before;   // the captured pass launched nothing
    }
    B200_CUDA_OK(cudaGraphLaunch(it->second, s));
    g_launches += cm->graph_launches[it->second];
    return B200_OK;
}

// the same through host buffers (what a ClickHouse worker holds): pinned or pageable queries in, the global top-k out on
// every rank; H2D, scan, all-gather, merge, D2H and the synchronise are all inside
extern "C" int b200_sharded_corpus_search_host(b200_comm *cm, b200_corpus *corpus, const float *queries, int64_t nq, int d, int k,
                                               int64_t id_offset, float *out_dis, int64_t *out_ids, void *stream, int use_graph) {
    if (!cm || !corpus || (!queries && nq > 0) || !out_dis || !out_ids || nq < 0 || k <= 0 || d <= 0 || !stream)
        return fail(B200_ERR_INVALID, "bad arguments (a non-NULL stream is required)");
    if (nq == 0) return B200_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    {
        std::lock_guard<std::mutex> lk(cm->mu);
        B200_CUDA_OK(cudaSetDevice(cm->device));
        const size_t need = (size_t)nq * d * 4 + (size_t)nq * k * 12 + 64;
        if (need > cm->host_stage_cap) {
            if (cm->host_stage) cudaFree(cm->host_stage);
            cm->host_stage = nullptr;
            for (auto &kv : cm->graphs) cudaGraphExecDestroy(kv.second);
            cm->graphs.clear();
            B200_CUDA_OK(cudaMalloc(&cm->host_stage, need + need / 4));
            cm->host_stage_cap = need + need / 4;
        }
    }
    float *d_q = reinterpret_cast<float *>(cm->host_stage);
    float *d_od = reinterpret_cast<float *>(reinterpret_cast<char *>(cm->host_stage) + round_up((size_t)nq * d * 4, 16));
    int64_t *d_oi = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(d_od) + round_up((size_t)nq * k * 4, 16));
    B200_CUDA_OK(cudaMemcpyAsync(d_q, queries, (size_t)nq * d * 4, cudaMemcpyHostToDevice, s));
    B200_TRY(b200_sharded_corpus_search(cm, corpus, d_q, nq, k, nullptr, id_offset, d_od, d_oi, stream, use_graph));
    B200_CUDA_OK(cudaMemcpyAsync(out_dis, d_od, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaMemcpyAsync(out_ids, d_oi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    return B200_OK;
}

// a vector index sharded by rows (every rank built its own index over its rows): search + all-gather + merge
extern "C" int b200_sharded_index_search(b200_comm *cm, b200_index *ix, int metric, const float *d_queries, int64_t nq, int k, const char *params,
                                         const uint8_t *d_alive_bits, int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, void *stream) {
    if (!cm || !ix || (!d_queries && nq > 0) || !d_out_dis || !d_out_ids || nq < 0 || k <= 0 || !stream)
        return fail(B200_ERR_INVALID, "bad arguments (a non-NULL stream is required)");
    if (nq == 0) return B200_OK;
    std::lock_guard<std::mutex> lk(cm->mu);
    B200_CUDA_OK(cudaSetDevice(cm->device));
    B200_TRY(comm_reserve(cm, nq, k));
    float *l_dis = reinterpret_cast<float *>(cm->send);
    int64_t *l_ids = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(cm->send) + (size_t)nq * k * 4);
    B200_TRY(b200_index_search_device(ix, d_queries, nq, k, params, 0, d_alive_bits, id_offset, l_dis, l_ids, stream));
    return b200_comm_gather_merge(cm, nq, k, metric == B200_METRIC_IP ? 1 : 0, d_out_dis, d_out_ids, stream);
}
