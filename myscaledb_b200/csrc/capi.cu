// capi.cu -- the C ABI of libb200search.so (declared in include/b200_search.h) and the host
// orchestration under it: device-resident corpora, path selection (memory-bound scan vs
// tcgen05 GEMM), query staging, partial-list merge.
//
// There is deliberately no CPU compute path in this file: every entry point either runs
// CUDA kernels on an sm_100 device or fails with B200_ERR_NO_DEVICE / B200_ERR_CUDA.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace b200 {

thread_local std::string g_error;
thread_local int64_t g_launches = 0;

void set_error(const std::string &msg) { g_error = msg; }
int fail(int code, const std::string &msg) {
    g_error = msg;
    return code;
}

static int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(B200_ERR_NO_DEVICE, std::string("no CUDA device visible (") +
                                            (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                                            "); libb200search has no CPU fallback");
    }
    int dev = 0;
    B200_CUDA_OK(cudaGetDevice(&dev));
    int major = 0;
    B200_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10)
        return fail(B200_ERR_NO_DEVICE, "device compute capability major is " + std::to_string(major) +
                                            "; this library carries sm_100a code only");
    return B200_OK;
}

static int num_sms() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B200_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B200_ERR_NOMEM, "cudaMalloc(" + std::to_string(want) + ") failed: " + cudaGetErrorString(e));
        }
        cap = want;
        return B200_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() {
        return reinterpret_cast<T *>(p);
    }
};

}  // namespace b200

using namespace b200;

struct b200_corpus {
    int metric = 0, dtype = 0, d = 0, d_pad = 0;
    int64_t cap = 0, n = 0;
    int64_t row_bytes = 0;
    void *data = nullptr;
    size_t data_cap_bytes = 0;   // allocation size of `data` (owned corpora)
    int64_t side_cap_rows = 0;   // rows the row_scale / row_bias arrays can hold
    bool owns = true;
    float *row_scale = nullptr;  // cosine: -1/||y||
    float *row_bias = nullptr;   // L2: ||y||^2 (GEMM path)
    int device = 0;
    int path = 0;
    int gemm_cta_group = 0;  // 0 auto, 1 force single-CTA MMA (A/B experiments)
    int sms = 148;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    // workspaces
    DevBuf w_raw, w_q32, w_qbf, w_qlo, w_qnorm, w_pk, w_pi, w_lk, w_li, w_alive, w_odis, w_oids, w_stage, w_prog;
    int sync_slack = 2;
    // fused single-launch path of the host entry point (small batches, scan kernel): mapped pinned staging + counters
    void *h_pin = nullptr;         // [queries 8 * d fp32 | dis 8 * k | ids 8 * k | flag]
    size_t h_pin_bytes = 0;
    unsigned int *d_tickets = nullptr;   // [8] + tiles_done, then (at +64 bytes) the device copy of the staged queries
    unsigned int fused_seq = 0;
    int d_tickets_d = 0;           // row length the query staging behind d_tickets was sized for
    int fused_enabled = 1;         // B200_FUSED_SCAN=0 disables (A/B)
    int rescore_l2 = 1;      // tensor-core L2: re-score the k winners exactly (B200_GEMM_RESCORE_L2=0 disables, A/B only)
    int gemm_multicast = 1;  // CTA pairs per cluster sharing each corpus tile: 1 auto (4, else 2), 2, 4; B200_GEMM_MULTICAST=0 disables
    int gemm_ts = 0;  // 0 streaming (default: faster at every measured d), 1 TS when d_pad <= 512, 2 TS whenever it fits
    // what the last tensor-core launch really was (tests assert on it): cta_group, pairs per cluster, TS form, grid, kernel id
    int last_cg = 0, last_mc = 0, last_ts = 0, last_grid = 0, last_kernel = 0;
    // optional CUDA-event timing of the dominant kernel (scan or GEMM) for the roofline report
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_used, ev_free;
    double timed_ms = 0;
    int64_t timed_launches = 0;
};

static void timing_begin(b200_corpus *c, cudaStream_t s, std::pair<cudaEvent_t, cudaEvent_t> &ev) {
    if (!c->timing) return;
    if (c->ev_free.empty()) {
        cudaEventCreate(&ev.first);
        cudaEventCreate(&ev.second);
    } else {
        ev = c->ev_free.back();
        c->ev_free.pop_back();
    }
    cudaEventRecord(ev.first, s);
}
// fold the recorded event pairs into the running totals (waits for them) and recycle the events
static void timing_drain(b200_corpus *c) {
    for (auto &ev : c->ev_used) {
        float ms = 0;
        if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
            c->timed_ms += ms;
            c->timed_launches++;
        }
        c->ev_free.push_back(ev);
    }
    c->ev_used.clear();
}
static void timing_end(b200_corpus *c, cudaStream_t s, std::pair<cudaEvent_t, cudaEvent_t> &ev) {
    if (!c->timing) return;
    cudaEventRecord(ev.second, s);
    c->ev_used.push_back(ev);
    if (c->ev_used.size() >= 1024) timing_drain(c);  // nobody asked for the totals for a while: keep the list bounded
}

static int corpus_alloc(b200_corpus *c, int64_t rows);
static int corpus_norms(b200_corpus *c, int64_t first, int64_t n);

namespace b200 {
// hooks for comm.cu
int corpus_metric(const b200_corpus *c) { return c->metric; }
bool corpus_timing_enabled(const b200_corpus *c) { return c->timing; }
// hooks for the index layer (ivf.cu): device view of the rows / in-place row normalisation
const void *corpus_device_rows(const b200_corpus *c) { return c->data; }
// append fp32 rows [n][d] that already live on the device (index build, centroid tables); asynchronous on s except for
// a reallocation, synchronised before returning so that the caller may reuse d_rows
int corpus_append_device(b200_corpus *c, const float *d_rows, int64_t n, cudaStream_t s) {
    if (!c || (!d_rows && n > 0) || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    if (c->dtype == B200_DTYPE_BIN) return fail(B200_ERR_UNSUPPORTED, "device append: float corpora only");
    if (n == 0) return B200_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA_OK(cudaSetDevice(c->device));
    B200_TRY(corpus_alloc(c, std::max(c->n + n, c->cap)));
    char *dst = reinterpret_cast<char *>(c->data) + c->n * c->row_bytes;
    if (c->dtype == B200_DTYPE_BF16) B200_CUDA_OK(launch_f32_to_bf16_rows(d_rows, c->d, dst, c->d_pad, n, s));
    else B200_CUDA_OK(launch_pad_rows_f32(d_rows, c->d, reinterpret_cast<float *>(dst), c->d_pad, n, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    B200_TRY(corpus_norms(c, c->n, n));
    B200_CUDA_OK(cudaStreamSynchronize(c->stream));
    c->n += n;
    return B200_OK;
}
int corpus_normalize_rows(b200_corpus *c) {
    if (c->dtype != B200_DTYPE_F32) return fail(B200_ERR_UNSUPPORTED, "normalise: fp32 corpora only");
    B200_CUDA_OK(cudaSetDevice(c->device));
    B200_CUDA_OK(launch_normalize_rows_f32(reinterpret_cast<float *>(c->data), c->d_pad, c->n, c->stream));
    B200_CUDA_OK(cudaStreamSynchronize(c->stream));
    return B200_OK;
}
}  // namespace b200

static bool is_float_metric(int m) { return m == B200_METRIC_L2 || m == B200_METRIC_IP || m == B200_METRIC_COSINE; }
static bool is_bin_metric(int m) { return m == B200_METRIC_HAMMING || m == B200_METRIC_JACCARD; }

static int pad_for(int dtype, int d) {
    if (dtype == B200_DTYPE_BF16) return (int)round_up(d, 64);  // 128-byte TMA/UMMA swizzle rows
    if (dtype == B200_DTYPE_F32) return (int)round_up(d, 4);    // 16-byte vector loads
    return d / 8;                                                // binary: bytes
}

extern "C" const char *b200_last_error(void) { return g_error.c_str(); }
extern "C" const char *b200_version(void) { return "b200search 0.1 (sm_100a)"; }

extern "C" int b200_device_count(int *out_n) {
    if (!out_n) return fail(B200_ERR_INVALID, "out_n is null");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *out_n = n;
    return B200_OK;
}

extern "C" int b200_set_device(int device) {
    B200_TRY(ensure_device());
    B200_CUDA_OK(cudaSetDevice(device));
    return B200_OK;
}

extern "C" int64_t b200_launch_count(int reset) {
    int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

extern "C" int b200_corpus_create(int metric, int dtype, int d, int64_t capacity_rows, b200_corpus **out) {
    if (!out) return fail(B200_ERR_INVALID, "out is null");
    *out = nullptr;
    if (d <= 0 || capacity_rows < 0) return fail(B200_ERR_INVALID, "bad d / capacity");
    if (dtype == B200_DTYPE_BIN) {
        if (!is_bin_metric(metric)) return fail(B200_ERR_INVALID, "binary corpus needs HAMMING or JACCARD");
        if (d % 8) return fail(B200_ERR_INVALID, "binary dimension must be a multiple of 8 bits");
    } else if (dtype == B200_DTYPE_F32 || dtype == B200_DTYPE_BF16) {
        if (!is_float_metric(metric)) return fail(B200_ERR_INVALID, "float corpus needs L2, IP or COSINE");
    } else {
        return fail(B200_ERR_INVALID, "unknown dtype");
    }
    B200_TRY(ensure_device());
    b200_corpus *c = new b200_corpus();
    c->metric = metric;
    c->dtype = dtype;
    c->d = d;
    c->d_pad = pad_for(dtype, d);
    c->row_bytes = dtype == B200_DTYPE_BIN ? c->d_pad : (int64_t)c->d_pad * (dtype == B200_DTYPE_BF16 ? 2 : 4);
    c->cap = capacity_rows;
    cudaGetDevice(&c->device);
    c->sms = num_sms();
    if (const char *ev = getenv("B200_GEMM_SYNC_SLACK")) c->sync_slack = atoi(ev);
    if (const char *ev = getenv("B200_GEMM_RESCORE_L2")) c->rescore_l2 = atoi(ev);
    if (const char *ev = getenv("B200_FUSED_SCAN")) c->fused_enabled = atoi(ev);
    if (const char *ev = getenv("B200_GEMM_TS")) c->gemm_ts = atoi(ev);
    if (const char *ev = getenv("B200_GEMM_MULTICAST")) c->gemm_multicast = atoi(ev);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return fail(B200_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    *out = c;
    return B200_OK;
}

static int corpus_alloc(b200_corpus *c, int64_t rows) {
    if (c->data && !c->owns) return fail(B200_ERR_INVALID, "corpus adopted device memory; cannot append");
    // +1 row of slack so that 16-byte vector loads of the last row never leave the allocation
    const size_t need = (size_t)(rows + 1) * c->row_bytes + 256;
    const bool want_scale = c->metric == B200_METRIC_COSINE, want_bias = c->metric == B200_METRIC_L2;
    if (!c->data || c->data_cap_bytes < need) {
        void *nd = nullptr;
        cudaError_t e = cudaMalloc(&nd, need);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B200_ERR_NOMEM, std::string("cudaMalloc corpus: ") + cudaGetErrorString(e));
        }
        if (c->data && c->n) {
            cudaMemcpyAsync(nd, c->data, (size_t)c->n * c->row_bytes, cudaMemcpyDeviceToDevice, c->stream);
            cudaStreamSynchronize(c->stream);
        }
        if (c->data) cudaFree(c->data);
        c->data = nd;
        c->data_cap_bytes = need;
    }
    if ((want_scale && (!c->row_scale || c->side_cap_rows < rows)) || (want_bias && (!c->row_bias || c->side_cap_rows < rows))) {
        float *ns = nullptr, *nb = nullptr;
        if (want_scale && cudaMalloc(&ns, (size_t)rows * 4 + 256) != cudaSuccess) return fail(B200_ERR_NOMEM, "cudaMalloc row_scale");
        if (want_bias && cudaMalloc(&nb, (size_t)rows * 4 + 256) != cudaSuccess) return fail(B200_ERR_NOMEM, "cudaMalloc row_bias");
        if (c->n) {
            if (ns && c->row_scale) cudaMemcpyAsync(ns, c->row_scale, (size_t)c->n * 4, cudaMemcpyDeviceToDevice, c->stream);
            if (nb && c->row_bias) cudaMemcpyAsync(nb, c->row_bias, (size_t)c->n * 4, cudaMemcpyDeviceToDevice, c->stream);
            cudaStreamSynchronize(c->stream);
        }
        if (c->row_scale) cudaFree(c->row_scale);
        if (c->row_bias) cudaFree(c->row_bias);
        c->row_scale = ns;
        c->row_bias = nb;
        c->side_cap_rows = rows;
    }
    c->cap = std::max(c->cap, rows);
    c->owns = true;
    return B200_OK;
}

static int corpus_norms(b200_corpus *c, int64_t first, int64_t n) {
    const char *rows = reinterpret_cast<const char *>(c->data) + first * c->row_bytes;
    if (c->metric == B200_METRIC_COSINE)
        B200_CUDA_OK(launch_row_norms(rows, c->dtype == B200_DTYPE_BF16, c->d_pad, n, 1, c->row_scale + first, c->stream));
    if (c->metric == B200_METRIC_L2)
        B200_CUDA_OK(launch_row_norms(rows, c->dtype == B200_DTYPE_BF16, c->d_pad, n, 0, c->row_bias + first, c->stream));
    return B200_OK;
}

extern "C" int b200_corpus_append(b200_corpus *c, const void *rows, int64_t n) {
    if (!c || (!rows && n > 0) || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    if (n == 0) return B200_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA_OK(cudaSetDevice(c->device));
    B200_TRY(corpus_alloc(c, std::max(c->n + n, c->cap)));
    char *dst = reinterpret_cast<char *>(c->data) + c->n * c->row_bytes;
    if (c->dtype == B200_DTYPE_BIN || (c->dtype == B200_DTYPE_F32 && c->d == c->d_pad)) {
        // rows already have the HBM layout: pageable host memory -> device through the pinned ring (ingest.cu)
        B200_TRY(staged_h2d(dst, rows, (size_t)n * c->row_bytes, c->device, c->stream));
    } else {
        // stage raw fp32 rows in chunks, then pad / convert on device
        const int64_t chunk = std::max<int64_t>(1, (int64_t)(256ll << 20) / ((int64_t)c->d * 4));
        for (int64_t off = 0; off < n; off += chunk) {
            const int64_t m = std::min(chunk, n - off);
            B200_TRY(c->w_raw.reserve((size_t)m * c->d * 4));
            B200_TRY(staged_h2d(c->w_raw.p, reinterpret_cast<const float *>(rows) + off * c->d, (size_t)m * c->d * 4, c->device,
                                c->stream));
            char *dd = dst + off * c->row_bytes;
            if (c->dtype == B200_DTYPE_BF16)
                B200_CUDA_OK(launch_f32_to_bf16_rows(c->w_raw.as<float>(), c->d, dd, c->d_pad, m, c->stream));
            else
                B200_CUDA_OK(launch_pad_rows_f32(c->w_raw.as<float>(), c->d, reinterpret_cast<float *>(dd), c->d_pad, m, c->stream));
            B200_CUDA_OK(cudaStreamSynchronize(c->stream));  // w_raw is reused
        }
    }
    if (c->dtype != B200_DTYPE_BIN) B200_TRY(corpus_norms(c, c->n, n));
    B200_CUDA_OK(cudaStreamSynchronize(c->stream));
    c->n += n;
    return B200_OK;
}

extern "C" int b200_corpus_memory_bytes(const b200_corpus *c, uint64_t *out_bytes) {
    if (!c || !out_bytes) return fail(B200_ERR_INVALID, "bad arguments");
    // rows (as allocated; adopted rows belong to the caller but still occupy HBM) + per-row side arrays
    uint64_t b = c->owns ? (uint64_t)c->data_cap_bytes : (uint64_t)c->n * (uint64_t)c->row_bytes;
    if (c->row_scale) b += (uint64_t)std::max(c->side_cap_rows, c->n) * 4;
    if (c->row_bias) b += (uint64_t)std::max(c->side_cap_rows, c->n) * 4;
    *out_bytes = b;
    return B200_OK;
}

extern "C" int b200_corpus_adopt_device(b200_corpus *c, const void *device_rows, int64_t n) {
    if (!c || !device_rows || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->data) return fail(B200_ERR_INVALID, "corpus already holds data");
    if (c->dtype != B200_DTYPE_BIN && c->d != c->d_pad)
        return fail(B200_ERR_INVALID, "adopted rows must already be padded (d % 64 == 0 for bf16, d % 4 == 0 for f32)");
    B200_CUDA_OK(cudaSetDevice(c->device));
    c->data = const_cast<void *>(device_rows);
    c->owns = false;
    c->n = n;
    c->cap = n;
    if (c->metric == B200_METRIC_COSINE) B200_CUDA_OK(cudaMalloc(&c->row_scale, (size_t)n * 4 + 256));
    if (c->metric == B200_METRIC_L2) B200_CUDA_OK(cudaMalloc(&c->row_bias, (size_t)n * 4 + 256));
    c->side_cap_rows = n;
    if (c->dtype != B200_DTYPE_BIN) B200_TRY(corpus_norms(c, 0, n));
    B200_CUDA_OK(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

extern "C" int b200_corpus_size(const b200_corpus *c, int64_t *out_rows) {
    if (!c || !out_rows) return fail(B200_ERR_INVALID, "null argument");
    *out_rows = c->n;
    return B200_OK;
}

namespace b200 {
// slots of a per-thread top-k list (kernels.h): k, unless B200_LIST_APPEND_MIN_K selects the append form for this k
int list_cap_for(int k) {
    static const int min_k = getenv("B200_LIST_APPEND_MIN_K") ? atoi(getenv("B200_LIST_APPEND_MIN_K")) : 0;
    // default: the two-level (tournament) form from k = 17 (k = 100: 25.7 ms per launch against 36.8 for the plain rescan, k = 64:
    // 18.7 against 22.3; k <= 16 keeps the plain form, which is what the k = 10 headline runs); B200_LIST_TOURN_MIN_K=0 turns it off
    static const int tourn_k = getenv("B200_LIST_TOURN_MIN_K") ? atoi(getenv("B200_LIST_TOURN_MIN_K")) : 17;
    if (min_k > 0 && k >= min_k) return list_cap_append(k);
    if (tourn_k > 0 && k >= tourn_k) return list_cap_tourn(k);
    return k;
}
}  // namespace b200

extern "C" int b200_corpus_set_path(b200_corpus *c, int path) {
    if (!c || path < 0 || path > 7) return fail(B200_ERR_INVALID, "path must be 0..7");
    std::lock_guard<std::mutex> lk(c->mu);
    // 0 auto | 1 scan | 2 tensor cores, the variant auto picks (streaming operands, TMA multicast in the largest cluster
    // the batch allows) | 3 single-CTA MMAs <1,1> | 4 CTA pairs without multicast <2,1> | 5 at most two pairs per
    // cluster <2,2> | 6 up to four pairs per cluster <2,4> (= 2, explicit) | 7 queries stationary in TMEM (TS form)
    c->path = path >= 2 ? 2 : path;
    c->gemm_cta_group = path == 3 ? 1 : 0;
    c->gemm_ts = path == 7 ? 2 : 0;
    c->gemm_multicast = path == 4 ? 0 : path == 5 ? 2 : 1;
    if (path == 0) {  // auto honours the environment overrides again
        if (const char *ev = getenv("B200_GEMM_TS")) c->gemm_ts = atoi(ev);
        if (const char *ev = getenv("B200_GEMM_MULTICAST")) c->gemm_multicast = atoi(ev);
    }
    return B200_OK;
}

extern "C" int b200_corpus_last_variant(b200_corpus *c, int *kernel, int *cta_group, int *pairs_per_cluster, int *grid) {
    if (!c) return fail(B200_ERR_INVALID, "null corpus");
    std::lock_guard<std::mutex> lk(c->mu);
    if (kernel) *kernel = c->last_kernel;
    if (cta_group) *cta_group = c->last_cg;
    if (pairs_per_cluster) *pairs_per_cluster = c->last_mc;
    if (grid) *grid = c->last_grid;
    return B200_OK;
}

extern "C" int b200_corpus_enable_timing(b200_corpus *c, int on) {
    if (!c) return fail(B200_ERR_INVALID, "null corpus");
    std::lock_guard<std::mutex> lk(c->mu);
    c->timing = on != 0;
    return B200_OK;
}

extern "C" int b200_corpus_kernel_time(b200_corpus *c, int reset, double *out_total_ms, int64_t *out_launches) {
    if (!c || !out_total_ms || !out_launches) return fail(B200_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA_OK(cudaSetDevice(c->device));
    timing_drain(c);
    *out_total_ms = c->timed_ms;
    *out_launches = c->timed_launches;
    if (reset) {
        c->timed_ms = 0;
        c->timed_launches = 0;
    }
    return B200_OK;
}

extern "C" int b200_corpus_free(b200_corpus *c) {
    if (!c) return B200_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->data && c->owns) cudaFree(c->data);
    if (c->row_scale) cudaFree(c->row_scale);
    if (c->row_bias) cudaFree(c->row_bias);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->d_tickets) cudaFree(c->d_tickets);
    for (DevBuf *b : {&c->w_raw, &c->w_q32, &c->w_qbf, &c->w_qlo, &c->w_qnorm, &c->w_pk, &c->w_pi, &c->w_lk, &c->w_li, &c->w_alive,
                      &c->w_odis, &c->w_oids, &c->w_stage, &c->w_prog})
        b->release();
    for (auto *v : {&c->ev_used, &c->ev_free})
        for (auto &ev : *v) {
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return B200_OK;
}

// ------------------------------------------------------------------------------------
// search core: everything on device, asynchronous on `s`
// d_queries: raw device fp32 [nq][d] (or bytes [nq][d/8] for binary corpora)
// ------------------------------------------------------------------------------------
static int search_core(b200_corpus *c, const void *d_queries, int64_t nq, int k, const uint8_t *d_alive, int64_t id_offset,
                       int ip_min_quirk, float *d_out_dis, int64_t *d_out_ids, cudaStream_t s) {
    if (k <= 0) return fail(B200_ERR_INVALID, "k must be positive");
    if (nq == 0) return B200_OK;
    if (c->n >= (int64_t)0xffffffffll) return fail(B200_ERR_UNSUPPORTED, "corpus shards are limited to 2^32 - 1 rows");
    const int sms = c->sms;

    if (c->dtype == B200_DTYPE_BIN) {
        if (k > 1024) return fail(B200_ERR_UNSUPPORTED, "k > 1024 not supported on the binary scan path");
        int blocks_x = (int)std::min<int64_t>(std::max<int64_t>(1, ceil_div(c->n, 256)), std::max<int64_t>(1, (2 * sms) / std::max<int64_t>(1, std::min<int64_t>(nq, 2 * sms))));
        B200_TRY(c->w_pk.reserve((size_t)nq * blocks_x * k * 4));
        B200_TRY(c->w_pi.reserve((size_t)nq * blocks_x * k * 4));
        BinaryScanParams bp{};
        bp.corpus = reinterpret_cast<const uint8_t *>(c->data);
        bp.queries = reinterpret_cast<const uint8_t *>(d_queries);
        bp.alive = d_alive;
        bp.part_keys = c->w_pk.as<float>();
        bp.part_ids = c->w_pi.as<uint32_t>();
        bp.n = c->n;
        bp.nq = nq;
        bp.nbytes = c->d_pad;
        bp.k = k;
        bp.jaccard = c->metric == B200_METRIC_JACCARD;
        B200_CUDA_OK(launch_binary_scan(bp, blocks_x, s));
        MergeParams mp{};
        mp.in_keys = bp.part_keys;
        mp.in_ids = bp.part_ids;
        mp.list_stride = k;
        mp.q_stride = (int64_t)blocks_x * k;
        mp.n_lists = blocks_x;
        mp.k_in = k;
        mp.k = k;
        mp.nq = nq;
        mp.out_mode = kOutKey;
        mp.id_offset = id_offset;
        mp.out_dis = d_out_dis;
        mp.out_ids = d_out_ids;
        B200_CUDA_OK(launch_topk_merge(mp, false, s));
        return B200_OK;
    }

    // ---- stage queries: pad to d_pad fp32, cosine -> normalise (VIWithDataPart.h:354-360)
    B200_TRY(c->w_q32.reserve((size_t)nq * c->d_pad * 4));
    float *q32 = c->w_q32.as<float>();
    B200_CUDA_OK(launch_pad_rows_f32(reinterpret_cast<const float *>(d_queries), c->d, q32, c->d_pad, nq, s));
    int path = c->path;
    // auto: when the batch goes to the tensor cores (bf16 rows: kind::f16 GEMM; fp32 rows: 3xTF32 split GEMM, same
    // accuracy class as the fp32 FMA scan).  Measured crossovers (profiles/r01_small_batch.log, 4M x 768): bf16 rows
    // 0.92 ms per <= 128-query batch (the HBM floor) vs 1.3 ms for ONE query on the scan; fp32 rows 3.6 ms vs 5.1 ms
    // for 8 queries on the scan.  L2 keeps faiss' own switch: below distance_compute_blas_threshold = 20 queries the
    // reference sums exact differences, from 20 up it uses ||x||^2 + ||y||^2 - 2xy like the GEMM kernels do (which
    // cancels badly for far-from-origin data), so L2 batches move to the tensor cores at 20.  Very large k stays on
    // the scan path, whose lists are warp-cooperative.
    if (path == 0) {
        const int64_t min_nq = c->metric == B200_METRIC_L2 ? 20 : (c->dtype == B200_DTYPE_BF16 ? 2 : 5);
        path = (nq >= min_nq && k <= (c->dtype == B200_DTYPE_BF16 ? 1024 : 256)) ? 2 : 1;
    }
    if (c->n == 0) path = 1;  // nothing to tile: the scan kernel exits at once and the merge emits the empty result
    // scan path: queries normalised in fp32 like the reference.  GEMM path: the bf16 operand
    // keeps the caller's values (normalising first would add a bf16 rounding of the unit
    // vector); the positive per-query factor 1/||q|| is applied when the result is emitted.
    if (c->metric == B200_METRIC_COSINE && path == 1) B200_CUDA_OK(launch_normalize_rows_f32(q32, c->d_pad, nq, s));

    const int out_mode_scan = c->metric == B200_METRIC_L2 ? kOutKey : c->metric == B200_METRIC_IP ? kOutNeg : kOutOnePlus;

    if (path == 1) {
        if (k > 2048) return fail(B200_ERR_UNSUPPORTED, "k > 2048 not supported on the scan path");
        int qt = nq == 1 ? 1 : nq <= 4 ? 4 : 8;
        while (qt > 1 && scan_smem_bytes(qt, c->d_pad, k) > 100 * 1024) qt = qt == 8 ? 4 : 1;
        if (scan_smem_bytes(qt, c->d_pad, k) > 200 * 1024)
            return fail(B200_ERR_UNSUPPORTED, "d * 4 + 64 * k exceeds the shared-memory budget of the scan kernel");
        const int elems = c->dtype == B200_DTYPE_BF16 ? 8 : 4;
        const int chunks = c->d_pad / elems;
        int group = 1;
        while (group < 32 && group < chunks) group <<= 1;
        const int64_t y_tiles = ceil_div(nq, qt);
        const int64_t rows_per_block_step = 8 * (32 / group);
        int64_t bx = std::max<int64_t>(1, (2 * sms) / std::min<int64_t>(y_tiles, 2 * sms));
        bx = std::min<int64_t>(bx, std::max<int64_t>(1, ceil_div(c->n, rows_per_block_step)));
        const int blocks_x = (int)bx;
        B200_TRY(c->w_pk.reserve((size_t)nq * blocks_x * k * 4));
        B200_TRY(c->w_pi.reserve((size_t)nq * blocks_x * k * 4));
        ScanParams sp{};
        sp.corpus = c->data;
        sp.queries = q32;
        sp.row_scale = c->metric == B200_METRIC_COSINE ? c->row_scale : nullptr;
        sp.alive = d_alive;
        sp.part_keys = c->w_pk.as<float>();
        sp.part_ids = c->w_pi.as<uint32_t>();
        sp.n = c->n;
        sp.nq = nq;
        sp.row_bytes = c->row_bytes;
        sp.d_pad = c->d_pad;
        sp.k = k;
        sp.group = group;
        sp.l2 = c->metric == B200_METRIC_L2;
        sp.bf16 = c->dtype == B200_DTYPE_BF16;
        std::pair<cudaEvent_t, cudaEvent_t> ev;
        timing_begin(c, s, ev);
        B200_CUDA_OK(launch_flat_scan(sp, qt, blocks_x, s));
        timing_end(c, s, ev);
        c->last_kernel = B200_KERNEL_SCAN; c->last_cg = 0; c->last_mc = 0; c->last_grid = blocks_x;
        MergeParams mp{};
        mp.in_keys = sp.part_keys;
        mp.in_ids = sp.part_ids;
        mp.list_stride = k;
        mp.q_stride = (int64_t)blocks_x * k;
        mp.n_lists = blocks_x;
        mp.k_in = k;
        mp.k = k;
        mp.nq = nq;
        mp.out_mode = out_mode_scan;
        mp.ip_min_quirk = ip_min_quirk && c->metric == B200_METRIC_IP;
        mp.id_offset = id_offset;
        mp.out_dis = d_out_dis;
        mp.out_ids = d_out_ids;
        B200_CUDA_OK(launch_topk_merge(mp, false, s));
        return B200_OK;
    }

    // ---- path 2: tcgen05 GEMM with fused top-k, <= 1024 queries (8 query tiles) per launch
    if (k > 1024) return fail(B200_ERR_UNSUPPORTED, "k > 1024 not supported on the GEMM path");
    const int64_t QCHUNK = 1024;
    for (int64_t qb = 0; qb < nq; qb += QCHUNK) {
        const int64_t nq_c = std::min(QCHUNK, nq - qb);
        const bool f32 = c->dtype == B200_DTYPE_F32;
        // >= 2 query tiles: CTA pairs (tcgen05 cta_group::2), query tiles padded to an even count;
        // up to 128 queries: one CTA per MMA (a pair would spend half of its rows on padding)
        const int cta_group = (nq_c > 128 && c->gemm_cta_group != 1) ? 2 : 1;
        const int nq_pad = (int)round_up(nq_c, 128 * cta_group);
        const int q_tiles = nq_pad / 128;
        if (f32) {
            // fp32 rows: queries split once per batch into TF32 hi / lo planes (ip_gemm_tf32x3_sm100.cu)
            B200_TRY(c->w_qbf.reserve((size_t)nq_pad * c->d_pad * 4));
            B200_TRY(c->w_qlo.reserve((size_t)nq_pad * c->d_pad * 4));
            B200_CUDA_OK(launch_split_tf32(q32 + qb * c->d_pad, nq_c, c->d_pad, c->w_qbf.as<float>(), c->w_qlo.as<float>(), nq_pad, s));
        } else {
            B200_TRY(c->w_qbf.reserve((size_t)nq_pad * c->d_pad * 2));
            B200_CUDA_OK(cudaMemsetAsync(c->w_qbf.p, 0, (size_t)nq_pad * c->d_pad * 2, s));
            B200_CUDA_OK(launch_f32_to_bf16_rows(q32 + qb * c->d_pad, c->d_pad, c->w_qbf.p, c->d_pad, nq_c, s));
        }
        const float *q_add = nullptr;
        if (c->metric == B200_METRIC_L2 || c->metric == B200_METRIC_COSINE) {
            // L2: ||q||^2 of the operand the MMA sees (bf16-rounded / fp32); cosine: -(1/||q||) (mode 1 stores the negative)
            B200_TRY(c->w_qnorm.reserve((size_t)nq_pad * 4));
            if (f32)
                B200_CUDA_OK(launch_row_norms(q32 + qb * c->d_pad, 0, c->d_pad, nq_c, c->metric == B200_METRIC_L2 ? 0 : 1,
                                              c->w_qnorm.as<float>(), s));
            else
                B200_CUDA_OK(launch_row_norms(c->w_qbf.p, 1, c->d_pad, nq_c, c->metric == B200_METRIC_L2 ? 0 : 1,
                                              c->w_qnorm.as<float>(), s));
            q_add = c->w_qnorm.as<float>();
        }
        // 2 or 4 CTA pairs per cluster share every corpus tile through TMA multicast when the query tiles allow it
        // (gemm_multicast: 0 off, 2 / 4 pairs per cluster, 1 = auto: 4 when q_tiles % 8 == 0 else 2)
        int pairs = 1;
        if (cta_group == 2 && c->gemm_multicast && !f32) {
            const int want = c->gemm_multicast == 1 ? 4 : c->gemm_multicast;
            if (want >= 4 && q_tiles % 8 == 0) pairs = 4;
            else if (want >= 2 && q_tiles % 4 == 0) pairs = 2;
        }
        int grid = gemm_topk_grid(q_tiles, c->n, sms);
        while (pairs > 1) {
            // persistent + paced kernel: never launch more clusters than can be co-resident
            const int maxc = gemm_topk_max_clusters(2, pairs, k);
            const int groups = q_tiles / (2 * pairs);
            const int per_group = maxc / groups;
            if (per_group >= 1) {
                grid = std::min(grid, per_group * groups * 2 * pairs);
                break;
            }
            pairs /= 2;
        }
        grid = (grid / q_tiles) * q_tiles;
        if (grid < q_tiles) grid = q_tiles;
        B200_TRY(c->w_pk.reserve((size_t)grid * 128 * k * 4));
        B200_TRY(c->w_pi.reserve((size_t)grid * 128 * k * 4));
        GemmTopkParams gp{};
        gp.corpus_bf16 = c->data;
        gp.queries_bf16 = c->w_qbf.p;
        gp.queries_lo = f32 ? c->w_qlo.p : nullptr;
        gp.row_scale = c->metric == B200_METRIC_COSINE ? c->row_scale : nullptr;
        gp.scale_const = c->metric == B200_METRIC_L2 ? -2.f : -1.f;
        gp.row_bias = c->metric == B200_METRIC_L2 ? c->row_bias : nullptr;
        gp.alive = d_alive;
        gp.part_keys = c->w_pk.as<float>();
        gp.part_ids = c->w_pi.as<uint32_t>();
        {  // global scratch for the per-thread lists, used when they do not fit in shared memory (large k)
            B200_TRY(c->w_lk.reserve((size_t)grid * 128 * list_cap_for(k) * 4));
            B200_TRY(c->w_li.reserve((size_t)grid * 128 * list_cap_for(k) * 4));
            gp.list_keys_gmem = c->w_lk.as<float>();
            gp.list_ids_gmem = c->w_li.as<uint32_t>();
        }
        gp.n = c->n;
        gp.nq_pad = nq_pad;
        gp.nq_valid = (int)nq_c;
        gp.d_pad = c->d_pad;
        gp.k = k;
        gp.q_tiles = q_tiles;
        gp.cta_group = cta_group;
        gp.pairs_per_cluster = pairs;
        if (const char *dv = getenv("B200_GEMM_DEBUG")) gp.debug = atoi(dv);
        if (q_tiles > cta_group && c->sync_slack > 0) {
            B200_TRY(c->w_prog.reserve((size_t)grid * 4));
            B200_CUDA_OK(cudaMemsetAsync(c->w_prog.p, 0, (size_t)grid * 4, s));
            gp.progress = c->w_prog.as<int>();
            gp.sync_slack = c->sync_slack;
        }
        const char *detail = nullptr;
        std::pair<cudaEvent_t, cudaEvent_t> ev;
        timing_begin(c, s, ev);
        // TS (queries stationary in TMEM): measured slower than streaming at d = 768 (N = 64 MMAs are bound
        // by the 64 B/clk TMEM->tensor-core operand path: 75 vs 32 cycles per MMA); auto-enabled only
        // where a 2 x 128-column accumulator ring fits (d_pad <= 512); gemm_ts = 2 forces it.
        const bool use_ts = !f32 && cta_group == 2 && k <= 30 && gemm_topk_ts_supported(c->d_pad, q_tiles) &&
                            (c->gemm_ts == 2 || (c->gemm_ts == 1 && c->d_pad <= 512));
        cudaError_t e = f32 ? launch_gemm3_topk(gp, grid, s, &detail)
                            : use_ts ? launch_gemm_topk_ts(gp, grid, s, &detail) : launch_gemm_topk(gp, grid, s, &detail);
        timing_end(c, s, ev);
        c->last_kernel = f32 ? B200_KERNEL_GEMM_TF32X3 : use_ts ? B200_KERNEL_GEMM_TS : B200_KERNEL_GEMM_BF16;
        c->last_cg = cta_group; c->last_mc = (f32 || use_ts) ? 1 : pairs; c->last_grid = grid;
        if (e != cudaSuccess)
            return fail(B200_ERR_CUDA, std::string("gemm_topk launch: ") + (detail ? detail : cudaGetErrorString(e)));
        MergeParams mp{};
        mp.in_keys = gp.part_keys;
        mp.in_ids = gp.part_ids;
        mp.list_stride = (int64_t)nq_pad * k;
        mp.q_stride = k;
        mp.n_lists = grid / q_tiles;
        mp.k_in = k;
        mp.k = k;
        mp.nq = nq_c;
        mp.out_mode = c->metric == B200_METRIC_L2 ? kOutAddQ : c->metric == B200_METRIC_IP ? kOutNeg : kOutCosQ;
        mp.q_add = q_add;
        mp.ip_min_quirk = ip_min_quirk && c->metric == B200_METRIC_IP;
        mp.id_offset = id_offset;
        mp.out_dis = d_out_dis + qb * k;
        mp.out_ids = d_out_ids + qb * k;
        B200_CUDA_OK(launch_topk_merge(mp, false, s));
        // L2: the winners' distances from the direct difference form (the expanded form above only RANKS; it cancels for
        // data far from the origin).  The query operand is what the caller passed (fp32), the rows what is stored.
        if (c->metric == B200_METRIC_L2 && k <= 1024 && c->rescore_l2)
            B200_CUDA_OK(launch_rescore_l2(c->data, c->dtype == B200_DTYPE_BF16, c->row_bytes, c->d_pad, q32 + qb * c->d_pad, nq_c, id_offset, k,
                                           d_out_dis + qb * k, d_out_ids + qb * k, s));
    }
    return B200_OK;
}

extern "C" int b200_corpus_search_device(b200_corpus *c, const float *d_queries, int64_t nq, int k,
                                         const uint8_t *d_alive_bits, int64_t id_offset, float *d_out_dis,
                                         int64_t *d_out_ids, void *stream) {
    if (!c || (!d_queries && nq > 0) || !d_out_dis || !d_out_ids || nq < 0)
        return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : c->stream;
    B200_TRY(search_core(c, d_queries, nq, k, d_alive_bits, id_offset, 0, d_out_dis, d_out_ids, s));
    if (!stream) B200_CUDA_OK(cudaStreamSynchronize(s));
    return B200_OK;
}

// One launch per call for small batches that go to the scan kernel (the reference's usual shape: one query x many parts from a
// ThreadPool, MergeTreeSelectWithHybridSearchProcessor.cpp:1212-1241).  The query is written to mapped pinned memory and
// read by the kernel directly, the last block to finish merges the partial lists and writes the result to mapped pinned
// memory, the host waits on a flag in that memory: no pad / merge launches, no H2D / D2H copies, no stream synchronise
// (round 1: 103 us per resident call against a 16 us kernel).
static int search_host_fused(b200_corpus *c, const float *queries, int64_t nq, int k, const uint8_t *alive_bits, int ip_min_quirk,
                             float *out_dis, int64_t *out_ids, bool *done) {
    *done = false;
    if (!c->fused_enabled || c->dtype == B200_DTYPE_BIN || nq > 8 || k > 1024 || c->n == 0) return B200_OK;
    int path = c->path;
    if (path == 0) {
        const int64_t min_nq = c->metric == B200_METRIC_L2 ? 20 : (c->dtype == B200_DTYPE_BF16 ? 2 : 5);
        path = (nq >= min_nq) ? 2 : 1;
    }
    if (path != 1) return B200_OK;
    int qt = nq == 1 ? 1 : nq <= 4 ? 4 : 8;
    while (qt > 1 && scan_smem_bytes(qt, c->d_pad, k) > 100 * 1024) qt = qt == 8 ? 4 : 1;
    if (scan_smem_bytes(qt, c->d_pad, k) > 200 * 1024) return B200_OK;   // the staged path reports the error
    cudaStream_t s = c->stream;
    const size_t need = (size_t)8 * c->d * 4 + (size_t)8 * k * 12 + 64;
    if (need > c->h_pin_bytes) {
        if (c->h_pin) cudaFreeHost(c->h_pin);
        c->h_pin = nullptr;
        c->h_pin_bytes = 0;
        if (cudaHostAlloc(&c->h_pin, need, cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return B200_OK;   // no pinned memory: the staged path still works
        }
        c->h_pin_bytes = need;
    }
    if (!c->d_tickets || c->d_tickets_d != c->d) {
        if (c->d_tickets) cudaFree(c->d_tickets);
        c->d_tickets = nullptr;
        B200_CUDA_OK(cudaMalloc(&c->d_tickets, 64 + (size_t)8 * c->d * 4));
        B200_CUDA_OK(cudaMemsetAsync(c->d_tickets, 0, 64, s));
        c->d_tickets_d = c->d;
    }
    float *d_q = reinterpret_cast<float *>(reinterpret_cast<char *>(c->d_tickets) + 64);
    char *hp = reinterpret_cast<char *>(c->h_pin);
    float *h_q = reinterpret_cast<float *>(hp);
    float *h_dis = reinterpret_cast<float *>(hp + (size_t)8 * c->d * 4);
    int64_t *h_ids = reinterpret_cast<int64_t *>(hp + (size_t)8 * c->d * 4 + (size_t)round_up(8 * k * 4, 8));
    volatile unsigned int *h_flag = reinterpret_cast<volatile unsigned int *>(hp + need - 16);
    void *dp = nullptr;
    B200_CUDA_OK(cudaHostGetDevicePointer(&dp, c->h_pin, 0));
    char *dpc = reinterpret_cast<char *>(dp);
    // the query goes pinned -> device with one small async copy (reading it from the kernel over PCIe, 4 bytes per thread and
    // block, cost 45 us of a 60 us kernel at 296 blocks); the RESULT is written to mapped host memory by one block
    const bool q_inline = nq * c->d <= 256;   // small queries ride in the kernel parameters: no copy at all
    if (!q_inline) {
        memcpy(h_q, queries, (size_t)nq * c->d * 4);
        B200_CUDA_OK(cudaMemcpyAsync(d_q, h_q, (size_t)nq * c->d * 4, cudaMemcpyHostToDevice, s));
    }
    const uint8_t *d_alive = nullptr;
    if (alive_bits) {
        const size_t ab = (size_t)ceil_div(c->n, 8);
        B200_TRY(c->w_alive.reserve(ab + 16));
        B200_CUDA_OK(cudaMemcpyAsync(c->w_alive.p, alive_bits, ab, cudaMemcpyHostToDevice, s));
        d_alive = c->w_alive.as<uint8_t>();
    }
    const int elems = c->dtype == B200_DTYPE_BF16 ? 8 : 4;
    const int chunks = c->d_pad / elems;
    int group = 1;
    while (group < 32 && group < chunks) group <<= 1;
    const int64_t y_tiles = ceil_div(nq, qt);
    const int64_t rows_per_block_step = 8 * (32 / group);
    int64_t bx = std::max<int64_t>(1, (2 * c->sms) / std::min<int64_t>(y_tiles, 2 * c->sms));
    bx = std::min<int64_t>(bx, std::max<int64_t>(1, ceil_div(c->n, rows_per_block_step)));
    const int blocks_x = (int)bx;
    B200_TRY(c->w_pk.reserve((size_t)nq * blocks_x * k * 4));
    B200_TRY(c->w_pi.reserve((size_t)nq * blocks_x * k * 4));
    ScanParams sp{};
    sp.corpus = c->data;
    sp.queries = d_q;
    sp.row_scale = c->metric == B200_METRIC_COSINE ? c->row_scale : nullptr;
    sp.alive = d_alive;
    sp.part_keys = c->w_pk.as<float>();
    sp.part_ids = c->w_pi.as<uint32_t>();
    sp.n = c->n;
    sp.nq = nq;
    sp.row_bytes = c->row_bytes;
    sp.d_pad = c->d_pad;
    sp.k = k;
    sp.group = group;
    sp.l2 = c->metric == B200_METRIC_L2;
    sp.bf16 = c->dtype == B200_DTYPE_BF16;
    sp.fused = 1;
    sp.q_dim = c->d;
    sp.cosine = c->metric == B200_METRIC_COSINE;
    sp.out_mode = c->metric == B200_METRIC_L2 ? kOutKey : c->metric == B200_METRIC_IP ? kOutNeg : kOutOnePlus;
    sp.ip_min_quirk = ip_min_quirk && c->metric == B200_METRIC_IP;
    sp.id_offset = 0;
    sp.tickets = c->d_tickets;
    sp.tiles_done = c->d_tickets + 8;
    sp.out_dis = reinterpret_cast<float *>(dpc + (reinterpret_cast<char *>(h_dis) - hp));
    sp.out_ids = reinterpret_cast<int64_t *>(dpc + (reinterpret_cast<char *>(h_ids) - hp));
    sp.done_flag = reinterpret_cast<volatile unsigned int *>(dpc + need - 16);
    sp.done_value = ++c->fused_seq ? c->fused_seq : ++c->fused_seq;   // never 0
    sp.q_inline = q_inline ? 1 : 0;
    if (q_inline) memcpy(sp.qinline, queries, (size_t)nq * c->d * 4);
    static unsigned long long *dbg_ts = nullptr;   // B200_FUSED_DEBUG_TS=1: phase stamps of the kernel, printed every 64th call
    static const bool dbg_on = getenv("B200_FUSED_DEBUG_TS") && atoi(getenv("B200_FUSED_DEBUG_TS"));
    if (dbg_on && !dbg_ts) cudaMallocManaged(&dbg_ts, 16 * sizeof(unsigned long long));
    sp.debug_ts = dbg_on ? dbg_ts : nullptr;
    *h_flag = 0;
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    timing_begin(c, s, ev);
    B200_CUDA_OK(launch_flat_scan(sp, qt, blocks_x, s));
    timing_end(c, s, ev);
    c->last_kernel = B200_KERNEL_SCAN; c->last_cg = 0; c->last_mc = 0; c->last_grid = blocks_x;
    // wait for the flag; look at the stream now and then so that a failed launch cannot hang the caller
    for (uint64_t spin = 1;; spin++) {
        if (*h_flag == sp.done_value) break;
        if ((spin & 0x3fff) == 0) {
            const cudaError_t qe = cudaStreamQuery(s);
            if (qe == cudaSuccess) {
                if (*h_flag == sp.done_value) break;
                return fail(B200_ERR_CUDA, "fused scan finished without publishing its result");
            }
            if (qe != cudaErrorNotReady) return fail(B200_ERR_CUDA, std::string("fused scan: ") + cudaGetErrorString(qe));
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    if (dbg_on && (c->fused_seq & 63) == 0) {
        cudaStreamSynchronize(s);
        fprintf(stderr, "fused ts (us since block 0 start): scan_done %.1f tail_start %.1f fence %.1f staged %.1f bounds %.1f compacted %.1f lists %.1f merged %.1f written %.1f sysfence %.1f survivors %llu\n",
                (dbg_ts[1] - dbg_ts[0]) * 1e-3, (dbg_ts[2] - dbg_ts[0]) * 1e-3, (dbg_ts[3] - dbg_ts[0]) * 1e-3, (dbg_ts[4] - dbg_ts[0]) * 1e-3,
                (dbg_ts[5] - dbg_ts[0]) * 1e-3, (dbg_ts[6] - dbg_ts[0]) * 1e-3, (dbg_ts[7] - dbg_ts[0]) * 1e-3, (dbg_ts[8] - dbg_ts[0]) * 1e-3,
                (dbg_ts[9] - dbg_ts[0]) * 1e-3, (dbg_ts[10] - dbg_ts[0]) * 1e-3, dbg_ts[15]);
    }
    memcpy(out_dis, h_dis, (size_t)nq * k * 4);
    memcpy(out_ids, h_ids, (size_t)nq * k * 8);
    *done = true;
    return B200_OK;
}

static int search_host(b200_corpus *c, const void *queries, int64_t nq, int k, const uint8_t *alive_bits, int ip_min_quirk,
                       float *out_dis, int64_t *out_ids) {
    if (!c || (!queries && nq > 0) || !out_dis || !out_ids || nq < 0 || k <= 0)
        return fail(B200_ERR_INVALID, "bad arguments");
    if (nq == 0) return B200_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    B200_CUDA_OK(cudaSetDevice(c->device));
    {
        bool done = false;
        B200_TRY(search_host_fused(c, reinterpret_cast<const float *>(queries), nq, k, alive_bits, ip_min_quirk, out_dis, out_ids, &done));
        if (done) return B200_OK;
    }
    cudaStream_t s = c->stream;
    const size_t q_bytes = c->dtype == B200_DTYPE_BIN ? (size_t)nq * c->d_pad : (size_t)nq * c->d * 4;
    B200_TRY(c->w_stage.reserve(q_bytes));
    B200_TRY(c->w_odis.reserve((size_t)nq * k * 4));
    B200_TRY(c->w_oids.reserve((size_t)nq * k * 8));
    B200_CUDA_OK(cudaMemcpyAsync(c->w_stage.p, queries, q_bytes, cudaMemcpyHostToDevice, s));
    const uint8_t *d_alive = nullptr;
    if (alive_bits) {
        const size_t ab = (size_t)ceil_div(c->n, 8);
        B200_TRY(c->w_alive.reserve(ab + 16));
        B200_CUDA_OK(cudaMemcpyAsync(c->w_alive.p, alive_bits, ab, cudaMemcpyHostToDevice, s));
        d_alive = c->w_alive.as<uint8_t>();
    }
    B200_TRY(search_core(c, c->w_stage.p, nq, k, d_alive, 0, ip_min_quirk, c->w_odis.as<float>(), c->w_oids.as<int64_t>(), s));
    B200_CUDA_OK(cudaMemcpyAsync(out_dis, c->w_odis.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaMemcpyAsync(out_ids, c->w_oids.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    return B200_OK;
}

extern "C" int b200_corpus_search(b200_corpus *c, const float *queries, int64_t nq, int k, const uint8_t *alive_bits,
                                  float *out_dis, int64_t *out_ids) {
    return search_host(c, queries, nq, k, alive_bits, 0, out_dis, out_ids);
}

// The host-buffer entry points (b200_flat_knn / b200_binary_knn / b200_part_scan) are called once per part or per
// mark by each ClickHouse worker thread: they reuse one scratch corpus per thread (device buffers, stream and
// workspaces grow-only) instead of paying cudaMalloc / cudaStreamCreate on every call.
struct ScratchCorpus {
    b200_corpus *c = nullptr;
    ~ScratchCorpus() {
        if (c) b200_corpus_free(c);
    }
};
static thread_local ScratchCorpus t_scratch;

// gives this thread's scratch corpus (device buffers sized for the largest part it has scanned) back to the driver
extern "C" int b200_thread_release(void) {
    if (t_scratch.c) {
        b200_corpus_free(t_scratch.c);
        t_scratch.c = nullptr;
    }
    return B200_OK;
}

static int scratch_corpus(int metric, int dtype, int d, int64_t rows, b200_corpus **out) {
    b200_corpus *c = t_scratch.c;
    int dev = 0;
    cudaGetDevice(&dev);
    if (c && (c->device != dev || !c->owns)) {
        b200_corpus_free(c);
        c = t_scratch.c = nullptr;
    }
    if (!c) {
        B200_TRY(b200_corpus_create(metric, dtype, d, rows, &c));
        t_scratch.c = c;
        *out = c;
        return B200_OK;
    }
    // re-dimension in place; corpus_alloc() keeps the device buffers whenever they are large enough
    c->metric = metric;
    c->dtype = dtype;
    c->d = d;
    c->d_pad = pad_for(dtype, d);
    c->row_bytes = dtype == B200_DTYPE_BIN ? c->d_pad : (int64_t)c->d_pad * (dtype == B200_DTYPE_BF16 ? 2 : 4);
    c->n = 0;
    c->cap = 0;
    c->path = 0;
    (void)rows;
    *out = c;
    return B200_OK;
}

static void fill_empty(int metric, int64_t n, float *dis, int64_t *ids, int quirk) {
    for (int64_t i = 0; i < n; i++) {
        ids[i] = -1;
        dis[i] = quirk ? FLT_MIN : (metric == B200_METRIC_IP ? -FLT_MAX : FLT_MAX);
    }
}

extern "C" int b200_flat_knn(int metric, const float *x, int64_t nx, const float *y, int64_t ny, int d, int k,
                             const uint8_t *alive_bits, float *out_dis, int64_t *out_ids) {
    if (!is_float_metric(metric)) return fail(B200_ERR_INVALID, "b200_flat_knn: metric must be L2, IP or COSINE");
    if (nx < 0 || ny < 0 || d <= 0 || k <= 0 || !out_dis || !out_ids) return fail(B200_ERR_INVALID, "bad arguments");
    B200_TRY(ensure_device());
    if (nx == 0) return B200_OK;
    if (ny == 0) {
        fill_empty(metric, nx * k, out_dis, out_ids, 0);
        return B200_OK;
    }
    b200_corpus *c = nullptr;
    B200_TRY(scratch_corpus(metric, B200_DTYPE_F32, d, ny, &c));
    int rc = b200_corpus_append(c, y, ny);
    if (rc == B200_OK) rc = search_host(c, x, nx, k, alive_bits, 0, out_dis, out_ids);
    return rc;
}

extern "C" int b200_binary_knn(int metric, const uint8_t *x, int64_t nx, const uint8_t *y, int64_t ny, int nbytes, int k,
                               const uint8_t *alive_bits, float *out_dis, int64_t *out_ids) {
    if (!is_bin_metric(metric)) return fail(B200_ERR_INVALID, "b200_binary_knn: metric must be HAMMING or JACCARD");
    if (nx < 0 || ny < 0 || nbytes <= 0 || k <= 0 || !out_dis || !out_ids) return fail(B200_ERR_INVALID, "bad arguments");
    B200_TRY(ensure_device());
    if (nx == 0) return B200_OK;
    if (ny == 0) {
        fill_empty(metric, nx * k, out_dis, out_ids, 0);
        return B200_OK;
    }
    b200_corpus *c = nullptr;
    B200_TRY(scratch_corpus(metric, B200_DTYPE_BIN, nbytes * 8, ny, &c));
    int rc = b200_corpus_append(c, y, ny);
    if (rc == B200_OK) rc = search_host(c, x, nx, k, alive_bits, 0, out_dis, out_ids);
    return rc;
}

extern "C" int b200_part_scan(int metric, const void *x, int64_t nx, const void *y, int64_t ny, int d, int k,
                              int64_t block_rows, const uint8_t *row_exists, const uint8_t *filter_bits, float *out_dis,
                              int64_t *out_ids) {
    (void)block_rows;
    if (nx < 0 || ny < 0 || d <= 0 || k <= 0 || !out_dis || !out_ids) return fail(B200_ERR_INVALID, "bad arguments");
    if (!is_float_metric(metric) && !is_bin_metric(metric)) return fail(B200_ERR_INVALID, "unknown metric");
    B200_TRY(ensure_device());
    if (nx == 0) return B200_OK;
    const int quirk = metric == B200_METRIC_IP;
    if (ny == 0) {
        fill_empty(metric, nx * k, out_dis, out_ids, quirk);
        return B200_OK;
    }
    // alive = filter (which already folds the lightweight-delete mask, MergeTreeVSManager.cpp:1040)
    // or, without filter, the _row_exists column
    std::vector<uint8_t> alive;
    const uint8_t *alive_ptr = filter_bits;
    if (!filter_bits && row_exists) {
        alive.assign((size_t)ceil_div(ny, 8), 0);
        for (int64_t i = 0; i < ny; i++)
            if (row_exists[i]) alive[i >> 3] |= (uint8_t)(1u << (i & 7));
        alive_ptr = alive.data();
    }
    b200_corpus *c = nullptr;
    const bool bin = is_bin_metric(metric);
    B200_TRY(scratch_corpus(metric, bin ? B200_DTYPE_BIN : B200_DTYPE_F32, d, ny, &c));
    int rc = b200_corpus_append(c, y, ny);
    if (rc == B200_OK) rc = search_host(c, x, nx, k, alive_ptr, quirk, out_dis, out_ids);
    return rc;
}

extern "C" int b200_topk_merge_device(const float *d_dis, const int64_t *d_ids, int n_lists, int64_t nq, int k,
                                      int descending, float *d_out_dis, int64_t *d_out_ids, void *stream) {
    return b200_topk_merge_device_strided(d_dis, d_ids, n_lists, nq * k, nq * k, nq, k, descending, d_out_dis, d_out_ids,
                                          stream);
}

extern "C" int b200_topk_merge_device_strided(const float *d_dis, const int64_t *d_ids, int n_lists,
                                              int64_t dis_list_stride, int64_t ids_list_stride, int64_t nq, int k,
                                              int descending, float *d_out_dis, int64_t *d_out_ids, void *stream) {
    return b200_topk_merge_device_ex(d_dis, d_ids, n_lists, dis_list_stride, ids_list_stride, nq, k, k, descending, 0, d_out_dis,
                                     d_out_ids, nullptr, stream);
}

extern "C" int b200_topk_merge_device_ex(const float *d_dis, const int64_t *d_ids, int n_lists, int64_t dis_list_stride,
                                         int64_t ids_list_stride, int64_t nq, int k_in, int k, int descending, int tie_mode,
                                         float *d_out_dis, int64_t *d_out_ids, int32_t *d_out_list, void *stream) {
    if (!d_dis || !d_ids || !d_out_dis || !d_out_ids || n_lists <= 0 || nq < 0 || k <= 0 || k_in <= 0 || tie_mode < 0 || tie_mode > 1)
        return fail(B200_ERR_INVALID, "bad arguments");
    if (k > 2048) return fail(B200_ERR_UNSUPPORTED, "k > 2048 not supported by the merge kernel");
    B200_TRY(ensure_device());
    if (nq == 0) return B200_OK;
    MergeParams mp{};
    mp.in_keys = d_dis;
    mp.in_ids = d_ids;
    mp.list_stride = dis_list_stride;
    mp.id_list_stride = ids_list_stride;
    mp.q_stride = k_in;
    mp.n_lists = n_lists;
    mp.k_in = k_in;
    mp.k = k;
    mp.nq = nq;
    mp.descending = descending;
    mp.tie_mode = tie_mode;
    mp.out_list = d_out_list;
    mp.out_dis = d_out_dis;
    mp.out_ids = d_out_ids;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    B200_CUDA_OK(launch_topk_merge(mp, true, s));
    if (!stream) B200_CUDA_OK(cudaStreamSynchronize(s));
    return B200_OK;
}
