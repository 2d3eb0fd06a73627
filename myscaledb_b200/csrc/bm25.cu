// bm25.cu -- K8: BM25 posting-list scorer behind TextSearch()/HybridSearch().
//
// Replaces TANTIVY::ffi_bm25_search / ffi_get_doc_freq / ffi_get_total_num_docs /
// ffi_get_total_num_tokens / ffi_index_multi_column_docs as called from
// TantivyIndexStore (reference: src/Storages/MergeTree/TantivyIndexStore.cpp:742, :900-998) and
// driven by MergeTreeTextSearchManager::textSearch
// (src/VectorIndex/Storages/MergeTreeTextSearchManager.cpp:69-279).
//
// The Rust crate (tantivy_search 0.1.0 over tantivy 0.21.1) is not in the reference tree; the
// scoring formula is tantivy's published BM25 (k1 1.2, b 0.75, fp32 idf, 1-byte field-norm code),
// the tokenizer is tantivy's "default" (alphanumeric runs, <= 40 bytes, lowercased).
//
// Layout (one index per part, resident in HBM): postings as two flat arrays (doc ordinal u32,
// term frequency u32) concatenated term after term, each list sorted by doc; one field-norm code
// byte per (field, doc); row id per doc.  The dictionary (term -> list) stays on the host: a
// query touches a handful of terms.
// Kernel: bm25_score_kernel, grid (slices, queries).  A thread takes one posting of one clause
// (field, term) of its query; the doc is OWNED by the first clause that contains it (binary
// search in the earlier lists), the owner adds the other clauses' contributions in clause order
// (binary search in the later lists) so sums are deterministic and equal the reference's
// term-at-a-time order; dead rows (alive bitmap) are skipped; survivors go through the same
// warp-cooperative top-k as the vector scan (key = -score, tie -> smaller doc).
// HBM-bound: sum over clauses of df * (8 B posting + 1 B field norm).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace b200 {

constexpr float kBm25K1 = 1.2f;
constexpr float kBm25B = 0.75f;
constexpr int kMaxClauses = 64;

static uint32_t g_fieldnorm[256];
static std::once_flag g_fieldnorm_once;
static void init_fieldnorm() {
    std::call_once(g_fieldnorm_once, [] {
        for (int i = 0; i < 40; i++) g_fieldnorm[i] = (uint32_t)i;
        uint64_t v = 40;
        for (int i = 40; i < 256; i++) {
            g_fieldnorm[i] = v > 0xffffffffull ? 0xffffffffu : (uint32_t)v;
            const int shift = (i - 40) / 8 + 1;
            v += shift >= 40 ? (1ull << 40) : (1ull << shift);
        }
    });
}
static uint8_t fieldnorm_to_id(uint32_t n) {
    int lo = 0, hi = 255;
    while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        if (g_fieldnorm[mid] <= n) lo = mid; else hi = mid - 1;
    }
    return (uint8_t)lo;
}

// tantivy "default" tokenizer
template <typename F>
static void tokenize_default(const char *text, F &&emit) {
    const size_t n = strlen(text);
    size_t i = 0;
    std::string tok;
    auto is_tok = [](unsigned char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 0x80; };
    while (i < n) {
        while (i < n && !is_tok((unsigned char)text[i])) i++;
        const size_t s = i;
        while (i < n && is_tok((unsigned char)text[i])) i++;
        const size_t len = i - s;
        if (len == 0 || len > 40) continue;
        tok.assign(text + s, len);
        for (auto &ch : tok)
            if (ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');
        emit(tok);
    }
}

struct TermList {
    std::vector<uint32_t> docs, tfs;
    uint64_t offset = 0;  // into the device arrays after commit
};

struct Clause {  // device-side description of one (field, term) of one query
    uint64_t offset;
    uint32_t df;
    float weight;
    uint32_t field;
    uint32_t cache;  // index of the 256-entry norm table
};

struct Bm25ScoreParams {
    const uint32_t *post_docs;
    const uint32_t *post_tfs;
    const uint8_t *fieldnorm;  // [n_fields][n_docs]
    const uint32_t *row_id;    // [n_docs]
    const uint8_t *alive;      // LSB-first over row ids, or null
    const Clause *clauses;     // all queries
    const uint32_t *clause_begin;  // [nq + 1]
    const float *caches;       // [n_caches][256]
    float *part_keys;          // [nq][gridDim.x][k]
    uint32_t *part_ids;
    uint32_t n_docs;
    int k;
    int operator_or;
};

// first index in docs[0..n) whose value is >= key, found by the whole warp: every round the 32 lanes probe 32 evenly
// spaced splitters of the remaining range (one ballot narrows it 33x), the last <= 32 entries are read side by side
__device__ __forceinline__ uint32_t warp_lower_bound(const uint32_t *docs, uint32_t n, uint32_t key, int lane) {
    uint32_t lo = 0, hi = n;  // the answer lies in [lo, hi]
    while (hi - lo > 32) {
        const uint64_t span = hi - lo;
        const uint32_t idx = lo + (uint32_t)(span * (uint32_t)(lane + 1) / 33);  // lo < idx < hi, ascending with the lane
        const unsigned less = __ballot_sync(0xffffffffu, docs[idx] < key);        // a prefix mask: the list is sorted
        const uint32_t cnt = __popc(less);
        const uint32_t nlo = cnt ? lo + (uint32_t)(span * cnt / 33) + 1 : lo;
        const uint32_t nhi = cnt < 32 ? lo + (uint32_t)(span * (cnt + 1) / 33) : hi;
        lo = nlo;
        hi = nhi;
    }
    const uint32_t j = lo + (uint32_t)lane;
    const unsigned less = __ballot_sync(0xffffffffu, j < hi && docs[j] < key);
    return lo + __popc(less);
}
// [wlo, whi): the positions in docs[0..n) of every value in [dmin, dmax]
__device__ __forceinline__ void warp_window(const uint32_t *docs, uint32_t n, uint32_t dmin, uint32_t dmax, int lane,
                                            uint32_t &wlo, uint32_t &whi) {
    wlo = warp_lower_bound(docs, n, dmin, lane);
    whi = wlo + warp_lower_bound(docs + wlo, n - wlo, dmax + 1u, lane);  // doc ordinals stay far below 2^32 - 1
}
// private binary search of one lane inside the warp's window
__device__ __forceinline__ bool window_find(const uint32_t *docs, uint32_t lo, uint32_t hi, uint32_t doc, uint32_t &pos) {
    const uint32_t end = hi;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (docs[mid] < doc) lo = mid + 1; else hi = mid;
    }
    pos = lo;
    return lo < end && docs[lo] == doc;
}

__global__ void __launch_bounds__(256) bm25_score_kernel(const Bm25ScoreParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *lk = reinterpret_cast<float *>(smem_raw);
    uint32_t *li = reinterpret_cast<uint32_t *>(lk + 8 * p.k);
    __shared__ Clause cl[kMaxClauses];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.y;
    const uint32_t c0 = p.clause_begin[q], nc = p.clause_begin[q + 1] - c0;
    for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) cl[i] = p.clauses[c0 + i];
    WarpTopK list;
    list.init(lk + (size_t)warp * p.k, li + (size_t)warp * p.k, p.k);
    for (int j = lane; j < p.k; j += 32) list.keys[j] = FLT_MAX;
    __syncthreads();

    for (uint32_t c = 0; c < nc; c++) {
        const uint32_t df = cl[c].df;
        const uint32_t *docs = p.post_docs + cl[c].offset;
        const uint32_t *tfs = p.post_tfs + cl[c].offset;
        // warp-uniform trip count
        for (uint32_t base = (blockIdx.x * 8 + warp) * 32; base < df; base += gridDim.x * 256) {
            const uint32_t i = base + lane;
            const bool valid = i < df;
            bool cand = false;
            float key = FLT_MAX;
            // the 32 postings of this step are consecutive in a sorted list: their counterparts in another clause's
            // list lie in one narrow window, which the warp brackets together (32-ary search, 3 probes for 32 K
            // postings) before each lane finishes with a short private binary search inside it
            const uint32_t doc = valid ? docs[i] : 0u;
            const uint32_t n_valid = min(32u, df - base);
            const uint32_t dmin = __shfl_sync(0xffffffffu, doc, 0);
            const uint32_t dmax = __shfl_sync(0xffffffffu, doc, (int)n_valid - 1);
            bool active = valid;  // no earlier clause owns this document
            uint32_t pos;
            for (uint32_t e = 0; e < c; e++) {
                if (!cl[e].df) continue;
                const uint32_t *od = p.post_docs + cl[e].offset;
                uint32_t wlo, whi;
                warp_window(od, cl[e].df, dmin, dmax, lane, wlo, whi);
                if (active && window_find(od, wlo, whi, doc, pos)) active = false;
            }
            bool all = active && (p.operator_or || c == 0);  // AND: only clause 0 can own a full match
            float score = 0.f;
            if (all) {
                const float tf = (float)tfs[i];
                const float norm = p.caches[(size_t)cl[c].cache * 256 + p.fieldnorm[(size_t)cl[c].field * p.n_docs + doc]];
                // explicit rn ops: no FMA contraction, so sums equal the reference's fp32 arithmetic bit for bit
                score = __fmul_rn(cl[c].weight, __fdiv_rn(tf, __fadd_rn(tf, norm)));
            }
            for (uint32_t e = c + 1; e < nc; e++) {
                if (!__any_sync(0xffffffffu, all)) break;
                if (!cl[e].df) {
                    if (!p.operator_or) all = false;
                    continue;
                }
                const uint32_t *od = p.post_docs + cl[e].offset;
                uint32_t wlo, whi;
                warp_window(od, cl[e].df, dmin, dmax, lane, wlo, whi);
                if (all) {
                    if (window_find(od, wlo, whi, doc, pos)) {
                        const float tf2 = (float)p.post_tfs[cl[e].offset + pos];
                        const float n2 = p.caches[(size_t)cl[e].cache * 256 + p.fieldnorm[(size_t)cl[e].field * p.n_docs + doc]];
                        score = __fadd_rn(score, __fmul_rn(cl[e].weight, __fdiv_rn(tf2, __fadd_rn(tf2, n2))));
                    } else if (!p.operator_or) {
                        all = false;
                    }
                }
            }
            if (all) {
                const uint32_t rid = p.row_id[doc];
                const bool live = !p.alive || ((p.alive[rid >> 3] >> (rid & 7)) & 1);
                key = -score;
                cand = live && list.passes(key, doc);
            }
            unsigned m = __ballot_sync(0xffffffffu, cand);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                list.insert(__shfl_sync(0xffffffffu, key, src), __shfl_sync(0xffffffffu, doc, src));
            }
        }
    }
    __syncthreads();
    block_rank_merge(lk, li, 8, p.k, p.k, p.part_keys + ((size_t)q * gridDim.x + blockIdx.x) * p.k,
                     p.part_ids + ((size_t)q * gridDim.x + blockIdx.x) * p.k);
}

// doc ordinal -> row id on the merged result
__global__ void bm25_finish_kernel(const float *dis, const int64_t *docs, const uint32_t *row_id, int64_t n, float *out_score,
                                   uint64_t *out_row, uint32_t *out_count, int k) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t d = docs[i];
    out_score[i] = d >= 0 ? dis[i] : 0.f;
    out_row[i] = d >= 0 ? (uint64_t)row_id[d] : ~0ull;
    if (d >= 0) atomicAdd(&out_count[i / k], 1u);
}

struct DevVec {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B200_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (cudaMalloc(&p, bytes + 256) != cudaSuccess) {
            cudaGetLastError();
            return fail(B200_ERR_NOMEM, "cudaMalloc failed in bm25");
        }
        cap = bytes;
        return B200_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace b200

using namespace b200;

struct b200_bm25 {
    uint32_t n_fields = 1;
    std::vector<std::unordered_map<std::string, uint32_t>> dict;  // per field: term -> list index
    std::vector<TermList> lists;
    std::vector<std::vector<uint32_t>> doc_len;  // [field][doc]
    std::vector<uint64_t> row_ids;
    std::vector<uint64_t> total_tokens;
    bool committed = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    DevVec d_docs, d_tfs, d_fn, d_rows, d_clauses, d_begin, d_caches, d_pk, d_pi, d_alive, d_odis, d_oids, d_score, d_row64, d_cnt;
};

static int bm25_device_ok() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(B200_ERR_NO_DEVICE, "no CUDA device visible; libb200search has no CPU fallback");
    }
    return B200_OK;
}

extern "C" int b200_bm25_create(uint32_t n_fields, b200_bm25 **out) {
    if (!out || n_fields == 0 || n_fields > 64) return fail(B200_ERR_INVALID, "bad arguments");
    *out = nullptr;
    B200_TRY(bm25_device_ok());
    init_fieldnorm();
    b200_bm25 *ix = new b200_bm25();
    ix->n_fields = n_fields;
    ix->dict.resize(n_fields);
    ix->doc_len.resize(n_fields);
    ix->total_tokens.assign(n_fields, 0);
    cudaGetDevice(&ix->device);
    if (cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ix;
        return fail(B200_ERR_CUDA, "cudaStreamCreate failed");
    }
    *out = ix;
    return B200_OK;
}

extern "C" int b200_bm25_free(b200_bm25 *ix) {
    if (!ix) return B200_OK;
    cudaSetDevice(ix->device);
    for (DevVec *v : {&ix->d_docs, &ix->d_tfs, &ix->d_fn, &ix->d_rows, &ix->d_clauses, &ix->d_begin, &ix->d_caches, &ix->d_pk,
                      &ix->d_pi, &ix->d_alive, &ix->d_odis, &ix->d_oids, &ix->d_score, &ix->d_row64, &ix->d_cnt})
        v->release();
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
    return B200_OK;
}

// ffi_index_multi_column_docs(path, row_id, column_names, column_docs): one call per row;
// texts[f] may be null; call b200_bm25_add_text again for further values of an Array(String) column.
extern "C" int b200_bm25_add_doc(b200_bm25 *ix, uint64_t row_id) {
    if (!ix) return fail(B200_ERR_INVALID, "null index");
    if (ix->committed) return fail(B200_ERR_INVALID, "index already committed");
    if (row_id >= 0xffffffffull) return fail(B200_ERR_UNSUPPORTED, "row ids are UInt32 part offsets");
    ix->row_ids.push_back(row_id);
    for (auto &dl : ix->doc_len) dl.push_back(0);
    return B200_OK;
}

extern "C" int b200_bm25_add_text(b200_bm25 *ix, uint32_t field, const char *text) {
    if (!ix || !text || field >= ix->n_fields) return fail(B200_ERR_INVALID, "bad arguments");
    if (ix->committed || ix->row_ids.empty()) return fail(B200_ERR_INVALID, "add_doc first / index already committed");
    const uint32_t doc = (uint32_t)ix->row_ids.size() - 1;
    tokenize_default(text, [&](const std::string &tok) {
        auto it = ix->dict[field].find(tok);
        uint32_t li;
        if (it == ix->dict[field].end()) {
            li = (uint32_t)ix->lists.size();
            ix->dict[field].emplace(tok, li);
            ix->lists.emplace_back();
        } else {
            li = it->second;
        }
        TermList &tl = ix->lists[li];
        if (!tl.docs.empty() && tl.docs.back() == doc) tl.tfs.back()++;
        else {
            tl.docs.push_back(doc);
            tl.tfs.push_back(1);
        }
        ix->doc_len[field][doc]++;
        ix->total_tokens[field]++;
    });
    return B200_OK;
}

// ffi_index_writer_commit: freeze and upload to HBM
extern "C" int b200_bm25_commit(b200_bm25 *ix) {
    if (!ix) return fail(B200_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (ix->committed) return B200_OK;
    B200_CUDA_OK(cudaSetDevice(ix->device));
    uint64_t total = 0;
    for (auto &tl : ix->lists) {
        tl.offset = total;
        total += tl.docs.size();
    }
    const size_t nd = ix->row_ids.size();
    std::vector<uint32_t> docs(total ? total : 1), tfs(total ? total : 1);
    for (auto &tl : ix->lists) {
        std::copy(tl.docs.begin(), tl.docs.end(), docs.begin() + tl.offset);
        std::copy(tl.tfs.begin(), tl.tfs.end(), tfs.begin() + tl.offset);
    }
    std::vector<uint8_t> fn((size_t)ix->n_fields * (nd ? nd : 1));
    for (uint32_t f = 0; f < ix->n_fields; f++)
        for (size_t d = 0; d < nd; d++) fn[(size_t)f * nd + d] = fieldnorm_to_id(ix->doc_len[f][d]);
    std::vector<uint32_t> rows(nd ? nd : 1);
    for (size_t d = 0; d < nd; d++) rows[d] = (uint32_t)ix->row_ids[d];
    B200_TRY(ix->d_docs.reserve(docs.size() * 4));
    B200_TRY(ix->d_tfs.reserve(tfs.size() * 4));
    B200_TRY(ix->d_fn.reserve(fn.size()));
    B200_TRY(ix->d_rows.reserve(rows.size() * 4));
    B200_CUDA_OK(cudaMemcpyAsync(ix->d_docs.p, docs.data(), docs.size() * 4, cudaMemcpyHostToDevice, ix->stream));
    B200_CUDA_OK(cudaMemcpyAsync(ix->d_tfs.p, tfs.data(), tfs.size() * 4, cudaMemcpyHostToDevice, ix->stream));
    B200_CUDA_OK(cudaMemcpyAsync(ix->d_fn.p, fn.data(), fn.size(), cudaMemcpyHostToDevice, ix->stream));
    B200_CUDA_OK(cudaMemcpyAsync(ix->d_rows.p, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, ix->stream));
    B200_CUDA_OK(cudaStreamSynchronize(ix->stream));
    for (auto &tl : ix->lists) {  // host copies of the postings are no longer needed (df stays)
        tl.tfs.clear();
        tl.tfs.shrink_to_fit();
    }
    ix->committed = true;
    return B200_OK;
}

extern "C" int b200_bm25_total_docs(const b200_bm25 *ix, uint64_t *out) {  // ffi_get_total_num_docs
    if (!ix || !out) return fail(B200_ERR_INVALID, "null argument");
    *out = ix->row_ids.size();
    return B200_OK;
}
extern "C" int b200_bm25_total_tokens(const b200_bm25 *ix, uint32_t field, uint64_t *out) {  // ffi_get_total_num_tokens
    if (!ix || !out || field >= ix->n_fields) return fail(B200_ERR_INVALID, "bad arguments");
    *out = ix->total_tokens[field];
    return B200_OK;
}
extern "C" int b200_bm25_doc_freq(const b200_bm25 *ix, uint32_t field, const char *term, uint64_t *out) {  // ffi_get_doc_freq
    if (!ix || !out || !term || field >= ix->n_fields) return fail(B200_ERR_INVALID, "bad arguments");
    auto it = ix->dict[field].find(term);
    *out = it == ix->dict[field].end() ? 0 : ix->lists[it->second].docs.size();
    return B200_OK;
}

// Tokenise a sentence into distinct lowercase terms, written NUL-separated into buf.
extern "C" int b200_bm25_query_terms(const char *sentence, char *buf, size_t buf_len, uint32_t *out_n) {
    if (!sentence || !buf || !out_n) return fail(B200_ERR_INVALID, "null argument");
    std::vector<std::string> terms;
    tokenize_default(sentence, [&](const std::string &t) {
        if (std::find(terms.begin(), terms.end(), t) == terms.end()) terms.push_back(t);
    });
    size_t off = 0;
    for (auto &t : terms) {
        if (off + t.size() + 1 > buf_len) return fail(B200_ERR_INVALID, "buffer too small");
        memcpy(buf + off, t.c_str(), t.size() + 1);
        off += t.size() + 1;
    }
    *out_n = (uint32_t)terms.size();
    return B200_OK;
}

// Batched ffi_bm25_search.  sentences[nq]; fields[n_fields_q] searched for every query.
// stats (nullable): table-wide statistics -- total_docs, total_tokens[n_fields] and, per query,
// doc_freq[q][fq * 64 + term] for its distinct terms in tokenisation order (up to 64 per query).
// Output per query: out_rows/out_scores [nq][topk] (score descending, ties -> smaller doc),
// out_counts[nq].
extern "C" int b200_bm25_search_batch(b200_bm25 *ix, const char *const *sentences, int64_t nq, const uint32_t *fields,
                                      uint32_t n_fields_q, uint32_t topk, const uint8_t *alive_bits, int use_filter,
                                      int operator_or, uint64_t stat_total_docs, const uint64_t *stat_total_tokens,
                                      const uint64_t *stat_doc_freq, uint64_t *out_rows, float *out_scores,
                                      uint32_t *out_counts) {
    if (!ix || !sentences || !fields || !out_rows || !out_scores || !out_counts || nq < 0 || n_fields_q == 0)
        return fail(B200_ERR_INVALID, "bad arguments");
    if (topk == 0 || topk > 2048) return fail(B200_ERR_UNSUPPORTED, "topk must be in 1..2048");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (!ix->committed) return fail(B200_ERR_INVALID, "commit the index before searching");
    if (nq == 0) return B200_OK;
    B200_CUDA_OK(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    const size_t nd = ix->row_ids.size();
    for (int64_t q = 0; q < nq; q++) out_counts[q] = 0;
    if (nd == 0) return B200_OK;

    // ---- host: tokenise, resolve terms, BM25 weights and norm tables (fp32, tantivy order of ops)
    std::vector<Clause> clauses;
    std::vector<uint32_t> begin(nq + 1, 0);
    std::vector<float> caches;
    for (int64_t q = 0; q < nq; q++) {
        begin[q] = (uint32_t)clauses.size();
        std::vector<std::string> terms;
        tokenize_default(sentences[q], [&](const std::string &t) {
            if (terms.size() < 64 && std::find(terms.begin(), terms.end(), t) == terms.end()) terms.push_back(t);
        });
        bool dead = false;  // AND with an unknown term matches nothing
        std::vector<Clause> mine;
        for (uint32_t fq = 0; fq < n_fields_q && !dead; fq++) {
            const uint32_t f = fields[fq];
            if (f >= ix->n_fields) return fail(B200_ERR_INVALID, "field out of range");
            const uint64_t N = stat_total_docs ? stat_total_docs : nd;
            const uint64_t T = stat_total_docs ? stat_total_tokens[f] : ix->total_tokens[f];
            const float avgdl = (float)T / (float)N;
            const uint32_t cache_idx = (uint32_t)(caches.size() / 256);
            for (int c = 0; c < 256; c++) caches.push_back(kBm25K1 * (1.0f - kBm25B + kBm25B * (float)g_fieldnorm[c] / avgdl));
            for (size_t t = 0; t < terms.size(); t++) {
                auto it = ix->dict[f].find(terms[t]);
                if (it == ix->dict[f].end()) {
                    if (!operator_or) dead = true;
                    continue;
                }
                const TermList &tl = ix->lists[it->second];
                const uint64_t n = stat_total_docs ? stat_doc_freq[(size_t)q * n_fields_q * 64 + fq * 64 + t] : tl.docs.size();
                const float x = ((float)(N - n) + 0.5f) / ((float)n + 0.5f);
                const float idf = logf(1.0f + x);
                Clause cl;
                cl.offset = tl.offset;
                cl.df = (uint32_t)tl.docs.size();
                cl.weight = idf * (1.0f + kBm25K1);
                cl.field = f;
                cl.cache = cache_idx;
                mine.push_back(cl);
            }
        }
        if (!operator_or && terms.empty()) dead = true;
        if (!dead) {
            if (mine.size() > (size_t)kMaxClauses) return fail(B200_ERR_UNSUPPORTED, "more than 64 (field, term) clauses in one query");
            clauses.insert(clauses.end(), mine.begin(), mine.end());
        }
    }
    begin[nq] = (uint32_t)clauses.size();
    if (clauses.empty()) return B200_OK;

    // ---- device
    const int k = (int)topk;
    uint32_t max_df = 1;
    for (auto &c : clauses) max_df = std::max(max_df, c.df);
    int bx = (int)std::min<int64_t>(std::max<int64_t>(1, ceil_div(max_df, 256 * 4)), std::max<int64_t>(1, 592 / std::max<int64_t>(1, std::min<int64_t>(nq, 592))));
    B200_TRY(ix->d_clauses.reserve(clauses.size() * sizeof(Clause)));
    B200_TRY(ix->d_begin.reserve(begin.size() * 4));
    B200_TRY(ix->d_caches.reserve(caches.size() * 4));
    B200_TRY(ix->d_pk.reserve((size_t)nq * bx * k * 4));
    B200_TRY(ix->d_pi.reserve((size_t)nq * bx * k * 4));
    B200_TRY(ix->d_odis.reserve((size_t)nq * k * 4));
    B200_TRY(ix->d_oids.reserve((size_t)nq * k * 8));
    B200_TRY(ix->d_score.reserve((size_t)nq * k * 4));
    B200_TRY(ix->d_row64.reserve((size_t)nq * k * 8));
    B200_TRY(ix->d_cnt.reserve((size_t)nq * 4));
    B200_CUDA_OK(cudaMemcpyAsync(ix->d_clauses.p, clauses.data(), clauses.size() * sizeof(Clause), cudaMemcpyHostToDevice, s));
    B200_CUDA_OK(cudaMemcpyAsync(ix->d_begin.p, begin.data(), begin.size() * 4, cudaMemcpyHostToDevice, s));
    B200_CUDA_OK(cudaMemcpyAsync(ix->d_caches.p, caches.data(), caches.size() * 4, cudaMemcpyHostToDevice, s));
    const uint8_t *d_alive = nullptr;
    if (use_filter && alive_bits) {
        uint64_t max_row = 0;
        for (auto r : ix->row_ids) max_row = std::max(max_row, r);
        const size_t ab = (size_t)(max_row / 8 + 1);
        B200_TRY(ix->d_alive.reserve(ab));
        B200_CUDA_OK(cudaMemcpyAsync(ix->d_alive.p, alive_bits, ab, cudaMemcpyHostToDevice, s));
        d_alive = reinterpret_cast<const uint8_t *>(ix->d_alive.p);
    }
    Bm25ScoreParams sp{};
    sp.post_docs = reinterpret_cast<const uint32_t *>(ix->d_docs.p);
    sp.post_tfs = reinterpret_cast<const uint32_t *>(ix->d_tfs.p);
    sp.fieldnorm = reinterpret_cast<const uint8_t *>(ix->d_fn.p);
    sp.row_id = reinterpret_cast<const uint32_t *>(ix->d_rows.p);
    sp.alive = d_alive;
    sp.clauses = reinterpret_cast<const Clause *>(ix->d_clauses.p);
    sp.clause_begin = reinterpret_cast<const uint32_t *>(ix->d_begin.p);
    sp.caches = reinterpret_cast<const float *>(ix->d_caches.p);
    sp.part_keys = reinterpret_cast<float *>(ix->d_pk.p);
    sp.part_ids = reinterpret_cast<uint32_t *>(ix->d_pi.p);
    sp.n_docs = (uint32_t)nd;
    sp.k = k;
    sp.operator_or = operator_or;
    const size_t smem = (size_t)8 * k * 8;
    B200_CUDA_OK(cudaFuncSetAttribute(bm25_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bm25_score_kernel<<<dim3(bx, (unsigned)nq), 256, smem, s>>>(sp);
    g_launches++;
    B200_CUDA_OK(cudaGetLastError());
    MergeParams mp{};
    mp.in_keys = sp.part_keys;
    mp.in_ids = sp.part_ids;
    mp.list_stride = k;
    mp.q_stride = (int64_t)bx * k;
    mp.n_lists = bx;
    mp.k_in = k;
    mp.k = k;
    mp.nq = nq;
    mp.out_mode = kOutNeg;
    mp.out_dis = reinterpret_cast<float *>(ix->d_odis.p);
    mp.out_ids = reinterpret_cast<int64_t *>(ix->d_oids.p);
    B200_CUDA_OK(launch_topk_merge(mp, false, s));
    B200_CUDA_OK(cudaMemsetAsync(ix->d_cnt.p, 0, (size_t)nq * 4, s));
    const int64_t tot = nq * k;
    bm25_finish_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, s>>>(mp.out_dis, mp.out_ids, sp.row_id, tot,
                                                                   reinterpret_cast<float *>(ix->d_score.p),
                                                                   reinterpret_cast<uint64_t *>(ix->d_row64.p),
                                                                   reinterpret_cast<uint32_t *>(ix->d_cnt.p), k);
    g_launches++;
    B200_CUDA_OK(cudaGetLastError());
    B200_CUDA_OK(cudaMemcpyAsync(out_scores, ix->d_score.p, (size_t)tot * 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaMemcpyAsync(out_rows, ix->d_row64.p, (size_t)tot * 8, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaMemcpyAsync(out_counts, ix->d_cnt.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    return B200_OK;
}

extern "C" int b200_bm25_search(b200_bm25 *ix, const char *sentence, const uint32_t *fields, uint32_t n_fields_q,
                                uint32_t topk, const uint8_t *alive_bits, int use_filter, int operator_or,
                                uint64_t stat_total_docs, const uint64_t *stat_total_tokens, const uint64_t *stat_doc_freq,
                                uint64_t *out_rows, float *out_scores, uint32_t *out_n) {
    const char *one[1] = {sentence};
    return b200_bm25_search_batch(ix, one, 1, fields, n_fields_q, topk, alive_bits, use_filter, operator_or, stat_total_docs,
                                  stat_total_tokens, stat_doc_freq, out_rows, out_scores, out_n);
}
