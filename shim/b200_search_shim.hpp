// b200_search_shim.hpp -- header-only C++ adapter that presents the names MyScaleDB's
// src/VectorIndex code calls (namespace Search::, faiss::knn_*, jaccard_knn, TANTIVY::ffi_*) and
// forwards them to the C ABI of libb200search.so (include/b200_search.h).
//
// The originals live in the un-vendored submodules contrib/search-index and
// rust/supercrate/libs/tantivy_search; the surface below is reconstructed from every call site
// in the reference, each member citing the call site it serves
// (paths relative to /root/reference/src).  Only what the hot path calls is provided.
#pragma once
#include <b200_search.h>

#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// VectorIndex/Common/VICommon.h:86-91
class SearchIndexException : public std::exception {
    int code_;
    std::string msg_;
public:
    SearchIndexException(int code, std::string msg) : code_(code), msg_(std::move(msg)) {}
    int getCode() const { return code_; }
    const char * what() const noexcept override { return msg_.c_str(); }
};

namespace Search {
inline void b200Check(int rc) { if (rc != B200_OK) throw SearchIndexException(rc, b200_last_error()); }

enum class DataType { FloatVector, BinaryVector };                    // VICommon.h:139-143
enum class Metric { L2, IP, Cosine, Hamming, Jaccard };               // MergeTreeVSManager.cpp:1560-1578
enum class IndexType { FLAT, BinaryFLAT, IVFFLAT, IVFPQ, MSTG };      // VICommon.h:178-180 (subset implemented here)
using idx_t = int64_t;

inline int toB200(Metric m) {
    switch (m) {
        case Metric::L2: return B200_METRIC_L2;
        case Metric::IP: return B200_METRIC_IP;
        case Metric::Cosine: return B200_METRIC_COSINE;
        case Metric::Hamming: return B200_METRIC_HAMMING;
        default: return B200_METRIC_JACCARD;
    }
}
inline const char * toB200(IndexType t) {
    switch (t) {
        case IndexType::IVFFLAT: return "IVFFLAT";
        case IndexType::IVFPQ: return "IVFPQ";
        case IndexType::MSTG: return "MSTG";
        default: return "FLAT";
    }
}

// MergeTreeVSManager.cpp:361-366, VIWithDataPart.cpp:645
struct Parameters : std::map<std::string, std::string> {
    void setParam(const std::string & k, const std::string & v) { (*this)[k] = v; }
    std::string toString() const {
        std::string s;
        for (auto & kv : *this) s += (s.empty() ? "" : ", ") + kv.first + "=" + kv.second;
        return s;
    }
};

// uses: MergeTreeVSManager.cpp:1060-1062,:1147,:1274,:1456,:1624; VIUtils.cpp:488; MergeTreeTextSearchManager.cpp:191-194
class DenseBitmap {
    size_t n_;
    std::vector<uint8_t> bits_;
public:
    explicit DenseBitmap(size_t n, bool init = false) : n_(n), bits_((n + 7) / 8, init ? 0xff : 0) {
        if (init && (n & 7)) bits_.back() = static_cast<uint8_t>((1u << (n & 7)) - 1);
    }
    size_t get_size() const { return n_; }
    size_t byte_size() const { return bits_.size(); }
    bool test(size_t i) const { return i < n_ && ((bits_[i >> 3] >> (i & 7)) & 1); }
    bool unsafe_test(size_t i) const { return (bits_[i >> 3] >> (i & 7)) & 1; }
    bool is_member(size_t i) const { return test(i); }
    void set(size_t i) { bits_[i >> 3] |= static_cast<uint8_t>(1u << (i & 7)); }
    void unset(size_t i) { bits_[i >> 3] &= static_cast<uint8_t>(~(1u << (i & 7))); }
    size_t count() const { size_t c = 0; for (auto b : bits_) c += __builtin_popcount(b); return c; }
    bool any() const { return count() > 0; }
    bool all() const { return count() == n_; }
    uint8_t * get_bitmap() { return bits_.data(); }                    // LSB-first bytes, consumed as-is by the kernels
    const uint8_t * get_bitmap() const { return bits_.data(); }
    std::vector<size_t> to_vector() const { std::vector<size_t> v; for (size_t i = 0; i < n_; ++i) if (unsafe_test(i)) v.push_back(i); return v; }
};
using DenseBitmapPtr = std::shared_ptr<DenseBitmap>;
inline DenseBitmapPtr intersectDenseBitmaps(DenseBitmapPtr a, DenseBitmapPtr b) {   // VIWithDataPart.cpp:908, :560
    if (!a) return b;
    if (!b) return a;
    auto r = std::make_shared<DenseBitmap>(a->get_size());
    for (size_t i = 0; i < r->byte_size(); ++i) r->get_bitmap()[i] = a->get_bitmap()[i] & b->get_bitmap()[i];
    return r;
}

template <class T> struct DataSet {                                    // VIWithDataPart.cpp:851, :923, :932
    T * data; int64_t n, dim;
    DataSet(T * d, int64_t n_, int64_t dim_) : data(d), n(n_), dim(dim_) {}
    T * getData() const { return data; }
    int64_t numData() const { return n; }
    int64_t dimension() const { return dim; }
};

// MergeTreeVSManager.cpp:456-461,:565-567; VIWithDataPart.cpp:61-65,:95
class SearchResult {
    int64_t nq_, k_, ncand_;
    std::vector<idx_t> ids_;
    std::vector<float> dis_;
public:
    SearchResult(int64_t nq, int64_t k) : nq_(nq), k_(k), ncand_(k), ids_(nq * k, -1), dis_(nq * k, 0.f) {}
    static std::shared_ptr<SearchResult> createTopKHolder(int64_t nq, int64_t k) { return std::make_shared<SearchResult>(nq, k); }
    idx_t * getResultIndices() { return ids_.data(); }
    float * getResultDistances() { return dis_.data(); }
    int64_t numQueries() const { return nq_; }
    int64_t getNumCandidates() const { return ncand_; }
    void setNumCandidates(int64_t c) { ncand_ = c; }
    int64_t topK() const { return k_; }
};

// createVectorIndex + VectorIndex (VIWithDataPart.cpp:416-430, :131, :853, :878, :926)
template <DataType DT>
class VectorIndex {
    b200_index * h_ = nullptr;
    b200_corpus * bin_ = nullptr;   // binary vectors: BinaryFLAT is a resident corpus
    int64_t n_ = 0;
    size_t dim_;
    bool two_stage_;
public:
    VectorIndex(const std::string & /*name*/, IndexType type, Metric metric, size_t dim, size_t total_vec, const Parameters & params,
                size_t /*max_threads*/ = 0, const std::string & /*cache_prefix*/ = "", std::function<bool()> /*cancel*/ = {})
        : dim_(dim), two_stage_(type == IndexType::MSTG) {
        if constexpr (DT == DataType::BinaryVector) b200Check(b200_corpus_create(toB200(metric), B200_DTYPE_BIN, int(dim), int64_t(total_vec), &bin_));
        else b200Check(b200_index_create(toB200(type), toB200(metric), int(dim), params.toString().c_str(), &h_));
    }
    ~VectorIndex() { if (h_) b200_index_free(h_); if (bin_) b200_corpus_free(bin_); }
    VectorIndex(const VectorIndex &) = delete;
    // build(IndexSourceDataReader*, n_threads, cancel): the reader's chunks are concatenated by the caller (VIPartReader.h:203-246)
    void build(const void * rows, int64_t n) {
        if constexpr (DT == DataType::BinaryVector) b200Check(b200_corpus_append(bin_, rows, n));
        else b200Check(b200_index_build(h_, static_cast<const float *>(rows), n));
        n_ = n;
    }
    template <class T>
    std::shared_ptr<SearchResult> search(std::shared_ptr<DataSet<T>> q, int32_t k, Parameters & params, bool first_stage_only = false,
                                         DenseBitmap * filter = nullptr) {
        auto res = SearchResult::createTopKHolder(q->numData(), k);
        const uint8_t * bits = filter ? filter->get_bitmap() : nullptr;
        if constexpr (DT == DataType::BinaryVector)
            b200Check(b200_corpus_search(bin_, reinterpret_cast<const float *>(q->getData()), q->numData(), k, bits,
                                         res->getResultDistances(), res->getResultIndices()));
        else {
            int64_t ncand = k;
            b200Check(b200_index_search(h_, q->getData(), q->numData(), k, params.toString().c_str(), first_stage_only ? 1 : 0, bits,
                                        res->getResultDistances(), res->getResultIndices(), &ncand));
            res->setNumCandidates(ncand);
        }
        return res;
    }
    // computeTopDistanceSubset(queries, first_stage_result, top_k) (VIWithDataPart.cpp:838-856)
    std::shared_ptr<SearchResult> computeTopDistanceSubset(std::shared_ptr<DataSet<float>> q, std::shared_ptr<SearchResult> first, int32_t top_k) {
        auto res = SearchResult::createTopKHolder(q->numData(), top_k);
        b200Check(b200_index_refine(h_, q->getData(), q->numData(), first->getResultIndices(), first->topK(), top_k,
                                    res->getResultDistances(), res->getResultIndices()));
        return res;
    }
    bool supportTwoStageSearch() const { return two_stage_; }
    bool ready() const { return n_ > 0; }
    size_t numData() const { return size_t(n_); }
};
}  // namespace Search

// ---- faiss entry points used by BruteForceSearch.h:77-105 -------------------------------------
namespace faiss {
struct float_maxheap_array_t { size_t nh, k; int64_t * ids; float * val; };
struct float_minheap_array_t { size_t nh, k; int64_t * ids; float * val; };
inline void knn_L2sqr(const float * x, const float * y, size_t d, size_t nx, size_t ny, float_maxheap_array_t * res, const void * = nullptr) {
    Search::b200Check(b200_flat_knn(B200_METRIC_L2, x, int64_t(nx), y, int64_t(ny), int(d), int(res->k), nullptr, res->val, res->ids));
}
inline void knn_inner_product(const float * x, const float * y, size_t d, size_t nx, size_t ny, float_minheap_array_t * res, const void * = nullptr) {
    Search::b200Check(b200_flat_knn(B200_METRIC_IP, x, int64_t(nx), y, int64_t(ny), int(d), int(res->k), nullptr, res->val, res->ids));
}
inline void hammings_knn_mc(const uint8_t * a, const uint8_t * b, size_t na, size_t nb, size_t k, size_t ncodes, int32_t * distances,
                            int64_t * labels, const void * = nullptr) {
    std::vector<float> d(na * k);
    Search::b200Check(b200_binary_knn(B200_METRIC_HAMMING, a, int64_t(na), b, int64_t(nb), int(ncodes), int(k), nullptr, d.data(), labels));
    for (size_t i = 0; i < na * k; ++i) distances[i] = static_cast<int32_t>(d[i]);   // int32 into the caller's buffer (:99)
}
}  // namespace faiss
inline void jaccard_knn(const uint8_t * a, const uint8_t * b, size_t na, size_t nb, size_t k, size_t ncodes, float * distances,
                        int64_t * labels, const void * = nullptr) {
    Search::b200Check(b200_binary_knn(B200_METRIC_JACCARD, a, int64_t(na), b, int64_t(nb), int(ncodes), int(k), nullptr, distances, labels));
}

// ---- TANTIVY::ffi_* used by Storages/MergeTree/TantivyIndexStore.cpp:654-998 --------------------
namespace TANTIVY {
struct RowIdWithScore { uint64_t row_id; float score; };
struct DocWithFreq { std::string term_str; uint32_t field_id; uint64_t doc_freq; };
struct FieldTokenNums { uint32_t field_id; uint64_t field_total_tokens; };
struct Statistics { std::vector<DocWithFreq> docs_freq; std::vector<FieldTokenNums> total_num_tokens; uint64_t total_num_docs = 0; };
struct FFIError { bool is_error = false; std::string message; };
template <class T> struct FFIResult { T result{}; FFIError error; };

// the Rust side keeps a process-global registry keyed by index directory; so does this adapter
inline std::map<std::string, b200_bm25 *> & registry() { static std::map<std::string, b200_bm25 *> r; return r; }

inline FFIResult<bool> ffi_create_index_with_parameter(const std::string & path, const std::vector<std::string> & columns, const std::string & /*json*/) {
    FFIResult<bool> r; b200_bm25 * h = nullptr;
    if (b200_bm25_create(uint32_t(columns.size()), &h) != B200_OK) { r.error = {true, b200_last_error()}; return r; }
    registry()[path] = h; r.result = true; return r;
}
inline FFIResult<bool> ffi_index_multi_column_docs(const std::string & path, uint64_t row_id, const std::vector<std::string> & /*columns*/,
                                                   const std::vector<std::string> & docs) {
    FFIResult<bool> r; auto * h = registry()[path];
    int rc = b200_bm25_add_doc(h, row_id);
    for (size_t f = 0; rc == B200_OK && f < docs.size(); ++f) rc = b200_bm25_add_text(h, uint32_t(f), docs[f].c_str());
    if (rc != B200_OK) r.error = {true, b200_last_error()}; else r.result = true;
    return r;
}
inline FFIResult<bool> ffi_index_writer_commit(const std::string & path) {
    FFIResult<bool> r; if (b200_bm25_commit(registry()[path]) != B200_OK) r.error = {true, b200_last_error()}; else r.result = true; return r;
}
inline FFIResult<uint64_t> ffi_get_total_num_docs(const std::string & path) {
    FFIResult<uint64_t> r; if (b200_bm25_total_docs(registry()[path], &r.result) != B200_OK) r.error = {true, b200_last_error()}; return r;
}
inline FFIResult<std::vector<FieldTokenNums>> ffi_get_total_num_tokens(const std::string & path, uint32_t n_fields = 1) {
    FFIResult<std::vector<FieldTokenNums>> r;
    for (uint32_t f = 0; f < n_fields; ++f) { uint64_t t = 0; if (b200_bm25_total_tokens(registry()[path], f, &t) != B200_OK) { r.error = {true, b200_last_error()}; break; } r.result.push_back({f, t}); }
    return r;
}
inline std::vector<std::string> queryTerms(const std::string & sentence) {
    std::vector<char> buf(4096); uint32_t n = 0; std::vector<std::string> out;
    if (b200_bm25_query_terms(sentence.c_str(), buf.data(), buf.size(), &n) != B200_OK) return out;
    const char * p = buf.data();
    for (uint32_t i = 0; i < n; ++i) { out.emplace_back(p); p += out.back().size() + 1; }
    return out;
}
inline FFIResult<std::vector<DocWithFreq>> ffi_get_doc_freq(const std::string & path, const std::string & sentence, uint32_t n_fields = 1) {
    FFIResult<std::vector<DocWithFreq>> r;
    for (auto & t : queryTerms(sentence)) for (uint32_t f = 0; f < n_fields; ++f) { uint64_t df = 0; b200_bm25_doc_freq(registry()[path], f, t.c_str(), &df); r.result.push_back({t, f, df}); }
    return r;
}
// ffi_bm25_search(path, sentence, column_names, topk, alive_bitmap, use_filter, enable_nlq, operator_or, statistics)  (:908/:939)
inline FFIResult<std::vector<RowIdWithScore>> ffi_bm25_search(const std::string & path, const std::string & sentence, const std::vector<uint32_t> & fields,
                                                              uint32_t topk, const std::vector<uint8_t> & alive_bitmap, bool use_filter, bool /*enable_nlq*/,
                                                              bool operator_or, const Statistics & st) {
    FFIResult<std::vector<RowIdWithScore>> r;
    std::vector<uint64_t> rows(topk), tok, df; std::vector<float> sc(topk); uint32_t n = 0;
    const auto terms = queryTerms(sentence);
    if (st.total_num_docs) {
        uint32_t maxf = 0; for (auto & t : st.total_num_tokens) maxf = std::max(maxf, t.field_id + 1);
        tok.assign(maxf, 0); for (auto & t : st.total_num_tokens) tok[t.field_id] = t.field_total_tokens;
        df.assign(fields.size() * 64, 0);
        for (size_t fi = 0; fi < fields.size(); ++fi) for (size_t ti = 0; ti < terms.size() && ti < 64; ++ti)
            for (auto & d : st.docs_freq) if (d.field_id == fields[fi] && d.term_str == terms[ti]) df[fi * 64 + ti] = d.doc_freq;
    }
    if (b200_bm25_search(registry()[path], sentence.c_str(), fields.data(), uint32_t(fields.size()), topk, use_filter ? alive_bitmap.data() : nullptr,
                         use_filter, operator_or, st.total_num_docs, tok.data(), df.data(), rows.data(), sc.data(), &n) != B200_OK) { r.error = {true, b200_last_error()}; return r; }
    for (uint32_t i = 0; i < n; ++i) r.result.push_back({rows[i], sc[i]});
    return r;
}
inline FFIResult<bool> ffi_free_index_reader(const std::string & path) { FFIResult<bool> r; auto it = registry().find(path); if (it != registry().end()) { b200_bm25_free(it->second); registry().erase(it); } r.result = true; return r; }
}  // namespace TANTIVY
