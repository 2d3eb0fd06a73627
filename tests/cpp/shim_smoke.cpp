// Compiles against shim/b200_search_shim.hpp exactly the way src/VectorIndex would and exercises the
// reference-side call shapes (BruteForceSearch.h:77-105, VIWithDataPart.cpp:926/:853,
// TantivyIndexStore.cpp:742-948) on the GPU.  Prints "SHIM OK" on success.
#include <b200_search_shim.hpp>
#include <cmath>
#include <cstdio>

int main() {
    try {
        // tryBruteForceSearch: 100 rows [n,n,n], query [0.1]*3 -> golden 00001: ids 0..9, 0.030000001 ...
        std::vector<float> y(300), q = {0.1f, 0.1f, 0.1f};
        for (int n = 0; n < 100; ++n) y[3 * n] = y[3 * n + 1] = y[3 * n + 2] = float(n);
        std::vector<int64_t> ids(10); std::vector<float> dis(10);
        faiss::float_maxheap_array_t res = {1, 10, ids.data(), dis.data()};
        faiss::knn_L2sqr(q.data(), y.data(), 3, 1, 100, &res, nullptr);
        for (int i = 0; i < 10; ++i) if (ids[i] != i) { std::printf("bad id %d -> %ld\n", i, long(ids[i])); return 1; }
        if (std::fabs(dis[0] - 0.030000001f) > 1e-7f || std::fabs(dis[9] - 237.62997f) > 1e-3f) { std::printf("bad distances\n"); return 1; }
        // FLAT index through Search::VectorIndex with a filter bitmap
        Search::Parameters params;
        Search::VectorIndex<Search::DataType::FloatVector> index("v1", Search::IndexType::FLAT, Search::Metric::L2, 3, 100, params);
        index.build(y.data(), 100);
        Search::DenseBitmap filter(100, true);
        filter.unset(0);
        auto qs = std::make_shared<Search::DataSet<float>>(q.data(), 1, 3);
        auto r = index.search(qs, 5, params, false, &filter);
        if (r->getResultIndices()[0] != 1 || r->getResultIndices()[4] != 5) { std::printf("filtered search wrong\n"); return 1; }
        auto r2 = index.computeTopDistanceSubset(qs, r, 3);
        if (r2->getResultIndices()[0] != 1 || std::fabs(r2->getResultDistances()[0] - 2.4299998f) > 1e-5f) { std::printf("refine wrong\n"); return 1; }
        // BM25 through the TANTIVY:: names
        const char * docs[3] = {"Ancient empires rise and fall", "Artistic expressions reflect heritages", "Ancient philosophies provide wisdom"};
        TANTIVY::ffi_create_index_with_parameter("/p", {"doc"}, "{}");
        for (uint64_t i = 0; i < 3; ++i) TANTIVY::ffi_index_multi_column_docs("/p", i, {"doc"}, {docs[i]});
        TANTIVY::ffi_index_writer_commit("/p");
        auto hits = TANTIVY::ffi_bm25_search("/p", "Ancient", {0}, 5, {}, false, false, true, TANTIVY::Statistics{});
        if (hits.error.is_error || hits.result.size() != 2) { std::printf("bm25 wrong: %s\n", hits.error.message.c_str()); return 1; }
        if (TANTIVY::ffi_get_total_num_docs("/p").result != 3) return 1;
        TANTIVY::ffi_free_index_reader("/p");
        std::printf("SHIM OK\n");
        return 0;
    } catch (const SearchIndexException & e) {
        std::printf("SearchIndexException %d: %s\n", e.getCode(), e.what());
        return 2;
    }
}
