/*
 * b200_search.h -- C ABI of libb200search.so, the B200-native (sm_100a) engine for
 * MyScaleDB's ANN / BM25 hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes only, no C++/torch types.
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference/src).  INTEGRATION.md shows the C++ shim a MyScaleDB maintainer
 * adds on the ClickHouse side (namespace Search:: / TANTIVY::ffi_* forwarding here).
 *
 * Conventions
 *   - every function returns B200_OK (0) or an error code; b200_last_error() gives the
 *     thread-local message (the reference's libraries throw SearchIndexException /
 *     return {result, error{is_error,message}}; the C++ shim re-throws, VICommon.h:75-104);
 *   - all entry points are re-entrant: ClickHouse calls them from one ThreadPool worker
 *     per part (MergeTreeSelectWithHybridSearchProcessor.cpp:1212-1241);
 *   - float vectors are row-major fp32 [n][d]; binary vectors row-major bytes [n][d/8];
 *   - alive / filter bitmaps are LSB-first bytes, bit=1 => row may be returned
 *     (Search::DenseBitmap::get_bitmap(), MergeTreeTextSearchManager.cpp:191-194);
 *   - results: out_dis[nq*k], out_ids[nq*k], best first; unfilled slots id = -1
 *     (faiss heap convention the callers test with `ids > -1`, MergeTreeVSManager.cpp:469-488);
 *   - tie rule: better score, then smaller row id;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     B200_ERR_NO_DEVICE.
 */
#ifndef B200_SEARCH_H
#define B200_SEARCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_INVALID 1     /* bad argument */
#define B200_ERR_CUDA 2        /* CUDA runtime / driver failure */
#define B200_ERR_UNSUPPORTED 3 /* valid request this build does not implement */
#define B200_ERR_NO_DEVICE 4   /* no sm_100 device visible */
#define B200_ERR_NOMEM 5
#define B200_ERR_NOT_FOUND 6   /* cache miss */

/* Search::Metric (MergeTreeVSManager.cpp:1560-1578) */
#define B200_METRIC_L2 0
#define B200_METRIC_IP 1
#define B200_METRIC_COSINE 2
#define B200_METRIC_HAMMING 3
#define B200_METRIC_JACCARD 4

/* device storage type of a resident corpus */
#define B200_DTYPE_F32 0
#define B200_DTYPE_BF16 1
#define B200_DTYPE_BIN 2

const char *b200_last_error(void);
const char *b200_version(void);
int b200_device_count(int *out_n);
int b200_set_device(int device);

/* ------------------------------------------------------------------------------------
 * Brute force from host buffers.
 * Replaces VectorIndex::tryBruteForceSearch<T> (VectorIndex/Common/BruteForceSearch.h:63-111)
 * as called by VIWithColumnInPart::searchWithoutIndex<T> (VectorIndex/Common/VIWithDataPart.h:342-382):
 * L2 = squared L2 ascending, IP descending, COSINE = normalise both (skip rows with
 * sum(x^2) < FLT_EPSILON) -> IP -> 1 - ip.  x, y are NOT modified.
 * ---------------------------------------------------------------------------------- */
int b200_flat_knn(int metric, const float *x, int64_t nx, const float *y, int64_t ny, int d, int k,
                  const uint8_t *alive_bits /*nullable*/, float *out_dis, int64_t *out_ids);

/* Binary vectors: HAMMING (faiss::hammings_knn_mc) / JACCARD (jaccard_knn),
 * BruteForceSearch.h:94-110.  Distances are returned as float VALUES (the reference
 * writes int32 Hamming distances into the float buffer, :99; the shim re-encodes). */
int b200_binary_knn(int metric, const uint8_t *x, int64_t nx, const uint8_t *y, int64_t ny, int nbytes, int k,
                    const uint8_t *alive_bits /*nullable*/, float *out_dis, int64_t *out_ids);

/* Whole-part brute force with the semantics of MergeTreeVSManager::vectorScanWithoutIndex<T>
 * + searchWrapper<T> (VectorIndex/Storages/MergeTreeVSManager.cpp:960-1679): per-mark
 * blocks merged with "earlier block wins ties", lightweight-delete mask row_exists
 * (1 byte per row, 0 = deleted, :1435-1460), PREWHERE filter bitmap (:1042-1330), and the
 * IP initial value numeric_limits<float>::min() (:1030-1037: IP scores <= FLT_MIN are never
 * returned).  The part column is DMA'd to HBM once and scanned in one pass; block_rows only
 * documents the mark size (results are independent of it under the tie rule).
 * y is fp32 [ny][d] for float metrics or bytes [ny][d/8] for binary metrics (d in bits). */
int b200_part_scan(int metric, const void *x, int64_t nx, const void *y, int64_t ny, int d, int k,
                   int64_t block_rows, const uint8_t *row_exists /*nullable*/, const uint8_t *filter_bits /*nullable*/,
                   float *out_dis, int64_t *out_ids);

/* ------------------------------------------------------------------------------------
 * Device-resident corpus = a FLAT vector index / cached part column in HBM.
 * Replaces Search::VectorIndex<...>(IndexType::FLAT)::{build, search} as reached through
 * VIWithColumnInPart::search (VectorIndex/Common/VIWithDataPart.cpp:858-957) and the
 * VICacheManager residency model (VectorIndex/Cache/VICacheManager.cpp:65-157).
 * ---------------------------------------------------------------------------------- */
typedef struct b200_corpus b200_corpus;

int b200_corpus_create(int metric, int dtype, int d, int64_t capacity_rows, b200_corpus **out);
/* append fp32 (or binary bytes for B200_DTYPE_BIN) rows from host memory; converted to the
 * corpus dtype on device; row norms are computed on device. */
int b200_corpus_append(b200_corpus *c, const void *rows, int64_t n);
/* adopt rows that already live in HBM in the corpus dtype (the rows must be COMPLETE when this is called: the row norms
 * are computed on the corpus' own stream, which is not ordered after the caller's streams), row-major [n][d] with
 * d % 64 == 0 for bf16 (d % 4 == 0 for f32); the memory stays owned by the caller. */
int b200_corpus_adopt_device(b200_corpus *c, const void *device_rows, int64_t n);
int b200_corpus_size(const b200_corpus *c, int64_t *out_rows);
int b200_corpus_free(b200_corpus *c);

/* search with host queries/results (H2D of queries and D2H of results inside the call) */
int b200_corpus_search(b200_corpus *c, const float *queries, int64_t nq, int k, const uint8_t *alive_bits /*nullable*/,
                       float *out_dis, int64_t *out_ids);
/* same, queries and results in device memory, asynchronous on `stream` (a cudaStream_t
 * passed as void*; NULL = the corpus' own stream followed by a synchronise).
 * id_offset is added to every returned id (shard base for multi-GPU merges). */
int b200_corpus_search_device(b200_corpus *c, const float *d_queries, int64_t nq, int k,
                              const uint8_t *d_alive_bits /*nullable*/, int64_t id_offset, float *d_out_dis,
                              int64_t *d_out_ids, void *stream);
/* Force a search path (tests, A/B measurements; production leaves 0):
 *   0 auto | 1 memory-bound scan kernel | 2 tensor cores, exactly the variant auto picks for a batch (operands
 *   streamed through shared memory, CTA pairs from 2 query tiles up, TMA multicast in clusters of 2 / 4 pairs when
 *   the batch has a multiple of 4 / 8 query tiles; fp32 corpora: the 3xTF32 kernel) | 3 single-CTA MMAs <1,1> |
 *   4 CTA pairs, no multicast <2,1> | 5 at most two pairs per cluster <2,2> | 6 up to four pairs per cluster <2,4>
 *   (same as 2, explicit) | 7 queries stationary in TMEM ("TS" form; k <= 30, d <= 768, bf16 corpora). */
int b200_corpus_set_path(b200_corpus *c, int path);
/* which kernel the last search on this corpus launched (so a test can prove it exercised the variant it meant to) */
#define B200_KERNEL_SCAN 1
#define B200_KERNEL_GEMM_BF16 2   /* gemm_topk_kernel<cta_group, pairs_per_cluster> */
#define B200_KERNEL_GEMM_TS 3     /* gemm_topk_ts_kernel */
#define B200_KERNEL_GEMM_TF32X3 4 /* gemm3_topk_kernel */
int b200_corpus_last_variant(b200_corpus *c, int *kernel, int *cta_group, int *pairs_per_cluster, int *grid);
/* CUDA-event timing of the dominant kernel (scan or GEMM) of every search on this corpus,
 * recorded on the launching stream; used by bench.py for the roofline report. */
int b200_corpus_enable_timing(b200_corpus *c, int on);
int b200_corpus_kernel_time(b200_corpus *c, int reset, double *out_total_ms, int64_t *out_launches);
/* number of kernels this library launched on the calling thread since the last reset */
int64_t b200_launch_count(int reset);
/* b200_flat_knn / b200_binary_knn / b200_part_scan keep one grow-only scratch corpus per calling thread (device rows,
 * workspaces, stream), sized for the largest part that thread has scanned; this frees it (it is also freed at thread
 * exit).  A ClickHouse worker calls it when it goes idle. */
int b200_thread_release(void);

/* K-way merge of per-part / per-GPU top-k lists on device.
 * Replaces MergeTreeBaseSearchManager::getTotalTopSearchResultImpl
 * (VectorIndex/Storages/MergeTreeBaseSearchManager.cpp:207-299) for the sharded path:
 * d_dis/d_ids are [n_lists][nq][k] (e.g. the NCCL all-gather buffer); output [nq][k].
 * descending = 1 for IP / BM25. */
int b200_topk_merge_device(const float *d_dis, const int64_t *d_ids, int n_lists, int64_t nq, int k, int descending,
                           float *d_out_dis, int64_t *d_out_ids, void *stream);

/* Same, for a packed all-gather buffer: list l starts at d_dis + l * dis_list_stride (floats) and
 * d_ids + l * ids_list_stride (int64s), so that one NCCL all-gather of
 * {float dis[nq*k]; int64 ids[nq*k]} per rank feeds the merge directly. */
int b200_topk_merge_device_strided(const float *d_dis, const int64_t *d_ids, int n_lists, int64_t dis_list_stride,
                                   int64_t ids_list_stride, int64_t nq, int k, int descending, float *d_out_dis,
                                   int64_t *d_out_ids, void *stream);
/* General form.  Input lists hold k_in entries per query (list l, query q at d_dis + l * dis_list_stride + q * k_in),
 * sorted best first, ids < 0 = unused slot; ids are full 64-bit values (shard base + row, any magnitude).
 * tie_mode 0: equal scores -> smaller id first (this library's contract everywhere else).
 * tie_mode 1: the reference's own order -- std::multimap insertion order of getTotalTopSearchResultImpl
 *   (list 0's entries first, each list in its own order): ascending walks return the earlier-inserted of two equal
 *   scores, the reverse walk used for IP / BM25 returns the later-inserted one (MergeTreeBaseSearchManager.cpp:271).
 *   d_out_list (nullable) receives the list (part) index of every output entry, -1 for unused slots. */
int b200_topk_merge_device_ex(const float *d_dis, const int64_t *d_ids, int n_lists, int64_t dis_list_stride,
                              int64_t ids_list_stride, int64_t nq, int k_in, int k, int descending, int tie_mode,
                              float *d_out_dis, int64_t *d_out_ids, int32_t *d_out_list /*nullable*/, void *stream);

/* ------------------------------------------------------------------------------------
 * Vector indexes.  Replaces Search::createVectorIndex / VectorIndex::{build, search,
 * computeTopDistanceSubset} (VectorIndex/Common/VIWithDataPart.cpp:416-430, :131, :926, :838-856).
 * type (the reference's index type names, VectorIndex/Common/VICommon.h:178-180 and the 2_vector_search tests):
 *   "FLAT"                      exact scan of the resident rows;
 *   "IVFFLAT"                   inverted lists holding bf16 rows, candidates re-ranked exactly against the fp32 rows;
 *   "IVFSQ"                     lists of 8-bit scalar-quantised rows (one byte per dimension);
 *   "IVFPQ"                     lists of m-byte product-quantiser codes of the residual (d / M in {1, 2, 4, 8}, d <= 320);
 *   "MSTG"                      closed source upstream; here the two-stage index of SURVEY 2.5 K6: bf16 lists + exact
 *                               fp32 second stage (supportTwoStageSearch, first_stage_only, computeTopDistanceSubset);
 *   "SCANN", "HNSWFLAT", "HNSWSQ", "HNSWPQ"   accepted and SERVED BY THE INVERTED-FILE ENGINE with the payload their
 *                               name implies (PQ + re-rank, bf16, 8-bit, PQ): there is no graph traversal and no
 *                               anisotropic quantiser in this library; the contract for every ANN type is recall against
 *                               FLAT, not traversal order (parity unpinned for ANN at large N).
 * params: the reference's key=value / JSON parameter string: "ncentroids=1024" (or nlist), "M=32", "nprobe=64",
 * "refine_factor=8" (candidates per returned row handed to the exact second stage; 1 = first-stage distances),
 * "keep_raw=0" (do not keep the fp32 rows: no second stage, half the memory).  Parts smaller than
 * max(2000, 8 * nlist) rows are served by an exact FLAT scan (the reference's fallback_to_flat, test 00029).
 *
 * Build = the reference's reader-driven build (VIPartReader train block / add blocks, VIWithDataPart.cpp:131):
 *   b200_index_reserve(total rows)  -- createVectorIndex's total_vec: sizes the page pool
 *   b200_index_train(sample)        -- coarse k-means (+ PQ codebooks / SQ ranges) on a sample
 *   b200_index_add(chunk) ...       -- any number of chunks, row ids continue from the rows already added
 *   b200_index_finalize()
 * b200_index_build(rows, n) does all four from one host array.  *_device variants take fp32 rows already in HBM.
 * ---------------------------------------------------------------------------------- */
typedef struct b200_index b200_index;
int b200_index_create(const char *type, int metric, int d, const char *params, b200_index **out);
int b200_index_build(b200_index *ix, const float *rows, int64_t n);
int b200_index_reserve(b200_index *ix, int64_t total_rows);
int b200_index_train(b200_index *ix, const float *rows, int64_t n);
int b200_index_train_device(b200_index *ix, const float *d_rows, int64_t n);
int b200_index_add(b200_index *ix, const float *rows, int64_t n);
int b200_index_add_device(b200_index *ix, const float *d_rows, int64_t n);
int b200_index_finalize(b200_index *ix);
int b200_index_info(const b200_index *ix, int64_t *n, int *nlist, int *m, int *uses_ivf);
/* first_stage_only (two-stage types): return the first-stage candidates with first-stage distances;
 * out_num_candidates receives the width the first stage ran with (SearchResult::getNumCandidates).
 * "exact_batch=1" in `params` answers by an exact pass over the fp32 rows instead (recall 1).
 * Tuning / A-B switches, also in `params` (defaults are chosen from the batch shape): "pages_per_chunk=N" (pages of a list one work
 * item streams), "shared_bound=0" (do not share a per-query bound between the work items of a launch), "coarse_path=1|2|3" (centroid
 * probe by the scan kernel / the tensor-core top-k / score tiles + warp select; default 3 for nprobe > 8). */
int b200_index_search(b200_index *ix, const float *queries, int64_t nq, int k, const char *params, int first_stage_only,
                      const uint8_t *alive_bits /*nullable*/, float *out_dis, int64_t *out_ids, int64_t *out_num_candidates);
/* same with device buffers, asynchronous on `stream` (NULL = the index's own stream, synchronised); id_offset is added to
 * every returned id (shard base for multi-GPU merges) */
int b200_index_search_device(b200_index *ix, const float *d_queries, int64_t nq, int k, const char *params, int first_stage_only,
                             const uint8_t *d_alive_bits /*nullable*/, int64_t id_offset, float *d_out_dis, int64_t *d_out_ids,
                             void *stream);
/* roofline inputs of the list scan: CUDA-event time of the grouped scan kernel since the last reset, bytes per list row,
 * and an upper bound of the work items of the last search */
int b200_index_phase_ms(b200_index *ix, double out_ms[5]);   /* last search: coarse | pairs+plan+gather | scan | merge | refine */
int b200_index_list_sizes(const b200_index *ix, uint32_t *out_sizes /*[nlist]*/, int capacity);
int b200_index_enable_timing(b200_index *ix, int on);
int b200_index_last_scan(b200_index *ix, int64_t *rows_streamed, int64_t *payload_row_bytes, int64_t *work_items,
                         double *kernel_ms_total, int64_t *kernel_launches, int reset);
/* computeTopDistanceSubset: exact distances of candidate ids [nq][ncand] (negative = unused) -> top-k */
int b200_index_refine(b200_index *ix, const float *queries, int64_t nq, const int64_t *cand_ids, int64_t ncand, int k,
                      float *out_dis, int64_t *out_ids);
/* VIWithColumnInPart::serialize / load (VIWithDataPart.cpp:451-525, :578-764): one self-describing file
 * ("B2IX" v2; the closed library's .vidx3 payload cannot be reproduced).  load validates every size it derives. */
int b200_index_save(b200_index *ix, const char *path);
int b200_index_load(const char *path, b200_index **out);
/* the same through the host's own streams (Search::IndexDataFileWriter / Reader over ClickHouse disks,
 * VectorIndex/Common/VectorIndexIO.h:33-164): the callbacks return 0 when every byte was written / read */
int b200_index_save_cb(b200_index *ix, int (*write)(void *ctx, const void *data, size_t bytes), void *ctx);
int b200_index_load_cb(int (*read)(void *ctx, void *data, size_t bytes), void *ctx, b200_index **out);
int b200_index_free(b200_index *ix);

/* ------------------------------------------------------------------------------------
 * Multi-GPU: parts / row ranges shard over the GPUs of one box, one communicator rank per GPU (one process per GPU, or
 * one host thread per device).  The reference merges per-part top-k lists on the host
 * (MergeTreeBaseSearchManager::getTotalTopSearchResultImpl, VectorIndex/Storages/MergeTreeBaseSearchManager.cpp:207-299)
 * and sums BM25 statistics over parts (ReadWithHybridSearch::getStatisticForTextSearch,
 * VectorIndex/Processors/ReadWithHybridSearch.cpp:89-209); across GPUs these are ONE ncclAllGather of the packed
 * per-shard top-k + the merge kernel, and ONE ncclAllReduce(sum) of a few counters.  NCCL is resolved with dlopen
 * (nccl_lib_path, $B200_NCCL_LIB, or the libnccl.so.2 already in the process); rank 0 creates the 128-byte unique id and
 * the host distributes it (any side channel: MPI, a file, torch.distributed.broadcast).
 * ---------------------------------------------------------------------------------- */
typedef struct b200_comm b200_comm;
int b200_comm_unique_id(const char *nccl_lib_path /*nullable*/, void *out_id_128_bytes);
int b200_comm_create(const char *nccl_lib_path /*nullable*/, const void *unique_id_128_bytes, int rank, int world, b200_comm **out);
int b200_comm_info(const b200_comm *c, int *rank, int *world);
int b200_comm_free(b200_comm *c);
/* building blocks: device buffers a shard search writes its [nq][k] result to, then all-gather + merge on `stream` */
int b200_comm_local_buffers(b200_comm *c, int64_t nq, int k, float **d_dis, int64_t **d_ids);
int b200_comm_gather_merge(b200_comm *c, int64_t nq, int k, int descending, float *d_out_dis, int64_t *d_out_ids, void *stream);
/* the same for lists produced on the host (per-shard BM25 top-k: scores descending, unused slots score -inf / id -1);
 * uploads, all-gathers, merges and returns the table-wide top-k to the host; synchronous */
int b200_comm_gather_merge_host(b200_comm *c, const float *h_dis, const int64_t *h_ids, int64_t nq, int k, int descending,
                                float *h_out_dis, int64_t *h_out_ids);
/* in-place sum over the ranks of n host-resident uint64 counters (total_docs, total_tokens[field], doc_freq[...]) */
int b200_comm_allreduce_sum_u64(b200_comm *c, uint64_t *host_counters, int64_t n);
/* whole steps.  Every rank passes its own shard and the same queries; every rank receives the global top-k.
 * id_offset = first global row id of this rank's shard.  `stream` must be a real stream.  use_graph != 0 replays the step
 * (query conversion, tensor-core scan, all-gather, merge) as ONE CUDA graph from the second call with the same arguments. */
int b200_sharded_corpus_search(b200_comm *cm, b200_corpus *corpus, const float *d_queries, int64_t nq, int k,
                               const uint8_t *d_alive_bits /*nullable*/, int64_t id_offset, float *d_out_dis, int64_t *d_out_ids,
                               void *stream, int use_graph);
/* host queries in, host results out (H2D, scan, all-gather, merge, D2H, synchronise inside) */
int b200_sharded_corpus_search_host(b200_comm *cm, b200_corpus *corpus, const float *queries, int64_t nq, int d, int k,
                                    int64_t id_offset, float *out_dis, int64_t *out_ids, void *stream, int use_graph);
int b200_sharded_index_search(b200_comm *cm, b200_index *ix, int metric, const float *d_queries, int64_t nq, int k, const char *params,
                              const uint8_t *d_alive_bits /*nullable*/, int64_t id_offset, float *d_out_dis, int64_t *d_out_ids,
                              void *stream);

/* ------------------------------------------------------------------------------------
 * HBM residency cache: the device-side VICacheManager (VectorIndex/Cache/VICacheManager.cpp:65-157, an
 * LRUResourceCache keyed by CacheKey = table path / part / index / column, VICacheObject.h:119-137, weighted by
 * getResourceUsage().memory_usage_bytes).  One process-wide cache; keys are the caller's CacheKey::toString().
 * Entries are pinned while a caller holds them (get / put return pinned); only unpinned entries are evicted,
 * least recently used first; an expired entry that is still pinned is freed at its last release.
 * ---------------------------------------------------------------------------------- */
#define B200_CACHE_CORPUS 0   /* b200_corpus *, freed with b200_corpus_free */
#define B200_CACHE_INDEX 1    /* b200_index *,  freed with b200_index_free */
#define B200_CACHE_BM25 2     /* b200_bm25 *,   freed with b200_bm25_free */
#define B200_CACHE_OPAQUE 3   /* caller-defined object, freed with the deleter passed to b200_cache_put_opaque */
/* VICacheManager::setCacheSize -> updateMaxWeight: shrinking evicts unpinned entries at once */
int b200_cache_set_capacity(uint64_t bytes);
/* VICacheManager::get: B200_OK (pinned handle in *handle, its kind in *kind) or B200_ERR_NOT_FOUND */
int b200_cache_get(const char *key, void **handle, int *kind);
/* VICacheManager::put / load (getOrSet): inserts `handle` weighing `bytes` and pins it; if the key is already
 * resident, *resident is the EXISTING pinned handle and `handle` stays the caller's (free it).  B200_ERR_NOMEM when
 * it does not fit even after evicting every unpinned entry (ownership stays with the caller). */
int b200_cache_put(const char *key, int kind, void *handle, uint64_t bytes, void **resident);
int b200_cache_put_opaque(const char *key, void *handle, uint64_t bytes, void (*deleter)(void *), void **resident);
/* drop one pin taken by get / put.  Pass the handle that get / put returned: if the key was expired while pinned and put
 * again (index rebuilt under the same CacheKey), two generations exist, and only the handle tells whose pin this is.
 * b200_cache_release(key) without a handle releases the live entry first and is only safe without such re-puts. */
int b200_cache_release_handle(const char *key, const void *handle);
int b200_cache_release(const char *key);
/* VICacheManager::forceExpire -> tryRemove: freed now, or at the last release if pinned */
int b200_cache_expire(const char *key);
/* removes every entry whose key starts with `prefix` (a dropped table or part: removeOldParts); returns how many */
int b200_cache_expire_prefix(const char *prefix, int64_t *out_removed);
/* VICacheManager::countItem and the CurrentMetrics counters */
int b200_cache_stats(uint64_t *capacity, uint64_t *used, uint64_t *items, uint64_t *hits, uint64_t *misses,
                     uint64_t *evictions);
/* getResourceUsage().memory_usage_bytes of a resident object (rows + side arrays / lists + codes), for `bytes` above */
int b200_corpus_memory_bytes(const b200_corpus *c, uint64_t *out_bytes);
int b200_index_memory_bytes(const b200_index *ix, uint64_t *out_bytes);

/* ------------------------------------------------------------------------------------
 * Filter bitmaps and decoupled-part row-id maps (VIWithMeta::{row_ids_map, inverted_row_ids_map,
 * inverted_row_sources_map}, VectorIndex/Cache/VICacheObject.h:40-117).
 * ---------------------------------------------------------------------------------- */
/* Device-resident filters (the filtered-search fast path): build the DenseBitmap in HBM from the surviving _part_offset values
 * of the PREWHERE pipeline (getFilterFromPipeline, ...SelectWithHybridSearchProcessor.cpp:906-934) or from the _row_exists
 * bytes of a lightweight delete (MergeTreeVSManager.cpp:1435-1460), intersect there, and pass the result as d_alive_bits to the
 * *_search_device / b200_sharded_* calls.  Buffers: (nbits + 7) / 8 bytes rounded up to a multiple of 4, 4-byte aligned. */
int b200_bitmap_from_offsets_device(const uint64_t *d_offsets, int64_t n, int64_t nbits, uint8_t *d_out_bits, void *stream);
int b200_bitmap_from_row_exists_device(const uint8_t *d_row_exists, int64_t n, uint8_t *d_out_bits, void *stream);
int b200_bitmap_and_device(const uint8_t *d_a, const uint8_t *d_b, int64_t nbits, uint8_t *d_out, void *stream);
/* Search::intersectDenseBitmaps (VIWithDataPart.cpp:908): out = a & b */
int b200_bitmap_and(const uint8_t *a, const uint8_t *b, int64_t nbits, uint8_t *out);
/* getRealBitmap (VectorIndex/Utils/VIUtils.cpp:479-497): filter over the merged part -> bitmap over this old part */
int b200_real_bitmap(const uint8_t *filter_bits, int64_t n_new_rows, const uint64_t *inverted_row_ids_map,
                     const uint8_t *inverted_row_sources_map, uint32_t own_id, int64_t total_vec, uint8_t *out_bits);
/* VIWithColumnInPart::transferToNewRowIds (VIWithDataPart.cpp:56-68), in place */
int b200_remap_labels(const uint64_t *row_ids_map, int64_t map_len, int64_t *labels, int64_t n);
/* VIWithColumnInPart::TransferToOldRowIds (VIWithDataPart.cpp:69-126): order-preserving filter + map */
int b200_transfer_to_old_row_ids(const int64_t *new_ids, const float *new_dis, int64_t num_candidates,
                                 const uint64_t *inverted_row_ids_map, const uint8_t *inverted_row_sources_map, int64_t map_len,
                                 uint32_t own_id, int64_t *out_ids, float *out_dis, int64_t *out_n);

/* ------------------------------------------------------------------------------------
 * BM25 full-text search (Boundary B).  Replaces the TANTIVY::ffi_* calls of TantivyIndexStore
 * (Storages/MergeTree/TantivyIndexStore.cpp): ffi_index_multi_column_docs :742,
 * ffi_index_writer_commit :824, ffi_bm25_search :908/:939, ffi_get_doc_freq :962,
 * ffi_get_total_num_docs :974, ffi_get_total_num_tokens :986.  One index per part, resident
 * in HBM; tantivy-0.21 BM25 and "default" tokenizer semantics; results are score-descending,
 * ties to the smaller doc (tantivy TopDocs).
 * ---------------------------------------------------------------------------------- */
typedef struct b200_bm25 b200_bm25;
int b200_bm25_create(uint32_t n_fields, b200_bm25 **out);
int b200_bm25_free(b200_bm25 *ix);
/* one row: add_doc(row_id) then add_text(field, text) per value (several per field for Array(String)) */
int b200_bm25_add_doc(b200_bm25 *ix, uint64_t row_id);
int b200_bm25_add_text(b200_bm25 *ix, uint32_t field, const char *text);
int b200_bm25_commit(b200_bm25 *ix);
/* ffi_load_index_reader / the index files of a part (TantivyIndexStore.cpp:646-686): one self-describing file ("B2TX" v1;
 * tantivy's segment files cannot be reproduced without the crate); load validates, uploads to HBM and returns a committed index */
int b200_bm25_save(b200_bm25 *ix, const char *path);
int b200_bm25_load(const char *path, b200_bm25 **out);
int b200_bm25_total_docs(const b200_bm25 *ix, uint64_t *out);
int b200_bm25_total_tokens(const b200_bm25 *ix, uint32_t field, uint64_t *out);
int b200_bm25_doc_freq(const b200_bm25 *ix, uint32_t field, const char *term, uint64_t *out);
/* distinct lowercase terms of a sentence, NUL-separated, in tokenisation order */
int b200_bm25_query_terms(const char *sentence, char *buf, size_t buf_len, uint32_t *out_n);
/* stat_*: table-wide statistics (ReadWithHybridSearch::getStatisticForTextSearch,
 * VectorIndex/Processors/ReadWithHybridSearch.cpp:89-209), used iff stat_total_docs > 0:
 * stat_total_tokens[n_fields of the index], stat_doc_freq[q][fq * 64 + term_index]. */
int b200_bm25_search(b200_bm25 *ix, const char *sentence, const uint32_t *fields, uint32_t n_fields_q, uint32_t topk,
                     const uint8_t *alive_bits /*over row ids*/, int use_filter, int operator_or, uint64_t stat_total_docs,
                     const uint64_t *stat_total_tokens, const uint64_t *stat_doc_freq, uint64_t *out_rows,
                     float *out_scores, uint32_t *out_n);
/* roofline inputs of the last batch on this index: scoring-kernel ms (CUDA events), wall ms inside the C call, postings walked */
int b200_bm25_last_timing(b200_bm25 *ix, double *kernel_ms, double *call_ms, uint64_t *postings);
int b200_bm25_search_batch(b200_bm25 *ix, const char *const *sentences, int64_t nq, const uint32_t *fields,
                           uint32_t n_fields_q, uint32_t topk, const uint8_t *alive_bits, int use_filter, int operator_or,
                           uint64_t stat_total_docs, const uint64_t *stat_total_tokens, const uint64_t *stat_doc_freq,
                           uint64_t *out_rows /*[nq][topk]*/, float *out_scores, uint32_t *out_counts /*[nq]*/);

/* ------------------------------------------------------------------------------------
 * Hybrid-search fusion, batched.  Replaces RankFusion / RelativeScoreFusion
 * (VectorIndex/Utils/HybridSearchUtils.cpp:164-274) + the final ordering of
 * MergeTreeHybridSearchManager::hybridSearch (VectorIndex/Storages/MergeTreeHybridSearchManager.cpp:108-171).
 * fusion_type 0 = RSF (fusion_weight, vector_scan_direction 1 asc / -1 desc), 1 = RRF (fusion_k).
 * Candidate lists are [nq][stride] arrays (already globally ordered) with per-query counts;
 * outputs [nq][top_k], fused score descending, ties in ascending (shard, part, label) order.
 * ---------------------------------------------------------------------------------- */
int b200_hybrid_fusion_batch(int fusion_type, int64_t nq, const uint32_t *vec_shard, const uint64_t *vec_part,
                             const uint64_t *vec_label, const float *vec_score, const uint32_t *vec_count, int64_t vec_stride,
                             const uint32_t *txt_shard, const uint64_t *txt_part, const uint64_t *txt_label,
                             const float *txt_score, const uint32_t *txt_count, int64_t txt_stride, float fusion_weight,
                             uint64_t fusion_k, int vector_scan_direction, uint32_t top_k, uint32_t *out_shard,
                             uint64_t *out_part, uint64_t *out_label, float *out_score, uint32_t *out_count);

#ifdef __cplusplus
}
#endif
#endif /* B200_SEARCH_H */
