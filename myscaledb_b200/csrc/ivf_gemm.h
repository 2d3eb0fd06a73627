// ivf_gemm.h -- work items and parameters of the grouped tensor-core IVF scan (ivf_gemm_sm100.cu), built by ivf.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum { IVF_PRODUCER_TMA = 0, IVF_PRODUCER_PQ = 1, IVF_PRODUCER_SQ8 = 2 };

// queries [q_begin, q_begin + q_count) of the list-sorted pair array x pages [page_begin, page_begin + page_count) of one list
struct IvfGemmItem {
    uint32_t q_begin;     // first row of the gathered query buffer = first pair of this item
    uint32_t q_count;     // 1..128
    uint32_t page_begin;  // absolute index into list_pages
    uint32_t page_count;  // >= 1
    uint32_t row_limit;   // rows of the list from the item's first page on: rows valid in tile j = min(256, row_limit - 256 j)
    uint32_t chunk;       // which row chunk of its list this item is (selects the partial list of every pair)
};

struct IvfGemmParams {
    const IvfGemmItem *items;
    const int *n_items_ptr;           // device scalar written by the planning kernel
    const uint32_t *list_pages;       // page ids, list after list
    const float *row_bias;            // [pool rows] L2: ||y||^2 (PQ: 2<c, r^> + ||r^||^2); null for IP / cosine
    const uint32_t *row_ids;          // [pool rows] row id inside the part
    const uint8_t *alive;             // LSB-first bitmap over row ids, or null
    const uint32_t *pair_part_base;   // [pairs] first partial list of pair i; chunk c of its list writes base + c
    float *part_keys;                 // [parts][k] unsorted
    uint32_t *part_ids;
    float *part_worst;                // [parts] worst kept key when the list is full, else FLT_MAX
    float scale_const;                // -1 IP / cosine, -2 L2
    float *list_keys_gmem;            // scratch when the per-thread lists do not fit in shared memory: [grid][list_cap_for(k)][128]
    uint32_t *list_ids_gmem;
    int d_pad, k;
    int producer;                     // IVF_PRODUCER_*
    // code payloads
    const uint8_t *codes;             // [pool rows][code_bytes]
    const void *codebook_bf16;        // PQ: [m][256][dsub] bf16
    int code_bytes, m, dsub, codebook_bytes;
    // filled in by the launcher
    // Per-query bound shared by all the work items of a launch (nprobe > 1): query_bound[q] is the smallest k-th key
    // (order-preserving u32 encoding, 0xffffffff = none yet) any FULL partial list of query q has published, in absolute key space
    // (item key + pair_const).  Items read it once per page and filter with it: far lists stop inserting once a near list is done.
    uint32_t *query_bound;            // [nq] or null
    const uint32_t *sorted_pair;      // [n_pairs] sorted position -> original pair index (query = pair / nprobe)
    const float *pair_const;          // [n_pairs]
    int nprobe;
    int stages, lists_in_smem, list_cap, codebook_smem_off, coop_smem_off, coop_enabled;
};

// queries_bf16: gathered query rows [n_query_rows][d_pad] (bf16); pool_bf16: page pool [pool_rows][d_pad] (bf16 payload only)
cudaError_t launch_ivf_gemm_topk(const IvfGemmParams &p, const void *queries_bf16, int64_t n_query_rows, const void *pool_bf16,
                                 int64_t pool_rows, int grid, cudaStream_t s, const char **err_detail);

}  // namespace b200
