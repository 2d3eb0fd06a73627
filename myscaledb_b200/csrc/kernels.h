// kernels.h -- parameter blocks and host launchers shared between the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

struct ScanParams {
    const void *corpus;      // [n][row_bytes]
    const float *queries;    // device fp32 [nq][d_pad] (cosine: pre-normalised)
    const float *row_scale;  // cosine: -1/||row|| (-1 for rows with sum sq < FLT_EPSILON), else null (= -1)
    const uint8_t *alive;    // LSB-first bitmap or null
    float *part_keys;        // [nq][gridDim.x][k]
    uint32_t *part_ids;
    int64_t n, nq;
    int64_t row_bytes;
    int d_pad, k, group;
    int l2, bf16;
    // fused single-launch form (small batches on a resident corpus: the call is latency-bound): raw queries [nq][q_dim]
    // are padded (and for cosine normalised) while they are staged, the LAST block of a query tile to finish merges every
    // block's partial list and writes the final result (mapped pinned host memory) -- one launch, no pad / merge kernels
    int fused;               // 0 = plain scan (queries pre-padded, partial lists only)
    int q_dim;               // fused: row length of `queries`
    int cosine;              // fused: normalise the staged queries (skip sum sq < FLT_EPSILON), output 1 + key
    int out_mode, ip_min_quirk;
    int64_t id_offset;
    unsigned int *tickets;   // fused: [gridDim.y] zeroed counters (the last block resets its own)
    float *out_dis;          // fused: [nq][k]
    int64_t *out_ids;
    volatile unsigned int *done_flag;  // fused, nullable: set to done_value once every query tile has been written
    unsigned int done_value;
    unsigned int *tiles_done;          // fused: one zeroed counter (reset by the last tile)
    int q_inline;                      // fused: the queries travel in the kernel parameters (nq * q_dim <= 256 floats)
    unsigned long long *debug_ts;      // fused, debugging (B200_FUSED_DEBUG_TS=1): %globaltimer stamps of block 0's start and the tail's phases
    int stage_cap;                     // fused: candidates (gridDim.x * k) the last block may stage in shared memory, 0 = none (set by the launcher)
    float qinline[256];
};

struct BinaryScanParams {
    const uint8_t *corpus;   // [n][nbytes]
    const uint8_t *queries;  // [nq][nbytes]
    const uint8_t *alive;
    float *part_keys;
    uint32_t *part_ids;
    int64_t n, nq;
    int nbytes, k, jaccard;
};

enum { kOutKey = 0, kOutNeg = 1, kOutOnePlus = 2, kOutAddQ = 3, kOutCosQ = 4 };

struct MergeParams {
    const void *in_keys;     // float
    const void *in_ids;      // uint32 (internal partials) or int64 (external lists)
    int64_t list_stride, q_stride;
    int64_t id_list_stride;  // 0 = same as list_stride (element units of the id type)
    int n_lists, k_in, k;
    int64_t nq;
    int descending;          // external only
    int tie_mode;            // external only: 0 = smaller 64-bit id first, 1 = the reference's multimap insertion order
    int32_t *out_list;       // external, tie_mode 1: source list of every output entry (nullable)
    int out_mode;            // kOut*
    int ip_min_quirk;        // part-scan IP: drop scores <= FLT_MIN
    const float *q_add;      // kOutAddQ: ||q||^2 per query; kOutCosQ: -(1/||q||) per query
    int64_t id_offset;
    float *out_dis;
    int64_t *out_ids;
};

size_t scan_smem_bytes(int qt, int d_pad, int k);
cudaError_t launch_flat_scan(const ScanParams &p, int qt, int blocks_x, cudaStream_t s);
cudaError_t launch_binary_scan(const BinaryScanParams &p, int blocks_x, cudaStream_t s);
cudaError_t launch_topk_merge(const MergeParams &p, bool external, cudaStream_t s);
// exact squared-L2 of the k winners of every query (direct differences, fp32) + re-order by (distance, id); k <= 1024
cudaError_t launch_rescore_l2(const void *corpus, int bf16, int64_t row_bytes, int d_pad, const float *queries, int64_t nq, int64_t id_offset,
                              int k, float *dis, int64_t *ids, cudaStream_t s);

// ---- tcgen05 bf16 GEMM + fused top-k (ip_gemm_sm100.cu) ---------------------------
struct GemmTopkParams {
    const void *corpus_bf16;   // [n][d_pad] bf16, d_pad % 64 == 0
    const void *queries_bf16;  // [nq_pad][d_pad] bf16, nq_pad % 128 == 0 (3xTF32 kernel: fp32 hi plane; corpus_bf16 = fp32 rows)
    const void *queries_lo;    // 3xTF32 kernel only: fp32 lo plane [nq_pad][d_pad]
    const float *row_scale;    // per corpus row multiplier a[j] or null (= scale_const)
    float scale_const;         // -1 for IP, -2 for L2
    const float *row_bias;     // per corpus row addend b[j] or null (=0)
    const uint8_t *alive;      // LSB-first bitmap or null
    float *part_keys;          // [gridDim.x][128][k]
    uint32_t *part_ids;
    float *list_keys_gmem;     // scratch for lists that do not fit in shared memory: [gridDim.x][list_cap_for(k)][128]
    uint32_t *list_ids_gmem;
    int64_t n;
    int nq_pad, d_pad, k;
    int nq_valid;              // queries actually in the batch (rows past it are padding)
    int q_tiles;               // nq_pad / 128
    int cta_group;             // 1: one CTA per MMA; 2: CTA pairs (cluster of 2), q_tiles must be even
    int pairs_per_cluster;     // 1, or 2: two CTA pairs share (TMA-multicast) every corpus tile; q_tiles % 4 == 0
    int *progress;             // [grid / q_tiles][q_tiles] zeroed pacing counters, or null
    int stages;                // smem ring depth (filled in by the launcher)
    int kps;                   // k-blocks per full/empty barrier stage (1 or 2; filled in by the launcher)
    int lists_in_smem;         // per-thread top-k lists in shared memory (else global scratch); set by the launcher
    int list_cap;              // slots per list: k (rescan mode) or list_cap_append(k) (append mode); set by the launcher
    int debug;                 // experiments only (B200_GEMM_DEBUG): 1 no epilogue, 2 TMEM loads only, 4 no TMA
    int sync_slack;            // tiles a CTA may run ahead of the slowest sharer of its corpus tiles
};
// per-thread top-k lists live in shared memory up to this k (the smem ring gets shallower: 6 stages up to k = 14,
// 5 up to 46, 4 up to 78, 3 up to 110, 2 up to 128 for CTA pairs); larger k uses global scratch
constexpr int kGemmSmemK = 128;
// per-thread top-k lists (gemm_common.cuh, ThreadTopK): k slots and a rescan per insert (the default), or an append buffer
// of 2k + 32 slots compacted in lock-step (B200_LIST_APPEND_MIN_K=<k>: lists of at least that k use it; measured in
// profiles/r02_list_modes.md -- it only pays when the doubled buffer still fits in shared memory, which it does not at k = 100)
__host__ __device__ inline int list_cap_append(int k) { return 2 * k + 32; }
// tournament form: k entries + one (key, id) slot per group of 8 (k <= 64) or 16 entries holding the group's worst
// (B200_LIST_TOURN_MIN_K); 16 keeps k = 100 at 107 slots, which still leaves the flat kernel a 3-stage operand ring
__host__ __device__ inline int list_tourn_group(int k) { return k <= 64 ? 8 : 16; }
__host__ __device__ inline int list_cap_tourn(int k) { return k + (k + list_tourn_group(k) - 1) / list_tourn_group(k); }
int list_cap_for(int k);   // capi.cu: k, or list_cap_append(k) when the environment asks for the append form
int gemm_topk_grid(int q_tiles, int64_t n, int num_sms);
// returns cudaSuccess or an error; tensor maps are encoded inside
cudaError_t launch_gemm_topk(const GemmTopkParams &p, int grid, cudaStream_t s, const char **err_detail);
int gemm_topk_max_clusters(int cta_group, int pairs_per_cluster, int k);
// queries-stationary-in-TMEM form (ip_gemm_ts_sm100.cu): CTA pairs, d_pad <= 768, even q_tiles
bool gemm_topk_ts_supported(int d_pad, int q_tiles);
int gemm_topk_ts_tile_rows(int d_pad);
cudaError_t launch_gemm_topk_ts(const GemmTopkParams &p, int grid, cudaStream_t s, const char **err_detail);

// ---- fp32 rows on the tensor cores with fp32-level accuracy (ip_gemm_tf32x3_sm100.cu): CTA pairs, even q_tiles
cudaError_t launch_split_tf32(const float *src, int64_t n_src, int d_pad, float *hi, float *lo, int64_t n_pad, cudaStream_t s);
cudaError_t launch_gemm3_topk(const GemmTopkParams &p, int grid, cudaStream_t s, const char **err_detail);

// ---- host ingest (ingest.cu): pageable host memory -> device through a multi-threaded pinned ring; returns a B200_* code
int staged_h2d(void *dst, const void *src, size_t bytes, int device, cudaStream_t s);

// ---- elementwise prep kernels (prep.cu) --------------------------------------------
cudaError_t launch_f32_to_bf16_rows(const float *src, int d, void *dst, int d_pad, int64_t n, cudaStream_t s);
cudaError_t launch_pad_rows_f32(const float *src, int d, float *dst, int d_pad, int64_t n, cudaStream_t s);
// per-row sum of squares (fp32 accumulate) of a [n][d_pad] corpus; mode 0: out = ss, mode 1: out = -(ss<eps ? 1 : 1/sqrt(ss))
cudaError_t launch_row_norms(const void *rows, int bf16, int d_pad, int64_t n, int mode, float *out, cudaStream_t s);
// normalise fp32 query rows in place (cosine), skipping rows with ss < FLT_EPSILON
cudaError_t launch_normalize_rows_f32(float *rows, int d_pad, int64_t n, cudaStream_t s);

}  // namespace b200
