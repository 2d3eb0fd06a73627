// ivf.cu -- K5/K6: the inverted-file index family, its streamed GPU build, and the exact second stage.
//
// Replaces Search::createVectorIndex / Search::VectorIndex<...>::{build, search, computeTopDistanceSubset, serialize, load}
// for the index types the reference reaches through VIWithColumnInPart (reference:
// src/VectorIndex/Common/VIWithDataPart.cpp:416-430 create, :131 build, :858-957 search, :838-856 second stage,
// :451-525 / :578-764 serialize / load) and MergeTreeVSManager::executeSecondStageVectorScan
// (src/VectorIndex/Storages/MergeTreeVSManager.cpp:510-630).  The reference's implementations live in the un-vendored
// search-index library (Faiss IVF*, hnswlib, ScaNN) and the closed-source MSTG; the algorithms here are the published
// IVF ones, the layout and the execution model are ours:
//   * coarse quantiser: nlist fp32 centroids, k-means on device (assignment = exact top-1 search of the centroid table
//     on the tensor cores, 3xTF32);
//   * PAGED inverted lists: a page = 256 consecutive rows of one list in a pre-reserved pool (exactly one tcgen05 tile);
//     `add` appends chunk after chunk (assign -> sort by list -> allocate pages by prefix sums -> scatter), nothing is
//     ever compacted or moved, so 100 M x 768 rows stream through a few GB of scratch (VIPartReader's chunked build);
//   * payload of a row: bf16 vector (IVFFLAT / MSTG-class first stage), one byte per dimension (IVFSQ) or m PQ codes of
//     the residual (IVFPQ / SCANN-class); + its row id and, for L2, the norm term of the expanded distance;
//   * search: coarse top-nprobe, then ALL (query, list) pairs of the batch are radix-sorted by list and cut into work
//     items (<= 128 queries x a run of pages) for the grouped tensor-core scan of ivf_gemm_sm100.cu, so a list is read
//     from HBM once per batch however many queries probe it; per-pair partial lists are merged per query;
//   * optional fp32 rows in id order (`keep_raw`) for the exact second stage (refine_kernel, warp per candidate).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"
#include "ivf_gemm.h"
#include "kernels.h"

namespace b200 {

constexpr int kPageRows = 256;

// ------------------------------------------------------------------------------------
// k-means assignment: for every point the nearest centroid under L2 (argmin ||c||^2 - 2 x.c).
// 64 points x 64 centroids per tile, K chunks of 16, 4x4 register micro-tiles.
// ------------------------------------------------------------------------------------
constexpr int KA_T = 64, KA_K = 16;

__global__ void __launch_bounds__(256) kmeans_assign_kernel(const float *__restrict__ x, int64_t n, int64_t x_stride, int d,
                                                            const float *__restrict__ c, int nc, const float *__restrict__ cnorm,
                                                            uint32_t *__restrict__ out_idx, float *__restrict__ out_dist) {
    __shared__ float xs[KA_K][KA_T + 4];
    __shared__ float cs[KA_K][KA_T + 4];
    __shared__ float best_d[KA_T][17];
    __shared__ uint32_t best_i[KA_T][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each 4 points x 4 centroids
    const int64_t p0 = (int64_t)blockIdx.x * KA_T;
    float run_d[4];
    uint32_t run_i[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
        run_d[a] = FLT_MAX;
        run_i[a] = 0;
    }
    for (int c0 = 0; c0 < nc; c0 += KA_T) {
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
        for (int k0 = 0; k0 < d; k0 += KA_K) {
            for (int i = threadIdx.x; i < KA_T * KA_K; i += 256) {
                const int r = i / KA_K, kk = i % KA_K;
                const int64_t pr = p0 + r;
                xs[kk][r] = (pr < n && k0 + kk < d) ? x[pr * x_stride + k0 + kk] : 0.f;
                const int cr = c0 + r;
                cs[kk][r] = (cr < nc && k0 + kk < d) ? c[(int64_t)cr * d + k0 + kk] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KA_K; kk++) {
                float xv[4], cv[4];
#pragma unroll
                for (int a = 0; a < 4; a++) xv[a] = xs[kk][ty * 4 + a];
#pragma unroll
                for (int b = 0; b < 4; b++) cv[b] = cs[kk][tx * 4 + b];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[a][b] = fmaf(xv[a], cv[b], acc[a][b]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int cr = c0 + tx * 4 + b;
                if (cr < nc) {
                    const float dist = cnorm[cr] - 2.f * acc[a][b];
                    if (dist < run_d[a]) {  // ascending centroid order inside a thread: ties keep the smaller id
                        run_d[a] = dist;
                        run_i[a] = (uint32_t)cr;
                    }
                }
            }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        best_d[ty * 4 + a][tx] = run_d[a];
        best_i[ty * 4 + a][tx] = run_i[a];
    }
    __syncthreads();
    if (threadIdx.x < KA_T) {
        const int r = threadIdx.x;
        float bd = FLT_MAX;
        uint32_t bi = 0;
        for (int t = 0; t < 16; t++) {
            const float dv = best_d[r][t];
            const uint32_t iv = best_i[r][t];
            if (dv < bd || (dv == bd && iv < bi)) {
                bd = dv;
                bi = iv;
            }
        }
        if (p0 + r < n) {
            out_idx[p0 + r] = bi;
            if (out_dist) out_dist[p0 + r] = bd;
        }
    }
}

__global__ void kmeans_accumulate_kernel(const float *x, int64_t n, int64_t x_stride, int d, const uint32_t *idx, float *sums,
                                         uint32_t *counts) {
    const int64_t total = n * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int j = (int)(i - r * d);
        atomicAdd(&sums[(int64_t)idx[r] * d + j], x[r * x_stride + j]);
        if (j == 0) atomicAdd(&counts[idx[r]], 1u);
    }
}

__global__ void kmeans_update_kernel(float *c, const float *sums, const uint32_t *counts, int nc, int d) {
    const int64_t total = (int64_t)nc * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t cnt = counts[i / d];
        if (cnt) c[i] = sums[i] / (float)cnt;  // empty cluster: keep the previous centroid
    }
}

__global__ void rows_sqnorm_kernel(const float *c, int nc, int d, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    float s = 0.f;
    for (int j = 0; j < d; j++) s = fmaf(c[(int64_t)i * d + j], c[(int64_t)i * d + j], s);
    out[i] = s;
}

__global__ void gather_rows_kernel(const float *x, int64_t x_stride, const int64_t *pick, int64_t np, int d, float *out) {
    const int64_t total = np * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        out[i] = x[pick[r] * x_stride + (i - r * d)];
    }
}

// residual sub-vectors of one sub-quantiser: out[r][t] = x[r][j*dsub + t] - centroid[list[r]][j*dsub + t]
__global__ void residual_sub_kernel(const float *x, int64_t n, int64_t x_stride, const float *cent, const uint32_t *list, int d,
                                    int j, int dsub, float *out) {
    const int64_t total = n * dsub;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / dsub;
        const int t = (int)(i - r * dsub);
        out[i] = x[r * x_stride + j * dsub + t] - cent[(int64_t)list[r] * d + j * dsub + t];
    }
}

// pairs[2 i] = an empty cluster, pairs[2 i + 1] = a large one: the empty one becomes a nudged copy of the large one
__global__ void kmeans_split_kernel(float *c, const int *pairs, int d) {
    const int dst = pairs[2 * blockIdx.x], src = pairs[2 * blockIdx.x + 1];
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        const float v = c[(size_t)src * d + j];
        const float eps = (j & 1) ? 1.f / 1024.f : -1.f / 1024.f;
        c[(size_t)dst * d + j] = v * (1.f + eps) + eps * 1e-3f;
        c[(size_t)src * d + j] = v * (1.f - eps) - eps * 1e-3f;
    }
}

__global__ void iota_kernel(uint32_t *v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

static inline int gridsz(int64_t work, int threads = 256) {
    int64_t b = ceil_div(work, threads);
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, 148 * 32));
}

// ------------------------------------------------------------------------------------
// build: chunk -> paged lists
// ------------------------------------------------------------------------------------
__global__ void assign_to_u32_kernel(const int64_t *ids, int64_t n, uint32_t *out, uint32_t *cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t l = ids[i] < 0 ? 0u : (uint32_t)ids[i];
    out[i] = l;
    atomicAdd(&cnt[l], 1u);
}

// Exclusive scan of per-thread partial sums across one 1024-thread CTA.
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t *total) {
    typedef cub::BlockScan<uint32_t, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    uint32_t excl, tot;
    Scan(tmp).ExclusiveSum(v, excl, tot);
    __syncthreads();
    if (total) *total = tot;
    return excl;
}

// One CTA plans a chunk's page allocation: per list the rows arriving (cnt), where its segment starts in the sorted
// chunk, how many NEW pages it needs and their ids (prefix sum on top of *pages_used), and stamps the new pages.
struct AddPlan {
    const uint32_t *cnt;        // [nlist] rows of this chunk per list
    uint32_t *seg_start;        // [nlist] out: first index of the list's segment in the list-sorted chunk
    uint32_t *new_base;         // [nlist] out: first new page id of the list
    uint32_t *first_new_seq;    // [nlist] out: sequence number (page index inside the list) of its first new page
    const uint32_t *list_len;   // [nlist] rows already in the list
    uint32_t *page_owner, *page_seq;
    uint32_t *pages_used;       // in/out (device scalar)
    uint32_t pool_pages;
    int nlist;
    int *overflow;              // out: set when the pool is exhausted
};

__global__ void __launch_bounds__(1024) add_plan_kernel(const AddPlan p) {
    const int per = (p.nlist + 1023) / 1024;
    const int l0 = threadIdx.x * per, l1 = min(p.nlist, l0 + per);
    uint32_t rows = 0, pages = 0;
    for (int l = l0; l < l1; l++) {
        const uint32_t len = p.list_len[l], c = p.cnt[l];
        rows += c;
        pages += (len + c + kPageRows - 1) / kPageRows - (len + kPageRows - 1) / kPageRows;
    }
    uint32_t tot_pages = 0;
    uint32_t row_off = block_exclusive_scan_1024(rows, nullptr);
    uint32_t page_off = block_exclusive_scan_1024(pages, &tot_pages);
    const uint32_t used = *p.pages_used;
    __syncthreads();
    const bool fits = used + tot_pages <= p.pool_pages;
    for (int l = l0; l < l1; l++) {
        const uint32_t len = p.list_len[l], c = p.cnt[l];
        const uint32_t first = (len + kPageRows - 1) / kPageRows;
        const uint32_t need = (len + c + kPageRows - 1) / kPageRows - first;
        p.seg_start[l] = row_off;
        p.new_base[l] = used + page_off;
        p.first_new_seq[l] = first;
        if (fits)
            for (uint32_t t = 0; t < need; t++) {
                p.page_owner[used + page_off + t] = (uint32_t)l;
                p.page_seq[used + page_off + t] = first + t;
            }
        row_off += c;
        page_off += need;
    }
    if (threadIdx.x == 0) {
        if (fits) *p.pages_used = used + tot_pages;
        else *p.overflow = 1;
    }
}

// One warp per row of the list-sorted chunk: convert / encode the row into its pool slot, record its id and norm term.
struct ScatterParams {
    const float *rows;          // chunk rows fp32 [n][stride] (cosine: already unit length)
    int64_t stride;
    const uint32_t *sorted_list;  // [n] list of sorted element i
    const uint32_t *sorted_row;   // [n] chunk row of sorted element i
    const uint32_t *seg_start, *new_base, *first_new_seq, *list_len, *tail_page;
    uint32_t id_base;
    int64_t n;
    int d, d_pad64;
    int l2;
    // bf16 payload
    __nv_bfloat16 *pool;
    // SQ8 payload
    const float *sq_lo, *sq_inv_step, *sq_step;   // per dimension
    // PQ payload
    const float *centroids;     // [nlist][d]
    const float *pq;            // [m][256][dsub] fp32 (nearest-centroid search)
    const __nv_bfloat16 *pq_bf16;  // values the scan kernel will see
    int m, dsub;
    uint8_t *codes;
    int code_bytes;
    float *row_bias;
    uint32_t *row_ids;
    int payload;
};

__device__ __forceinline__ uint32_t pool_row_of(const ScatterParams &p, uint32_t l, uint32_t pos) {
    const uint32_t seq = pos / kPageRows;
    const uint32_t page = seq < p.first_new_seq[l] ? p.tail_page[l] : p.new_base[l] + (seq - p.first_new_seq[l]);
    return page * kPageRows + (pos % kPageRows);
}

__global__ void __launch_bounds__(256) scatter_rows_kernel(const ScatterParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp_global; i < p.n; i += nwarps) {
        const uint32_t l = p.sorted_list[i], r = p.sorted_row[i];
        const uint32_t pos = p.list_len[l] + ((uint32_t)i - p.seg_start[l]);
        const uint32_t slot = pool_row_of(p, l, pos);
        const float *x = p.rows + (int64_t)r * p.stride;
        float acc = 0.f;
        if (p.payload == IVF_PRODUCER_TMA) {
            // k-block-major pages: element (row r of the page, dim j) at ((page * KB + j / 64) * 256 + r) * 64 + j % 64, so
            // that the 256 x 64 tile of one k-block is 32 KB contiguous in HBM (one DRAM-friendly TMA box per tile)
            const uint32_t page = slot / kPageRows, r_in = slot % kPageRows;
            const int kbc = p.d_pad64 / 64;
            __nv_bfloat16 *pbase = p.pool + (size_t)page * kPageRows * p.d_pad64;
            for (int j = lane; j < p.d_pad64; j += 32) {
                const __nv_bfloat16 b = __float2bfloat16_rn(j < p.d ? x[j] : 0.f);
                (void)kbc;
                pbase[((size_t)(j >> 6) * kPageRows + r_in) * 64 + (j & 63)] = b;
                const float v = __bfloat162float(b);
                acc = fmaf(v, v, acc);
            }
        } else if (p.payload == IVF_PRODUCER_SQ8) {
            uint8_t *dst = p.codes + (size_t)slot * p.code_bytes;
            for (int j = lane; j < p.code_bytes; j += 32) {
                uint32_t code = 128;  // padding decodes to 0
                if (j < p.d) {
                    const float t = rintf((x[j] - p.sq_lo[j]) * p.sq_inv_step[j]);
                    code = (uint32_t)fminf(fmaxf(t, 0.f), 255.f);
                    const float v = p.sq_lo[j] + (float)code * p.sq_step[j];   // the value this code stands for
                    acc = fmaf(v, v, acc);
                }
                dst[j] = (uint8_t)code;
            }
        } else {
            // PQ on the residual x - centroid[l]: lane handles sub-quantisers lane, lane + 32, ...
            uint8_t *dst = p.codes + (size_t)slot * p.code_bytes;
            const float *c = p.centroids + (size_t)l * p.d;
            for (int j = lane; j < p.code_bytes; j += 32) {
                uint32_t best = 0;
                if (j < p.m) {
                    const float *cb = p.pq + (size_t)j * 256 * p.dsub;
                    float bd = FLT_MAX;
                    for (int e = 0; e < 256; e++) {
                        float s = 0.f;
                        for (int t = 0; t < p.dsub; t++) {
                            const float u = (x[j * p.dsub + t] - c[j * p.dsub + t]) - cb[e * p.dsub + t];
                            s = fmaf(u, u, s);
                        }
                        if (s < bd) {
                            bd = s;
                            best = (uint32_t)e;
                        }
                    }
                    // norm term of the expanded L2 with the bf16 codebook values the scan sees: 2 <c, r^> + ||r^||^2
                    const __nv_bfloat16 *rb = p.pq_bf16 + ((size_t)j * 256 + best) * p.dsub;
                    for (int t = 0; t < p.dsub; t++) {
                        const float rv = __bfloat162float(rb[t]);
                        acc = fmaf(rv, rv + 2.f * c[j * p.dsub + t], acc);
                    }
                }
                dst[j] = (uint8_t)best;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            p.row_ids[slot] = p.id_base + r;
            if (p.row_bias) p.row_bias[slot] = acc;
        }
    }
}

__global__ void add_commit_kernel(const uint32_t *cnt, const uint32_t *new_base, const uint32_t *first_new_seq, uint32_t *list_len,
                                  uint32_t *tail_page, int nlist) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist || cnt[l] == 0) return;
    const uint32_t len = list_len[l] + cnt[l];
    const uint32_t last_seq = (len - 1) / kPageRows;
    if (last_seq >= first_new_seq[l]) tail_page[l] = new_base[l] + (last_seq - first_new_seq[l]);
    list_len[l] = len;
}

__global__ void page_keys_kernel(const uint32_t *owner, const uint32_t *seq, uint32_t n, uint64_t *keys, uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = ((uint64_t)owner[i] << 32) | seq[i];
    vals[i] = i;
}

__global__ void list_pages_scan_kernel(const uint32_t *list_len, int nlist, uint32_t *list_page_off, uint32_t *neg_len) {
    // single CTA: list_page_off = exclusive scan of ceil(len / 256); neg_len = ~len (sort key for "longest list first")
    const int per = (nlist + 1023) / 1024;
    const int l0 = threadIdx.x * per, l1 = min(nlist, l0 + per);
    uint32_t pages = 0;
    for (int l = l0; l < l1; l++) pages += (list_len[l] + kPageRows - 1) / kPageRows;
    uint32_t tot = 0;
    uint32_t off = block_exclusive_scan_1024(pages, &tot);
    for (int l = l0; l < l1; l++) {
        list_page_off[l] = off;
        off += (list_len[l] + kPageRows - 1) / kPageRows;
        neg_len[l] = ~list_len[l];
    }
    if (threadIdx.x == 0) list_page_off[nlist] = tot;
}

// per-dimension min / max of the training sample (SQ8)
__global__ void dim_minmax_kernel(const float *x, int64_t n, int64_t stride, int d, float *lo, float *hi) {
    const int j = blockIdx.x;
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
        const float v = x[r * stride + j];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    __shared__ float smn[256], smx[256];
    smn[threadIdx.x] = mn;
    smx[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + o]);
            smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        lo[j] = smn[0];
        hi[j] = smx[0];
    }
}

// ------------------------------------------------------------------------------------
// search: (query, list) pairs -> work items
// ------------------------------------------------------------------------------------
__global__ void pairs_make_kernel(const int64_t *probe, int64_t n_pairs, int nlist, uint32_t *keys, uint32_t *vals, uint32_t *cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int64_t l = probe[i];
    const uint32_t key = (l < 0 || l >= nlist) ? (uint32_t)nlist : (uint32_t)l;   // invalid probes sort behind every list
    keys[i] = key;
    vals[i] = (uint32_t)i;
    if (key < (uint32_t)nlist) atomicAdd(&cnt[key], 1u);
}

struct SearchPlan {
    const uint32_t *cnt;          // [nlist] pairs per list
    const uint32_t *list_len, *list_page_off, *list_order;   // list_order: lists by decreasing length
    uint32_t *pair_start;         // [nlist] out: first sorted pair of the list
    uint32_t *part_off;           // [nlist] out: first partial list of the list's first pair
    uint32_t *n_chunks;           // [nlist] out
    IvfGemmItem *items;           // out
    int *n_items, *n_parts;       // out (device scalars)
    unsigned long long *scan_rows;  // out: list rows the scan kernel will stream (every query tile of a list reads the whole list)
    int nlist, max_items;
    uint32_t pages_per_chunk;
};

__global__ void __launch_bounds__(1024) search_plan_kernel(const SearchPlan p) {
    const int per = (p.nlist + 1023) / 1024;
    const int l0 = threadIdx.x * per, l1 = min(p.nlist, l0 + per);
    // pass 1 in list-id order: pair_start (the sort order of the pairs)
    uint32_t pairs = 0;
    for (int l = l0; l < l1; l++) pairs += p.cnt[l];
    uint32_t poff = block_exclusive_scan_1024(pairs, nullptr);
    for (int l = l0; l < l1; l++) {
        p.pair_start[l] = poff;
        poff += p.cnt[l];
    }
    // pass 2 in decreasing-length order: items and partial lists (big lists first = LPT-like static schedule)
    uint32_t items = 0, parts = 0;
    unsigned long long rows_local = 0;
    for (int o = l0; o < l1; o++) {
        const uint32_t l = p.list_order[o];
        const uint32_t c = p.cnt[l], len = p.list_len[l];
        const uint32_t pages = (len + kPageRows - 1) / kPageRows;
        const uint32_t nch = (c && pages) ? (pages + p.pages_per_chunk - 1) / p.pages_per_chunk : 0;
        p.n_chunks[l] = nch;
        items += ((c + 127) / 128) * nch;
        parts += c * nch;
        if (nch) rows_local += (unsigned long long)((c + 127) / 128) * len;
    }
    if (threadIdx.x == 0) *p.scan_rows = 0;
    __syncthreads();
    if (rows_local) atomicAdd(p.scan_rows, rows_local);
    uint32_t tot_items = 0, tot_parts = 0;
    uint32_t ioff = block_exclusive_scan_1024(items, &tot_items);
    uint32_t paoff = block_exclusive_scan_1024(parts, &tot_parts);
    __syncthreads();
    for (int o = l0; o < l1; o++) {
        const uint32_t l = p.list_order[o];
        const uint32_t c = p.cnt[l], len = p.list_len[l], nch = p.n_chunks[l];
        p.part_off[l] = paoff;
        paoff += c * nch;
        if (!nch) continue;
        const uint32_t pages = (len + kPageRows - 1) / kPageRows;
        const uint32_t q0 = p.pair_start[l];
        for (uint32_t qt = 0; qt * 128 < c; qt++)
            for (uint32_t ch = 0; ch < nch; ch++) {
                if (ioff < (uint32_t)p.max_items) {
                    IvfGemmItem it;
                    it.q_begin = q0 + qt * 128;
                    it.q_count = min(128u, c - qt * 128);
                    it.page_begin = p.list_page_off[l] + ch * p.pages_per_chunk;
                    it.page_count = min(p.pages_per_chunk, pages - ch * p.pages_per_chunk);
                    it.row_limit = len - ch * p.pages_per_chunk * kPageRows;
                    it.chunk = ch;
                    p.items[ioff] = it;
                }
                ioff++;
            }
    }
    if (threadIdx.x == 0) {
        *p.n_items = (int)min(tot_items, (uint32_t)p.max_items);
        *p.n_parts = (int)tot_parts;
    }
}

// ------------------------------------------------------------------------------------
// Coarse probe for nprobe > 8: the centroid table is small (nlist x d fp32, L2-resident) and nprobe is a large k for it --
// the fused top-k kernels keep one k-entry list per query and, with only nlist / workers rows per list, almost every row is
// an insert (measured 21 ms for 10 000 queries x 4 096 centroids x 96, nprobe 32, on either path).  Here the ranking keys
// ||c||^2 - 2 <x, c> of a chunk of queries are written out by a plain fp32 tile kernel (64 x 64 tiles, 4 x 4 per thread) and
// one warp per query selects its nprobe smallest with a sorted warp list: scores are read once, coalesced, and an insert is
// O(nprobe / 32).
// ------------------------------------------------------------------------------------
constexpr int kCoarseTile = 64, kCoarseTK = 16;

__global__ void __launch_bounds__(256) coarse_scores_kernel(const float *x, int64_t ldx, const float *cent, const float *cnorm, int64_t nq,
                                                            int nl, int d, float *out /*[nq][nl]*/) {
    __shared__ float sx[kCoarseTK][kCoarseTile + 4], sc[kCoarseTK][kCoarseTile + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t q0 = (int64_t)blockIdx.y * kCoarseTile;
    const int c0 = blockIdx.x * kCoarseTile;
    float acc[4][4] = {};
    // loads: thread t fetches element (row t / 16 + 16 r, column t % 16) of each 64 x 16 operand tile, r = 0..3
    const int lr = threadIdx.x >> 4, lc = threadIdx.x & 15;
    for (int k0 = 0; k0 < d; k0 += kCoarseTK) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int row = lr + 16 * r;
            const int64_t q = q0 + row;
            const int c = c0 + row;
            const int kk = k0 + lc;
            sx[lc][row] = (q < nq && kk < d) ? x[q * ldx + kk] : 0.f;
            sc[lc][row] = (c < nl && kk < d) ? cent[(size_t)c * d + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kCoarseTK; kk++) {
            const float4 a = *reinterpret_cast<const float4 *>(&sx[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&sc[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t q = q0 + ty * 4 + i;
        if (q >= nq) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c0 + tx * 4 + j;
            if (c < nl) out[q * nl + c] = fmaf(-2.f, acc[i][j], cnorm[c]);
        }
    }
}

__global__ void __launch_bounds__(256) coarse_select_kernel(const float *scores, int64_t nq, int nl, int k, float *out_key, int64_t *out_ids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *lk = reinterpret_cast<float *>(smem_raw) + (size_t)warp * k;
    uint32_t *li = reinterpret_cast<uint32_t *>(reinterpret_cast<float *>(smem_raw) + (size_t)8 * k) + (size_t)warp * k;
    const int64_t q = (int64_t)blockIdx.x * 8 + warp;
    if (q >= nq) return;
    WarpTopK list;
    list.init(lk, li, k);
    const float *row = scores + q * nl;
    for (int c0 = 0; c0 < nl; c0 += 32) {
        const int c = c0 + lane;
        const float key = c < nl ? row[c] : FLT_MAX;
        unsigned m = __ballot_sync(0xffffffffu, c < nl && list.passes(key, (uint32_t)c));
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            list.insert(__shfl_sync(0xffffffffu, key, src), (uint32_t)(c0 + src));
        }
    }
    __syncwarp();
    for (int j = lane; j < k; j += 32) {
        out_key[q * k + j] = j < list.n ? lk[j] : FLT_MAX;
        out_ids[q * k + j] = j < list.n ? (int64_t)li[j] : -1;
    }
}

// One warp per sorted pair: gather (and for SQ8 scale) the query into the bf16 A-operand buffer, record where the pair's
// partial lists start, the inverse permutation, and the pair's additive constant (PQ: ||q - c||^2 or -<q, c>).
struct PairFill {
    const uint32_t *sorted_list, *sorted_pair;   // [n_pairs]
    const uint32_t *pair_start, *part_off, *n_chunks;
    const float *queries;        // [nq][d_pad] fp32, prepared (cosine: unit length)
    const float *sq_step;        // SQ8: per-dimension step (null otherwise)
    const float *centroids;      // PQ: [nlist][d] (null otherwise)
    __nv_bfloat16 *qbuf;         // [n_pairs][d_pad64]
    uint32_t *inv;               // [n_pairs] original pair -> sorted position
    uint32_t *pair_part_base;    // [n_pairs]
    float *pair_const;           // [n_pairs]
    int64_t n_pairs;
    int nprobe, nlist, d, d_pad, d_pad64, l2;
};

__global__ void __launch_bounds__(256) pair_fill_kernel(const PairFill p) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= p.n_pairs) return;
    const uint32_t l = p.sorted_list[i], pr = p.sorted_pair[i];
    if (lane == 0) p.inv[pr] = (uint32_t)i;
    if (l >= (uint32_t)p.nlist) {   // invalid probe: no work item reads this row
        if (lane == 0) {
            p.pair_part_base[i] = 0;
            p.pair_const[i] = 0.f;
        }
        return;
    }
    const uint32_t q = pr / (uint32_t)p.nprobe;
    const float *x = p.queries + (size_t)q * p.d_pad;
    __nv_bfloat16 *dst = p.qbuf + (size_t)i * p.d_pad64;
    const float *c = p.centroids ? p.centroids + (size_t)l * p.d : nullptr;
    float acc = 0.f;
    for (int j = lane; j < p.d_pad64; j += 32) {
        float v = j < p.d ? x[j] : 0.f;
        if (c && j < p.d) {
            const float cv = c[j];
            acc = p.l2 ? fmaf(v - cv, v - cv, acc) : fmaf(-v, cv, acc);
        }
        if (p.sq_step && j < p.d) v *= p.sq_step[j];
        dst[j] = __float2bfloat16_rn(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        p.pair_part_base[i] = p.part_off[l] + ((uint32_t)i - p.pair_start[l]) * p.n_chunks[l];
        p.pair_const[i] = acc;
    }
}

// per-query constants of the expanded distances: qc[q] = L2 ? ||q||^2 (- 2 <q, mid> for SQ8) : (- <q, mid> for SQ8, else 0)
__global__ void query_const_kernel(const float *queries, int64_t nq, int d, int d_pad, const float *sq_mid, int l2, int round_bf16, float *qc) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    const float *x = queries + (size_t)q * d_pad;
    float nn = 0.f, qm = 0.f;
    for (int j = lane; j < d; j += 32) {
        // bf16 payload: the scan multiplies bf16-rounded queries, so ||q||^2 is taken of the same values
        const float v = round_bf16 ? __bfloat162float(__float2bfloat16_rn(x[j])) : x[j];
        nn = fmaf(v, v, nn);
        if (sq_mid) qm = fmaf(x[j], sq_mid[j], qm);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
        qm += __shfl_xor_sync(0xffffffffu, qm, o);
    }
    if (lane == 0) qc[q] = l2 ? nn - 2.f * qm : -qm;
}

// One CTA per query: the partial lists of its nprobe pairs (x chunks) -> top-k, real distances.
struct IvfMerge {
    const uint32_t *inv, *pair_part_base, *sorted_list, *n_chunks;
    const float *pair_const, *query_const;
    const float *part_keys, *part_worst;
    const uint32_t *part_ids;
    float *out_dis;
    int64_t *out_ids;
    int64_t id_offset;
    int nprobe, nlist, k_part, k, metric;   // metric: B200_METRIC_*
};

__global__ void __launch_bounds__(256) ivf_merge_kernel(const IvfMerge p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *lk = reinterpret_cast<float *>(smem_raw);                      // [8][k] + merged [k]
    uint32_t *li = reinterpret_cast<uint32_t *>(lk + (size_t)9 * p.k);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = blockIdx.x;
    WarpTopK list;
    list.init(lk + (size_t)warp * p.k, li + (size_t)warp * p.k, p.k);
    for (int j = lane; j < p.k; j += 32) list.keys[j] = FLT_MAX;
    __syncwarp();
    // bound: a FULL partial list's worst key (+ its pair constant) bounds the query's k_part-th key from above
    if (p.k_part >= p.k) {
        __shared__ float bound_s[8];
        float b = FLT_MAX;
        for (int pr = 0; pr < p.nprobe; pr++) {
            const uint32_t i = p.inv[q * p.nprobe + pr];
            const uint32_t l = p.sorted_list[i];
            if (l >= (uint32_t)p.nlist) continue;
            const uint32_t nch = p.n_chunks[l];
            for (uint32_t ch = threadIdx.x; ch < nch; ch += blockDim.x) {
                const float w = p.part_worst[p.pair_part_base[i] + ch];
                if (w < FLT_MAX) b = fminf(b, w + p.pair_const[i]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
        if (lane == 0) bound_s[warp] = b;
        __syncthreads();
        b = bound_s[0];
#pragma unroll
        for (int w = 1; w < 8; w++) b = fminf(b, bound_s[w]);
        if (b < FLT_MAX) {
            list.thr_key = b;
            list.thr_id = kNoId;
        }
    }
    for (int pr = 0; pr < p.nprobe; pr++) {
        const uint32_t i = p.inv[q * p.nprobe + pr];
        const uint32_t l = p.sorted_list[i];
        if (l >= (uint32_t)p.nlist) continue;
        const int64_t ncand = (int64_t)p.n_chunks[l] * p.k_part;
        const size_t base = (size_t)p.pair_part_base[i] * p.k_part;
        const float pc = p.pair_const[i];
        for (int64_t c0 = (int64_t)warp * 32; c0 < ncand; c0 += 256) {
            const int64_t c = c0 + lane;
            float key = FLT_MAX;
            uint32_t id = kNoId;
            bool cand = false;
            if (c < ncand) {
                id = p.part_ids[base + c];
                key = p.part_keys[base + c] + pc;
                cand = id != kNoId && list.passes(key, id);
            }
            unsigned m = __ballot_sync(0xffffffffu, cand);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                list.insert(__shfl_sync(0xffffffffu, key, src), __shfl_sync(0xffffffffu, id, src));
            }
        }
    }
    __syncthreads();
    float *fk = lk + (size_t)8 * p.k;
    uint32_t *fi = li + (size_t)8 * p.k;
    block_rank_merge(lk, li, 8, p.k, p.k, fk, fi);
    __syncthreads();
    const float qc = p.query_const ? p.query_const[q] : 0.f;
    for (int j = threadIdx.x; j < p.k; j += blockDim.x) {
        float dis;
        int64_t id = -1;
        if (fi[j] != kNoId) {
            const float key = fk[j] + qc;
            id = (int64_t)fi[j] + p.id_offset;
            dis = p.metric == B200_METRIC_L2 ? fmaxf(key, 0.f) : p.metric == B200_METRIC_IP ? -key : 1.f + key;
        } else {
            dis = p.metric == B200_METRIC_IP ? -FLT_MAX : FLT_MAX;
        }
        p.out_dis[q * p.k + j] = dis;
        p.out_ids[q * p.k + j] = id;
    }
}

// ------------------------------------------------------------------------------------
// exact second stage: one CTA per query, one warp per candidate row (random 4*d-byte gathers)
// ------------------------------------------------------------------------------------
struct RefineParams {
    const float *queries;  // [nq][d_pad]
    const float *rows;     // [n][d_pad]
    const int64_t *cand;   // [nq][ncand], negative = empty
    float *out_dis;        // [nq][k]
    int64_t *out_ids;
    int64_t n;
    int d_pad, ncand, k;
    int l2;                // else inner product
    int cosine;            // output 1 - ip
    int64_t id_offset;     // added to every returned id (shard base)
};

__global__ void __launch_bounds__(256) refine_kernel(const RefineParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *qs = reinterpret_cast<float *>(smem_raw);
    float *lk = qs + p.d_pad;
    uint32_t *li = reinterpret_cast<uint32_t *>(lk + 8 * p.k);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = blockIdx.x;
    for (int i = threadIdx.x; i < p.d_pad; i += 256) qs[i] = p.queries[q * p.d_pad + i];
    WarpTopK list;
    list.init(lk + (size_t)warp * p.k, li + (size_t)warp * p.k, p.k);
    for (int j = lane; j < p.k; j += 32) list.keys[j] = FLT_MAX;
    __syncthreads();
    for (int c = warp; c < p.ncand; c += 8) {
        const int64_t id = p.cand[q * p.ncand + c];
        if (id < 0 || id >= p.n) continue;  // warp-uniform
        // duplicates in the candidate set would be returned twice; the first stage never produces them
        const float4 *row = reinterpret_cast<const float4 *>(p.rows + (size_t)id * p.d_pad);
        float acc = 0.f;
        for (int cc = lane; cc < p.d_pad / 4; cc += 32) {
            const float4 y = row[cc];
            const float4 x = reinterpret_cast<const float4 *>(qs)[cc];
            if (p.l2) {
                float t = x.x - y.x; acc = fmaf(t, t, acc);
                t = x.y - y.y; acc = fmaf(t, t, acc);
                t = x.z - y.z; acc = fmaf(t, t, acc);
                t = x.w - y.w; acc = fmaf(t, t, acc);
            } else {
                acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
                acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        list.insert(p.l2 ? acc : -acc, (uint32_t)id);
    }
    __syncthreads();
    float *fk = lk + (size_t)8 * p.k * 2;  // merged list, behind the 8 warp lists (keys + ids)
    uint32_t *fi = reinterpret_cast<uint32_t *>(fk + p.k);
    block_rank_merge(lk, li, 8, p.k, p.k, fk, fi);
    __syncthreads();
    for (int j = threadIdx.x; j < p.k; j += blockDim.x) {
        const bool have = fi[j] != kNoId;
        const float key = have ? fk[j] : 0.f;
        p.out_ids[q * p.k + j] = have ? (int64_t)fi[j] + p.id_offset : -1;
        p.out_dis[q * p.k + j] = !have ? (p.l2 || p.cosine ? FLT_MAX : -FLT_MAX) : p.l2 ? key : p.cosine ? 1.f + key : -key;
    }
}

}  // namespace b200

using namespace b200;

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
enum { IDX_FLAT = 0, IDX_IVFFLAT = 1, IDX_IVFPQ = 2, IDX_MSTG = 3, IDX_IVFSQ = 4, IDX_SCANN = 5, IDX_HNSWFLAT = 6, IDX_HNSWSQ = 7, IDX_HNSWPQ = 8 };
static const char *kTypeNames[] = {"FLAT", "IVFFLAT", "IVFPQ", "MSTG", "IVFSQ", "SCANN", "HNSWFLAT", "HNSWSQ", "HNSWPQ"};

struct DevArr {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return B200_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return fail(B200_ERR_NOMEM, "cudaMalloc(" + std::to_string(want) + ") failed (index workspace)");
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
    T *as() const { return reinterpret_cast<T *>(p); }
};

struct b200_index {
    int type = IDX_FLAT, metric = B200_METRIC_L2, d = 0, d_pad = 0, d_pad64 = 0;
    int nlist = 0, m = 0, dsub = 0;
    int default_nprobe = 32, refine_factor = 4;
    int payload = IVF_PRODUCER_TMA;
    int keep_raw = -1;              // -1 auto (yes), 0 no fp32 rows (first-stage distances only), 1 yes
    int code_bytes = 0;
    int64_t n = 0, reserved = 0;
    bool trained = false, built = false, use_ivf = false;
    b200_corpus *raw = nullptr;     // fp32 rows in id order (cosine: unit vectors), metric L2 or IP
    b200_corpus *coarse = nullptr;  // centroid table as a FLAT corpus (L2)
    float *d_centroids = nullptr;   // [nlist][d]
    float *d_cnorm = nullptr;       // [nlist] ||c||^2 (coarse probe)
    float *d_pq = nullptr;          // [m][256][dsub] fp32
    __nv_bfloat16 *d_pq_bf16 = nullptr;
    float *d_sq = nullptr;          // [4][d]: lo, step, 1/step, mid
    // paged lists
    uint32_t pool_pages = 0, pages_used = 0;
    void *d_pool = nullptr;         // bf16 [pool_pages * 256][d_pad64]  |  codes [pool_pages * 256][code_bytes]
    float *d_row_bias = nullptr;
    uint32_t *d_row_ids = nullptr;
    uint32_t *d_list_len = nullptr, *d_tail_page = nullptr, *d_page_owner = nullptr, *d_page_seq = nullptr, *d_pages_used = nullptr;
    int *d_flag = nullptr;
    uint32_t *d_list_page_off = nullptr, *d_list_pages = nullptr, *d_list_order = nullptr;   // after finalize
    std::vector<uint32_t> list_len;  // host copy after finalize
    uint32_t max_list_pages = 0;
    int device = 0, sms = 148;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    // workspaces (grow-only)
    DevArr w_rows, w_assign_i, w_assign_d, w_u32a, w_u32b, w_u32c, w_u32d, w_cnt, w_plan, w_sort, w_q, w_qraw, w_probe, w_pd, w_items,
        w_qbuf, w_inv, w_ppb, w_pconst, w_qconst, w_qb, w_cs, w_pk, w_pi, w_pw, w_lk, w_li, w_alive, w_od, w_oi, w_cand, w_host_q;
    // statistics of the last search (tests, bench roofline): rows x payload bytes the scan kernel was asked to stream
    int64_t last_scan_rows = 0, last_items = 0;
    bool timing = false, timed_pending = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_ph[6] = {};      // phase boundaries of the last search: start | coarse | pairs+plan+gather | scan | merge | refine
    double phase_ms[5] = {0, 0, 0, 0, 0};
    double timed_ms = 0;
    int64_t timed_launches = 0;
};

static void timing_collect(b200_index *ix);

// internal hooks into capi.cu
extern "C" int b200_corpus_search_device(b200_corpus *c, const float *d_queries, int64_t nq, int k, const uint8_t *d_alive_bits,
                                         int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, void *stream);
extern "C" int b200_corpus_set_path(b200_corpus *c, int path);
namespace b200 {
const void *corpus_device_rows(const b200_corpus *c);
int corpus_normalize_rows(b200_corpus *c);
int corpus_append_device(b200_corpus *c, const float *d_rows, int64_t n, cudaStream_t s);
}

static int parse_int_param(const char *json, const char *key, int defv) {
    if (!json) return defv;
    const size_t kl = strlen(key);
    for (const char *p = strstr(json, key); p; p = strstr(p + 1, key)) {
        // whole-word match: "m" must not hit "nprobe_m..." or the tail of "num"
        const bool left_ok = p == json || !(isalnum((unsigned char)p[-1]) || p[-1] == '_');
        const char *e = p + kl;
        const bool right_ok = !(isalnum((unsigned char)*e) || *e == '_');
        if (!left_ok || !right_ok) continue;
        while (*e && (*e == '"' || *e == ':' || *e == '=' || *e == ' ' || *e == '\'')) e++;
        if (!(*e >= '0' && *e <= '9')) continue;
        return atoi(e);
    }
    return defv;
}

extern "C" int b200_index_create(const char *type, int metric, int d, const char *params, b200_index **out) {
    if (!type || !out || d <= 0) return fail(B200_ERR_INVALID, "bad arguments");
    *out = nullptr;
    if (metric != B200_METRIC_L2 && metric != B200_METRIC_IP && metric != B200_METRIC_COSINE)
        return fail(B200_ERR_INVALID, "float indexes take L2, IP or COSINE");
    std::string t(type);
    for (auto &ch : t) ch = (char)toupper((unsigned char)ch);
    int ty = -1;
    for (int i = 0; i < 9; i++)
        if (t == kTypeNames[i]) ty = i;
    if (ty < 0) return fail(B200_ERR_UNSUPPORTED, "index type " + t + " is not implemented (FLAT, IVFFLAT, IVFSQ, IVFPQ, SCANN, MSTG, HNSWFLAT, HNSWSQ, HNSWPQ)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B200_ERR_NO_DEVICE, "no CUDA device visible; libb200search has no CPU fallback");
    }
    b200_index *ix = new b200_index();
    ix->type = ty;
    ix->metric = metric;
    ix->d = d;
    ix->d_pad = (int)round_up(d, 4);
    ix->d_pad64 = (int)round_up(d, 64);
    ix->nlist = parse_int_param(params, "ncentroids", parse_int_param(params, "nlist", 0));
    ix->m = parse_int_param(params, "M", parse_int_param(params, "m", 0));
    ix->default_nprobe = parse_int_param(params, "nprobe", 32);
    // payload of a list row.  The graph types of the reference (hnswlib) and ScaNN have no graph / anisotropic quantiser
    // here: they are SERVED by the inverted-file engine with the payload their suffix names (recall contract).
    switch (ty) {
        case IDX_IVFPQ: case IDX_SCANN: case IDX_HNSWPQ: ix->payload = IVF_PRODUCER_PQ; break;
        case IDX_IVFSQ: case IDX_HNSWSQ: ix->payload = IVF_PRODUCER_SQ8; break;
        default: ix->payload = IVF_PRODUCER_TMA;
    }
    // candidates re-ranked exactly per returned row when fp32 rows are kept; plain IVFPQ / IVFSQ return first-stage
    // (ADC) distances like Faiss unless asked (refine_factor > 1)
    const int dflt_refine = (ty == IDX_IVFPQ || ty == IDX_IVFSQ) ? 1 : (ix->payload == IVF_PRODUCER_PQ ? 16 : 4);
    ix->refine_factor = parse_int_param(params, "refine_factor", parse_int_param(params, "reorder_k_factor", dflt_refine));
    ix->keep_raw = parse_int_param(params, "keep_raw", -1);
    cudaGetDevice(&ix->device);
    cudaDeviceGetAttribute(&ix->sms, cudaDevAttrMultiProcessorCount, ix->device);
    if (cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ix;
        return fail(B200_ERR_CUDA, "cudaStreamCreate failed");
    }
    *out = ix;
    return B200_OK;
}

extern "C" int b200_index_free(b200_index *ix) {
    if (!ix) return B200_OK;
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    if (ix->raw) b200_corpus_free(ix->raw);
    if (ix->coarse) b200_corpus_free(ix->coarse);
    for (void *p : {(void *)ix->d_centroids, (void *)ix->d_cnorm, (void *)ix->d_pq, (void *)ix->d_pq_bf16, (void *)ix->d_sq, ix->d_pool, (void *)ix->d_row_bias,
                    (void *)ix->d_row_ids, (void *)ix->d_list_len, (void *)ix->d_tail_page, (void *)ix->d_page_owner, (void *)ix->d_page_seq,
                    (void *)ix->d_pages_used, (void *)ix->d_flag, (void *)ix->d_list_page_off, (void *)ix->d_list_pages, (void *)ix->d_list_order})
        if (p) cudaFree(p);
    for (DevArr *a : {&ix->w_rows, &ix->w_assign_i, &ix->w_assign_d, &ix->w_u32a, &ix->w_u32b, &ix->w_u32c, &ix->w_u32d, &ix->w_cnt, &ix->w_plan,
                      &ix->w_sort, &ix->w_q, &ix->w_qraw, &ix->w_probe, &ix->w_pd, &ix->w_items, &ix->w_qbuf, &ix->w_inv, &ix->w_ppb, &ix->w_pconst, &ix->w_qb, &ix->w_cs,
                      &ix->w_qconst, &ix->w_pk, &ix->w_pi, &ix->w_pw, &ix->w_lk, &ix->w_li, &ix->w_alive, &ix->w_od, &ix->w_oi, &ix->w_cand,
                      &ix->w_host_q})
        a->release();
    if (ix->ev0) cudaEventDestroy(ix->ev0);
    if (ix->ev1) cudaEventDestroy(ix->ev1);
    for (auto &e : ix->ev_ph)
        if (e) cudaEventDestroy(e);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
    return B200_OK;
}

static size_t payload_row_bytes(const b200_index *ix) {
    return ix->payload == IVF_PRODUCER_TMA ? (size_t)ix->d_pad64 * 2 : (size_t)ix->code_bytes;
}

// k-means on device rows x [n][stride]; centroids written to d_c [nc][d].  Assignment: exact top-1 search of the centroid
// table with the FLAT engine (tensor cores from 20 rows up) when the table is large, the tiled fp32 kernel otherwise.
static int kmeans_device(const float *x, int64_t n, int64_t stride, int d, int nc, int iters, float *d_c, cudaStream_t s) {
    std::vector<int64_t> pick(nc);
    for (int i = 0; i < nc; i++) pick[i] = (int64_t)((double)i * (double)n / (double)nc);
    int64_t *d_pick = nullptr;
    float *d_sums = nullptr, *d_cn = nullptr, *d_dis = nullptr;
    uint32_t *d_cnt = nullptr, *d_idx = nullptr;
    int64_t *d_idx64 = nullptr;
    B200_CUDA_OK(cudaMalloc(&d_pick, (size_t)nc * 8));
    B200_CUDA_OK(cudaMalloc(&d_sums, (size_t)nc * d * 4));
    B200_CUDA_OK(cudaMalloc(&d_cn, (size_t)nc * 4));
    B200_CUDA_OK(cudaMalloc(&d_cnt, (size_t)nc * 4));
    B200_CUDA_OK(cudaMalloc(&d_idx, (size_t)n * 4));
    B200_CUDA_OK(cudaMemcpyAsync(d_pick, pick.data(), (size_t)nc * 8, cudaMemcpyHostToDevice, s));
    gather_rows_kernel<<<gridsz((int64_t)nc * d), 256, 0, s>>>(x, stride, d_pick, nc, d, d_c);
    g_launches++;
    const bool big = (double)n * nc * d > 2e11 && stride == d;   // tensor-core assignment pays from ~0.2 TFLOP per iteration
    b200_corpus *table = nullptr;
    if (big) {
        B200_CUDA_OK(cudaMalloc(&d_idx64, (size_t)n * 8));
        B200_CUDA_OK(cudaMalloc(&d_dis, (size_t)n * 4));
    }
    int rc = B200_OK;
    for (int it = 0; it < iters && rc == B200_OK; it++) {
        if (big) {
            if (table) b200_corpus_free(table);
            table = nullptr;
            rc = b200_corpus_create(B200_METRIC_L2, B200_DTYPE_F32, d, nc, &table);
            if (rc == B200_OK) rc = corpus_append_device(table, d_c, nc, s);
            if (rc == B200_OK) rc = b200_corpus_search_device(table, x, n, 1, nullptr, 0, d_dis, d_idx64, s);
            if (rc != B200_OK) break;
            B200_CUDA_OK(cudaMemsetAsync(d_cnt, 0, (size_t)nc * 4, s));
            assign_to_u32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(d_idx64, n, d_idx, d_cnt);
            g_launches++;
        } else {
            rows_sqnorm_kernel<<<(unsigned)ceil_div(nc, 256), 256, 0, s>>>(d_c, nc, d, d_cn);
            kmeans_assign_kernel<<<(unsigned)ceil_div(n, KA_T), 256, 0, s>>>(x, n, stride, d, d_c, nc, d_cn, d_idx, nullptr);
            g_launches += 2;
        }
        B200_CUDA_OK(cudaMemsetAsync(d_sums, 0, (size_t)nc * d * 4, s));
        B200_CUDA_OK(cudaMemsetAsync(d_cnt, 0, (size_t)nc * 4, s));
        kmeans_accumulate_kernel<<<gridsz(n * d), 256, 0, s>>>(x, n, stride, d, d_idx, d_sums, d_cnt);
        kmeans_update_kernel<<<gridsz((int64_t)nc * d), 256, 0, s>>>(d_c, d_sums, d_cnt, nc, d);
        g_launches += 2;
        // empty clusters take half of the currently largest ones (both copies nudged apart; the next assignment splits the
        // members) -- without this a strided initialisation leaves a third of well-separated clusters without a centroid
        if (it + 1 < iters && nc >= 2) {
            std::vector<uint32_t> h_cnt(nc);
            B200_CUDA_OK(cudaMemcpyAsync(h_cnt.data(), d_cnt, (size_t)nc * 4, cudaMemcpyDeviceToHost, s));
            B200_CUDA_OK(cudaStreamSynchronize(s));
            std::vector<int> empties, order(nc);
            for (int i = 0; i < nc; i++) {
                order[i] = i;
                if (h_cnt[i] == 0) empties.push_back(i);
            }
            if (!empties.empty()) {
                std::partial_sort(order.begin(), order.begin() + std::min<size_t>(nc, empties.size()), order.end(),
                                  [&](int a, int b) { return h_cnt[a] > h_cnt[b]; });
                std::vector<int> pairs;
                for (size_t e = 0; e < empties.size() && e < (size_t)nc; e++) {
                    const int src = order[e];
                    if (h_cnt[src] < 2) break;
                    pairs.push_back(empties[e]);
                    pairs.push_back(src);
                }
                if (!pairs.empty()) {
                    int *d_pairs = nullptr;
                    B200_CUDA_OK(cudaMalloc(&d_pairs, pairs.size() * 4));
                    B200_CUDA_OK(cudaMemcpyAsync(d_pairs, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice, s));
                    kmeans_split_kernel<<<(unsigned)(pairs.size() / 2), 128, 0, s>>>(d_c, d_pairs, d);
                    g_launches++;
                    B200_CUDA_OK(cudaStreamSynchronize(s));
                    cudaFree(d_pairs);
                }
            }
        }
    }
    if (rc == B200_OK) {
        B200_CUDA_OK(cudaGetLastError());
        B200_CUDA_OK(cudaStreamSynchronize(s));
    }
    if (table) b200_corpus_free(table);
    for (void *p : {(void *)d_pick, (void *)d_sums, (void *)d_cn, (void *)d_cnt, (void *)d_idx, (void *)d_idx64, (void *)d_dis})
        if (p) cudaFree(p);
    return rc;
}

// total rows the index will hold (Search::createVectorIndex's total_vec, VIWithDataPart.cpp:416-430): sizes the page pool
extern "C" int b200_index_reserve(b200_index *ix, int64_t total_rows) {
    if (!ix || total_rows < 0) return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (ix->trained || ix->n) return fail(B200_ERR_INVALID, "reserve comes before train / add");
    ix->reserved = total_rows;
    return B200_OK;
}

static int upload_coarse(b200_index *ix, cudaStream_t s) {
    if (ix->coarse) b200_corpus_free(ix->coarse);
    ix->coarse = nullptr;
    // L2 for every metric (unit vectors under cosine; IP indexes probe by L2 too, like Faiss's default quantiser)
    B200_TRY(b200_corpus_create(B200_METRIC_L2, B200_DTYPE_F32, ix->d, ix->nlist, &ix->coarse));
    if (ix->d_cnorm) cudaFree(ix->d_cnorm);
    ix->d_cnorm = nullptr;
    B200_CUDA_OK(cudaMalloc(&ix->d_cnorm, (size_t)ix->nlist * 4));
    B200_CUDA_OK(launch_row_norms(ix->d_centroids, 0, ix->d, ix->nlist, 0, ix->d_cnorm, s));
    return corpus_append_device(ix->coarse, ix->d_centroids, ix->nlist, s);
}

// Search::VectorIndex::train: coarse quantiser (+ PQ codebooks / SQ ranges) from a sample already on the device,
// rows fp32 [n][d] contiguous.  Decides FLAT fallback for small parts (the reference's fallback_to_flat, test 00029).
static int train_device_locked(b200_index *ix, const float *d_rows, int64_t n) {
    if (ix->trained) return fail(B200_ERR_INVALID, "index already trained");
    if (ix->built) return fail(B200_ERR_INVALID, "index already built");
    cudaStream_t s = ix->stream;
    const int d = ix->d;
    const int64_t total = ix->reserved > 0 ? ix->reserved : n;
    const bool want_ivf = ix->type != IDX_FLAT;
    if (want_ivf && ix->nlist <= 0)
        ix->nlist = (int)std::max<int64_t>(1, std::min<int64_t>(65536, (int64_t)(4.0 * sqrt((double)std::max<int64_t>(total, 1)))));
    ix->use_ivf = want_ivf && total >= std::max<int64_t>(2000, 8ll * ix->nlist) && n >= ix->nlist;
    if (!ix->use_ivf) {
        ix->keep_raw = 1;
        ix->trained = true;
        return B200_OK;
    }
    if (ix->keep_raw < 0) ix->keep_raw = 1;
    const int nl = ix->nlist;
    // training rows: unit length under cosine
    const float *x = d_rows;
    if (ix->metric == B200_METRIC_COSINE) {
        B200_TRY(ix->w_rows.reserve((size_t)n * d * 4));
        B200_CUDA_OK(cudaMemcpyAsync(ix->w_rows.p, d_rows, (size_t)n * d * 4, cudaMemcpyDeviceToDevice, s));
        B200_CUDA_OK(launch_normalize_rows_f32(ix->w_rows.as<float>(), d, n, s));
        x = ix->w_rows.as<float>();
    }
    B200_CUDA_OK(cudaMalloc(&ix->d_centroids, (size_t)nl * d * 4));
    B200_TRY(kmeans_device(x, n, d, d, nl, 10, ix->d_centroids, s));
    B200_TRY(upload_coarse(ix, s));
    if (ix->payload == IVF_PRODUCER_SQ8) {
        ix->code_bytes = (int)round_up(d, 16);
        B200_CUDA_OK(cudaMalloc(&ix->d_sq, (size_t)4 * d * 4));
        float *lo = ix->d_sq, *hi = ix->d_sq + d;
        dim_minmax_kernel<<<d, 256, 0, s>>>(x, n, d, d, lo, hi);
        g_launches++;
        std::vector<float> h((size_t)4 * d);
        B200_CUDA_OK(cudaMemcpyAsync(h.data(), ix->d_sq, (size_t)2 * d * 4, cudaMemcpyDeviceToHost, s));
        B200_CUDA_OK(cudaStreamSynchronize(s));
        for (int j = 0; j < d; j++) {
            const float l = h[j], u = h[d + j];
            const float step = u > l ? (u - l) / 255.f : 1.f;
            h[d + j] = step;
            h[2 * d + j] = 1.f / step;
            h[3 * d + j] = l + 128.f * step;   // value of code 128 = the zero of the offset-binary code the scan decodes
        }
        B200_CUDA_OK(cudaMemcpyAsync(ix->d_sq, h.data(), (size_t)4 * d * 4, cudaMemcpyHostToDevice, s));
    }
    if (ix->payload == IVF_PRODUCER_PQ) {
        if (ix->m <= 0) {  // default: sub-vectors of <= 8 dims
            ix->m = d;
            for (int cand : {8, 4, 2, 1})
                if (d % cand == 0) { ix->m = d / cand; break; }
        }
        if (d % ix->m) return fail(B200_ERR_INVALID, "PQ M must divide the dimension");
        const int m = ix->m, dsub = d / m;
        if (dsub != 1 && dsub != 2 && dsub != 4 && dsub != 8)
            return fail(B200_ERR_UNSUPPORTED, "PQ sub-vector length d / M must be 1, 2, 4 or 8 (codes are decoded into tensor-core tiles)");
        if ((size_t)256 * ix->d_pad64 * 2 > 160 * 1024)
            return fail(B200_ERR_UNSUPPORTED, "PQ codebook (512 B x d) must fit in shared memory next to the operand ring: d <= 320");
        ix->dsub = dsub;
        ix->code_bytes = (int)round_up(m, 16);
        B200_CUDA_OK(cudaMalloc(&ix->d_pq, (size_t)m * 256 * dsub * 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_pq_bf16, (size_t)m * 256 * dsub * 2));
        // residuals of (a sample of) the training rows, one sub-quantiser at a time
        const int64_t ns = std::min<int64_t>(n, 65536);
        int64_t *d_a = nullptr;
        float *d_ad = nullptr, *d_res = nullptr;
        uint32_t *d_l = nullptr, *d_c32 = nullptr;
        B200_CUDA_OK(cudaMalloc(&d_a, (size_t)ns * 8));
        B200_CUDA_OK(cudaMalloc(&d_ad, (size_t)ns * 4));
        B200_CUDA_OK(cudaMalloc(&d_l, (size_t)ns * 4));
        B200_CUDA_OK(cudaMalloc(&d_c32, (size_t)nl * 4));
        B200_CUDA_OK(cudaMalloc(&d_res, (size_t)ns * dsub * 4));
        // the first ns rows of a strided view
        const int64_t step = std::max<int64_t>(1, n / ns);
        float *d_samp = nullptr;
        B200_CUDA_OK(cudaMalloc(&d_samp, (size_t)ns * d * 4));
        B200_CUDA_OK(cudaMemcpy2DAsync(d_samp, (size_t)d * 4, x, (size_t)step * d * 4, (size_t)d * 4, ns, cudaMemcpyDeviceToDevice, s));
        int rc = b200_corpus_search_device(ix->coarse, d_samp, ns, 1, nullptr, 0, d_ad, d_a, s);
        if (rc == B200_OK) {
            cudaMemsetAsync(d_c32, 0, (size_t)nl * 4, s);
            assign_to_u32_kernel<<<(unsigned)ceil_div(ns, 256), 256, 0, s>>>(d_a, ns, d_l, d_c32);
            g_launches++;
            for (int j = 0; j < m && rc == B200_OK; j++) {
                residual_sub_kernel<<<gridsz(ns * dsub), 256, 0, s>>>(d_samp, ns, d, ix->d_centroids, d_l, d, j, dsub, d_res);
                g_launches++;
                rc = kmeans_device(d_res, ns, dsub, dsub, 256, 8, ix->d_pq + (size_t)j * 256 * dsub, s);
            }
        }
        if (rc == B200_OK) {
            cudaError_t e = launch_f32_to_bf16_rows(ix->d_pq, dsub, ix->d_pq_bf16, dsub, (int64_t)m * 256, s);
            if (e != cudaSuccess) rc = fail(B200_ERR_CUDA, cudaGetErrorString(e));
        }
        cudaStreamSynchronize(s);
        for (void *p : {(void *)d_a, (void *)d_ad, (void *)d_l, (void *)d_c32, (void *)d_res, (void *)d_samp}) cudaFree(p);
        B200_TRY(rc);
    }
    // ---- page pool: every list wastes less than one page
    {
        const int64_t pages = ceil_div(total, kPageRows) + nl;
        if (pages * kPageRows >= (int64_t)0xffffffffll) return fail(B200_ERR_UNSUPPORTED, "an index shard is limited to 2^32 - 1 pool rows");
        ix->pool_pages = (uint32_t)pages;
        const size_t rows = (size_t)pages * kPageRows;
        if (cudaMalloc(&ix->d_pool, rows * payload_row_bytes(ix) + 256) != cudaSuccess) {
            cudaGetLastError();
            return fail(B200_ERR_NOMEM, "cudaMalloc of the page pool failed (" + std::to_string(rows * payload_row_bytes(ix)) + " bytes)");
        }
        B200_CUDA_OK(cudaMemsetAsync(ix->d_pool, 0, rows * payload_row_bytes(ix), s));
        B200_CUDA_OK(cudaMalloc(&ix->d_row_ids, rows * 4));
        if (ix->metric == B200_METRIC_L2) B200_CUDA_OK(cudaMalloc(&ix->d_row_bias, rows * 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_list_len, (size_t)nl * 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_tail_page, (size_t)nl * 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_page_owner, (size_t)pages * 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_page_seq, (size_t)pages * 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_pages_used, 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_flag, 32));
        B200_CUDA_OK(cudaMemsetAsync(ix->d_list_len, 0, (size_t)nl * 4, s));
        B200_CUDA_OK(cudaMemsetAsync(ix->d_tail_page, 0, (size_t)nl * 4, s));
        B200_CUDA_OK(cudaMemsetAsync(ix->d_pages_used, 0, 4, s));
        B200_CUDA_OK(cudaMemsetAsync(ix->d_flag, 0, 32, s));
    }
    B200_CUDA_OK(cudaStreamSynchronize(s));
    ix->trained = true;
    return B200_OK;
}

extern "C" int b200_index_train_device(b200_index *ix, const float *d_rows, int64_t n) {
    if (!ix || (!d_rows && n > 0) || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    return train_device_locked(ix, d_rows, n);
}

extern "C" int b200_index_train(b200_index *ix, const float *rows, int64_t n) {
    if (!ix || (!rows && n > 0) || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    B200_TRY(ix->w_host_q.reserve((size_t)std::max<int64_t>(n, 1) * ix->d * 4));
    B200_TRY(staged_h2d(ix->w_host_q.p, rows, (size_t)n * ix->d * 4, ix->device, ix->stream));
    B200_CUDA_OK(cudaStreamSynchronize(ix->stream));
    int rc = train_device_locked(ix, ix->w_host_q.as<float>(), n);
    ix->w_host_q.release();
    return rc;
}

// Search::VectorIndex::add of one chunk already on the device (fp32 [n][d] contiguous); row ids continue from ix->n
static int add_device_locked(b200_index *ix, const float *d_rows, int64_t n) {
    if (!ix->trained) return fail(B200_ERR_INVALID, "train the index before adding rows");
    if (ix->built) return fail(B200_ERR_INVALID, "index already finalized");
    if (n == 0) return B200_OK;
    cudaStream_t s = ix->stream;
    const int d = ix->d, nl = ix->nlist;
    if (ix->n + n >= (int64_t)0xffffffffll) return fail(B200_ERR_UNSUPPORTED, "an index shard is limited to 2^32 - 1 rows");
    const float *x = d_rows;
    if (ix->metric == B200_METRIC_COSINE) {
        B200_TRY(ix->w_rows.reserve((size_t)n * d * 4));
        B200_CUDA_OK(cudaMemcpyAsync(ix->w_rows.p, d_rows, (size_t)n * d * 4, cudaMemcpyDeviceToDevice, s));
        B200_CUDA_OK(launch_normalize_rows_f32(ix->w_rows.as<float>(), d, n, s));
        x = ix->w_rows.as<float>();
    }
    if (ix->keep_raw == 1) {
        if (!ix->raw) {
            const int raw_metric = ix->metric == B200_METRIC_L2 ? B200_METRIC_L2 : B200_METRIC_IP;
            B200_TRY(b200_corpus_create(raw_metric, B200_DTYPE_F32, d, std::max<int64_t>(ix->reserved, n), &ix->raw));
        }
        B200_TRY(corpus_append_device(ix->raw, x, n, s));
    }
    if (!ix->use_ivf) {
        ix->n += n;
        return B200_OK;
    }
    // ---- assign -> (list, row) sorted by list
    B200_TRY(ix->w_assign_i.reserve((size_t)n * 8));
    B200_TRY(ix->w_assign_d.reserve((size_t)n * 4));
    B200_TRY(ix->w_u32a.reserve((size_t)n * 4));
    B200_TRY(ix->w_u32b.reserve((size_t)n * 4));
    B200_TRY(ix->w_u32c.reserve((size_t)n * 4));
    B200_TRY(ix->w_u32d.reserve((size_t)n * 4));
    B200_TRY(ix->w_cnt.reserve((size_t)nl * 4));
    B200_TRY(ix->w_plan.reserve((size_t)nl * 4 * 3));
    B200_TRY(b200_corpus_search_device(ix->coarse, x, n, 1, nullptr, 0, ix->w_assign_d.as<float>(), ix->w_assign_i.as<int64_t>(), s));
    B200_CUDA_OK(cudaMemsetAsync(ix->w_cnt.p, 0, (size_t)nl * 4, s));
    assign_to_u32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(ix->w_assign_i.as<int64_t>(), n, ix->w_u32a.as<uint32_t>(), ix->w_cnt.as<uint32_t>());
    iota_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(ix->w_u32b.as<uint32_t>(), n);
    g_launches += 2;
    {
        int bits = 1;
        while ((1 << bits) < nl) bits++;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ix->w_u32a.as<uint32_t>(), ix->w_u32c.as<uint32_t>(), ix->w_u32b.as<uint32_t>(),
                                        ix->w_u32d.as<uint32_t>(), (int)n, 0, bits, s);
        B200_TRY(ix->w_sort.reserve(tmp_bytes + 256));
        cub::DeviceRadixSort::SortPairs(ix->w_sort.p, tmp_bytes, ix->w_u32a.as<uint32_t>(), ix->w_u32c.as<uint32_t>(), ix->w_u32b.as<uint32_t>(),
                                        ix->w_u32d.as<uint32_t>(), (int)n, 0, bits, s);
        g_launches++;
    }
    uint32_t *seg_start = ix->w_plan.as<uint32_t>(), *new_base = seg_start + nl, *first_new = new_base + nl;
    AddPlan ap{};
    ap.cnt = ix->w_cnt.as<uint32_t>();
    ap.seg_start = seg_start;
    ap.new_base = new_base;
    ap.first_new_seq = first_new;
    ap.list_len = ix->d_list_len;
    ap.page_owner = ix->d_page_owner;
    ap.page_seq = ix->d_page_seq;
    ap.pages_used = ix->d_pages_used;
    ap.pool_pages = ix->pool_pages;
    ap.nlist = nl;
    ap.overflow = ix->d_flag;
    add_plan_kernel<<<1, 1024, 0, s>>>(ap);
    g_launches++;
    ScatterParams sp{};
    sp.rows = x;
    sp.stride = d;
    sp.sorted_list = ix->w_u32c.as<uint32_t>();
    sp.sorted_row = ix->w_u32d.as<uint32_t>();
    sp.seg_start = seg_start;
    sp.new_base = new_base;
    sp.first_new_seq = first_new;
    sp.list_len = ix->d_list_len;
    sp.tail_page = ix->d_tail_page;
    sp.id_base = (uint32_t)ix->n;
    sp.n = n;
    sp.d = d;
    sp.d_pad64 = ix->d_pad64;
    sp.l2 = ix->metric == B200_METRIC_L2;
    sp.pool = reinterpret_cast<__nv_bfloat16 *>(ix->d_pool);
    if (ix->d_sq) {
        sp.sq_lo = ix->d_sq;
        sp.sq_step = ix->d_sq + d;
        sp.sq_inv_step = ix->d_sq + 2 * d;
    }
    sp.centroids = ix->d_centroids;
    sp.pq = ix->d_pq;
    sp.pq_bf16 = ix->d_pq_bf16;
    sp.m = ix->m;
    sp.dsub = ix->dsub;
    sp.codes = reinterpret_cast<uint8_t *>(ix->d_pool);
    sp.code_bytes = ix->code_bytes;
    sp.row_bias = ix->d_row_bias;
    sp.row_ids = ix->d_row_ids;
    sp.payload = ix->payload;
    int over = 0;
    B200_CUDA_OK(cudaMemcpyAsync(&over, ix->d_flag, 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    if (over) return fail(B200_ERR_NOMEM, "page pool exhausted: more rows added than b200_index_reserve() announced");
    scatter_rows_kernel<<<gridsz(n * 32), 256, 0, s>>>(sp);
    add_commit_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, s>>>(ix->w_cnt.as<uint32_t>(), new_base, first_new, ix->d_list_len, ix->d_tail_page, nl);
    g_launches += 2;
    B200_CUDA_OK(cudaGetLastError());
    B200_CUDA_OK(cudaStreamSynchronize(s));
    ix->n += n;
    return B200_OK;
}

extern "C" int b200_index_add_device(b200_index *ix, const float *d_rows, int64_t n) {
    if (!ix || (!d_rows && n > 0) || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    // bounded scratch: sub-chunks of <= 1 M rows
    for (int64_t off = 0; off < n; off += (1 << 20))
        B200_TRY(add_device_locked(ix, d_rows + off * ix->d, std::min<int64_t>(1 << 20, n - off)));
    return B200_OK;
}

extern "C" int b200_index_add(b200_index *ix, const float *rows, int64_t n) {
    if (!ix || (!rows && n > 0) || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    const int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(1 << 20, (int64_t)(1ll << 30) / ((int64_t)ix->d * 4)));
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t mrows = std::min(chunk, n - off);
        B200_TRY(ix->w_host_q.reserve((size_t)mrows * ix->d * 4));
        B200_TRY(staged_h2d(ix->w_host_q.p, rows + off * ix->d, (size_t)mrows * ix->d * 4, ix->device, ix->stream));
        B200_CUDA_OK(cudaStreamSynchronize(ix->stream));
        B200_TRY(add_device_locked(ix, ix->w_host_q.as<float>(), mrows));
    }
    return B200_OK;
}

static int finalize_locked(b200_index *ix) {
    if (ix->built) return B200_OK;
    if (!ix->trained) return fail(B200_ERR_INVALID, "index not trained");
    cudaStream_t s = ix->stream;
    if (ix->use_ivf) {
        const int nl = ix->nlist;
        B200_CUDA_OK(cudaMemcpy(&ix->pages_used, ix->d_pages_used, 4, cudaMemcpyDeviceToHost));
        const uint32_t np = ix->pages_used;
        B200_CUDA_OK(cudaMalloc(&ix->d_list_page_off, (size_t)(nl + 1) * 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_list_pages, (size_t)std::max<uint32_t>(np, 1) * 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_list_order, (size_t)nl * 4));
        uint64_t *keys = nullptr, *keys_out = nullptr;
        uint32_t *vals = nullptr, *neg = nullptr, *neg_out = nullptr, *iota = nullptr;
        B200_CUDA_OK(cudaMalloc(&keys, (size_t)std::max<uint32_t>(np, 1) * 8));
        B200_CUDA_OK(cudaMalloc(&keys_out, (size_t)std::max<uint32_t>(np, 1) * 8));
        B200_CUDA_OK(cudaMalloc(&vals, (size_t)std::max<uint32_t>(np, 1) * 4));
        B200_CUDA_OK(cudaMalloc(&neg, (size_t)nl * 4));
        B200_CUDA_OK(cudaMalloc(&neg_out, (size_t)nl * 4));
        B200_CUDA_OK(cudaMalloc(&iota, (size_t)nl * 4));
        if (np) {
            page_keys_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, s>>>(ix->d_page_owner, ix->d_page_seq, np, keys, vals);
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_out, vals, ix->d_list_pages, (int)np, 0, 64, s);
            B200_TRY(ix->w_sort.reserve(tb + 256));
            cub::DeviceRadixSort::SortPairs(ix->w_sort.p, tb, keys, keys_out, vals, ix->d_list_pages, (int)np, 0, 64, s);
            g_launches += 2;
        }
        list_pages_scan_kernel<<<1, 1024, 0, s>>>(ix->d_list_len, nl, ix->d_list_page_off, neg);
        iota_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, s>>>(iota, nl);
        {
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, neg, neg_out, iota, ix->d_list_order, nl, 0, 32, s);
            B200_TRY(ix->w_sort.reserve(tb + 256));
            cub::DeviceRadixSort::SortPairs(ix->w_sort.p, tb, neg, neg_out, iota, ix->d_list_order, nl, 0, 32, s);
        }
        g_launches += 3;
        ix->list_len.resize(nl);
        B200_CUDA_OK(cudaMemcpyAsync(ix->list_len.data(), ix->d_list_len, (size_t)nl * 4, cudaMemcpyDeviceToHost, s));
        B200_CUDA_OK(cudaStreamSynchronize(s));
        for (void *p : {(void *)keys, (void *)keys_out, (void *)vals, (void *)neg, (void *)neg_out, (void *)iota}) cudaFree(p);
        ix->max_list_pages = 0;
        for (int l = 0; l < nl; l++) ix->max_list_pages = std::max<uint32_t>(ix->max_list_pages, (ix->list_len[l] + kPageRows - 1) / kPageRows);
    }
    // build scratch is not needed any more
    for (DevArr *a : {&ix->w_rows, &ix->w_assign_i, &ix->w_assign_d, &ix->w_u32a, &ix->w_u32b, &ix->w_u32c, &ix->w_u32d, &ix->w_plan, &ix->w_host_q})
        a->release();
    if (!ix->raw) {  // an index without a single row still answers (empty results)
        const int raw_metric = ix->metric == B200_METRIC_L2 ? B200_METRIC_L2 : B200_METRIC_IP;
        if (!ix->use_ivf) B200_TRY(b200_corpus_create(raw_metric, B200_DTYPE_F32, ix->d, 0, &ix->raw));
    }
    ix->built = true;
    return B200_OK;
}

extern "C" int b200_index_finalize(b200_index *ix) {
    if (!ix) return fail(B200_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    return finalize_locked(ix);
}

// VIWithColumnInPart::buildIndex -> Search::VectorIndex::build (VIWithDataPart.cpp:131): one-shot build from host rows =
// reserve + train on a strided sample (<= 256 rows per list, the reference's train block) + add in chunks + finalize
extern "C" int b200_index_build(b200_index *ix, const float *rows, int64_t n) {
    if (!ix || (!rows && n > 0) || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    {
        std::lock_guard<std::mutex> lk(ix->mu);
        if (ix->built || ix->trained) return fail(B200_ERR_INVALID, "index already built");
        ix->reserved = n;
    }
    int nl = ix->nlist;
    if (ix->type != IDX_FLAT && nl <= 0) nl = (int)std::max<int64_t>(1, std::min<int64_t>(65536, (int64_t)(4.0 * sqrt((double)std::max<int64_t>(n, 1)))));
    const int64_t ns = std::min<int64_t>(n, std::max<int64_t>(256ll * std::max(nl, 1), 65536));
    if (ns == n || ix->type == IDX_FLAT) {
        B200_TRY(b200_index_train(ix, rows, ix->type == IDX_FLAT ? 0 : n));
    } else {
        std::vector<float> sample((size_t)ns * ix->d);
        for (int64_t i = 0; i < ns; i++) {
            const int64_t r = (int64_t)((double)i * (double)n / (double)ns);
            memcpy(sample.data() + i * ix->d, rows + r * ix->d, (size_t)ix->d * 4);
        }
        B200_TRY(b200_index_train(ix, sample.data(), ns));
    }
    B200_TRY(b200_index_add(ix, rows, n));
    return b200_index_finalize(ix);
}

extern "C" int b200_index_memory_bytes(const b200_index *ix, uint64_t *out_bytes) {
    if (!ix || !out_bytes) return fail(B200_ERR_INVALID, "bad arguments");
    uint64_t b = 0, t = 0;
    if (ix->raw && b200_corpus_memory_bytes(ix->raw, &t) == B200_OK) b += t;
    if (ix->coarse && b200_corpus_memory_bytes(ix->coarse, &t) == B200_OK) b += t;
    if (ix->use_ivf) {
        const uint64_t rows = (uint64_t)ix->pool_pages * kPageRows;
        b += (uint64_t)ix->nlist * ix->d * 4 + rows * (payload_row_bytes(ix) + 4 + (ix->d_row_bias ? 4 : 0)) + (uint64_t)ix->pool_pages * 12;
        if (ix->d_pq) b += (uint64_t)ix->m * 256 * ix->dsub * 6;
    }
    *out_bytes = b;
    return B200_OK;
}

extern "C" int b200_index_info(const b200_index *ix, int64_t *n, int *nlist, int *m, int *uses_ivf) {
    if (!ix) return fail(B200_ERR_INVALID, "null index");
    if (n) *n = ix->n;
    if (nlist) *nlist = ix->nlist;
    if (m) *m = ix->m;
    if (uses_ivf) *uses_ivf = ix->use_ivf ? 1 : 0;
    return B200_OK;
}

extern "C" int b200_index_last_scan(b200_index *ix, int64_t *rows_streamed, int64_t *payload_row_bytes_out, int64_t *work_items,
                                    double *kernel_ms_total, int64_t *kernel_launches, int reset) {
    if (!ix) return fail(B200_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lk(ix->mu);
    cudaSetDevice(ix->device);
    timing_collect(ix);
    if (ix->d_flag && ix->use_ivf) {
        unsigned long long r = 0;
        if (cudaMemcpy(&r, ix->d_flag + 4, 8, cudaMemcpyDeviceToHost) == cudaSuccess) ix->last_scan_rows = (int64_t)r;
    }
    if (rows_streamed) *rows_streamed = ix->last_scan_rows;
    if (payload_row_bytes_out) *payload_row_bytes_out = (int64_t)payload_row_bytes(ix);
    if (work_items) *work_items = ix->last_items;
    if (kernel_ms_total) *kernel_ms_total = ix->timed_ms;
    if (kernel_launches) *kernel_launches = ix->timed_launches;
    if (reset) {
        ix->timed_ms = 0;
        ix->timed_launches = 0;
    }
    return B200_OK;
}

// milliseconds of the phases of the LAST list search (timing enabled): coarse probe | pair sort + plan + query gather |
// grouped scan | per-query merge | exact second stage
extern "C" int b200_index_phase_ms(b200_index *ix, double out_ms[5]) {
    if (!ix || !out_ms) return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(ix->mu);
    cudaSetDevice(ix->device);
    timing_collect(ix);
    for (int i = 0; i < 5; i++) out_ms[i] = ix->phase_ms[i];
    return B200_OK;
}

// rows per inverted list (diagnostics: balance of the coarse quantiser); out_sizes[nlist]
extern "C" int b200_index_list_sizes(const b200_index *ix, uint32_t *out_sizes, int capacity) {
    if (!ix || !out_sizes) return fail(B200_ERR_INVALID, "bad arguments");
    if (!ix->built || !ix->use_ivf) return fail(B200_ERR_INVALID, "no inverted lists (FLAT index or not finalized)");
    if (capacity < ix->nlist) return fail(B200_ERR_INVALID, "buffer too small");
    memcpy(out_sizes, ix->list_len.data(), (size_t)ix->nlist * 4);
    return B200_OK;
}

extern "C" int b200_index_enable_timing(b200_index *ix, int on) {
    if (!ix) return fail(B200_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->timing = on != 0;
    if (on && !ix->ev0) {
        cudaEventCreate(&ix->ev0);
        cudaEventCreate(&ix->ev1);
        for (auto &e : ix->ev_ph) cudaEventCreate(&e);
    }
    return B200_OK;
}

// computeTopDistanceSubset (VIWithDataPart.cpp:838-856): exact distances of a candidate id set -> top-k
static int refine_device(b200_index *ix, const float *d_q /*[nq][d_pad] prepared*/, int64_t nq, const int64_t *d_cand, int ncand,
                         int k, int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, cudaStream_t s) {
    RefineParams rp{};
    rp.queries = d_q;
    rp.rows = reinterpret_cast<const float *>(corpus_device_rows(ix->raw));
    rp.cand = d_cand;
    rp.out_dis = d_out_dis;
    rp.out_ids = d_out_ids;
    rp.n = ix->n;
    rp.d_pad = ix->d_pad;
    rp.ncand = ncand;
    rp.k = k;
    rp.l2 = ix->metric == B200_METRIC_L2;
    rp.cosine = ix->metric == B200_METRIC_COSINE;
    rp.id_offset = id_offset;
    const size_t smem = (size_t)ix->d_pad * 4 + (size_t)9 * k * 8;
    B200_CUDA_OK(cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    refine_kernel<<<(unsigned)nq, 256, smem, s>>>(rp);
    g_launches++;
    B200_CUDA_OK(cudaGetLastError());
    return B200_OK;
}

// d_queries_raw: device fp32 [nq][d] -> ix->w_q [nq][d_pad] (cosine: unit length)
static int prepare_queries_device(b200_index *ix, const float *d_queries_raw, int64_t nq, cudaStream_t s) {
    B200_TRY(ix->w_q.reserve((size_t)nq * ix->d_pad * 4));
    B200_CUDA_OK(launch_pad_rows_f32(d_queries_raw, ix->d, ix->w_q.as<float>(), ix->d_pad, nq, s));
    if (ix->metric == B200_METRIC_COSINE) B200_CUDA_OK(launch_normalize_rows_f32(ix->w_q.as<float>(), ix->d_pad, nq, s));
    return B200_OK;
}

__global__ void cosine_finish_kernel(const float *q, int64_t nq, int d, int d_pad, int k, float *dis, const int64_t *ids) {
    // raw rows are unit vectors searched under IP with the prepared (unit) queries: distance = 1 - ip
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * k) return;
    dis[i] = ids[i] >= 0 ? 1.f - dis[i] : FLT_MAX;
}

// The whole search on the device, asynchronous on s.  d_queries: fp32 [nq][d]; outputs [nq][k].
static int search_device_locked(b200_index *ix, const float *d_queries, int64_t nq, int k, const char *params, int first_stage_only,
                                const uint8_t *d_alive, int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, int64_t *out_num_candidates,
                                cudaStream_t s) {
    if (out_num_candidates) *out_num_candidates = k;
    if (nq == 0) return B200_OK;
    B200_TRY(prepare_queries_device(ix, d_queries, nq, s));
    const float *d_q = ix->w_q.as<float>();
    const int force_exact = parse_int_param(params, "exact_batch", 0);
    if (!ix->use_ivf || force_exact == 1) {
        if (!ix->raw) return fail(B200_ERR_INVALID, "exact search needs the fp32 rows (keep_raw=0 index)");
        // FLAT / fallback-to-flat: exact scan of the raw rows.  The raw corpus wants [nq][d] rows: strip the padding again.
        B200_TRY(ix->w_qraw.reserve((size_t)nq * ix->d * 4));
        if (ix->d == ix->d_pad) B200_CUDA_OK(cudaMemcpyAsync(ix->w_qraw.p, d_q, (size_t)nq * ix->d * 4, cudaMemcpyDeviceToDevice, s));
        else B200_CUDA_OK(cudaMemcpy2DAsync(ix->w_qraw.p, (size_t)ix->d * 4, d_q, (size_t)ix->d_pad * 4, (size_t)ix->d * 4, nq, cudaMemcpyDeviceToDevice, s));
        B200_TRY(b200_corpus_search_device(ix->raw, ix->w_qraw.as<float>(), nq, k, d_alive, id_offset, d_out_dis, d_out_ids, s));
        if (ix->metric == B200_METRIC_COSINE) {
            cosine_finish_kernel<<<(unsigned)ceil_div(nq * k, 256), 256, 0, s>>>(d_q, nq, ix->d, ix->d_pad, k, d_out_dis, d_out_ids);
            g_launches++;
        }
        return B200_OK;
    }
    if (k > 1024) return fail(B200_ERR_UNSUPPORTED, "k > 1024 on IVF indexes");
    const int nl = ix->nlist;
    int nprobe = parse_int_param(params, "nprobe", ix->default_nprobe);
    nprobe = std::max(1, std::min(nprobe, nl));
    const int refine_factor = std::max(1, parse_int_param(params, "refine_factor", parse_int_param(params, "reorder_k_factor", ix->refine_factor)));
    const bool two_stage = ix->raw && refine_factor > 1 && !first_stage_only;
    const int k1 = two_stage ? std::min(1024, k * refine_factor) : k;
    if (out_num_candidates) *out_num_candidates = k1;
    const int64_t n_pairs = nq * nprobe;
    if (n_pairs >= (int64_t)1 << 31) return fail(B200_ERR_UNSUPPORTED, "nq * nprobe must stay below 2^31");

    if (ix->timing) cudaEventRecord(ix->ev_ph[0], s);
    // ---- coarse probe: nprobe nearest centroids per query (exact FLAT search of the centroid table)
    B200_TRY(ix->w_qraw.reserve((size_t)nq * ix->d * 4));
    B200_TRY(ix->w_probe.reserve((size_t)n_pairs * 8));
    B200_TRY(ix->w_pd.reserve((size_t)n_pairs * 4));
    if (ix->d == ix->d_pad) B200_CUDA_OK(cudaMemcpyAsync(ix->w_qraw.p, d_q, (size_t)nq * ix->d * 4, cudaMemcpyDeviceToDevice, s));
    else B200_CUDA_OK(cudaMemcpy2DAsync(ix->w_qraw.p, (size_t)ix->d * 4, d_q, (size_t)ix->d_pad * 4, (size_t)ix->d * 4, nq, cudaMemcpyDeviceToDevice, s));
    {
        // The centroid table is small and nprobe is a large k for it: the tensor-core path keeps one k-list per query lane
        // and never gets a selective threshold when k / nlist is a few percent (measured 9 ms for 10 000 x 4096 x 96, k = 32).
        // The scan kernel's warp lists cost O(k / 32) per insert: use it when its estimated time (FMA-bound at ~1.4 TB/s of
        // table bytes per 8-query pass) undercuts ~1 us per (query, 32 probes) of the tensor-core path.
        const double t_scan = (double)ceil_div(nq, 8) * nl * ix->d_pad * 4.0 / 1.4e12;
        const double t_gemm = 1e-6 * (double)nq * std::max(1.0, nprobe / 32.0) + 30e-6;
        bool use_scan = nprobe > 8 && t_scan < t_gemm;
        // nprobe > 8: full ranking keys + warp select (coarse_scores_kernel / coarse_select_kernel above)
        int coarse_path = nprobe > 8 && nprobe <= 1024 ? 3 : use_scan ? 1 : 2;
        if (const int forced = parse_int_param(params, "coarse_path", 0)) coarse_path = forced;   // A/B: 1 scan kernel, 2 tensor-core path, 3 select
        if (coarse_path == 3) {
            const int64_t chunk = std::max<int64_t>(64, std::min<int64_t>(nq, ((int64_t)64 << 20) / std::max(1, nl)));   // <= 256 MB of keys
            B200_TRY(ix->w_cs.reserve((size_t)chunk * nl * 4));
            const size_t sel_smem = (size_t)8 * nprobe * 8;
            if (sel_smem > 48 * 1024) B200_CUDA_OK(cudaFuncSetAttribute(coarse_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
            for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
                const int64_t nqc = std::min(chunk, nq - q0);
                coarse_scores_kernel<<<dim3((unsigned)ceil_div(nl, kCoarseTile), (unsigned)ceil_div(nqc, kCoarseTile)), 256, 0, s>>>(
                    d_q + q0 * ix->d_pad, ix->d_pad, ix->d_centroids, ix->d_cnorm, nqc, nl, ix->d, ix->w_cs.as<float>());
                coarse_select_kernel<<<(unsigned)ceil_div(nqc, 8), 256, sel_smem, s>>>(ix->w_cs.as<float>(), nqc, nl, nprobe, ix->w_pd.as<float>() + q0 * nprobe,
                                                                                       ix->w_probe.as<int64_t>() + q0 * nprobe);
                g_launches += 2;
            }
            B200_CUDA_OK(cudaGetLastError());
        } else {
            b200_corpus_set_path(ix->coarse, coarse_path == 1 ? 1 : 0);
            const int rc = b200_corpus_search_device(ix->coarse, ix->w_qraw.as<float>(), nq, nprobe, nullptr, 0, ix->w_pd.as<float>(), ix->w_probe.as<int64_t>(), s);
            b200_corpus_set_path(ix->coarse, 0);
            B200_TRY(rc);
        }
    }

    if (ix->timing) cudaEventRecord(ix->ev_ph[1], s);
    // ---- pairs sorted by list
    B200_TRY(ix->w_u32a.reserve((size_t)n_pairs * 4));
    B200_TRY(ix->w_u32b.reserve((size_t)n_pairs * 4));
    B200_TRY(ix->w_u32c.reserve((size_t)n_pairs * 4));
    B200_TRY(ix->w_u32d.reserve((size_t)n_pairs * 4));
    B200_TRY(ix->w_cnt.reserve((size_t)nl * 4));
    B200_CUDA_OK(cudaMemsetAsync(ix->w_cnt.p, 0, (size_t)nl * 4, s));
    pairs_make_kernel<<<(unsigned)ceil_div(n_pairs, 256), 256, 0, s>>>(ix->w_probe.as<int64_t>(), n_pairs, nl, ix->w_u32a.as<uint32_t>(),
                                                                       ix->w_u32b.as<uint32_t>(), ix->w_cnt.as<uint32_t>());
    g_launches++;
    {
        int bits = 1;
        while ((1 << bits) < nl + 1) bits++;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, ix->w_u32a.as<uint32_t>(), ix->w_u32c.as<uint32_t>(), ix->w_u32b.as<uint32_t>(),
                                        ix->w_u32d.as<uint32_t>(), (int)n_pairs, 0, bits, s);
        B200_TRY(ix->w_sort.reserve(tb + 256));
        cub::DeviceRadixSort::SortPairs(ix->w_sort.p, tb, ix->w_u32a.as<uint32_t>(), ix->w_u32c.as<uint32_t>(), ix->w_u32b.as<uint32_t>(),
                                        ix->w_u32d.as<uint32_t>(), (int)n_pairs, 0, bits, s);
        g_launches++;
    }
    // ---- work items.  Long lists are cut into chunks of pages so that even a single query fills the SMs; the cut is
    //      chosen from host-side knowledge only (no device -> host round trip on the query path).
    // Lists are cut into chunks of `ppc` pages: an item never streams more than 16 pages (bounds the tail of the static
    // round-robin schedule), and small batches are split further so that every SM gets ~4 items.  Finer is NOT better: every item
    // pays one cold start of its top-k lists (sweep at 100 M x 768, nprobe 1: 4-page items 0.64 of the HBM peak, 8 to 16-page
    // items 0.77-0.78; nprobe 4: 0.33 vs 0.45; profiles/r02_gpu17_*).  The estimate uses host-side knowledge only (no device ->
    // host round trip on the query path): probed lists <= min(pairs, nlist), their length size-biased.
    const double avg_pages = std::max(1.0, (double)ix->pages_used / std::max(1, nl));
    const double est_lists = std::min<double>((double)n_pairs, (double)nl);
    const double est_pages = est_lists * std::min<double>(ix->max_list_pages ? ix->max_list_pages : 1, 1.5 * avg_pages);
    uint32_t ppc;
    {
        const double want_items = 4.0 * ix->sms;
        // Lists probed by more than 16 queries run on per-lane top-k lists, whose cold start is paid per item: longer items pay
        // (cfg-4 shape, ~78 queries per list: 21.5 / 18.6 / 18.5 / 16.5 ms at 8 / 16 / 32 / 48 pages per item, profiles/r02_gpu32.log);
        // cooperative items (<= 16 queries) keep the 16-page cap (sweep at 100 M x 768 above).
        const double q_per_list = (double)n_pairs / std::max(1.0, est_lists);
        const double cap = q_per_list > 16.0 ? 48.0 : 16.0;
        ppc = (uint32_t)std::min(cap, std::max(8.0, std::ceil(est_pages / want_items)));
        if (est_lists * 2 < want_items) ppc = (uint32_t)std::max(2.0, std::min<double>(ppc, std::ceil(1.5 * avg_pages * est_lists / want_items)));   // a handful of queries
        ppc = std::max<uint32_t>(ppc, (ix->max_list_pages + 63) / 64);   // at most 64 chunks per list
        ppc = std::max<uint32_t>(ppc, 1);
        if (const int forced = parse_int_param(params, "pages_per_chunk", 0)) ppc = (uint32_t)forced;
    }
    const uint32_t max_chunks = (ix->max_list_pages + ppc - 1) / ppc;
    const int64_t max_items = std::min<int64_t>((int64_t)n_pairs * max_chunks, (int64_t)(ceil_div(n_pairs, 128) + nl) * max_chunks);
    const int64_t max_parts = n_pairs * max_chunks;
    if (max_parts * k1 >= (int64_t)1 << 32) return fail(B200_ERR_UNSUPPORTED, "nq * nprobe * chunks * k too large for one batch; split the batch");
    B200_TRY(ix->w_items.reserve((size_t)max_items * sizeof(IvfGemmItem) + 64));
    B200_TRY(ix->w_plan.reserve((size_t)nl * 4 * 3));
    uint32_t *pair_start = ix->w_plan.as<uint32_t>(), *part_off = pair_start + nl, *n_chunks = part_off + nl;
    int *d_counts = ix->d_flag + 1;   // n_items, n_parts
    SearchPlan pl{};
    pl.cnt = ix->w_cnt.as<uint32_t>();
    pl.list_len = ix->d_list_len;
    pl.list_page_off = ix->d_list_page_off;
    pl.list_order = ix->d_list_order;
    pl.pair_start = pair_start;
    pl.part_off = part_off;
    pl.n_chunks = n_chunks;
    pl.items = ix->w_items.as<IvfGemmItem>();
    pl.n_items = d_counts;
    pl.n_parts = d_counts + 1;
    pl.scan_rows = reinterpret_cast<unsigned long long *>(ix->d_flag + 4);
    pl.nlist = nl;
    pl.max_items = (int)std::min<int64_t>(max_items, INT32_MAX);
    pl.pages_per_chunk = ppc;
    search_plan_kernel<<<1, 1024, 0, s>>>(pl);
    g_launches++;
    // ---- gather queries, per-pair bookkeeping
    B200_TRY(ix->w_qbuf.reserve(((size_t)n_pairs + 128) * ix->d_pad64 * 2));
    // the 128 rows behind the last pair are read by the last items' A tiles (TMEM lanes without a query): keep them finite
    B200_CUDA_OK(cudaMemsetAsync(ix->w_qbuf.as<char>() + (size_t)n_pairs * ix->d_pad64 * 2, 0, (size_t)128 * ix->d_pad64 * 2, s));
    B200_TRY(ix->w_inv.reserve((size_t)n_pairs * 4));
    B200_TRY(ix->w_ppb.reserve((size_t)n_pairs * 4));
    B200_TRY(ix->w_pconst.reserve((size_t)n_pairs * 4));
    B200_TRY(ix->w_qconst.reserve((size_t)nq * 4));
    PairFill pf{};
    pf.sorted_list = ix->w_u32c.as<uint32_t>();
    pf.sorted_pair = ix->w_u32d.as<uint32_t>();
    pf.pair_start = pair_start;
    pf.part_off = part_off;
    pf.n_chunks = n_chunks;
    pf.queries = d_q;
    pf.sq_step = ix->payload == IVF_PRODUCER_SQ8 ? ix->d_sq + ix->d : nullptr;
    pf.centroids = ix->payload == IVF_PRODUCER_PQ ? ix->d_centroids : nullptr;
    pf.qbuf = ix->w_qbuf.as<__nv_bfloat16>();
    pf.inv = ix->w_inv.as<uint32_t>();
    pf.pair_part_base = ix->w_ppb.as<uint32_t>();
    pf.pair_const = ix->w_pconst.as<float>();
    pf.n_pairs = n_pairs;
    pf.nprobe = nprobe;
    pf.nlist = nl;
    pf.d = ix->d;
    pf.d_pad = ix->d_pad;
    pf.d_pad64 = ix->d_pad64;
    pf.l2 = ix->metric == B200_METRIC_L2;
    pair_fill_kernel<<<(unsigned)ceil_div(n_pairs * 32, 256), 256, 0, s>>>(pf);
    query_const_kernel<<<(unsigned)ceil_div(nq * 32, 256), 256, 0, s>>>(d_q, nq, ix->d, ix->d_pad, ix->payload == IVF_PRODUCER_SQ8 ? ix->d_sq + 3 * ix->d : nullptr,
                                                                        ix->metric == B200_METRIC_L2, ix->payload == IVF_PRODUCER_TMA, ix->w_qconst.as<float>());
    g_launches += 2;
    // ---- the grouped tensor-core scan
    B200_TRY(ix->w_pk.reserve((size_t)max_parts * k1 * 4));
    B200_TRY(ix->w_pi.reserve((size_t)max_parts * k1 * 4));
    B200_TRY(ix->w_pw.reserve((size_t)max_parts * 4));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ix->sms, max_items));
    IvfGemmParams gp{};
    gp.items = ix->w_items.as<IvfGemmItem>();
    gp.n_items_ptr = d_counts;
    gp.list_pages = ix->d_list_pages;
    gp.row_bias = ix->d_row_bias;
    gp.row_ids = ix->d_row_ids;
    gp.alive = d_alive;
    gp.pair_part_base = ix->w_ppb.as<uint32_t>();
    gp.part_keys = ix->w_pk.as<float>();
    gp.part_ids = ix->w_pi.as<uint32_t>();
    gp.part_worst = ix->w_pw.as<float>();
    {   // shared per-query bound across the items of this launch (ivf_gemm.h); nothing to share with one list per query
        static const bool bound_on = !(getenv("B200_IVF_BOUND") && atoi(getenv("B200_IVF_BOUND")) == 0);
        if (bound_on && nprobe > 1 && parse_int_param(params, "shared_bound", 1) != 0) {   // shared_bound=0: A/B switch
            B200_TRY(ix->w_qb.reserve((size_t)nq * 4));
            B200_CUDA_OK(cudaMemsetAsync(ix->w_qb.p, 0xff, (size_t)nq * 4, s));
            gp.query_bound = ix->w_qb.as<uint32_t>();
            gp.sorted_pair = ix->w_u32d.as<uint32_t>();
            gp.pair_const = ix->w_pconst.as<float>();
            gp.nprobe = nprobe;
        }
    }
    gp.scale_const = ix->metric == B200_METRIC_L2 ? -2.f : -1.f;
    gp.d_pad = ix->d_pad64;
    gp.k = k1;
    gp.producer = ix->payload;
    gp.codes = reinterpret_cast<const uint8_t *>(ix->d_pool);
    gp.codebook_bf16 = ix->d_pq_bf16;
    gp.code_bytes = ix->code_bytes;
    gp.m = ix->m;
    gp.dsub = ix->dsub;
    gp.codebook_bytes = ix->payload == IVF_PRODUCER_PQ ? ix->m * 256 * ix->dsub * 2 : 0;
    if (k1 > kGemmSmemK || true) {  // global scratch for lists that do not fit in shared memory (the launcher decides)
        B200_TRY(ix->w_lk.reserve((size_t)grid * 128 * list_cap_for(k1) * 4));
        B200_TRY(ix->w_li.reserve((size_t)grid * 128 * list_cap_for(k1) * 4));
        gp.list_keys_gmem = ix->w_lk.as<float>();
        gp.list_ids_gmem = ix->w_li.as<uint32_t>();
    }
    const char *detail = nullptr;
    if (ix->timing) {
        cudaEventRecord(ix->ev_ph[2], s);
        cudaEventRecord(ix->ev0, s);
    }
    cudaError_t e = launch_ivf_gemm_topk(gp, ix->w_qbuf.p, n_pairs + 128, ix->d_pool, (int64_t)ix->pool_pages * kPageRows, grid, s, &detail);
    if (ix->timing) {
        cudaEventRecord(ix->ev1, s);
        cudaEventRecord(ix->ev_ph[3], s);
        ix->timed_pending = true;
    }
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, std::string("ivf_gemm_topk launch: ") + (detail ? detail : cudaGetErrorString(e)));
    ix->last_items = max_items;
    // ---- per-query merge of the partial lists
    float *m_dis = d_out_dis;
    int64_t *m_ids = d_out_ids;
    if (two_stage) {
        B200_TRY(ix->w_od.reserve((size_t)nq * k1 * 4));
        B200_TRY(ix->w_oi.reserve((size_t)nq * k1 * 8));
        m_dis = ix->w_od.as<float>();
        m_ids = ix->w_oi.as<int64_t>();
    }
    IvfMerge mg{};
    mg.inv = ix->w_inv.as<uint32_t>();
    mg.pair_part_base = ix->w_ppb.as<uint32_t>();
    mg.sorted_list = ix->w_u32c.as<uint32_t>();
    mg.n_chunks = n_chunks;
    mg.pair_const = ix->w_pconst.as<float>();
    mg.query_const = ix->w_qconst.as<float>();
    mg.part_keys = gp.part_keys;
    mg.part_worst = gp.part_worst;
    mg.part_ids = gp.part_ids;
    mg.out_dis = m_dis;
    mg.out_ids = m_ids;
    mg.id_offset = two_stage ? 0 : id_offset;
    mg.nprobe = nprobe;
    mg.nlist = nl;
    mg.k_part = k1;
    mg.k = k1;
    mg.metric = ix->metric;
    {
        const size_t smem = (size_t)9 * k1 * 8;
        B200_CUDA_OK(cudaFuncSetAttribute(ivf_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ivf_merge_kernel<<<(unsigned)nq, 256, smem, s>>>(mg);
        g_launches++;
        B200_CUDA_OK(cudaGetLastError());
    }
    if (ix->timing) cudaEventRecord(ix->ev_ph[4], s);
    if (two_stage) B200_TRY(refine_device(ix, d_q, nq, m_ids, k1, k, id_offset, d_out_dis, d_out_ids, s));
    if (ix->timing) cudaEventRecord(ix->ev_ph[5], s);
    return B200_OK;
}

static void timing_collect(b200_index *ix) {
    if (!ix->timing || !ix->timed_pending) return;
    float ms = 0;
    if (cudaEventSynchronize(ix->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, ix->ev0, ix->ev1) == cudaSuccess) {
        ix->timed_ms += ms;
        ix->timed_launches++;
    }
    if (cudaEventSynchronize(ix->ev_ph[5]) == cudaSuccess)
        for (int i = 0; i < 5; i++) {
            float pm = 0;
            if (cudaEventElapsedTime(&pm, ix->ev_ph[i], ix->ev_ph[i + 1]) == cudaSuccess) ix->phase_ms[i] = pm;
        }
    ix->timed_pending = false;
}

// Search::VectorIndex::search with device buffers, asynchronous on `stream` (NULL: the index's stream, synchronised)
extern "C" int b200_index_search_device(b200_index *ix, const float *d_queries, int64_t nq, int k, const char *params, int first_stage_only,
                                        const uint8_t *d_alive_bits, int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, void *stream) {
    if (!ix || (!d_queries && nq > 0) || !d_out_dis || !d_out_ids || nq < 0 || k <= 0) return fail(B200_ERR_INVALID, "bad arguments");
    if (!ix->built) return fail(B200_ERR_INVALID, "index not built");
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    timing_collect(ix);
    cudaStream_t s = stream ? reinterpret_cast<cudaStream_t>(stream) : ix->stream;
    B200_TRY(search_device_locked(ix, d_queries, nq, k, params, first_stage_only, d_alive_bits, id_offset, d_out_dis, d_out_ids, nullptr, s));
    if (!stream) B200_CUDA_OK(cudaStreamSynchronize(s));
    return B200_OK;
}

// Search::VectorIndex::search(queries, k, params, first_stage_only, filter) (VIWithDataPart.cpp:926), host buffers
extern "C" int b200_index_search(b200_index *ix, const float *queries, int64_t nq, int k, const char *params, int first_stage_only,
                                 const uint8_t *alive_bits, float *out_dis, int64_t *out_ids, int64_t *out_num_candidates) {
    if (!ix || (!queries && nq > 0) || !out_dis || !out_ids || nq < 0 || k <= 0) return fail(B200_ERR_INVALID, "bad arguments");
    if (!ix->built) return fail(B200_ERR_INVALID, "index not built");
    if (out_num_candidates) *out_num_candidates = k;
    if (nq == 0) return B200_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    timing_collect(ix);
    cudaStream_t s = ix->stream;
    B200_TRY(ix->w_host_q.reserve((size_t)nq * ix->d * 4));
    B200_TRY(ix->w_cand.reserve((size_t)nq * k * 12 + 16));
    B200_CUDA_OK(cudaMemcpyAsync(ix->w_host_q.p, queries, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, s));
    const uint8_t *d_alive = nullptr;
    if (alive_bits) {
        const size_t ab = (size_t)ceil_div(ix->n, 8);
        B200_TRY(ix->w_alive.reserve(ab + 16));
        B200_CUDA_OK(cudaMemcpyAsync(ix->w_alive.p, alive_bits, ab, cudaMemcpyHostToDevice, s));
        d_alive = ix->w_alive.as<uint8_t>();
    }
    float *r_d = ix->w_cand.as<float>();
    int64_t *r_i = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(ix->w_cand.p) + (size_t)round_up(nq * k * 4, 8));
    B200_TRY(search_device_locked(ix, ix->w_host_q.as<float>(), nq, k, params, first_stage_only, d_alive, 0, r_d, r_i, out_num_candidates, s));
    B200_CUDA_OK(cudaMemcpyAsync(out_dis, r_d, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaMemcpyAsync(out_ids, r_i, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    timing_collect(ix);
    return B200_OK;
}

// computeTopDistanceSubset: queries [nq][d], candidates [nq][ncand] (negative = unused) -> exact top-k
extern "C" int b200_index_refine(b200_index *ix, const float *queries, int64_t nq, const int64_t *cand_ids, int64_t ncand, int k,
                                 float *out_dis, int64_t *out_ids) {
    if (!ix || !queries || !cand_ids || !out_dis || !out_ids || nq < 0 || ncand <= 0 || k <= 0)
        return fail(B200_ERR_INVALID, "bad arguments");
    if (!ix->built) return fail(B200_ERR_INVALID, "index not built");
    if (!ix->raw) return fail(B200_ERR_INVALID, "this index keeps no fp32 rows (keep_raw=0): no second stage");
    if (k > 1024) return fail(B200_ERR_UNSUPPORTED, "k > 1024 in refine");
    if (nq == 0) return B200_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    B200_TRY(ix->w_host_q.reserve((size_t)nq * ix->d * 4));
    B200_CUDA_OK(cudaMemcpyAsync(ix->w_host_q.p, queries, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, s));
    B200_TRY(prepare_queries_device(ix, ix->w_host_q.as<float>(), nq, s));
    B200_TRY(ix->w_probe.reserve((size_t)nq * ncand * 8));
    B200_TRY(ix->w_od.reserve((size_t)nq * k * 4));
    B200_TRY(ix->w_oi.reserve((size_t)nq * k * 8));
    B200_CUDA_OK(cudaMemcpyAsync(ix->w_probe.p, cand_ids, (size_t)nq * ncand * 8, cudaMemcpyHostToDevice, s));
    B200_TRY(refine_device(ix, ix->w_q.as<float>(), nq, ix->w_probe.as<int64_t>(), (int)ncand, k, 0, ix->w_od.as<float>(), ix->w_oi.as<int64_t>(), s));
    B200_CUDA_OK(cudaMemcpyAsync(out_dis, ix->w_od.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaMemcpyAsync(out_ids, ix->w_oi.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    return B200_OK;
}

// ------------------------------------------------------------------------------------
// persistence: VIWithColumnInPart::serialize / load (reference: src/VectorIndex/Common/VIWithDataPart.cpp:451-525,
// :578-764) write `<idx>-*.vidx3` through Search::IndexDataFileWriter; the on-disk format of the closed library
// is not reproducible, so this is our own single-file layout ("B2IX" v2): header, fp32 rows (when kept), then the
// quantisers and the pages in list order.  Loading re-uploads to HBM (pages become consecutive) and validates the header.
// ------------------------------------------------------------------------------------
namespace {
struct IxHeader {
    char magic[4];
    uint32_t version;
    int32_t type, metric, d, nlist, m, dsub, default_nprobe, refine_factor, payload, has_raw, use_ivf, code_bytes;
    int64_t n;
    uint32_t pages_used, reserved0;
};
// the byte stream is the caller's: a FILE (b200_index_save / _load) or the host's own stream objects
// (b200_index_save_cb / _load_cb: Search::IndexDataFileWriter / Reader over ClickHouse disks, VectorIndexIO.h:33-164)
struct Io {
    int (*write)(void *, const void *, size_t) = nullptr;
    int (*read)(void *, void *, size_t) = nullptr;
    void *ctx = nullptr;
};
bool wr(Io *f, const void *p, size_t bytes) { return bytes == 0 || (f->write && f->write(f->ctx, p, bytes) == 0); }
bool rd(Io *f, void *p, size_t bytes) { return bytes == 0 || (f->read && f->read(f->ctx, p, bytes) == 0); }
int file_write(void *ctx, const void *p, size_t bytes) { return fwrite(p, 1, bytes, reinterpret_cast<FILE *>(ctx)) == bytes ? 0 : 1; }
int file_read(void *ctx, void *p, size_t bytes) { return fread(p, 1, bytes, reinterpret_cast<FILE *>(ctx)) == bytes ? 0 : 1; }
}  // namespace

static int index_save_io(b200_index *ix, Io *f) {
    if (!ix->built) return fail(B200_ERR_INVALID, "index not built");
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    IxHeader h{};
    memcpy(h.magic, "B2IX", 4);
    h.version = 2;
    h.type = ix->type; h.metric = ix->metric; h.d = ix->d; h.nlist = ix->nlist; h.m = ix->m; h.dsub = ix->dsub;
    h.default_nprobe = ix->default_nprobe; h.refine_factor = ix->refine_factor; h.payload = ix->payload; h.has_raw = ix->raw ? 1 : 0;
    h.use_ivf = ix->use_ivf ? 1 : 0; h.code_bytes = ix->code_bytes; h.n = ix->n; h.pages_used = ix->pages_used;
    bool ok = wr(f, &h, sizeof(h));
    try {
        if (ok && ix->raw) {  // fp32 rows, unpadded (cosine indexes hold unit vectors; they are written as stored)
            const int64_t chunk = std::max<int64_t>(1, (64ll << 20) / ((int64_t)ix->d_pad * 4));
            std::vector<float> buf((size_t)chunk * ix->d_pad);
            const float *rows = reinterpret_cast<const float *>(corpus_device_rows(ix->raw));
            for (int64_t off = 0; ok && off < ix->n; off += chunk) {
                const int64_t mrows = std::min(chunk, ix->n - off);
                if (cudaMemcpy(buf.data(), rows + off * ix->d_pad, (size_t)mrows * ix->d_pad * 4, cudaMemcpyDeviceToHost) != cudaSuccess) ok = false;
                for (int64_t r = 0; ok && r < mrows; r++) ok = wr(f, buf.data() + r * ix->d_pad, (size_t)ix->d * 4);
            }
        }
        if (ok && ix->use_ivf) {
            std::vector<char> tmp;
            auto dump = [&](const void *dptr, size_t bytes) {
                tmp.resize(bytes);
                if (bytes && cudaMemcpy(tmp.data(), dptr, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
                return wr(f, tmp.data(), bytes);
            };
            ok = dump(ix->d_centroids, (size_t)ix->nlist * ix->d * 4) && wr(f, ix->list_len.data(), (size_t)ix->nlist * 4);
            if (ok && ix->d_pq) ok = dump(ix->d_pq, (size_t)ix->m * 256 * ix->dsub * 4);
            if (ok && ix->d_sq) ok = dump(ix->d_sq, (size_t)4 * ix->d * 4);
            std::vector<uint32_t> pages(ix->pages_used);
            if (ok && ix->pages_used)
                ok = cudaMemcpy(pages.data(), ix->d_list_pages, (size_t)ix->pages_used * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
            const size_t pb = (size_t)kPageRows * payload_row_bytes(ix);
            for (uint32_t i = 0; ok && i < ix->pages_used; i++) {
                const size_t row0 = (size_t)pages[i] * kPageRows;
                ok = dump(reinterpret_cast<const char *>(ix->d_pool) + (size_t)pages[i] * pb, pb) && dump(ix->d_row_ids + row0, kPageRows * 4);
                if (ok && ix->d_row_bias) ok = dump(ix->d_row_bias + row0, kPageRows * 4);
            }
        }
    } catch (const std::bad_alloc &) {
        ok = false;
    }
    return ok ? B200_OK : fail(B200_ERR_INVALID, "index serialisation: write failed");
}

extern "C" int b200_index_save(b200_index *ix, const char *path) {
    if (!ix || !path) return fail(B200_ERR_INVALID, "bad arguments");
    FILE *fp = fopen(path, "wb");
    if (!fp) return fail(B200_ERR_INVALID, std::string("cannot open ") + path);
    Io io;
    io.write = file_write;
    io.ctx = fp;
    int rc = index_save_io(ix, &io);
    if (fclose(fp) != 0 && rc == B200_OK) rc = fail(B200_ERR_INVALID, std::string("write failed: ") + path);
    return rc;
}

extern "C" int b200_index_save_cb(b200_index *ix, int (*write)(void *ctx, const void *data, size_t bytes), void *ctx) {
    if (!ix || !write) return fail(B200_ERR_INVALID, "bad arguments");
    Io io;
    io.write = write;
    io.ctx = ctx;
    return index_save_io(ix, &io);
}

static int index_load_io(Io *f, b200_index **out) {
    *out = nullptr;
    IxHeader h{};
    if (!rd(f, &h, sizeof(h)) || memcmp(h.magic, "B2IX", 4) != 0 || h.version != 2)
        return fail(B200_ERR_INVALID, "not a B2IX v2 index file");
    // a truncated or corrupt file must fail here, not in a kernel: every size below is derived from these fields
    const bool sane = h.type >= 0 && h.type <= 8 && h.metric >= 0 && h.metric <= 2 && h.d > 0 && h.d <= (1 << 16) && h.n >= 0 &&
                      h.n < (int64_t)0xffffffffll && h.payload >= 0 && h.payload <= 2 &&
                      (!h.use_ivf || (h.nlist > 0 && h.nlist <= (1 << 24) && (uint64_t)h.pages_used <= (uint64_t)h.n / kPageRows + (uint64_t)h.nlist + 1)) &&
                      (h.payload != IVF_PRODUCER_PQ || !h.use_ivf || (h.m > 0 && h.dsub > 0 && h.m * h.dsub == h.d && h.code_bytes >= h.m && h.code_bytes % 16 == 0)) &&
                      (h.payload != IVF_PRODUCER_SQ8 || !h.use_ivf || (h.code_bytes >= h.d && h.code_bytes % 16 == 0)) && (h.has_raw || h.use_ivf);
    if (!sane) return fail(B200_ERR_INVALID, "corrupt index header");
    b200_index *ix = nullptr;
    int rc = b200_index_create(kTypeNames[h.type], h.metric, h.d, "", &ix);
    if (rc != B200_OK) return rc;
    auto bail = [&](const std::string &msg) {
        b200_index_free(ix);
        return fail(B200_ERR_INVALID, msg);
    };
    try {
        ix->nlist = h.nlist; ix->m = h.m; ix->dsub = h.dsub; ix->default_nprobe = h.default_nprobe; ix->refine_factor = h.refine_factor;
        ix->payload = h.payload; ix->use_ivf = h.use_ivf != 0; ix->code_bytes = h.code_bytes; ix->keep_raw = h.has_raw;
        const int raw_metric = h.metric == B200_METRIC_L2 ? B200_METRIC_L2 : B200_METRIC_IP;
        if (h.has_raw) {
            if (b200_corpus_create(raw_metric, B200_DTYPE_F32, h.d, h.n, &ix->raw) != B200_OK) return bail(b200_last_error());
            const int64_t chunk = std::max<int64_t>(1, (64ll << 20) / ((int64_t)h.d * 4));
            std::vector<float> buf((size_t)chunk * h.d);
            for (int64_t off = 0; off < h.n; off += chunk) {
                const int64_t mrows = std::min(chunk, h.n - off);
                if (!rd(f, buf.data(), (size_t)mrows * h.d * 4)) return bail("truncated index file (rows)");
                if (b200_corpus_append(ix->raw, buf.data(), mrows) != B200_OK) return bail(b200_last_error());
            }
        }
        ix->n = h.n;
        if (ix->use_ivf) {
            std::vector<char> tmp;
            auto slurp = [&](void **dptr, size_t bytes) {
                tmp.resize(bytes);
                if (!rd(f, tmp.data(), bytes)) return false;
                if (cudaMalloc(dptr, bytes + 256) != cudaSuccess) return false;
                return cudaMemcpy(*dptr, tmp.data(), bytes, cudaMemcpyHostToDevice) == cudaSuccess;
            };
            const int nl = h.nlist;
            ix->list_len.resize(nl);
            if (!slurp((void **)&ix->d_centroids, (size_t)nl * h.d * 4) || !rd(f, ix->list_len.data(), (size_t)nl * 4))
                return bail("truncated index file (quantiser)");
            uint64_t total = 0, pages = 0;
            std::vector<uint32_t> page_off(nl + 1, 0);
            for (int l = 0; l < nl; l++) {
                total += ix->list_len[l];
                page_off[l] = (uint32_t)pages;
                pages += (ix->list_len[l] + kPageRows - 1) / kPageRows;
                ix->max_list_pages = std::max<uint32_t>(ix->max_list_pages, (ix->list_len[l] + kPageRows - 1) / kPageRows);
            }
            page_off[nl] = (uint32_t)pages;
            if (total != (uint64_t)h.n || pages != h.pages_used) return bail("corrupt index file (list lengths do not add up)");
            if (h.payload == IVF_PRODUCER_PQ) {
                if (!slurp((void **)&ix->d_pq, (size_t)h.m * 256 * h.dsub * 4)) return bail("truncated index file (codebook)");
                if (cudaMalloc(&ix->d_pq_bf16, (size_t)h.m * 256 * h.dsub * 2) != cudaSuccess) return bail("cudaMalloc failed");
                if (launch_f32_to_bf16_rows(ix->d_pq, h.dsub, ix->d_pq_bf16, h.dsub, (int64_t)h.m * 256, ix->stream) != cudaSuccess)
                    return bail("codebook conversion failed");
            }
            if (h.payload == IVF_PRODUCER_SQ8 && !slurp((void **)&ix->d_sq, (size_t)4 * h.d * 4)) return bail("truncated index file (SQ ranges)");
            ix->pool_pages = ix->pages_used = h.pages_used;
            const size_t pb = (size_t)kPageRows * payload_row_bytes(ix), rows = (size_t)std::max<uint32_t>(h.pages_used, 1) * kPageRows;
            if (cudaMalloc(&ix->d_pool, rows * payload_row_bytes(ix) + 256) != cudaSuccess || cudaMalloc(&ix->d_row_ids, rows * 4) != cudaSuccess ||
                (h.metric == B200_METRIC_L2 && cudaMalloc(&ix->d_row_bias, rows * 4) != cudaSuccess))
                return bail("cudaMalloc of the page pool failed");
            std::vector<char> page(pb);
            std::vector<uint32_t> ids(kPageRows);
            std::vector<float> bias(kPageRows);
            uint32_t pg = 0;
            for (int l = 0; l < nl; l++) {
                const uint32_t np = (ix->list_len[l] + kPageRows - 1) / kPageRows;
                for (uint32_t t = 0; t < np; t++, pg++) {
                    if (!rd(f, page.data(), pb) || !rd(f, ids.data(), kPageRows * 4) || (ix->d_row_bias && !rd(f, bias.data(), kPageRows * 4)))
                        return bail("truncated index file (pages)");
                    const uint32_t valid = std::min<uint32_t>(kPageRows, ix->list_len[l] - t * kPageRows);
                    for (uint32_t r = 0; r < valid; r++)
                        if (ids[r] >= (uint64_t)h.n) return bail("corrupt index file (row id out of range)");
                    const size_t row0 = (size_t)pg * kPageRows;
                    if (cudaMemcpy(reinterpret_cast<char *>(ix->d_pool) + (size_t)pg * pb, page.data(), pb, cudaMemcpyHostToDevice) != cudaSuccess ||
                        cudaMemcpy(ix->d_row_ids + row0, ids.data(), kPageRows * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
                        (ix->d_row_bias && cudaMemcpy(ix->d_row_bias + row0, bias.data(), kPageRows * 4, cudaMemcpyHostToDevice) != cudaSuccess))
                        return bail("H2D failed");
                }
            }
            std::vector<uint32_t> iota_pages(std::max<uint32_t>(h.pages_used, 1)), order(nl);
            for (uint32_t i = 0; i < h.pages_used; i++) iota_pages[i] = i;
            for (int l = 0; l < nl; l++) order[l] = (uint32_t)l;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ix->list_len[a] > ix->list_len[b]; });
            auto up = [&](uint32_t **dptr, const std::vector<uint32_t> &v, size_t count) {
                return cudaMalloc(dptr, std::max<size_t>(count, 1) * 4) == cudaSuccess &&
                       cudaMemcpy(*dptr, v.data(), count * 4, cudaMemcpyHostToDevice) == cudaSuccess;
            };
            if (!up(&ix->d_list_pages, iota_pages, h.pages_used) || !up(&ix->d_list_page_off, page_off, (size_t)nl + 1) ||
                !up(&ix->d_list_order, order, nl) || !up(&ix->d_list_len, ix->list_len, nl) || cudaMalloc(&ix->d_flag, 32) != cudaSuccess)
                return bail("cudaMalloc failed");
            cudaMemset(ix->d_flag, 0, 32);
            if (upload_coarse(ix, ix->stream) != B200_OK) return bail(b200_last_error());
            if (cudaStreamSynchronize(ix->stream) != cudaSuccess) return bail("upload failed");
        }
    } catch (const std::bad_alloc &) {
        return bail("out of host memory while loading the index");
    }
    ix->trained = true;
    ix->built = true;
    *out = ix;
    return B200_OK;
}

extern "C" int b200_index_load(const char *path, b200_index **out) {
    if (!path || !out) return fail(B200_ERR_INVALID, "bad arguments");
    *out = nullptr;
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail(B200_ERR_INVALID, std::string("cannot open ") + path);
    Io io;
    io.read = file_read;
    io.ctx = fp;
    const int rc = index_load_io(&io, out);
    fclose(fp);
    return rc;
}

extern "C" int b200_index_load_cb(int (*read)(void *ctx, void *data, size_t bytes), void *ctx, b200_index **out) {
    if (!read || !out) return fail(B200_ERR_INVALID, "bad arguments");
    Io io;
    io.read = read;
    io.ctx = ctx;
    return index_load_io(&io, out);
}
