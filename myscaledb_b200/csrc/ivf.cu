// ivf.cu -- K5/K6: inverted-file indexes (IVFFLAT, IVFPQ), their GPU trainer, and the exact
// second-stage refine.
//
// Replaces Search::VectorIndex<...>::{build, search, computeTopDistanceSubset} for the index types
// the reference reaches through VIWithColumnInPart::search / computeTopDistanceSubset
// (reference: src/VectorIndex/Common/VIWithDataPart.cpp:858-957, :838-856) and
// MergeTreeVSManager::executeSecondStageVectorScan (src/VectorIndex/Storages/MergeTreeVSManager.cpp:510-630).
// The reference's implementations live in the un-vendored search-index library (Faiss IVF*) and the
// closed-source MSTG; the algorithms here are the published IVF-ADC ones, the layout is ours:
//   * coarse quantiser: nlist centroids (fp32), k-means trained on device;
//   * inverted lists: row ids sorted by list (stable radix sort => ascending id inside a list),
//     IVFPQ additionally m one-byte codes per row in list order (residual encoded, 256 centroids
//     per sub-quantiser);
//   * the raw fp32 rows stay resident (a b200_corpus) -- IVFFLAT scans them through the id lists,
//     the two-stage ("MSTG"-type) search re-ranks its candidates against them exactly.
// Kernels: kmeans_assign_kernel (tiled fp32 GEMM-like argmin), ivf_flat_scan_kernel and
// ivfpq_scan_kernel (one CTA per (query, probed list); PQ look-up table in shared memory;
// HBM-bound on the code / row bytes of the probed lists), refine_kernel (warp per candidate).
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"
#include "kernels.h"

namespace b200 {

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

__global__ void scatter_codes_kernel(const uint32_t *code_j, const uint32_t *pos_of_row, int64_t n, int m, int j, uint8_t *codes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) codes[(int64_t)pos_of_row[i] * m + j] = (uint8_t)code_j[i];
}

__global__ void invert_perm_kernel(const uint32_t *sorted_rows, int64_t n, uint32_t *pos_of_row) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos_of_row[sorted_rows[i]] = (uint32_t)i;
}

__global__ void iota_kernel(uint32_t *v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

// ------------------------------------------------------------------------------------
// search kernels: one CTA per (query, probe)
// ------------------------------------------------------------------------------------
struct IvfScanParams {
    const float *queries;      // [nq][d_pad]
    const int64_t *probe;      // [nq][nprobe] list ids (-1 = none)
    const uint32_t *list_off;  // [nlist + 1]
    const uint32_t *list_ids;  // row ids in list order
    const uint8_t *alive;
    // IVFFLAT
    const float *rows;         // raw store [n][d_pad]
    // IVFPQ
    const float *centroids;    // [nlist][d]
    const float *pq;           // [m][256][dsub]
    const uint8_t *codes;      // [n][m] list order
    float *part_keys;          // [nq][nprobe][k]
    uint32_t *part_ids;
    int d, d_pad, m, dsub, nprobe, k;
    int l2;
};

__global__ void __launch_bounds__(256) ivf_flat_scan_kernel(const IvfScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *qs = reinterpret_cast<float *>(smem_raw);           // [d_pad]
    float *lk = qs + p.d_pad;                                  // [8][k]
    uint32_t *li = reinterpret_cast<uint32_t *>(lk + 8 * p.k);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = blockIdx.y;
    const int pr = blockIdx.x;
    for (int i = threadIdx.x; i < p.d_pad; i += 256) qs[i] = p.queries[q * p.d_pad + i];
    WarpTopK list;
    list.init(lk + (size_t)warp * p.k, li + (size_t)warp * p.k, p.k);
    for (int j = lane; j < p.k; j += 32) list.keys[j] = FLT_MAX;
    __syncthreads();
    const int64_t l = p.probe[q * p.nprobe + pr];
    if (l >= 0) {
        const uint32_t beg = p.list_off[l], end = p.list_off[l + 1];
        for (uint32_t r0 = beg + warp; r0 < end; r0 += 8) {  // one row per warp step; trip count is warp-uniform
            const uint32_t id = p.list_ids[r0];
            if (p.alive && !((p.alive[id >> 3] >> (id & 7)) & 1)) continue;
            const float4 *row = reinterpret_cast<const float4 *>(p.rows + (size_t)id * p.d_pad);
            float acc = 0.f;
            for (int c = lane; c < p.d_pad / 4; c += 32) {
                const float4 y = row[c];
                const float4 x = reinterpret_cast<const float4 *>(qs)[c];
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
            list.insert(p.l2 ? acc : -acc, id);
        }
    }
    __syncthreads();
    // the 8 warp lists -> this (query, list)'s partial top-k
    block_rank_merge(lk, li, 8, p.k, p.k, p.part_keys + ((size_t)q * p.nprobe + pr) * p.k,
                     p.part_ids + ((size_t)q * p.nprobe + pr) * p.k);
}

__global__ void __launch_bounds__(256) ivfpq_scan_kernel(const IvfScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *lut = reinterpret_cast<float *>(smem_raw);          // [m][256]
    float *lk = lut + (size_t)p.m * 256;                       // [8][k]
    uint32_t *li = reinterpret_cast<uint32_t *>(lk + 8 * p.k);
    __shared__ float bias_s, bias_w[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q = blockIdx.y;
    const int pr = blockIdx.x;
    WarpTopK list;
    list.init(lk + (size_t)warp * p.k, li + (size_t)warp * p.k, p.k);
    for (int j = lane; j < p.k; j += 32) list.keys[j] = FLT_MAX;
    float *rs = reinterpret_cast<float *>(li + 8 * p.k);  // [d]: residual q - centroid (L2) or the query (IP)
    const int64_t l = p.probe[q * p.nprobe + pr];
    if (l >= 0) {
        const float *qv = p.queries + q * p.d_pad;
        const float *cv = p.centroids + l * p.d;
        float b = 0.f;
        for (int i = threadIdx.x; i < p.d; i += 256) {
            const float x = qv[i], c = cv[i];
            rs[i] = p.l2 ? x - c : x;
            b = fmaf(x, c, b);
        }
        if (!p.l2) {  // IP: the q . centroid term is the same for every row of the list
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
            if (lane == 0) bias_w[warp] = b;
        }
    }
    __syncthreads();
    if (l >= 0) {
        // look-up table on the residual (L2) / on the query (IP): one thread per (sub-quantiser, centroid) entry, its
        // dsub codebook floats read as whole 16-byte words (a warp reads one contiguous span of the codebook)
        for (int e = threadIdx.x; e < p.m * 256; e += 256) {
            const int j = e >> 8;
            const float *cw = p.pq + (size_t)e * p.dsub;
            const float *r = rs + j * p.dsub;
            float s = 0.f;
            if ((p.dsub & 3) == 0) {
                for (int t = 0; t < p.dsub; t += 4) {
                    const float4 w = *reinterpret_cast<const float4 *>(cw + t);
                    if (p.l2) {
                        float u = r[t] - w.x; s = fmaf(u, u, s);
                        u = r[t + 1] - w.y; s = fmaf(u, u, s);
                        u = r[t + 2] - w.z; s = fmaf(u, u, s);
                        u = r[t + 3] - w.w; s = fmaf(u, u, s);
                    } else {
                        s = fmaf(r[t], w.x, s); s = fmaf(r[t + 1], w.y, s);
                        s = fmaf(r[t + 2], w.z, s); s = fmaf(r[t + 3], w.w, s);
                    }
                }
            } else {
                for (int t = 0; t < p.dsub; t++) {
                    if (p.l2) {
                        const float u = r[t] - cw[t];
                        s = fmaf(u, u, s);
                    } else {
                        s = fmaf(r[t], cw[t], s);
                    }
                }
            }
            lut[e] = s;
        }
        if (threadIdx.x == 0) {
            float b = 0.f;
            if (!p.l2)
                for (int w = 0; w < 8; w++) b += bias_w[w];
            bias_s = b;
        }
    }
    __syncthreads();
    if (l >= 0) {
        const uint32_t beg = p.list_off[l], end = p.list_off[l + 1];
        const float bias = bias_s;
        for (uint32_t r0 = beg + warp * 32; r0 < end; r0 += 256) {
            const uint32_t r = r0 + lane;
            bool ok = r < end;
            uint32_t id = 0;
            float key = FLT_MAX;
            if (ok) {
                id = p.list_ids[r];
                if (p.alive) ok = (p.alive[id >> 3] >> (id & 7)) & 1;
            }
            if (ok) {
                const uint8_t *code = p.codes + (size_t)r * p.m;
                float s = 0.f;
                if ((p.m & 15) == 0) {
                    const uint4 *cq = reinterpret_cast<const uint4 *>(code);
                    for (int j16 = 0; j16 < p.m / 16; j16++) {
                        const uint4 w4 = cq[j16];
                        const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int h = 0; h < 4; h++) {
                            const float *lt = lut + (size_t)(j16 * 16 + h * 4) * 256;
                            s += lt[ww[h] & 255];
                            s += lt[256 + ((ww[h] >> 8) & 255)];
                            s += lt[512 + ((ww[h] >> 16) & 255)];
                            s += lt[768 + (ww[h] >> 24)];
                        }
                    }
                } else if ((p.m & 3) == 0) {
                    const uint32_t *cw = reinterpret_cast<const uint32_t *>(code);
                    for (int j4 = 0; j4 < p.m / 4; j4++) {
                        const uint32_t w = cw[j4];
                        s += lut[(j4 * 4 + 0) * 256 + (w & 255)];
                        s += lut[(j4 * 4 + 1) * 256 + ((w >> 8) & 255)];
                        s += lut[(j4 * 4 + 2) * 256 + ((w >> 16) & 255)];
                        s += lut[(j4 * 4 + 3) * 256 + (w >> 24)];
                    }
                } else {
                    for (int j = 0; j < p.m; j++) s += lut[j * 256 + code[j]];
                }
                key = p.l2 ? s : -(s + bias);
                ok = list.passes(key, id);
            }
            unsigned mk = __ballot_sync(0xffffffffu, ok);
            while (mk) {
                const int src = __ffs(mk) - 1;
                mk &= mk - 1;
                list.insert(__shfl_sync(0xffffffffu, key, src), __shfl_sync(0xffffffffu, id, src));
            }
        }
    }
    __syncthreads();
    // the 8 warp lists -> this (query, list)'s partial top-k
    block_rank_merge(lk, li, 8, p.k, p.k, p.part_keys + ((size_t)q * p.nprobe + pr) * p.k,
                     p.part_ids + ((size_t)q * p.nprobe + pr) * p.k);
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
        p.out_ids[q * p.k + j] = have ? (int64_t)fi[j] : -1;
        p.out_dis[q * p.k + j] = !have ? (p.l2 || p.cosine ? FLT_MAX : -FLT_MAX) : p.l2 ? key : p.cosine ? 1.f + key : -key;
    }
}

// candidates of stage 1 (int64 ids, -1 padded) -> as is; helper to widen u32 partial ids is the merge kernel.

static inline int gridsz(int64_t work, int threads = 256) {
    int64_t b = ceil_div(work, threads);
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, 148 * 32));
}

}  // namespace b200

using namespace b200;

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
enum { IDX_FLAT = 0, IDX_IVFFLAT = 1, IDX_IVFPQ = 2, IDX_MSTG = 3 };

struct b200_index {
    int type = IDX_FLAT, metric = B200_METRIC_L2, d = 0, d_pad = 0;
    int nlist = 0, m = 0, dsub = 0;
    int default_nprobe = 32, refine_factor = 8;
    int64_t n = 0;
    bool built = false;
    bool use_ivf = false;  // false => small part, FLAT fallback (reference: fallback_to_flat, test 00029)
    b200_corpus *raw = nullptr;     // fp32 rows (cosine: unit vectors), metric L2 or IP
    b200_corpus *coarse = nullptr;  // centroids as a tiny FLAT corpus (L2)
    float *d_centroids = nullptr, *d_pq = nullptr;
    uint32_t *d_list_off = nullptr, *d_list_ids = nullptr;
    uint8_t *d_codes = nullptr;
    std::vector<uint32_t> list_off;
    double biased_list_rows = 0;    // sum(size^2) / n over the inverted lists (planner input)
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    // workspaces
    void *w_q = nullptr, *w_probe = nullptr, *w_pd = nullptr, *w_pk = nullptr, *w_pi = nullptr, *w_alive = nullptr, *w_od = nullptr,
         *w_oi = nullptr, *w_cand = nullptr, *w_cd = nullptr;
    size_t c_q = 0, c_probe = 0, c_pd = 0, c_pk = 0, c_pi = 0, c_alive = 0, c_od = 0, c_oi = 0, c_cand = 0, c_cd = 0;
};

// internal hooks into capi.cu
extern "C" int b200_corpus_search_device(b200_corpus *c, const float *d_queries, int64_t nq, int k, const uint8_t *d_alive_bits,
                                         int64_t id_offset, float *d_out_dis, int64_t *d_out_ids, void *stream);
namespace b200 {
const void *corpus_device_rows(const b200_corpus *c);
int corpus_normalize_rows(b200_corpus *c);
}

static int wsr(void **p, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return B200_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    if (cudaMalloc(p, bytes + bytes / 4 + 256) != cudaSuccess) {
        cudaGetLastError();
        return fail(B200_ERR_NOMEM, "cudaMalloc failed (index workspace)");
    }
    *cap = bytes + bytes / 4;
    return B200_OK;
}

static int parse_int_param(const char *json, const char *key, int defv) {
    if (!json) return defv;
    const char *p = strstr(json, key);
    if (!p) return defv;
    p += strlen(key);
    while (*p && (*p == '"' || *p == ':' || *p == '=' || *p == ' ' || *p == '\'')) p++;
    if (!(*p >= '0' && *p <= '9')) return defv;
    return atoi(p);
}

extern "C" int b200_index_create(const char *type, int metric, int d, const char *params, b200_index **out) {
    if (!type || !out || d <= 0) return fail(B200_ERR_INVALID, "bad arguments");
    *out = nullptr;
    if (metric != B200_METRIC_L2 && metric != B200_METRIC_IP && metric != B200_METRIC_COSINE)
        return fail(B200_ERR_INVALID, "float indexes take L2, IP or COSINE");
    std::string t(type);
    for (auto &ch : t) ch = (char)toupper((unsigned char)ch);
    int ty;
    if (t == "FLAT") ty = IDX_FLAT;
    else if (t == "IVFFLAT") ty = IDX_IVFFLAT;
    else if (t == "IVFPQ") ty = IDX_IVFPQ;
    else if (t == "MSTG") ty = IDX_MSTG;  // our two-stage stand-in: IVFPQ candidates + exact refine
    else return fail(B200_ERR_UNSUPPORTED, "index type " + t + " is not implemented (FLAT, IVFFLAT, IVFPQ, MSTG)");
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
    ix->nlist = parse_int_param(params, "ncentroids", parse_int_param(params, "nlist", 0));
    ix->m = parse_int_param(params, "M", parse_int_param(params, "m", 0));
    ix->default_nprobe = parse_int_param(params, "nprobe", 32);
    ix->refine_factor = parse_int_param(params, "refine_factor", 8);
    cudaGetDevice(&ix->device);
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
    if (ix->raw) b200_corpus_free(ix->raw);
    if (ix->coarse) b200_corpus_free(ix->coarse);
    for (void *p : {(void *)ix->d_centroids, (void *)ix->d_pq, (void *)ix->d_list_off, (void *)ix->d_list_ids, (void *)ix->d_codes,
                    ix->w_q, ix->w_probe, ix->w_pd, ix->w_pk, ix->w_pi, ix->w_alive, ix->w_od, ix->w_oi, ix->w_cand, ix->w_cd})
        if (p) cudaFree(p);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
    return B200_OK;
}

static void set_biased_list_rows(b200_index *ix) {
    double ss = 0;
    for (int l = 0; l < ix->nlist; l++) {
        const double sz = (double)ix->list_off[l + 1] - (double)ix->list_off[l];
        ss += sz * sz;
    }
    ix->biased_list_rows = ix->n > 0 ? ss / (double)ix->n : 0;
}

// k-means on device rows x [n][stride]; centroids written to d_c [nc][d]
static int kmeans_device(const float *x, int64_t n, int64_t stride, int d, int nc, int iters, float *d_c, cudaStream_t s) {
    // init: nc points at a fixed stride through the data (deterministic)
    std::vector<int64_t> pick(nc);
    for (int i = 0; i < nc; i++) pick[i] = (int64_t)((double)i * (double)n / (double)nc);
    int64_t *d_pick = nullptr;
    float *d_sums = nullptr, *d_cn = nullptr;
    uint32_t *d_cnt = nullptr, *d_idx = nullptr;
    B200_CUDA_OK(cudaMalloc(&d_pick, (size_t)nc * 8));
    B200_CUDA_OK(cudaMalloc(&d_sums, (size_t)nc * d * 4));
    B200_CUDA_OK(cudaMalloc(&d_cn, (size_t)nc * 4));
    B200_CUDA_OK(cudaMalloc(&d_cnt, (size_t)nc * 4));
    B200_CUDA_OK(cudaMalloc(&d_idx, (size_t)n * 4));
    B200_CUDA_OK(cudaMemcpyAsync(d_pick, pick.data(), (size_t)nc * 8, cudaMemcpyHostToDevice, s));
    gather_rows_kernel<<<gridsz((int64_t)nc * d), 256, 0, s>>>(x, stride, d_pick, nc, d, d_c);
    g_launches++;
    for (int it = 0; it < iters; it++) {
        rows_sqnorm_kernel<<<(unsigned)ceil_div(nc, 256), 256, 0, s>>>(d_c, nc, d, d_cn);
        kmeans_assign_kernel<<<(unsigned)ceil_div(n, KA_T), 256, 0, s>>>(x, n, stride, d, d_c, nc, d_cn, d_idx, nullptr);
        B200_CUDA_OK(cudaMemsetAsync(d_sums, 0, (size_t)nc * d * 4, s));
        B200_CUDA_OK(cudaMemsetAsync(d_cnt, 0, (size_t)nc * 4, s));
        kmeans_accumulate_kernel<<<gridsz(n * d), 256, 0, s>>>(x, n, stride, d, d_idx, d_sums, d_cnt);
        kmeans_update_kernel<<<gridsz((int64_t)nc * d), 256, 0, s>>>(d_c, d_sums, d_cnt, nc, d);
        g_launches += 4;
    }
    B200_CUDA_OK(cudaGetLastError());
    B200_CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(d_pick);
    cudaFree(d_sums);
    cudaFree(d_cn);
    cudaFree(d_cnt);
    cudaFree(d_idx);
    return B200_OK;
}

// VIWithColumnInPart::buildIndex -> Search::VectorIndex::build (VIWithDataPart.cpp:131): one-shot build from host rows
extern "C" int b200_index_build(b200_index *ix, const float *rows, int64_t n) {
    if (!ix || (!rows && n > 0) || n < 0) return fail(B200_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (ix->built) return fail(B200_ERR_INVALID, "index already built");
    B200_CUDA_OK(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    const int raw_metric = ix->metric == B200_METRIC_L2 ? B200_METRIC_L2 : B200_METRIC_IP;
    B200_TRY(b200_corpus_create(raw_metric, B200_DTYPE_F32, ix->d, n, &ix->raw));
    if (n) B200_TRY(b200_corpus_append(ix->raw, rows, n));
    if (ix->metric == B200_METRIC_COSINE && n) B200_TRY(corpus_normalize_rows(ix->raw));
    ix->n = n;
    // small parts fall back to FLAT (the reference does the same for tiny parts, test 00029)
    const bool want_ivf = ix->type != IDX_FLAT;
    if (want_ivf && ix->nlist <= 0) ix->nlist = (int)std::max<int64_t>(1, std::min<int64_t>(65536, (int64_t)(4.0 * sqrt((double)std::max<int64_t>(n, 1)))));
    ix->use_ivf = want_ivf && n >= std::max<int64_t>(2000, 8ll * ix->nlist);
    if (!ix->use_ivf) {
        ix->built = true;
        return B200_OK;
    }
    const float *x = reinterpret_cast<const float *>(corpus_device_rows(ix->raw));
    const int d = ix->d, nl = ix->nlist;
    const int64_t stride = ix->d_pad;
    // ---- coarse quantiser
    B200_CUDA_OK(cudaMalloc(&ix->d_centroids, (size_t)nl * d * 4));
    {
        const int64_t ns = std::min<int64_t>(n, 256ll * nl);  // training sample: a strided subset, gathered
        float *d_sample = nullptr;
        const float *train = x;
        int64_t train_stride = stride;
        if (ns < n) {
            std::vector<int64_t> pick(ns);
            for (int64_t i = 0; i < ns; i++) pick[i] = (int64_t)((double)i * (double)n / (double)ns);
            int64_t *d_pick = nullptr;
            B200_CUDA_OK(cudaMalloc(&d_pick, (size_t)ns * 8));
            B200_CUDA_OK(cudaMalloc(&d_sample, (size_t)ns * d * 4));
            B200_CUDA_OK(cudaMemcpyAsync(d_pick, pick.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, s));
            gather_rows_kernel<<<gridsz(ns * d), 256, 0, s>>>(x, stride, d_pick, ns, d, d_sample);
            g_launches++;
            B200_CUDA_OK(cudaStreamSynchronize(s));
            cudaFree(d_pick);
            train = d_sample;
            train_stride = d;
        }
        int rc = kmeans_device(train, ns, train_stride, d, nl, 10, ix->d_centroids, s);
        if (d_sample) cudaFree(d_sample);
        B200_TRY(rc);
    }
    // ---- assign every row, sort rows by list (stable => ascending id inside a list)
    uint32_t *d_list = nullptr, *d_rows_in = nullptr, *d_list_sorted = nullptr;
    float *d_cn = nullptr;
    B200_CUDA_OK(cudaMalloc(&d_list, (size_t)n * 4));
    B200_CUDA_OK(cudaMalloc(&d_rows_in, (size_t)n * 4));
    B200_CUDA_OK(cudaMalloc(&d_list_sorted, (size_t)n * 4));
    B200_CUDA_OK(cudaMalloc(&ix->d_list_ids, (size_t)n * 4));
    B200_CUDA_OK(cudaMalloc(&d_cn, (size_t)nl * 4));
    rows_sqnorm_kernel<<<(unsigned)ceil_div(nl, 256), 256, 0, s>>>(ix->d_centroids, nl, d, d_cn);
    kmeans_assign_kernel<<<(unsigned)ceil_div(n, KA_T), 256, 0, s>>>(x, n, stride, d, ix->d_centroids, nl, d_cn, d_list, nullptr);
    iota_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(d_rows_in, n);  // one thread per row (not grid-stride)
    g_launches += 3;
    {
        size_t tmp_bytes = 0;
        int bits = 1;
        while ((1 << bits) < nl) bits++;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_list, d_list_sorted, d_rows_in, ix->d_list_ids, (int)n, 0, bits, s);
        void *tmp = nullptr;
        B200_CUDA_OK(cudaMalloc(&tmp, tmp_bytes + 256));
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_list, d_list_sorted, d_rows_in, ix->d_list_ids, (int)n, 0, bits, s);
        g_launches++;
        B200_CUDA_OK(cudaStreamSynchronize(s));
        cudaFree(tmp);
    }
    {
        std::vector<uint32_t> ls(n);
        B200_CUDA_OK(cudaMemcpy(ls.data(), d_list_sorted, (size_t)n * 4, cudaMemcpyDeviceToHost));
        ix->list_off.assign(nl + 1, 0);
        for (int64_t i = 0; i < n; i++) ix->list_off[ls[i] + 1]++;
        for (int l = 0; l < nl; l++) ix->list_off[l + 1] += ix->list_off[l];
        set_biased_list_rows(ix);
        B200_CUDA_OK(cudaMalloc(&ix->d_list_off, (size_t)(nl + 1) * 4));
        B200_CUDA_OK(cudaMemcpy(ix->d_list_off, ix->list_off.data(), (size_t)(nl + 1) * 4, cudaMemcpyHostToDevice));
    }
    // coarse quantiser as a FLAT corpus for the probe search (L2 for every metric: unit vectors under cosine;
    // IP indexes probe by L2 too, like Faiss's default quantiser)
    B200_TRY(b200_corpus_create(B200_METRIC_L2, B200_DTYPE_F32, d, nl, &ix->coarse));
    {
        std::vector<float> hc((size_t)nl * d);
        B200_CUDA_OK(cudaMemcpy(hc.data(), ix->d_centroids, hc.size() * 4, cudaMemcpyDeviceToHost));
        B200_TRY(b200_corpus_append(ix->coarse, hc.data(), nl));
    }
    // ---- product quantiser on residuals
    if (ix->type == IDX_IVFPQ || ix->type == IDX_MSTG) {
        if (ix->m <= 0) {  // default: sub-vectors of <= 8 dims
            ix->m = d;
            for (int cand : {8, 4, 2, 1}) if (d % cand == 0) { ix->m = d / cand; break; }
        }
        if (d % ix->m) return fail(B200_ERR_INVALID, "PQ M must divide the dimension");
        const int m = ix->m, dsub = d / m;
        ix->dsub = dsub;
        B200_CUDA_OK(cudaMalloc(&ix->d_pq, (size_t)m * 256 * dsub * 4));
        B200_CUDA_OK(cudaMalloc(&ix->d_codes, (size_t)n * m));
        uint32_t *d_pos = nullptr, *d_code = nullptr;
        float *d_res = nullptr, *d_cbn = nullptr;
        B200_CUDA_OK(cudaMalloc(&d_pos, (size_t)n * 4));
        B200_CUDA_OK(cudaMalloc(&d_code, (size_t)n * 4));
        B200_CUDA_OK(cudaMalloc(&d_res, (size_t)n * dsub * 4));
        B200_CUDA_OK(cudaMalloc(&d_cbn, 256 * 4));
        invert_perm_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(ix->d_list_ids, n, d_pos);
        g_launches++;
        const int64_t ns = std::min<int64_t>(n, 65536);
        for (int j = 0; j < m; j++) {
            residual_sub_kernel<<<gridsz(n * dsub), 256, 0, s>>>(x, n, stride, ix->d_centroids, d_list, d, j, dsub, d_res);
            g_launches++;
            float *cb = ix->d_pq + (size_t)j * 256 * dsub;
            // train on the first ns residuals of a strided view: stride (n / ns) rows
            const int64_t step = std::max<int64_t>(1, n / ns);
            B200_TRY(kmeans_device(d_res, ns, step * dsub, dsub, 256, 8, cb, s));
            rows_sqnorm_kernel<<<1, 256, 0, s>>>(cb, 256, dsub, d_cbn);
            kmeans_assign_kernel<<<(unsigned)ceil_div(n, KA_T), 256, 0, s>>>(d_res, n, dsub, dsub, cb, 256, d_cbn, d_code, nullptr);
            scatter_codes_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(d_code, d_pos, n, m, j, ix->d_codes);
            g_launches += 3;
        }
        B200_CUDA_OK(cudaGetLastError());
        B200_CUDA_OK(cudaStreamSynchronize(s));
        cudaFree(d_pos);
        cudaFree(d_code);
        cudaFree(d_res);
        cudaFree(d_cbn);
    }
    B200_CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(d_list);
    cudaFree(d_rows_in);
    cudaFree(d_list_sorted);
    cudaFree(d_cn);
    ix->built = true;
    return B200_OK;
}

extern "C" int b200_index_memory_bytes(const b200_index *ix, uint64_t *out_bytes) {
    if (!ix || !out_bytes) return fail(B200_ERR_INVALID, "bad arguments");
    uint64_t b = 0, t = 0;
    if (ix->raw && b200_corpus_memory_bytes(ix->raw, &t) == B200_OK) b += t;
    if (ix->coarse && b200_corpus_memory_bytes(ix->coarse, &t) == B200_OK) b += t;
    if (ix->use_ivf) {
        b += (uint64_t)ix->nlist * ix->d * 4 + (uint64_t)(ix->nlist + 1) * 4 + (uint64_t)ix->n * 4;  // centroids, offsets, ids
        if (ix->d_pq) b += (uint64_t)ix->m * 256 * ix->dsub * 4 + (uint64_t)ix->n * ix->m;          // codebook, codes
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

// computeTopDistanceSubset (VIWithDataPart.cpp:838-856): exact distances of a candidate id set -> top-k
static int refine_device(b200_index *ix, const float *d_q /*[nq][d_pad] prepared*/, int64_t nq, const int64_t *d_cand, int ncand,
                         int k, float *d_out_dis, int64_t *d_out_ids, cudaStream_t s) {
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
    const size_t smem = (size_t)ix->d_pad * 4 + (size_t)9 * k * 8;
    B200_CUDA_OK(cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    refine_kernel<<<(unsigned)nq, 256, smem, s>>>(rp);
    g_launches++;
    B200_CUDA_OK(cudaGetLastError());
    return B200_OK;
}

static int prepare_queries(b200_index *ix, const float *queries, int64_t nq, cudaStream_t s) {
    B200_TRY(wsr(&ix->w_cd, &ix->c_cd, (size_t)nq * ix->d * 4));
    B200_TRY(wsr(&ix->w_q, &ix->c_q, (size_t)nq * ix->d_pad * 4));
    B200_CUDA_OK(cudaMemcpyAsync(ix->w_cd, queries, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, s));
    B200_CUDA_OK(launch_pad_rows_f32(reinterpret_cast<const float *>(ix->w_cd), ix->d, reinterpret_cast<float *>(ix->w_q), ix->d_pad, nq, s));
    if (ix->metric == B200_METRIC_COSINE) B200_CUDA_OK(launch_normalize_rows_f32(reinterpret_cast<float *>(ix->w_q), ix->d_pad, nq, s));
    return B200_OK;
}

// Batch planner.  An IVF probe costs per query; the exact pass costs per 8-query scan pass or per 256-query tensor-core
// tile.  Rates are measured ones (profiles/r01_ivf_latency.log):
//  * IVF: one CTA per (query, list) walks its list at ~8 rows/us (PQ, 96-byte codes) or ~25 rows/us (IVFFLAT rows of
//    3 KB); 296 CTAs run at once; ~0.2 ms of fixed work (coarse probe, merges, refine).  The list a query lands in is
//    size-biased: expected rows = sum(size^2) / n, which is what `biased_list_rows` holds (clustered data makes it
//    10x the mean).
//  * exact: one query 5 TB/s, up to 8 queries 2.5 TB/s per pass, 16+ queries 850 TFLOP/s of tf32 MMA work at three
//    products per term (ip_gemm_tf32x3_sm100.cu), never below the one-pass HBM time.
// `exact_batch=0|1` in the search parameters overrides the choice.
static bool exact_batch_is_cheaper(const b200_index *ix, int64_t nq, const char *params) {
    const int force = parse_int_param(params, "exact_batch", -1);
    if (force >= 0) return force != 0;
    int nprobe = parse_int_param(params, "nprobe", ix->default_nprobe);
    nprobe = std::max(1, std::min(nprobe, ix->nlist));
    const double list_rows = ix->biased_list_rows > 0 ? ix->biased_list_rows : (double)ix->n / ix->nlist;
    const double t_list = list_rows / (ix->type == IDX_IVFFLAT ? 25e6 : 8e6);
    const double waves = std::max(1.0, (double)nq * nprobe / 296.0);
    const double t_ivf = 0.2e-3 + t_list * waves;
    const double row_bytes = (double)ix->n * ix->d_pad * 4;
    double t_exact;
    if (nq >= 16) t_exact = std::max(row_bytes / 5e12, (double)round_up(nq, 256) * ix->n * ix->d_pad * 6.0 / 850e12);
    else if (nq == 1) t_exact = row_bytes / 5e12;
    else t_exact = (double)ceil_div(nq, 8) * row_bytes / 2.5e12;
    t_exact += 0.1e-3;
    return t_exact < t_ivf;
}

// Search::VectorIndex::search(queries, k, params, first_stage_only, filter) (VIWithDataPart.cpp:926)
extern "C" int b200_index_search(b200_index *ix, const float *queries, int64_t nq, int k, const char *params, int first_stage_only,
                                 const uint8_t *alive_bits, float *out_dis, int64_t *out_ids, int64_t *out_num_candidates) {
    if (!ix || (!queries && nq > 0) || !out_dis || !out_ids || nq < 0 || k <= 0) return fail(B200_ERR_INVALID, "bad arguments");
    if (!ix->built) return fail(B200_ERR_INVALID, "index not built");
    if (out_num_candidates) *out_num_candidates = k;
    if (nq == 0) return B200_OK;
    if (!ix->use_ivf || exact_batch_is_cheaper(ix, nq, params)) {
        // FLAT / fallback-to-flat, or a batch large enough that one exact pass over the raw rows on the tensor cores
        // (3xTF32, ip_gemm_tf32x3_sm100.cu) costs less than nq list probes: exact scan of the raw rows, recall 1
        int rc = b200_corpus_search(ix->raw, queries, nq, k, alive_bits, out_dis, out_ids);
        if (rc == B200_OK && ix->metric == B200_METRIC_COSINE) {
            // raw rows are unit vectors searched under IP with unnormalised queries: finish the cosine
            for (int64_t q = 0; q < nq; q++) {
                float s32 = 0;
                for (int j = 0; j < ix->d; j++) s32 += queries[q * ix->d + j] * queries[q * ix->d + j];
                const float nf = s32 < FLT_EPSILON ? 1.f : sqrtf(s32);
                for (int j = 0; j < k; j++)
                    if (out_ids[q * k + j] >= 0) out_dis[q * k + j] = 1.f - out_dis[q * k + j] / nf;
                    else out_dis[q * k + j] = FLT_MAX;
            }
        }
        return rc;
    }
    if (k > 1024) return fail(B200_ERR_UNSUPPORTED, "k > 1024 on IVF indexes");
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    int nprobe = parse_int_param(params, "nprobe", ix->default_nprobe);
    nprobe = std::max(1, std::min(nprobe, ix->nlist));
    const bool two_stage = ix->type == IDX_MSTG && !first_stage_only;
    const int refine_factor = parse_int_param(params, "refine_factor", ix->refine_factor);
    const int k1 = two_stage ? std::min(1024, std::max(k, k * refine_factor)) : k;
    if (out_num_candidates) *out_num_candidates = k1;
    B200_TRY(prepare_queries(ix, queries, nq, s));
    const float *d_q = reinterpret_cast<const float *>(ix->w_q);
    // coarse probe: nprobe nearest centroids (exact FLAT scan of the centroid table)
    B200_TRY(wsr(&ix->w_probe, &ix->c_probe, (size_t)nq * nprobe * 8));
    B200_TRY(wsr(&ix->w_pd, &ix->c_pd, (size_t)nq * nprobe * 4));
    B200_TRY(wsr(&ix->w_cd, &ix->c_cd, (size_t)nq * ix->d * 4));
    {
        // the coarse corpus wants raw [nq][d] device queries: for cosine use the normalised ones (strip padding)
        float *d_raw = reinterpret_cast<float *>(ix->w_cd);
        if (ix->d == ix->d_pad) B200_CUDA_OK(cudaMemcpyAsync(d_raw, d_q, (size_t)nq * ix->d * 4, cudaMemcpyDeviceToDevice, s));
        else B200_CUDA_OK(cudaMemcpy2DAsync(d_raw, (size_t)ix->d * 4, d_q, (size_t)ix->d_pad * 4, (size_t)ix->d * 4, nq, cudaMemcpyDeviceToDevice, s));
        B200_TRY(b200_corpus_search_device(ix->coarse, d_raw, nq, nprobe, nullptr, 0, reinterpret_cast<float *>(ix->w_pd),
                                           reinterpret_cast<int64_t *>(ix->w_probe), s));
    }
    const uint8_t *d_alive = nullptr;
    if (alive_bits) {
        const size_t ab = (size_t)ceil_div(ix->n, 8);
        B200_TRY(wsr(&ix->w_alive, &ix->c_alive, ab + 16));
        B200_CUDA_OK(cudaMemcpyAsync(ix->w_alive, alive_bits, ab, cudaMemcpyHostToDevice, s));
        d_alive = reinterpret_cast<const uint8_t *>(ix->w_alive);
    }
    B200_TRY(wsr(&ix->w_pk, &ix->c_pk, (size_t)nq * nprobe * k1 * 4));
    B200_TRY(wsr(&ix->w_pi, &ix->c_pi, (size_t)nq * nprobe * k1 * 4));
    B200_TRY(wsr(&ix->w_od, &ix->c_od, (size_t)nq * k1 * 4));
    B200_TRY(wsr(&ix->w_oi, &ix->c_oi, (size_t)nq * k1 * 8));
    IvfScanParams sp{};
    sp.queries = d_q;
    sp.probe = reinterpret_cast<const int64_t *>(ix->w_probe);
    sp.list_off = ix->d_list_off;
    sp.list_ids = ix->d_list_ids;
    sp.alive = d_alive;
    sp.rows = reinterpret_cast<const float *>(corpus_device_rows(ix->raw));
    sp.centroids = ix->d_centroids;
    sp.pq = ix->d_pq;
    sp.codes = ix->d_codes;
    sp.part_keys = reinterpret_cast<float *>(ix->w_pk);
    sp.part_ids = reinterpret_cast<uint32_t *>(ix->w_pi);
    sp.d = ix->d;
    sp.d_pad = ix->d_pad;
    sp.m = ix->m;
    sp.dsub = ix->dsub;
    sp.nprobe = nprobe;
    sp.k = k1;
    sp.l2 = ix->metric == B200_METRIC_L2;
    const dim3 grid(nprobe, (unsigned)nq);
    if (ix->type == IDX_IVFFLAT) {
        const size_t smem = (size_t)ix->d_pad * 4 + (size_t)8 * k1 * 8;
        B200_CUDA_OK(cudaFuncSetAttribute(ivf_flat_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ivf_flat_scan_kernel<<<grid, 256, smem, s>>>(sp);
    } else {
        const size_t smem = (size_t)ix->m * 256 * 4 + (size_t)8 * k1 * 8 + (size_t)ix->d * 4;
        if (smem > 220 * 1024) return fail(B200_ERR_UNSUPPORTED, "PQ look-up table (m * 1 KB) + top-k lists exceed shared memory");
        B200_CUDA_OK(cudaFuncSetAttribute(ivfpq_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ivfpq_scan_kernel<<<grid, 256, smem, s>>>(sp);
    }
    g_launches++;
    B200_CUDA_OK(cudaGetLastError());
    MergeParams mp{};
    mp.in_keys = sp.part_keys;
    mp.in_ids = sp.part_ids;
    mp.list_stride = k1;
    mp.q_stride = (int64_t)nprobe * k1;
    mp.n_lists = nprobe;
    mp.k_in = k1;
    mp.k = k1;
    mp.nq = nq;
    mp.out_mode = ix->metric == B200_METRIC_L2 ? kOutKey : ix->metric == B200_METRIC_IP ? kOutNeg : kOutOnePlus;
    mp.out_dis = reinterpret_cast<float *>(ix->w_od);
    mp.out_ids = reinterpret_cast<int64_t *>(ix->w_oi);
    B200_CUDA_OK(launch_topk_merge(mp, false, s));
    float *d_fd = mp.out_dis;
    int64_t *d_fi = mp.out_ids;
    if (two_stage) {
        B200_TRY(wsr(&ix->w_cand, &ix->c_cand, (size_t)nq * k * 12));
        float *r_d = reinterpret_cast<float *>(ix->w_cand);
        int64_t *r_i = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(ix->w_cand) + (size_t)round_up(nq * k * 4, 8));
        B200_TRY(refine_device(ix, d_q, nq, d_fi, k1, k, r_d, r_i, s));
        d_fd = r_d;
        d_fi = r_i;
    }
    const int kout = two_stage ? k : k1;
    if (kout != k) return fail(B200_ERR_INVALID, "internal: output width mismatch");
    B200_CUDA_OK(cudaMemcpyAsync(out_dis, d_fd, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaMemcpyAsync(out_ids, d_fi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    return B200_OK;
}

// computeTopDistanceSubset: queries [nq][d], candidates [nq][ncand] (negative = unused) -> exact top-k
extern "C" int b200_index_refine(b200_index *ix, const float *queries, int64_t nq, const int64_t *cand_ids, int64_t ncand, int k,
                                 float *out_dis, int64_t *out_ids) {
    if (!ix || !queries || !cand_ids || !out_dis || !out_ids || nq < 0 || ncand <= 0 || k <= 0)
        return fail(B200_ERR_INVALID, "bad arguments");
    if (!ix->built) return fail(B200_ERR_INVALID, "index not built");
    if (k > 1024) return fail(B200_ERR_UNSUPPORTED, "k > 1024 in refine");
    if (nq == 0) return B200_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    B200_TRY(prepare_queries(ix, queries, nq, s));
    B200_TRY(wsr(&ix->w_probe, &ix->c_probe, (size_t)nq * ncand * 8));
    B200_TRY(wsr(&ix->w_od, &ix->c_od, (size_t)nq * k * 4));
    B200_TRY(wsr(&ix->w_oi, &ix->c_oi, (size_t)nq * k * 8));
    B200_CUDA_OK(cudaMemcpyAsync(ix->w_probe, cand_ids, (size_t)nq * ncand * 8, cudaMemcpyHostToDevice, s));
    B200_TRY(refine_device(ix, reinterpret_cast<const float *>(ix->w_q), nq, reinterpret_cast<const int64_t *>(ix->w_probe), (int)ncand, k,
                           reinterpret_cast<float *>(ix->w_od), reinterpret_cast<int64_t *>(ix->w_oi), s));
    B200_CUDA_OK(cudaMemcpyAsync(out_dis, ix->w_od, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaMemcpyAsync(out_ids, ix->w_oi, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    return B200_OK;
}

// ------------------------------------------------------------------------------------
// persistence: VIWithColumnInPart::serialize / load (reference: src/VectorIndex/Common/VIWithDataPart.cpp:451-525,
// :578-764) write `<idx>-*.vidx3` through Search::IndexDataFileWriter; the on-disk format of the closed library
// is not reproducible, so this is our own single-file layout ("B2IX" v1): header, raw rows, then the IVF/PQ
// structures.  Loading re-uploads to HBM and recomputes the row norms on device.
// ------------------------------------------------------------------------------------
#include <cstdio>

namespace {
struct IxHeader {
    char magic[4];
    uint32_t version;
    int32_t type, metric, d, d_pad, nlist, m, dsub, default_nprobe, refine_factor, use_ivf;
    int64_t n;
};
bool wr(FILE *f, const void *p, size_t bytes) { return bytes == 0 || fwrite(p, 1, bytes, f) == bytes; }
bool rd(FILE *f, void *p, size_t bytes) { return bytes == 0 || fread(p, 1, bytes, f) == bytes; }
}  // namespace

extern "C" int b200_index_save(b200_index *ix, const char *path) {
    if (!ix || !path) return fail(B200_ERR_INVALID, "bad arguments");
    if (!ix->built) return fail(B200_ERR_INVALID, "index not built");
    std::lock_guard<std::mutex> lk(ix->mu);
    B200_CUDA_OK(cudaSetDevice(ix->device));
    FILE *f = fopen(path, "wb");
    if (!f) return fail(B200_ERR_INVALID, std::string("cannot open ") + path);
    IxHeader h{};
    memcpy(h.magic, "B2IX", 4);
    h.version = 1;
    h.type = ix->type; h.metric = ix->metric; h.d = ix->d; h.d_pad = ix->d_pad; h.nlist = ix->nlist; h.m = ix->m; h.dsub = ix->dsub;
    h.default_nprobe = ix->default_nprobe; h.refine_factor = ix->refine_factor; h.use_ivf = ix->use_ivf ? 1 : 0; h.n = ix->n;
    bool ok = wr(f, &h, sizeof(h));
    // raw rows, unpadded (cosine indexes hold unit vectors; they are written as stored)
    {
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
        ok = dump(ix->d_centroids, (size_t)ix->nlist * ix->d * 4) && wr(f, ix->list_off.data(), (size_t)(ix->nlist + 1) * 4) &&
             dump(ix->d_list_ids, (size_t)ix->n * 4);
        if (ok && ix->d_pq) ok = dump(ix->d_pq, (size_t)ix->m * 256 * ix->dsub * 4) && dump(ix->d_codes, (size_t)ix->n * ix->m);
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? B200_OK : fail(B200_ERR_INVALID, std::string("write failed: ") + path);
}

extern "C" int b200_index_load(const char *path, b200_index **out) {
    if (!path || !out) return fail(B200_ERR_INVALID, "bad arguments");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(B200_ERR_INVALID, std::string("cannot open ") + path);
    IxHeader h{};
    if (!rd(f, &h, sizeof(h)) || memcmp(h.magic, "B2IX", 4) != 0 || h.version != 1) {
        fclose(f);
        return fail(B200_ERR_INVALID, "not a B2IX v1 index file");
    }
    static const char *names[] = {"FLAT", "IVFFLAT", "IVFPQ", "MSTG"};
    if (h.type < 0 || h.type > 3) {
        fclose(f);
        return fail(B200_ERR_INVALID, "corrupt index header");
    }
    b200_index *ix = nullptr;
    int rc = b200_index_create(names[h.type], h.metric, h.d, "", &ix);
    if (rc != B200_OK) {
        fclose(f);
        return rc;
    }
    ix->nlist = h.nlist; ix->m = h.m; ix->dsub = h.dsub; ix->default_nprobe = h.default_nprobe; ix->refine_factor = h.refine_factor;
    ix->use_ivf = h.use_ivf != 0; ix->n = h.n;
    auto bail = [&](const std::string &msg) {
        fclose(f);
        b200_index_free(ix);
        return fail(B200_ERR_INVALID, msg);
    };
    const int raw_metric = h.metric == B200_METRIC_L2 ? B200_METRIC_L2 : B200_METRIC_IP;
    if (b200_corpus_create(raw_metric, B200_DTYPE_F32, h.d, h.n, &ix->raw) != B200_OK) return bail(b200_last_error());
    {
        const int64_t chunk = std::max<int64_t>(1, (64ll << 20) / ((int64_t)h.d * 4));
        std::vector<float> buf((size_t)chunk * h.d);
        for (int64_t off = 0; off < h.n; off += chunk) {
            const int64_t mrows = std::min(chunk, h.n - off);
            if (!rd(f, buf.data(), (size_t)mrows * h.d * 4)) return bail("truncated index file (rows)");
            if (b200_corpus_append(ix->raw, buf.data(), mrows) != B200_OK) return bail(b200_last_error());
        }
    }
    if (ix->use_ivf) {
        std::vector<char> tmp;
        auto slurp = [&](void **dptr, size_t bytes) {
            tmp.resize(bytes);
            if (!rd(f, tmp.data(), bytes)) return false;
            if (cudaMalloc(dptr, bytes + 256) != cudaSuccess) return false;
            return cudaMemcpy(*dptr, tmp.data(), bytes, cudaMemcpyHostToDevice) == cudaSuccess;
        };
        ix->list_off.resize(h.nlist + 1);
        if (!slurp((void **)&ix->d_centroids, (size_t)h.nlist * h.d * 4) || !rd(f, ix->list_off.data(), (size_t)(h.nlist + 1) * 4) ||
            !slurp((void **)&ix->d_list_ids, (size_t)h.n * 4))
            return bail("truncated index file (lists)");
        if (cudaMalloc(&ix->d_list_off, (size_t)(h.nlist + 1) * 4) != cudaSuccess ||
            cudaMemcpy(ix->d_list_off, ix->list_off.data(), (size_t)(h.nlist + 1) * 4, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail("cudaMalloc failed");
        ix->nlist = h.nlist;
        ix->n = h.n;
        set_biased_list_rows(ix);
        if (h.type == IDX_IVFPQ || h.type == IDX_MSTG)
            if (!slurp((void **)&ix->d_pq, (size_t)h.m * 256 * h.dsub * 4) || !slurp((void **)&ix->d_codes, (size_t)h.n * h.m))
                return bail("truncated index file (codes)");
        std::vector<float> hc((size_t)h.nlist * h.d);
        if (cudaMemcpy(hc.data(), ix->d_centroids, hc.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return bail("D2H failed");
        if (b200_corpus_create(B200_METRIC_L2, B200_DTYPE_F32, h.d, h.nlist, &ix->coarse) != B200_OK ||
            b200_corpus_append(ix->coarse, hc.data(), h.nlist) != B200_OK)
            return bail(b200_last_error());
    }
    fclose(f);
    ix->built = true;
    *out = ix;
    return B200_OK;
}
