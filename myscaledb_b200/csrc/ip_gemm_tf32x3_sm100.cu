// ip_gemm_tf32x3_sm100.cu -- K2b: batched fp32 queries x fp32 corpus on the tensor cores with
// fp32-level accuracy ("3xTF32"), top-k fused into the epilogue like ip_gemm_sm100.cu.
//
// Why: the reference stores vectors as Float32 (Array(Float32) columns) and serves batches
// (batch_distance(), Search::VectorIndex::search with nq > 1) through faiss' BLAS sgemm path
// (VectorIndex/Common/BruteForceSearch.h:77-88; knn_L2sqr / knn_inner_product for nx >= 20).
// A bf16 copy of the corpus would change the results; the CUDA-core scan re-reads the corpus
// once per 8 queries.  Here every fp32 value is split on the fly into two TF32 numbers
//     x = hi + lo,   hi = x with the low 13 mantissa bits cleared,  lo = tf32(x - hi)
// and  q.y = q_hi.y_hi + q_lo.y_hi + q_hi.y_lo  (+ q_lo.y_lo ~ 2^-22, dropped) is accumulated in
// fp32 by three tcgen05.mma.kind::tf32 per k-step: error ~2^-21 relative per product, the same
// class as an fp32 FMA chain in another summation order (tests/test_gpu_flat.py bounds it).
//
// One CTA PAIR (cluster of 2, cta_group::2) computes D[256 queries x 256 corpus rows]; each CTA
// stages its own 128 queries (hi and lo planes, pre-split once per batch by split_tf32_kernel)
// and HALF of the corpus tile as raw fp32 by TMA.  Warp roles per CTA (320 threads):
//   warp 0     TMA producer            warp 1     TMEM allocator + MMA issuer (leader CTA)
//   warps 2-5  top-k epilogue (one thread = one query = one TMEM lane, as in ip_gemm_sm100.cu)
//   warps 6-9  converters: rewrite the landed corpus k-block in place as y_hi and write y_lo to
//              a second buffer at the same (swizzled) offsets, fence.proxy.async, then signal the
//              leader's MMA warp.  16 KB per k-block per CTA: ~160 warp-instructions against
//              12 MMAs x 128 cycles, so the tensor pipe is the limiter (1/6 of the bf16 rate:
//              tf32 runs at half rate and there are three products).
// Smem ring: 3 stages x (q_hi 16 K + q_lo 16 K + y 16 K + y_lo 16 K) = 192 KB per CTA.
// Barriers per stage: full (leader; both CTAs' query planes landed), bfull (local; this CTA's
// corpus half landed), conv (leader; 2 x 4 converter warps done), empty (both; MMAs retired).
// Ranked key, filters, partial lists and the merge are those of ip_gemm_sm100.cu.
// Batches of up to 128 queries use the single-CTA form (CG = 1: M = 128, the whole 256-row corpus tile per CTA,
// 2 stages of 96 KB) so that no MMA row is spent on padding.
#include "gemm_common.cuh"

namespace b200 {
namespace gemm3 {

using namespace gemm;

constexpr int BK3 = 32;                       // fp32 per k-block = one 128-byte swizzle row
constexpr int UMMA_K3 = 8;                    // tf32 MMA K
constexpr int A3_BYTES = BM * BK3 * 4;        // 16 KB per plane
constexpr int CONV_WARPS = 4;
constexpr int NUM_THREADS3 = 64 + EPI_THREADS + CONV_WARPS * 32;  // 320
constexpr int SMEM_LIMIT = 232448;

// CG = 2: CTA pair, M = 256 (two query tiles), each CTA stages half of the 256-row corpus tile, 3 stages of 64 KB.
// CG = 1: one CTA, M = 128 (batches of up to 128 queries: a pair would waste half of its MMA rows), the whole
//         corpus tile per CTA, 2 stages of 96 KB.
template <int CG>
struct Cfg3 {
    static constexpr int STAGES = CG == 2 ? 3 : 2;
    static constexpr int B_ROWS = BN / CG;                  // corpus rows staged by one CTA
    static constexpr int B_BYTES = B_ROWS * BK3 * 4;        // 16 KB / 32 KB
    static constexpr int STAGE_BYTES = 2 * A3_BYTES + 2 * B_BYTES;
    static constexpr int OFF_AHI = 0;
    static constexpr int OFF_ALO = STAGES * A3_BYTES;
    static constexpr int OFF_B = 2 * STAGES * A3_BYTES;
    static constexpr int OFF_BLO = OFF_B + STAGES * B_BYTES;
    static constexpr int OFF_SIDE = STAGES * STAGE_BYTES;
    static constexpr int OFF_BAR = OFF_SIDE + 2 * BN * 4;
    static constexpr int OFF_SCRATCH = OFF_BAR + 256;
    static constexpr int OFF_LIST = OFF_SCRATCH + SCRATCH_BYTES;
    static constexpr int CONV_ITERS = B_BYTES / 16 / (CONV_WARPS * 32);  // 16-byte words per converter thread and stage
    static bool lists_fit(int k) { return OFF_LIST + k * EPI_THREADS * 8 + SMEM_ALIGN_SLACK <= SMEM_LIMIT; }
};

// kind::tf32, A = B = tf32 (K-major), D = f32, M = 128 * CG, N = 256
__device__ __forceinline__ constexpr uint32_t make_idesc_tf32(int cg) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * cg) >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32_cg1(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

// x = hi + lo with both parts exactly representable in TF32 (so the result does not depend on how the
// tensor core would round a raw fp32 operand).  hi by truncation: FLT_MAX (the reference's padding value for
// empty rows, MergeTreeVSManager.cpp:1380) stays finite.
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t &hi, uint32_t &lo) {
    hi = x & 0xffffe000u;
    const float r = __uint_as_float(x) - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__global__ void split_tf32_kernel(const float *__restrict__ src, int64_t n_src, int d_pad, float *__restrict__ hi,
                                  float *__restrict__ lo, int64_t n_pad) {
    const int64_t total = n_pad * d_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d_pad;
        uint32_t h = 0, l = 0;
        if (r < n_src) split_tf32(__float_as_uint(src[i]), h, l);
        hi[i] = __uint_as_float(h);
        lo[i] = __uint_as_float(l);
    }
}

template <int CG>
__global__ void __launch_bounds__(NUM_THREADS3, 1)
gemm3_topk_kernel(const __grid_constant__ CUtensorMap map_qhi, const __grid_constant__ CUtensorMap map_qlo,
                  const __grid_constant__ CUtensorMap map_c, const GemmTopkParams p) {
    using C = Cfg3<CG>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *sAhi = smem + C::OFF_AHI;
    unsigned char *sAlo = smem + C::OFF_ALO;
    unsigned char *sB = smem + C::OFF_B;
    unsigned char *sBlo = smem + C::OFF_BLO;
    float *side_scale = reinterpret_cast<float *>(smem + C::OFF_SIDE);
    float *side_bias = side_scale + BN;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
    uint64_t *bfull_bar = full_bar + STAGES;
    uint64_t *conv_bar = bfull_bar + STAGES;
    uint64_t *empty_bar = conv_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint64_t *tmem_empty_bar = tmem_full_bar + ACC_STAGES;
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;  // 0 / 1
    const uint32_t half = cta_rank & 1;
    const bool is_leader = half == 0;

    const int qt = blockIdx.x % p.q_tiles;
    const int worker = blockIdx.x / p.q_tiles;
    const int W = gridDim.x / p.q_tiles;
    const int64_t n_tiles = (p.n + BN - 1) / BN;
    const int kb_count = (p.d_pad + BK3 - 1) / BK3;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_qhi)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_qlo)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&bfull_bar[i], 1);
            mbar_init(&conv_bar[i], CG * CONV_WARPS);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < ACC_STAGES; i++) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 4 * CG);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                         "n"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                         "n"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer (every CTA) =====================
        int stage = 0;
        uint32_t phase = 0;
        int ordinal = 0;
        bool pacing = p.progress != nullptr;
        for (int64_t t = worker; t < n_tiles; t += W, ordinal++) {
            // pacing across the CTAs (pairs) that stream the same corpus tiles for other query tiles (see ip_gemm_sm100.cu)
            if (pacing && is_leader) {
                int ok = 1;
                if (lane == 0) {
                    volatile int *prog = p.progress + (size_t)worker * p.q_tiles;
                    prog[qt] = ordinal + 1;
                    int spins = 0;
                    for (int g = 0; g < p.q_tiles; g += CG)
                        while (prog[g] < ordinal + 1 - p.sync_slack && spins < 256) {
                            __nanosleep(200);
                            spins++;
                        }
                    if (spins >= 256) prog[qt] = 0x7fffffff;
                    ok = spins < 256;
                }
                pacing = __shfl_sync(0xffffffffu, ok, 0) != 0;
            }
            __syncwarp();
            for (int kb = 0; kb < kb_count; kb++) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (CG == 1) {
                        mbar_arrive_expect_tx(&full_bar[stage], 2 * A3_BYTES);
                        tma_load_2d(&map_qhi, &full_bar[stage], sAhi + stage * A3_BYTES, kb * BK3, qt * BM);
                        tma_load_2d(&map_qlo, &full_bar[stage], sAlo + stage * A3_BYTES, kb * BK3, qt * BM);
                    } else {
                        if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * 2 * A3_BYTES);
                        tma_load_2d_cg2(&map_qhi, &full_bar[stage], sAhi + stage * A3_BYTES, kb * BK3, qt * BM);
                        tma_load_2d_cg2(&map_qlo, &full_bar[stage], sAlo + stage * A3_BYTES, kb * BK3, qt * BM);
                    }
                    mbar_arrive_expect_tx(&bfull_bar[stage], C::B_BYTES);
                    tma_load_2d(&map_c, &bfull_bar[stage], sB + stage * C::B_BYTES, kb * BK3, (int)(t * BN + half * C::B_ROWS));
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (is_leader) {
            constexpr uint32_t idesc = make_idesc_tf32(CG);
            const uint64_t ahi0 = make_smem_desc(smem_u32(sAhi));
            const uint64_t alo0 = make_smem_desc(smem_u32(sAlo));
            const uint64_t b0 = make_smem_desc(smem_u32(sB));
            const uint64_t blo0 = make_smem_desc(smem_u32(sBlo));
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int64_t t = worker; t < n_tiles; t += W) {
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < kb_count; kb++) {
                    mbar_wait(&full_bar[stage], phase);   // query planes (of both CTAs)
                    mbar_wait(&conv_bar[stage], phase);   // corpus (halves) landed and split
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ahi = ahi0 + (uint64_t)(stage * (A3_BYTES >> 4));
                        const uint64_t alo = alo0 + (uint64_t)(stage * (A3_BYTES >> 4));
                        const uint64_t b = b0 + (uint64_t)(stage * (C::B_BYTES >> 4));
                        const uint64_t blo = blo0 + (uint64_t)(stage * (C::B_BYTES >> 4));
#pragma unroll
                        for (int k = 0; k < BK3 / UMMA_K3; k++) {
                            const uint64_t off = (uint64_t)(k * (UMMA_K3 * 4 >> 4));
                            const uint32_t acc0 = (kb | k) != 0 ? 1u : 0u;  // small terms first
                            if (CG == 1) {
                                umma_tf32_cg1(tmem_d, alo + off, b + off, idesc, acc0);
                                umma_tf32_cg1(tmem_d, ahi + off, blo + off, idesc, 1u);
                                umma_tf32_cg1(tmem_d, ahi + off, b + off, idesc, 1u);
                            } else {
                                umma_tf32_cg2(tmem_d, alo + off, b + off, idesc, acc0);
                                umma_tf32_cg2(tmem_d, ahi + off, blo + off, idesc, 1u);
                                umma_tf32_cg2(tmem_d, ahi + off, b + off, idesc, 1u);
                            }
                        }
                        if (CG == 1) {
                            umma_commit(&empty_bar[stage]);
                            if (kb == kb_count - 1) umma_commit(&tmem_full_bar[as]);
                        } else {
                            umma_commit_cg2(&empty_bar[stage], 3);
                            if (kb == kb_count - 1) umma_commit_cg2(&tmem_full_bar[as], 3);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++as == ACC_STAGES) {
                    as = 0;
                    aphase ^= 1;
                }
            }
        }
    } else if (warp >= 6) {
        // ===================== converters: y -> (y_hi in place, y_lo) =====================
        const int ct = threadIdx.x - (64 + EPI_THREADS);  // 0..127
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t t = worker; t < n_tiles; t += W) {
            for (int kb = 0; kb < kb_count; kb++) {
                mbar_wait(&bfull_bar[stage], phase);
                uint4 *b = reinterpret_cast<uint4 *>(sB + stage * C::B_BYTES);
                uint4 *bl = reinterpret_cast<uint4 *>(sBlo + stage * C::B_BYTES);
#pragma unroll
                for (int i0 = 0; i0 < C::CONV_ITERS; i0 += 8) {
                    uint4 raw[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) raw[i] = b[ct + (i0 + i) * (CONV_WARPS * 32)];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        uint4 h, l;
                        split_tf32(raw[i].x, h.x, l.x);
                        split_tf32(raw[i].y, h.y, l.y);
                        split_tf32(raw[i].z, h.z, l.z);
                        split_tf32(raw[i].w, h.w, l.w);
                        b[ct + (i0 + i) * (CONV_WARPS * 32)] = h;
                        bl[ct + (i0 + i) * (CONV_WARPS * 32)] = l;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA's reads
                __syncwarp();
                if (lane == 0) {
                    if (is_leader) mbar_arrive(&conv_bar[stage]);
                    else mbar_arrive_remote(&conv_bar[stage], 0);
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue: fused top-k (same scheme as ip_gemm_sm100.cu) =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int et = threadIdx.x - 64;
        const bool use_side = p.row_scale || p.row_bias || p.alive || p.scale_const != -1.f;
        float *scratch = reinterpret_cast<float *>(smem + C::OFF_SCRATCH) + et;
        ThreadTopK list;
        list.n = 0;
        list.worst = 0;
        // rows past the batch (zero padding up to the tile size) must never pay for the slow path: nothing beats -FLT_MAX
        list.thr_key = (qt * BM + row < p.nq_valid) ? FLT_MAX : -FLT_MAX;
        list.thr_id = 0;
        if (p.lists_in_smem)
            list_bind(list, reinterpret_cast<float *>(smem + C::OFF_LIST),
                      reinterpret_cast<uint32_t *>(smem + C::OFF_LIST + (size_t)p.list_cap * EPI_THREADS * 4), row, p.k, p.list_cap);
        else
            list_bind(list, p.list_keys_gmem + (size_t)blockIdx.x * p.list_cap * EPI_THREADS,
                      p.list_ids_gmem + (size_t)blockIdx.x * p.list_cap * EPI_THREADS, row, p.k, p.list_cap);
        int as = 0;
        uint32_t aphase = 0;
        for (int64_t t = worker; t < n_tiles; t += W) {
            const int64_t n0 = t * BN;
            if (use_side) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int c = et; c < BN; c += EPI_THREADS) {
                    const int64_t r = n0 + c;
                    bool ok = r < p.n;
                    if (ok && p.alive) ok = (p.alive[r >> 3] >> (r & 7)) & 1;
                    side_scale[c] = ok ? (p.row_scale ? p.row_scale[r] : p.scale_const) : 0.f;
                    side_bias[c] = ok ? (p.row_bias ? p.row_bias[r] : 0.f) : __int_as_float(0x7f800000);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
            const bool tail = n0 + BN > p.n;
            float va[32], vb[32];
            __syncwarp();
            tmem_ld32_issue(taddr, va);
            tmem_ld_wait();
#pragma unroll 1
            for (int chunk = 0; chunk < BN / 32; chunk += 2) {
                __syncwarp();
                tmem_ld32_issue(taddr + (chunk + 1) * 32, vb);
                epilogue_chunk(list, va, use_side, side_scale + chunk * 32, side_bias + chunk * 32, (uint32_t)(n0 + chunk * 32),
                               tail, p.n, scratch);
                tmem_ld_wait();
                __syncwarp();
                if (chunk + 2 < BN / 32) tmem_ld32_issue(taddr + (chunk + 2) * 32, va);
                epilogue_chunk(list, vb, use_side, side_scale + (chunk + 1) * 32, side_bias + (chunk + 1) * 32,
                               (uint32_t)(n0 + (chunk + 1) * 32), tail, p.n, scratch);
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (is_leader) mbar_arrive(&tmem_empty_bar[as]);
                else mbar_arrive_remote(&tmem_empty_bar[as], 0);
            }
            if (++as == ACC_STAGES) {
                as = 0;
                aphase ^= 1;
            }
        }
        float *ok = p.part_keys + ((size_t)blockIdx.x * BM + row) * p.k;
        uint32_t *oi = p.part_ids + ((size_t)blockIdx.x * BM + row) * p.k;
        list_publish(list, ok, oi);
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// 2-D fp32 tensor map over row-major [rows][d_pad], box = [box_rows][32 floats], 128-byte swizzle; the last
// k-block may hang over d_pad (TMA zero-fills), so fp32 corpora keep their 16-byte row padding
static bool encode_rows_map_f32(CUtensorMap *map, const void *base, int64_t rows, int d_pad, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)d_pad * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK3, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gemm3

cudaError_t launch_split_tf32(const float *src, int64_t n_src, int d_pad, float *hi, float *lo, int64_t n_pad, cudaStream_t s) {
    if (n_pad == 0) return cudaSuccess;
    int64_t b = ceil_div(n_pad * d_pad, 256);
    if (b > 148 * 8) b = 148 * 8;
    gemm3::split_tf32_kernel<<<(int)b, 256, 0, s>>>(src, n_src, d_pad, hi, lo, n_pad);
    g_launches++;
    return cudaGetLastError();
}

template <int CG>
static cudaError_t launch3_cg(const CUtensorMap &map_qhi, const CUtensorMap &map_qlo, const CUtensorMap &map_c,
                              const GemmTopkParams &p_in, int grid, cudaStream_t s) {
    using namespace gemm3;
    using C = Cfg3<CG>;
    GemmTopkParams p = p_in;
    p.list_cap = list_cap_for(p.k);
    p.lists_in_smem = C::lists_fit(p.list_cap) ? 1 : 0;
    p.stages = C::STAGES;
    const size_t smem = (size_t)C::OFF_LIST + (p.lists_in_smem ? (size_t)p.list_cap * EPI_THREADS * 8 : 0) + SMEM_ALIGN_SLACK;
    auto kern = gemm3_topk_kernel<CG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS3);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, map_qhi, map_qlo, map_c, p);
    g_launches++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_gemm3_topk(const GemmTopkParams &p, int grid, cudaStream_t s, const char **err_detail) {
    using namespace gemm3;
    *err_detail = nullptr;
    const int cg = p.cta_group == 2 ? 2 : 1;
    if (p.q_tiles % cg != 0 || grid % p.q_tiles != 0 || !p.queries_lo) {
        *err_detail = "gemm3: q_tiles must be a multiple of cta_group, grid a multiple of q_tiles, queries_lo set";
        return cudaErrorInvalidValue;
    }
    CUtensorMap map_qhi, map_qlo, map_c;
    if (!encode_rows_map_f32(&map_qhi, p.queries_bf16, p.nq_pad, p.d_pad, BM) ||
        !encode_rows_map_f32(&map_qlo, p.queries_lo, p.nq_pad, p.d_pad, BM) ||
        !encode_rows_map_f32(&map_c, p.corpus_bf16, p.n, p.d_pad, BN / cg)) {
        *err_detail = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
    }
    return cg == 2 ? launch3_cg<2>(map_qhi, map_qlo, map_c, p, grid, s) : launch3_cg<1>(map_qhi, map_qlo, map_c, p, grid, s);
}

}  // namespace b200
