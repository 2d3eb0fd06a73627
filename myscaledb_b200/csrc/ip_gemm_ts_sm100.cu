// ip_gemm_ts_sm100.cu -- K2, "queries stationary in tensor memory" form of the batched inner
// product + fused top-k (see ip_gemm_sm100.cu for the operand-streaming form and the epilogue).
//
// Why: profiles/r01_summary.md -- after HBM traffic was brought to the algorithmic minimum the
// streaming form was bound by L2->SM movement: 123 GB per launch, HALF of it the query tiles that
// every CTA re-streams for every corpus tile (a CTA's queries never change).  Here each CTA of a
// pair writes its 128 queries ONCE into TMEM (tcgen05.st, 128 lanes x d/2 columns of packed bf16
// pairs) and issues tcgen05.mma with the A operand taken from TMEM (".ts" form,
// cta_group::2, M = 256).  Only corpus rows flow through shared memory:
//     L2->SM bytes per CTA per 64-wide k-block: 32 KB  ->  BN/2 rows * 128 B  (4 KB at BN = 64)
// TMEM budget (512 columns): d_pad/2 columns for the queries, the rest for a 2-stage accumulator
// ring of BN columns each: d_pad <= 512 -> BN = 128, d_pad <= 768 -> BN = 64; wider vectors use
// the streaming kernel.
//
// Roles: warp 0 TMA producer (both CTAs, each loads its half of the corpus tile), warp 1 TMEM
// allocator + MMA issuer (leader CTA), warps 2..5 load the queries into TMEM, then run the
// top-k epilogue for their own 128 lanes.  A smem stage holds TWO 64-wide k-blocks so that the
// issuing lane has 8 MMAs per barrier round trip.
#include "gemm_common.cuh"

namespace b200 {
namespace gemm {

constexpr int TS_KB_PER_STAGE = 2;

template <int TBN>
struct TsCfg {
    static constexpr int ROWS = TBN / 2;                            // corpus rows staged by one CTA per tile
    static constexpr int KB_BYTES = ROWS * BK * 2;                  // one k-block of one CTA
    static constexpr int STAGE_BYTES = KB_BYTES * TS_KB_PER_STAGE;  // 8 KB (BN=64) / 16 KB (BN=128)
    static constexpr int STAGES = TBN == 64 ? 16 : 10;
    static constexpr int OFF_B = 0;
    static constexpr int OFF_SIDE = OFF_B + STAGES * STAGE_BYTES;   // scale[TBN], bias[TBN]
    static constexpr int OFF_BAR = OFF_SIDE + 2 * TBN * 4;
    static constexpr int OFF_SCRATCH = OFF_BAR + 512;
    static constexpr int OFF_LIST = OFF_SCRATCH + SCRATCH_BYTES;
};

__device__ __forceinline__ void umma_ts_cg2(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// kind::f16, A (TMEM) = B (smem, K-major) = bf16, D = f32, M = 256 (pair), N = n
__device__ __forceinline__ constexpr uint32_t make_idesc_ts(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int TBN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_topk_ts_kernel(const __grid_constant__ CUtensorMap map_c, const GemmTopkParams p) {
    using C = TsCfg<TBN>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *sB = smem + C::OFF_B;
    float *side_scale = reinterpret_cast<float *>(smem + C::OFF_SIDE);
    float *side_bias = side_scale + TBN;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint64_t *tmem_empty_bar = tmem_full_bar + ACC_STAGES;
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;

    const int qt = blockIdx.x % p.q_tiles;      // this CTA's query tile (pair = qt, qt ^ 1)
    const int worker = blockIdx.x / p.q_tiles;
    const int W = gridDim.x / p.q_tiles;
    const int64_t n_tiles = (p.n + TBN - 1) / TBN;
    const int kb_count = p.d_pad / BK;
    const int st_count = (kb_count + TS_KB_PER_STAGE - 1) / TS_KB_PER_STAGE;  // smem stages per tile
    const uint32_t a_cols = (uint32_t)(p.d_pad / 2);                          // TMEM columns holding the queries

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < ACC_STAGES; i++) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 4 * 2);  // one arrival per epilogue warp of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    // ---- queries -> TMEM: thread (quarter, lane) owns TMEM lane = query row; 32 columns = 64 bf16
    if (warp >= 2) {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(p.queries_bf16) +
                                                           ((size_t)qt * BM + row) * p.d_pad);
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        for (uint32_t c0 = 0; c0 < a_cols; c0 += 32) {
            uint32_t r[32];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint4 v = src[c0 / 4 + j];  // 4 columns (8 bf16) per 16-byte load
                r[4 * j + 0] = v.x;
                r[4 * j + 1] = v.y;
                r[4 * j + 2] = v.z;
                r[4 * j + 3] = v.w;
            }
            __syncwarp();
            tmem_st32(tlane + c0, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();  // both CTAs' queries are in TMEM, barriers initialised cluster-wide
    tc_fence_after();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        int64_t ordinal = 0;
        bool pacing = p.progress != nullptr;
        constexpr int TILES_PER_UNIT = 256 / TBN;  // pacing granularity: 256 corpus rows
        for (int64_t t = worker; t < n_tiles; t += W, ordinal++) {
            if (pacing && is_leader && (ordinal % TILES_PER_UNIT) == 0) {
                const int unit = (int)(ordinal / TILES_PER_UNIT);
                int ok = 1;
                if (lane == 0) {
                    volatile int *prog = p.progress + (size_t)worker * p.q_tiles;
                    prog[qt] = unit + 1;
                    int spins = 0;
                    for (int g = 0; g < p.q_tiles; g += 2)
                        while (prog[g] < unit + 1 - p.sync_slack && spins < 256) {
                            __nanosleep(200);
                            spins++;
                        }
                    if (spins >= 256) prog[qt] = 0x7fffffff;
                    ok = spins < 256;
                }
                pacing = __shfl_sync(0xffffffffu, ok, 0) != 0;
            }
            __syncwarp();
            for (int s = 0; s < st_count; s++) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    const int kb0 = s * TS_KB_PER_STAGE;
                    const int nkb = min(TS_KB_PER_STAGE, kb_count - kb0);
                    // the leader's barrier collects both CTAs' bytes
                    if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(nkb * C::KB_BYTES * 2));
                    for (int i = 0; i < nkb; i++)
                        tma_load_2d_cg2(&map_c, &full_bar[stage], sB + stage * C::STAGE_BYTES + i * C::KB_BYTES,
                                        (kb0 + i) * BK, (int)(t * TBN + cta_rank * C::ROWS));
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (is_leader) {
            constexpr uint32_t idesc = make_idesc_ts(TBN);
            const uint64_t bdesc0 = make_smem_desc(smem_u32(sB));
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int64_t t = worker; t < n_tiles; t += W) {
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + a_cols + (uint32_t)(as * TBN);
                for (int s = 0; s < st_count; s++) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const int kb0 = s * TS_KB_PER_STAGE;
                        const int nkb = min(TS_KB_PER_STAGE, kb_count - kb0);
                        const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (C::STAGE_BYTES >> 4));
                        for (int i = 0; i < nkb; i++) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; k++) {
                                const uint32_t acc = ((kb0 + i) | k) != 0 ? 1u : 0u;
                                // A: 16 bf16 = 8 TMEM columns per k-step
                                umma_ts_cg2(tmem_d, tmem_base + (uint32_t)((kb0 + i) * (BK / 2) + k * (UMMA_K / 2)),
                                            bdesc + (uint64_t)(i * (C::KB_BYTES >> 4) + k * (UMMA_K * 2 >> 4)), idesc, acc);
                            }
                        }
                        umma_commit_cg2(&empty_bar[stage]);
                        if (s == st_count - 1) umma_commit_cg2(&tmem_full_bar[as]);
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
    } else {
        // ===================== epilogue: fused top-k =====================
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
            const int64_t n0 = t * TBN;
            if (use_side) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int c = et; c < TBN; c += EPI_THREADS) {
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
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a_cols + (uint32_t)(as * TBN);
            const bool tail = n0 + TBN > p.n;
            if (!(p.debug & 1)) {
                float va[32], vb[32];
                __syncwarp();
                tmem_ld32_issue(taddr, va);
                tmem_ld_wait();
#pragma unroll 1
                for (int chunk = 0; chunk < TBN / 32; chunk += 2) {
                    __syncwarp();
                    tmem_ld32_issue(taddr + (chunk + 1) * 32, vb);
                    epilogue_chunk(list, va, use_side, side_scale + chunk * 32, side_bias + chunk * 32,
                                   (uint32_t)(n0 + chunk * 32), tail, p.n, scratch);
                    tmem_ld_wait();
                    __syncwarp();
                    if (chunk + 2 < TBN / 32) tmem_ld32_issue(taddr + (chunk + 2) * 32, va);
                    epilogue_chunk(list, vb, use_side, side_scale + (chunk + 1) * 32, side_bias + (chunk + 1) * 32,
                                   (uint32_t)(n0 + (chunk + 1) * 32), tail, p.n, scratch);
                    tmem_ld_wait();
                }
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
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

template <int TBN>
static cudaError_t launch_ts(const CUtensorMap &map_c, const GemmTopkParams &p_in, int grid, cudaStream_t s) {
    GemmTopkParams p = p_in;
    p.list_cap = list_cap_for(p.k);
    p.lists_in_smem = p.list_cap <= 64 ? 1 : 0;
    size_t smem = TsCfg<TBN>::OFF_LIST + SMEM_ALIGN_SLACK;
    if (p.lists_in_smem) smem += (size_t)p.list_cap * EPI_THREADS * 8;
    cudaError_t e = cudaFuncSetAttribute(gemm_topk_ts_kernel<TBN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, gemm_topk_ts_kernel<TBN>, map_c, p);
    g_launches++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace gemm

bool gemm_topk_ts_supported(int d_pad, int q_tiles) { return d_pad <= 768 && q_tiles >= 2 && q_tiles % 2 == 0; }
int gemm_topk_ts_tile_rows(int d_pad) { return d_pad <= 512 ? 128 : 64; }

cudaError_t launch_gemm_topk_ts(const GemmTopkParams &p, int grid, cudaStream_t s, const char **err_detail) {
    *err_detail = nullptr;
    if (!gemm_topk_ts_supported(p.d_pad, p.q_tiles) || grid % p.q_tiles != 0) {
        *err_detail = "TS kernel needs d_pad <= 768, an even number of query tiles and grid % q_tiles == 0";
        return cudaErrorInvalidValue;
    }
    const int tbn = gemm_topk_ts_tile_rows(p.d_pad);
    CUtensorMap map_c;
    if (!gemm::encode_rows_map(&map_c, p.corpus_bf16, p.n, p.d_pad, tbn / 2)) {
        *err_detail = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
    }
    return tbn == 128 ? gemm::launch_ts<128>(map_c, p, grid, s) : gemm::launch_ts<64>(map_c, p, grid, s);
}

}  // namespace b200
