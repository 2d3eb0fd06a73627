// ip_gemm_sm100.cu -- K2: batched multi-query x corpus inner product as a dense bf16 GEMM
// on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA),
// with the top-k selection FUSED into the epilogue: the [nq x N] score matrix is never
// written to memory.
//
// Replaces faiss::knn_inner_product / knn_L2sqr for nx >= 20 (the BLAS sgemm path) reached
// from tryBruteForceSearch (reference: VectorIndex/Common/BruteForceSearch.h:77-88) and the
// FLAT Search::VectorIndex::search scan (VectorIndex/Common/VIWithDataPart.cpp:926).
//
// Shape of one CTA tile: D[128 queries (TMEM lanes) x 256 corpus rows (TMEM columns)],
// K-loop over d in blocks of 64 bf16 (= one 128-byte swizzle atom row).
//   warp 0      : TMA producer (one lane)        -- A = query tile, B = corpus tile
//   warp 1      : TMEM allocator + MMA issuer (one lane), tcgen05.mma cta_group::1 kind::f16
//   warps 2..5  : epilogue; thread t owns TMEM lane t = one query, streams the 256 scores of
//                 the tile through a register threshold test and keeps a private sorted
//                 top-k in shared memory (k <= 30) or global scratch (larger k)
// Pipelines: smem ring (full/empty mbarriers, 4 stages of 48 KB), 2-stage TMEM accumulator ring.
// CG = 2 variant (used whenever there are >= 2 query tiles): a CTA PAIR (cluster of 2,
// tcgen05.mma cta_group::2) computes D[256 queries x 256 rows]; each CTA stages its own 128
// queries and HALF of the corpus tile (6 stages of 32 KB), which takes the operand traffic
// through shared memory from 192 B/clk (the measured 67 % tensor-pipe ceiling of the 1-CTA
// form, profiles/r01_gemm_topk_cg1.md) to 128 B/clk per SM and the L2->SM traffic from 48 KB
// to 32 KB per k-block.  The leader CTA issues the MMAs; both CTAs run a TMA producer and the
// top-k epilogue for their own 128 TMEM lanes.
// Each CTA owns ONE query tile for its whole life (blockIdx % q_tiles) and walks the corpus
// tiles worker, worker+W, ...; thresholds therefore live in registers for the whole kernel.
// The key that is ranked is  acc * row_scale[j] + row_bias[j]  (smaller = better):
//   IP: -acc | L2: ||y||^2 - 2 acc (+||q||^2 added at merge) | cosine: -acc / ||y||
//   filtered / out-of-range rows: scale 0, bias +inf.
//
// Tensor-bound: 2 * 128 * 256 * d FLOP per tile; algorithmic HBM bytes = N * d * 2 once.
#include <cstdlib>

#include "gemm_common.cuh"

namespace b200 {
namespace gemm {

// MC = CTA pairs per cluster that consume the SAME corpus tile (different query tiles): the tile is
// fetched from L2 once per cluster, each CTA loading 1/MC of its half and multicasting it.  The
// measured limiter of the MC = 1 form is L2->SM bandwidth (~9.8 TB/s, profiles/r01_summary.md).
template <int CG, int MC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c,
                 const GemmTopkParams p) {
    using C = Cfg<CG>;
    const int STAGES = p.stages;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = smem + C::off_a();
    unsigned char *sB = smem + C::off_b(STAGES);
    float *side_scale = reinterpret_cast<float *>(smem + C::off_side(STAGES));
    float *side_bias = side_scale + BN;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + C::off_bar(STAGES));
    uint64_t *empty_bar = full_bar + MAX_STAGES;
    uint64_t *tmem_full_bar = empty_bar + MAX_STAGES;
    uint64_t *tmem_empty_bar = tmem_full_bar + ACC_STAGES;
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;   // 0 .. CG * MC - 1
    const uint32_t half = cta_rank & 1;                           // which half of the pair's M = 256 / N = 256
    const uint32_t pair_in_cluster = cta_rank >> 1;
    const uint32_t leader_rank = cta_rank & ~1u;
    const bool is_leader = half == 0;                             // leader of its pair: issues the MMAs
    const bool paces = cta_rank == 0;                             // one CTA per cluster talks to the pacing counters

    // tile schedule: a CTA (CG=1) or CTA pair (CG=2) owns CG query tiles for its whole life and
    // walks the corpus tiles worker, worker + W, ...  (blockIdx = worker * q_tiles + qt)
    const int qt = blockIdx.x % p.q_tiles;
    const int worker = blockIdx.x / p.q_tiles;
    const int W = gridDim.x / p.q_tiles;  // the host launches a multiple of q_tiles CTAs
    const int64_t n_tiles = (p.n + BN - 1) / BN;
    const int kb_count = p.d_pad / BK;
    const int KPS = p.kps;            // k-blocks per barrier stage (1 or 2; kb_count % KPS == 0)
    const int NST = STAGES / KPS;     // barrier stages

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], MC);  // one commit per pair whose operands live in (or were sent from) this CTA
        }
        for (int i = 0; i < ACC_STAGES; i++) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 4 * CG);  // leader's: one arrival per epilogue warp of both CTAs
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
        // ===================== TMA producer (both CTAs of a pair) =====================
        // the whole warp walks the loop (uniform control flow); one elected lane issues
        int stage = 0;
        uint32_t phase = 0;
        int ordinal = 0;
        bool pacing = p.progress != nullptr;
        for (int64_t t = worker; t < n_tiles; t += W, ordinal++) {
            // Pacing: the CTAs (pairs) that stream the SAME corpus tiles for different query tiles
            // stay within `sync_slack` tiles of each other, so a tile is fetched from HBM once and
            // served to the others from L2 (without it ncu shows 2.9-3.6x the algorithmic DRAM
            // bytes, profiles/r01_*).  Only pacing, no data dependency: plain volatile counters.
            // Bounded wait (~50 us): if the sharers are not co-resident (another kernel holds SMs)
            // pacing is dropped instead of risking a co-residency deadlock.
            if (pacing && paces) {
                int ok = 1;
                if (lane == 0) {
                    volatile int *prog = p.progress + (size_t)worker * p.q_tiles;
                    prog[qt] = ordinal + 1;
                    int spins = 0;
                    for (int g = 0; g < p.q_tiles; g += CG * MC)
                        while (prog[g] < ordinal + 1 - p.sync_slack && spins < 256) {
                            __nanosleep(200);
                            spins++;
                        }
                    if (spins >= 256) prog[qt] = 0x7fffffff;  // never hold anybody back again
                    ok = spins < 256;
                }
                pacing = __shfl_sync(0xffffffffu, ok, 0) != 0;
            }
            __syncwarp();
            // KPS k-blocks share one full / empty barrier pair (slot = stage * KPS + sub): half the barrier round trips
            // and polls of the producer and issuer warps at KPS = 2 for the same bytes in flight
            for (int kb0 = 0; kb0 < kb_count; kb0 += KPS) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (p.debug & 4) {  // experiment: no TMA traffic, operands are whatever is in smem
                    if (is_leader && elect_one()) mbar_arrive(&full_bar[stage]);
                } else if (elect_one()) {
                    if (CG == 1 || is_leader) mbar_arrive_expect_tx(&full_bar[stage], C::TX_BYTES * KPS);
                    for (int sub = 0; sub < KPS; sub++) {
                        const int kb = kb0 + sub, slot = stage * KPS + sub;
                        if (CG == 1) {
                            tma_load_2d(&map_q, &full_bar[stage], sA + slot * A_BYTES, kb * BK, qt * BM);
                            tma_load_2d(&map_c, &full_bar[stage], sB + slot * C::B_BYTES, kb * BK, (int)(t * BN));
                        } else {
                            tma_load_2d_cg2(&map_q, &full_bar[stage], sA + slot * A_BYTES, kb * BK, qt * BM);
                            if (MC == 1) {
                                tma_load_2d_cg2(&map_c, &full_bar[stage], sB + slot * C::B_BYTES, kb * BK,
                                                (int)(t * BN + half * C::B_ROWS));
                            } else {
                                // this CTA fetches slice `pair_in_cluster` of its half and multicasts it to the CTAs of
                                // the same half in every pair of the cluster (ranks half, half + 2, ...)
                                constexpr int SLICE_ROWS = C::B_ROWS / MC;
                                uint16_t mask = 0;
                                for (int pp = 0; pp < MC; pp++) mask |= (uint16_t)(1u << (pp * 2 + half));
                                tma_load_2d_cg2_mc(&map_c, &full_bar[stage],
                                                   sB + slot * C::B_BYTES + pair_in_cluster * (SLICE_ROWS * BK * 2), kb * BK,
                                                   (int)(t * BN + half * C::B_ROWS + pair_in_cluster * SLICE_ROWS), mask);
                            }
                        }
                    }
                }
                __syncwarp();
                if (++stage == NST) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (is_leader) {
            constexpr uint32_t idesc = make_idesc(CG);
            // descriptors of stage 0; other stages / k-steps are plain adds on the 14-bit
            // (address >> 4) field, which cannot carry: shared addresses stay below 2^18
            const uint64_t adesc0 = make_smem_desc(smem_u32(sA));
            const uint64_t bdesc0 = make_smem_desc(smem_u32(sB));
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            for (int64_t t = worker; t < n_tiles; t += W) {
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                for (int kb0 = 0; kb0 < kb_count; kb0 += KPS) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        for (int sub = 0; sub < KPS; sub++) {
                            const int slot = stage * KPS + sub;
                            const uint64_t adesc = adesc0 + (uint64_t)(slot * (A_BYTES >> 4));
                            const uint64_t bdesc = bdesc0 + (uint64_t)(slot * (C::B_BYTES >> 4));
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; k++) {
                                const uint32_t acc = (kb0 | sub | k) != 0 ? 1u : 0u;
                                if (CG == 1) umma(tmem_d, adesc + k * (UMMA_K * 2 >> 4), bdesc + k * (UMMA_K * 2 >> 4), idesc, acc);
                                else umma_cg2(tmem_d, adesc + k * (UMMA_K * 2 >> 4), bdesc + k * (UMMA_K * 2 >> 4), idesc, acc);
                            }
                        }
                        // smem slots free (in both CTAs) once these MMAs retire
                        if (CG == 1) umma_commit(&empty_bar[stage]);
                        else umma_commit_cg2(&empty_bar[stage], (uint16_t)((1u << (CG * MC)) - 1));  // every CTA of the cluster
                        // accumulator ready for the epilogue (of both CTAs)
                        if (kb0 + KPS >= kb_count) {
                            if (CG == 1) umma_commit(&tmem_full_bar[as]);
                            else umma_commit_cg2(&tmem_full_bar[as], (uint16_t)(3u << leader_rank));  // own pair
                        }
                    }
                    __syncwarp();
                    if (++stage == NST) {
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
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;          // query row inside the tile
        const int et = threadIdx.x - 64;              // 0..127 among epilogue threads
        const bool use_side = p.row_scale || p.row_bias || p.alive || p.scale_const != -1.f;
        float *scratch = reinterpret_cast<float *>(smem + C::off_scratch(STAGES)) + et;
        ThreadTopK list;
        list.n = 0;
        list.worst = 0;
        // rows past the batch (zero padding up to the tile size) must never pay for the slow path: nothing beats -FLT_MAX
        list.thr_key = (qt * BM + row < p.nq_valid) ? FLT_MAX : -FLT_MAX;
        list.thr_id = 0;
        if (p.lists_in_smem)
            list_bind(list, reinterpret_cast<float *>(smem + C::off_list(STAGES)),
                      reinterpret_cast<uint32_t *>(smem + C::off_list(STAGES) + (size_t)p.list_cap * EPI_THREADS * 4), row, p.k, p.list_cap);
        else
            list_bind(list, p.list_keys_gmem + (size_t)blockIdx.x * p.list_cap * EPI_THREADS,
                      p.list_ids_gmem + (size_t)blockIdx.x * p.list_cap * EPI_THREADS, row, p.k, p.list_cap);
        int as = 0;
        uint32_t aphase = 0;
        for (int64_t t = worker; t < n_tiles; t += W) {
            const int64_t n0 = t * BN;
            if (use_side) {
                asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's readers done
                for (int c = et; c < BN; c += EPI_THREADS) {
                    const int64_t r = n0 + c;
                    bool ok = r < p.n;
                    if (ok && p.alive) ok = (p.alive[r >> 3] >> (r & 7)) & 1;
                    side_scale[c] = ok ? (p.row_scale ? p.row_scale[r] : p.scale_const) : 0.f;
                    side_bias[c] = ok ? (p.row_bias ? p.row_bias[r] : 0.f) : __int_as_float(0x7f800000);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
#if defined(B200_EPI_SINGLE_POLL) && B200_EPI_SINGLE_POLL
            // experiment (make EXTRA=-DB200_EPI_SINGLE_POLL=1): one lane per epilogue warp polls the accumulator barrier
            if (lane == 0) mbar_wait(&tmem_full_bar[as], aphase);
            __syncwarp();
#else
            mbar_wait(&tmem_full_bar[as], aphase);
#endif
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN);
            // Chunks of 32 columns; the TMEM load of chunk c+1 is in flight while chunk c is
            // filtered.  A chunk is first reduced to its best key (31 FMNMX); the per-element
            // test only runs for the rare chunk that can beat the current k-th key.
            const bool tail = n0 + BN > p.n;
            if (p.debug & 1) {  // experiment: epilogue does nothing
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 1 || is_leader) mbar_arrive(&tmem_empty_bar[as]);
                    else mbar_arrive_remote(&tmem_empty_bar[as], leader_rank);
                }
                if (++as == ACC_STAGES) {
                    as = 0;
                    aphase ^= 1;
                }
                continue;
            }
            float va[32], vb[32];
            __syncwarp();
            tmem_ld32_issue(taddr, va);
            tmem_ld_wait();
#pragma unroll 1
            for (int chunk = 0; chunk < BN / 32; chunk += 2) {
                __syncwarp();
                tmem_ld32_issue(taddr + (chunk + 1) * 32, vb);
                if (!(p.debug & 2))
                    epilogue_chunk(list, va, use_side, side_scale + chunk * 32, side_bias + chunk * 32,
                                   (uint32_t)(n0 + chunk * 32), tail, p.n, scratch);
                else if (va[3] == 12345.678f) list.n = 0;
                tmem_ld_wait();
                __syncwarp();
                if (chunk + 2 < BN / 32) tmem_ld32_issue(taddr + (chunk + 2) * 32, va);
                if (!(p.debug & 2))
                    epilogue_chunk(list, vb, use_side, side_scale + (chunk + 1) * 32, side_bias + (chunk + 1) * 32,
                                   (uint32_t)(n0 + (chunk + 1) * 32), tail, p.n, scratch);
                else if (vb[3] == 12345.678f) list.n = 0;
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 1 || is_leader) mbar_arrive(&tmem_empty_bar[as]);
                else mbar_arrive_remote(&tmem_empty_bar[as], leader_rank);
            }
            if (++as == ACC_STAGES) {
                as = 0;
                aphase ^= 1;
            }
        }
        // publish this CTA's per-query partial list
        float *ok = p.part_keys + ((size_t)blockIdx.x * BM + row) * p.k;
        uint32_t *oi = p.part_ids + ((size_t)blockIdx.x * BM + row) * p.k;
        list_publish(list, ok, oi);
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();  // nobody leaves while the peer may still signal us
    if (warp == 1) {
        tc_fence_after();
        if (CG == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

template <int CG, int MC>
static cudaError_t launch_cg(const CUtensorMap &map_q, const CUtensorMap &map_c, const GemmTopkParams &p_in, int grid,
                             cudaStream_t s) {
    GemmTopkParams p = p_in;
    // Per-thread lists sit in shared memory whenever they fit, even when that squeezes the operand ring: every insert rescans
    // the list, and from global scratch that is k L2 round trips (measured at k = 64 / 100: 22 / 36 ms per launch with the lists
    // in shared memory, 72 / 126 ms in global scratch; B200_GEMM_LIST_SMEM_MIN_STAGES raises the bar for experiments).
    static const int min_stages = getenv("B200_GEMM_LIST_SMEM_MIN_STAGES") ? atoi(getenv("B200_GEMM_LIST_SMEM_MIN_STAGES")) : 2;
    p.list_cap = list_cap_for(p.k);
    p.lists_in_smem = (Cfg<CG>::lists_fit(p.list_cap) && Cfg<CG>::stages_for(p.list_cap) >= min_stages) ? 1 : 0;
    const int k_smem = p.lists_in_smem ? p.list_cap : 0;
    p.stages = Cfg<CG>::stages_for(k_smem);
    {
        static const int env_kps = getenv("B200_GEMM_KPS") ? atoi(getenv("B200_GEMM_KPS")) : 1;
        p.kps = (env_kps == 2 && (p.d_pad / BK) % 2 == 0 && p.stages >= 4) ? 2 : 1;
    }
    const size_t smem = (size_t)Cfg<CG>::off_list(p.stages) + (size_t)k_smem * EPI_THREADS * 8 + SMEM_ALIGN_SLACK;
    auto kern = gemm_topk_kernel<CG, MC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG * MC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, map_q, map_c, p);
    g_launches++;
    return e != cudaSuccess ? e : cudaGetLastError();
}

// how many clusters of CG * MC CTAs of this kernel can be co-resident (persistent grid upper bound)
template <int CG, int MC>
static int max_clusters(int k) {
    static const int min_stages = getenv("B200_GEMM_LIST_SMEM_MIN_STAGES") ? atoi(getenv("B200_GEMM_LIST_SMEM_MIN_STAGES")) : 2;
    const int cap = list_cap_for(k);
    const int k_smem = (Cfg<CG>::lists_fit(cap) && Cfg<CG>::stages_for(cap) >= min_stages) ? cap : 0;
    const int stages = Cfg<CG>::stages_for(k_smem);
    const size_t smem = (size_t)Cfg<CG>::off_list(stages) + (size_t)k_smem * EPI_THREADS * 8 + SMEM_ALIGN_SLACK;
    auto kern = gemm_topk_kernel<CG, MC>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CG * MC * 64);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG * MC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

}  // namespace gemm

int gemm_topk_grid(int q_tiles, int64_t n, int num_sms) {
    const int64_t n_tiles = ceil_div(n, gemm::BN);
    int64_t g = (int64_t)q_tiles * (n_tiles < 1 ? 1 : n_tiles);
    if (g > num_sms) g = num_sms;
    if (g < q_tiles) g = q_tiles;
    return (int)g;
}

int gemm_topk_max_clusters(int cta_group, int pairs_per_cluster, int k) {
    if (cta_group == 2 && pairs_per_cluster == 4) return gemm::max_clusters<2, 4>(k);
    if (cta_group == 2 && pairs_per_cluster == 2) return gemm::max_clusters<2, 2>(k);
    if (cta_group == 2) return gemm::max_clusters<2, 1>(k);
    return gemm::max_clusters<1, 1>(k);
}

cudaError_t launch_gemm_topk(const GemmTopkParams &p, int grid, cudaStream_t s, const char **err_detail) {
    *err_detail = nullptr;
    const int cg = p.cta_group == 2 ? 2 : 1;
    const int mc = (cg == 2 && (p.pairs_per_cluster == 2 || p.pairs_per_cluster == 4)) ? p.pairs_per_cluster : 1;
    if (grid % p.q_tiles != 0 || p.q_tiles % (cg * mc) != 0) {
        *err_detail = "grid must be a multiple of q_tiles, q_tiles a multiple of the cluster size";
        return cudaErrorInvalidValue;
    }
    CUtensorMap map_q, map_c;
    if (!gemm::encode_rows_map(&map_q, p.queries_bf16, p.nq_pad, p.d_pad, gemm::BM) ||
        !gemm::encode_rows_map(&map_c, p.corpus_bf16, p.n, p.d_pad, gemm::BN / cg / mc)) {
        *err_detail = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
    }
    if (mc == 4) return gemm::launch_cg<2, 4>(map_q, map_c, p, grid, s);
    if (mc == 2) return gemm::launch_cg<2, 2>(map_q, map_c, p, grid, s);
    return cg == 2 ? gemm::launch_cg<2, 1>(map_q, map_c, p, grid, s) : gemm::launch_cg<1, 1>(map_q, map_c, p, grid, s);
}

}  // namespace b200
