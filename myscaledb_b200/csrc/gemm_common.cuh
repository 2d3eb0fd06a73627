// gemm_common.cuh -- PTX wrappers (mbarrier, TMA, tcgen05), the per-thread top-k list and the
// chunk filter shared by the tcgen05 kernels (ip_gemm_sm100.cu, ip_gemm_ts_sm100.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace b200 {
namespace gemm {


constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int ACC_STAGES = 2;
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;           // 16 KB
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_ALIGN_SLACK = 1024;
constexpr int MAX_STAGES = 6;

// CG = CTAs per MMA (cta_group): 1 or 2
constexpr int SCRATCH_BYTES = 32 * EPI_THREADS * 4;  // epilogue slow-path scratch [32][128] floats

template <int CG>
struct Cfg {
    static constexpr int B_ROWS = BN / CG;                 // corpus rows staged by one CTA
    static constexpr int B_BYTES = B_ROWS * BK * 2;        // 32 KB / 16 KB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // per CTA
    static constexpr int TX_BYTES = STAGE_BYTES * CG;      // what the (leader's) full barrier expects
    // smem layout for a ring of `stages` stages (runtime: 6/5 for pairs, 4/3 single, by top-k list size)
    __host__ __device__ static constexpr int off_a() { return 0; }
    __host__ __device__ static constexpr int off_b(int stages) { return stages * A_BYTES; }
    __host__ __device__ static constexpr int off_side(int stages) { return stages * STAGE_BYTES; }  // scale[256], bias[256]
    __host__ __device__ static constexpr int off_bar(int stages) { return off_side(stages) + 2 * BN * 4; }
    __host__ __device__ static constexpr int off_scratch(int stages) { return off_bar(stages) + 256; }
    __host__ __device__ static constexpr int off_list(int stages) { return off_scratch(stages) + SCRATCH_BYTES; }
    // deepest ring that leaves room for the per-thread lists (k <= kGemmSmemK) in 227 KB
    __host__ __device__ static int stages_for(int k_smem) {
        const int max_stages = CG == 1 ? 4 : 6;
        int st = max_stages;
        while (st > 2 && off_list(st) + k_smem * EPI_THREADS * 8 + SMEM_ALIGN_SLACK > 232448) st--;
        return st;
    }
    __host__ __device__ static bool lists_fit(int k) {
        return off_list(2) + k * EPI_THREADS * 8 + SMEM_ALIGN_SLACK <= 232448;   // k: slots per list
    }
};

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok != 0;
}
// Experiment knob (default 0 = plain spin, identical code): `make EXTRA=-DB200_MBAR_BACKOFF_NS=64` inserts a nanosleep
// between failed try_waits.  The committed ncu capture attributes ~30 % of the bf16 kernel's issued instructions to these
// spin loops; on a power-capped kernel that is worth measuring.
#ifndef B200_MBAR_BACKOFF_NS
#define B200_MBAR_BACKOFF_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    while (!mbar_try_wait(addr, parity)) {
#if B200_MBAR_BACKOFF_NS > 0
        __nanosleep(B200_MBAR_BACKOFF_NS);
#endif
    }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 2-CTA form: dst in this CTA, completion signalled on the LEADER CTA's mbarrier (the shared
// window address carries the CTA rank in bit 24; clearing it names the even CTA of the pair).
__device__ __forceinline__ void tma_load_2d_cg2(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}

// 2-CTA + multicast: the box is written at the same CTA-relative offset in every CTA of cta_mask and
// completes bytes on the mbarrier at `bar`'s offset in the LEADER of each destination's pair.
__device__ __forceinline__ void tma_load_2d_cg2_mc(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 remote;\n\t"
        "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remote];\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
// one lane of a converged warp (keeps the surrounding control flow warp-uniform, so the
// compiler holds descriptors / barrier addresses in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// K-major, 128-byte swizzled operand tile: rows of 64 bf16 (128 B), 8-row atoms of 1024 B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address
    d |= (uint64_t)1 << 16;                      // leading byte offset (unused for SW128 K-major)
    d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}

// kind::f16, A = B = bf16 (K-major), D = f32, M = 128 * CG, N = 256
__device__ __forceinline__ constexpr uint32_t make_idesc(int cg) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * cg) >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// arrive (once all prior MMAs retire) on the barrier at this offset in every CTA of `mask` (cluster ranks)
__device__ __forceinline__ void umma_commit_cg2(uint64_t *bar, uint16_t mask = 3) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
          "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
          "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr)
        : "memory");
}
// all tcgen05.ld issued by this thread have landed in their registers
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Per-thread top-k as an UNSORTED buffer (element j at [j * EPI_THREADS]: bank-conflict free in smem,
// coalesced in global scratch) plus the position of its current worst element.  An insert overwrites
// the worst slot and rescans the k slots with independent loads (~60 cycles at k = 10); the first
// version kept the list sorted and paid a dependent load-compare-store chain per shifted element
// (~400 cycles per insert, 1.7 ms of start-up "insert storm" per launch, profiles/r01_summary.md).
// The buffer is sorted once, when the CTA publishes its partial list.
struct ThreadTopK {
    float *keys;       // entry j of this thread's list: keys[j * stride]
    uint32_t *ids;
    int k, n, worst;
    int cap;           // slots of the buffer: == k -> "rescan" mode, > k (>= k + 32) -> "append" mode, see below
    int stride;        // EPI_THREADS: lists interleaved (entry j of all 128 lists side by side); 1: each list contiguous
    int coop;          // warp-uniform: inserts are done by the whole warp, one (lane, candidate) at a time (needs stride 1)
    int append;        // cap >= 2k + 32: append form
    int tourn;         // cap == list_cap_tourn(k): slots [k, cap) hold the worst (key, id) of every group of 8 / 16 entries; `worst` = a group
    float thr_key;     // key of the current worst kept element (FLT_MAX while n < k)
    uint32_t thr_id;
};

// Layout and insert form (cap == k only; the append form keeps the interleaved layout):
//  * interleaved lists, each lane inserts into its own list -- the default at every k;
//  * contiguous lists and COOPERATIVE inserts (k >= B200_LIST_COOP_MIN_K, off by default): the accepted candidate of one lane
//    is broadcast, lane 0 overwrites that list's worst entry, all 32 lanes rescan the list (k / 32 entries each) and a shuffle
//    arg-max yields the new worst.  Measured (profiles/r02_list_modes.md): it does NOT pay.  A slow-path event carries ~5
//    candidates spread over the lanes; the one-lane form retires them in ~1.5 parallel rounds of one k-entry rescan each, the
//    cooperative form in ~5 serial steps of ~100 instructions -- 15.6 vs 13.2 ms per launch at k = 17 / 16, 17.9 vs 14.4 at
//    k = 30, equal at k = 100.  Kept for the unit test and the record.
#ifndef B200_LIST_COOP_MIN_K
#define B200_LIST_COOP_MIN_K (1 << 30)
#endif
constexpr int kListCoopMinK = B200_LIST_COOP_MIN_K;
// Element j of a list sits at keys[j * LIST_STRIDE(t)].  Production builds have no contiguous (cooperative) lists, so the stride is
// the compile-time constant EPI_THREADS and the rescans address their entries with immediate offsets; a run-time stride cost the
// default form 16 % at k = 30 and 36 % at k = 100 (profiles/r02_gpu35.log vs r02_gpu22.log).
#if B200_LIST_COOP_MIN_K >= (1 << 30)
#define LIST_STRIDE(t) EPI_THREADS
#else
#define LIST_STRIDE(t) ((t).stride)
#endif
__device__ __forceinline__ void list_bind(ThreadTopK &t, float *keys_base, uint32_t *ids_base, int row, int k, int cap) {
    t.k = k;
    t.cap = cap;
    t.append = cap >= 2 * k + 32 ? 1 : 0;
    t.tourn = (!t.append && cap > k && cap == list_cap_tourn(k)) ? 1 : 0;
    t.coop = (cap == k && k >= kListCoopMinK) ? 1 : 0;
    t.stride = t.coop ? 1 : EPI_THREADS;
    t.keys = keys_base + (t.coop ? (size_t)row * cap : (size_t)row);
    t.ids = ids_base + (t.coop ? (size_t)row * cap : (size_t)row);
}

// Two ways to keep the k best:
//  * rescan (cap == k, the default): an accepted candidate overwrites the worst entry and the list is rescanned for the new
//    worst -- O(k) dependent-free loads per insert; the threshold is always exact.
//  * append (cap >= 2k + 32, opt-in): an accepted candidate is stored behind the others (two stores, nothing to wait for);
//    when some lane of the warp is within 32 slots of the end, EVERY lane of the warp compacts its own buffer in lock-step
//    (quickselect for the k-th entry, then one partition pass) and tightens its threshold.  Measured on a synthetic stream
//    shaped like the flat kernel's epilogue (tests/cuda/list_perf.cu, profiles/r02_list_modes.md): in shared memory it
//    halves the list cost at k = 100 (+8 ms against +21 ms per launch) and is equal at k <= 30, but it needs twice the
//    slots -- 200 KB at k = 100, which the operand ring does not leave -- and from global scratch it is no better than
//    the rescan form in shared memory.  The slack must be real: with k + 32 slots every slow-path event compacts (3x slower).
// (slot counts: list_cap_for / list_cap_append in kernels.h)

struct ListThr {
    float key;
    uint32_t id;
};

// counters for tests/cuda/list_append_test.cu --perf (compiled in only there)
#ifdef B200_LIST_STATS
__device__ unsigned long long g_list_stats[4];   // slow-path events, compaction calls (warp level), quickselect rounds, appended entries
#define B200_LIST_STAT(i, v) atomicAdd(&g_list_stats[i], (unsigned long long)(v))
#else
#define B200_LIST_STAT(i, v)
#endif

// Keep the k best of this thread's n (> k) entries in slots [0, k) and return the k-th (the new threshold).  Entries are
// distinct (unique ids), so (key, id) is a strict total order and exactly one entry has rank k.  Quickselect without
// moving data: the pivot's rank is counted in one pass, which also picks the next pivot on either side pseudo-randomly
// (smallest multiplicative hash of the slot), so sorted input does not degrade it.  All lanes of a warp run this together.
static __device__ __noinline__ ListThr list_compact(float *keys, uint32_t *ids, int k, int n, int stride) {
    float lo_k = 0.f, hi_k = 0.f;          // open interval (lo, hi) that still contains the rank-k entry
    uint32_t lo_i = 0, hi_i = 0;
    bool have_lo = false, have_hi = false;
    float pk = keys[(n - 1) * stride];
    uint32_t pi = ids[(n - 1) * stride];
    uint32_t salt = 0x9E3779B1u;
    for (;;) {
        int rank = 0;
        float ck_lo = 0.f, ck_hi = 0.f;
        uint32_t ci_lo = 0, ci_hi = 0, h_lo = 0xffffffffu, h_hi = 0xffffffffu;
        for (int j = 0; j < n; j++) {
            const float kj = keys[j * stride];
            const uint32_t ij = ids[j * stride];
            const uint32_t h = ((uint32_t)j + 1u) * salt;
            if (better(kj, ij, pk, pi)) {
                rank++;
                if ((!have_lo || better(lo_k, lo_i, kj, ij)) && h <= h_lo) {
                    h_lo = h;
                    ck_lo = kj;
                    ci_lo = ij;
                }
            } else if (better(pk, pi, kj, ij)) {
                if ((!have_hi || better(kj, ij, hi_k, hi_i)) && h <= h_hi) {
                    h_hi = h;
                    ck_hi = kj;
                    ci_hi = ij;
                }
            }
        }
        rank++;  // the pivot itself
        B200_LIST_STAT(2, 1);
        if (rank == k) break;
        if (rank < k) {          // the answer is worse than the pivot
            lo_k = pk; lo_i = pi; have_lo = true;
            pk = ck_hi; pi = ci_hi;
        } else {
            hi_k = pk; hi_i = pi; have_hi = true;
            pk = ck_lo; pi = ci_lo;
        }
        salt = salt * 0x85EBCA6Bu + 0xC2B2AE35u;
        salt |= 1u;
    }
    // partition: kept entries found behind slot k fill the slots of dropped entries in front of it
    int dst = 0;
    for (int j = k; j < n; j++) {
        const float kj = keys[j * stride];
        const uint32_t ij = ids[j * stride];
        if (!better(pk, pi, kj, ij)) {   // kj <= pivot: kept
            while (!better(pk, pi, keys[dst * stride], ids[dst * stride])) dst++;   // skip kept entries
            keys[dst * stride] = kj;
            ids[dst * stride] = ij;
            dst++;
        }
    }
    ListThr r;
    r.key = pk;
    r.id = pi;
    return r;
}

__device__ __forceinline__ void list_compact_if_over(ThreadTopK &t) {
    if (t.n > t.k) {
        const ListThr r = list_compact(t.keys, t.ids, t.k, t.n, LIST_STRIDE(t));
        t.n = t.k;
        t.thr_key = r.key;
        t.thr_id = r.id;
    }
}

__device__ __forceinline__ void list_insert(ThreadTopK &t, float key, uint32_t id) {
    if (!better(key, id, t.thr_key, t.thr_id)) return;
    if (t.append) {   // append mode: the caller keeps n + 32 <= cap before every chunk (epilogue_chunk)
        B200_LIST_STAT(3, 1);
        t.keys[t.n * LIST_STRIDE(t)] = key;
        t.ids[t.n * LIST_STRIDE(t)] = id;
        t.n++;
        return;
    }
    if (t.tourn) {
        // Two-level form: the worst entry of each group of G = 8 / 16 is cached behind the list, so an insert rescans ONE group (to
        // find the slot of the entry it evicts and that group's new worst) and the group worsts: 2 (G + k / G) loads instead of 2 k.
        const int G = list_tourn_group(t.k);
        const int ng = (t.k + G - 1) / G;
        float *gk = t.keys + (size_t)t.k * LIST_STRIDE(t);
        uint32_t *gi = t.ids + (size_t)t.k * LIST_STRIDE(t);
        if (t.n < t.k) {
            t.keys[t.n * LIST_STRIDE(t)] = key;
            t.ids[t.n * LIST_STRIDE(t)] = id;
            t.n++;
            if (t.n < t.k) return;
            for (int g = 0; g < ng; g++) {   // the list just became full: every group's worst, once
                const int base = g * G, end = base + G < t.k ? base + G : t.k;
                float wk = t.keys[base * LIST_STRIDE(t)];
                uint32_t wi = t.ids[base * LIST_STRIDE(t)];
                for (int j = base + 1; j < end; j++) {
                    const float kj = t.keys[j * LIST_STRIDE(t)];
                    const uint32_t ij = t.ids[j * LIST_STRIDE(t)];
                    if (better(wk, wi, kj, ij)) {
                        wk = kj;
                        wi = ij;
                    }
                }
                gk[g * LIST_STRIDE(t)] = wk;
                gi[g * LIST_STRIDE(t)] = wi;
            }
        } else {
            const int g = t.worst;   // the group that holds the evicted entry (= the current threshold)
            const int base = g * G, end = base + G < t.k ? base + G : t.k;
            float wk = 0.f;
            uint32_t wi = 0;
            int slot = -1;
            bool have = false;
            for (int j = base; j < end; j++) {
                float kj = t.keys[j * LIST_STRIDE(t)];
                uint32_t ij = t.ids[j * LIST_STRIDE(t)];
                if (slot < 0 && kj == t.thr_key && ij == t.thr_id) {
                    slot = j;
                    kj = key;
                    ij = id;
                }
                if (!have || better(wk, wi, kj, ij)) {
                    wk = kj;
                    wi = ij;
                    have = true;
                }
            }
            if (slot < 0) slot = base;   // cannot happen (ids are unique and the threshold is an entry of this group); never write out of range
            t.keys[slot * LIST_STRIDE(t)] = key;
            t.ids[slot * LIST_STRIDE(t)] = id;
            gk[g * LIST_STRIDE(t)] = wk;
            gi[g * LIST_STRIDE(t)] = wi;
        }
        float wk = gk[0];
        uint32_t wi = gi[0];
        int wg = 0;
        for (int g = 1; g < ng; g++) {
            const float kg = gk[g * LIST_STRIDE(t)];
            const uint32_t ig = gi[g * LIST_STRIDE(t)];
            if (better(wk, wi, kg, ig)) {
                wk = kg;
                wi = ig;
                wg = g;
            }
        }
        t.worst = wg;
        t.thr_key = wk;
        t.thr_id = wi;
        return;
    }
    if (t.n < t.k) {
        t.keys[t.n * LIST_STRIDE(t)] = key;
        t.ids[t.n * LIST_STRIDE(t)] = id;
        t.n++;
        if (t.n < t.k) return;
    } else {
        t.keys[t.worst * LIST_STRIDE(t)] = key;
        t.ids[t.worst * LIST_STRIDE(t)] = id;
    }
    // rescan for the worst (largest key, ties -> larger id)
    float wk = t.keys[0];
    uint32_t wi = t.ids[0];
    int wp = 0;
    for (int j = 1; j < t.k; j++) {
        const float kj = t.keys[j * LIST_STRIDE(t)];
        const uint32_t ij = t.ids[j * LIST_STRIDE(t)];
        if (better(wk, wi, kj, ij)) {
            wk = kj;
            wi = ij;
            wp = j;
        }
    }
    t.worst = wp;
    t.thr_key = wk;
    t.thr_id = wi;
}

// sort the n kept entries best-first (insertion sort, once per kernel) and publish them
static __device__ __noinline__ void list_publish(ThreadTopK &t, float *out_keys, uint32_t *out_ids) {
    list_compact_if_over(t);
    for (int i = 1; i < t.n; i++) {
        const float ki = t.keys[i * LIST_STRIDE(t)];
        const uint32_t ii = t.ids[i * LIST_STRIDE(t)];
        int j = i;
        while (j > 0 && better(ki, ii, t.keys[(j - 1) * LIST_STRIDE(t)], t.ids[(j - 1) * LIST_STRIDE(t)])) {
            t.keys[j * LIST_STRIDE(t)] = t.keys[(j - 1) * LIST_STRIDE(t)];
            t.ids[j * LIST_STRIDE(t)] = t.ids[(j - 1) * LIST_STRIDE(t)];
            j--;
        }
        t.keys[j * LIST_STRIDE(t)] = ki;
        t.ids[j * LIST_STRIDE(t)] = ii;
    }
    for (int j = 0; j < t.k; j++) {
        out_keys[j] = j < t.n ? t.keys[j * LIST_STRIDE(t)] : FLT_MAX;
        out_ids[j] = j < t.n ? t.ids[j * LIST_STRIDE(t)] : kNoId;
    }
}

// v[j] = v[j] * scale[j] + bias[j] for the 32 columns of a chunk; scale / bias are the per-tile side arrays in SHARED memory.
// Explicit ld.shared: through generic pointers the compiler emitted LD.E.128 (generic path, scoreboarded like a global load).
__device__ __forceinline__ void side_fma32(float (&v)[32], const float *scale, const float *bias) {
    const uint32_t sa = smem_u32(scale), ba = smem_u32(bias);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float s0, s1, s2, s3, b0, b1, b2, b3;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(s0), "=f"(s1), "=f"(s2), "=f"(s3) : "r"(sa + j * 4));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(ba + j * 4));
        v[j] = fmaf(v[j], s0, b0);
        v[j + 1] = fmaf(v[j + 1], s1, b1);
        v[j + 2] = fmaf(v[j + 2], s2, b2);
        v[j + 3] = fmaf(v[j + 3], s3, b3);
    }
}

// Cooperative insert (t.coop): called by the whole warp; lane `src` contributes the candidate and owns the list.
__device__ __forceinline__ void list_insert_coop(ThreadTopK &t, int src, float key, uint32_t id) {
    const int lane = threadIdx.x & 31;
    key = __shfl_sync(0xffffffffu, key, src);
    id = __shfl_sync(0xffffffffu, id, src);
    const float thr_key = __shfl_sync(0xffffffffu, t.thr_key, src);
    const uint32_t thr_id = __shfl_sync(0xffffffffu, t.thr_id, src);
    if (!better(key, id, thr_key, thr_id)) return;   // warp-uniform
    const unsigned long long kp = __shfl_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(t.keys), src);
    const unsigned long long ip = __shfl_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(t.ids), src);
    float *lk = reinterpret_cast<float *>((uintptr_t)kp);
    uint32_t *li = reinterpret_cast<uint32_t *>((uintptr_t)ip);
    int n = __shfl_sync(0xffffffffu, t.n, src);
    const int k = t.k;
    if (n < k) {
        if (lane == 0) {
            lk[n] = key;
            li[n] = id;
        }
        n++;
        if (n < k) {
            if (lane == src) t.n = n;
            return;
        }
    } else {
        const int worst = __shfl_sync(0xffffffffu, t.worst, src);
        if (lane == 0) {
            lk[worst] = key;
            li[worst] = id;
        }
    }
    __syncwarp();
    // the new worst: every lane scans k / 32 entries, then a shuffle arg-max over (key, id)
    float wk = 0.f;
    uint32_t wi = 0;
    int wp = -1;
    for (int j = lane; j < k; j += 32) {
        const float kj = lk[j];
        const uint32_t ij = li[j];
        if (wp < 0 || better(wk, wi, kj, ij)) {
            wk = kj;
            wi = ij;
            wp = j;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, wk, o);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, wi, o);
        const int op = __shfl_xor_sync(0xffffffffu, wp, o);
        if (op >= 0 && (wp < 0 || better(wk, wi, ok, oi))) {
            wk = ok;
            wi = oi;
            wp = op;
        }
    }
    if (lane == src) {
        t.n = n;
        t.worst = wp;
        t.thr_key = wk;
        t.thr_id = wi;
    }
    __syncwarp();
}

// Filter one chunk of 32 accumulator columns of this thread's query row.
// Fast path (steady state): reduce the chunk to its best key with FMNMX3 trees, one warp vote,
// done.  Slow path (some lane of the warp can improve its list; frequent only during the first
// tiles of a launch): every lane parks its 32 keys in a shared-memory scratch column and the WARP
// loops while any lane still has a candidate bit, each lane popping its own lowest bit and doing
// an inlined insert.  Compared with per-element calls under divergence this removed a fixed
// ~1.3 ms "insert storm" per launch (profiles/r01_summary.md).
// scratch: this thread's column of a [32][EPI_THREADS] float array.
__device__ __forceinline__ void epilogue_chunk(ThreadTopK &list, float (&v)[32], bool use_side, const float *scale,
                                               const float *bias, uint32_t id0, bool tail, int64_t n, float *scratch,
                                               float ext_bound = FLT_MAX /* a valid upper bound of the k-th key known from elsewhere */) {
    float thr = fminf(list.thr_key, ext_bound);
    bool mine;
    if (use_side) {
        side_fma32(v, scale, bias);  // 16 broadcast LDS.128
        float m0 = fminf(v[0], v[1]), m1 = fminf(v[2], v[3]), m2 = fminf(v[4], v[5]), m3 = fminf(v[6], v[7]);
#pragma unroll
        for (int j = 8; j < 32; j += 8) {
            m0 = fminf(m0, fminf(v[j], v[j + 1]));
            m1 = fminf(m1, fminf(v[j + 2], v[j + 3]));
            m2 = fminf(m2, fminf(v[j + 4], v[j + 5]));
            m3 = fminf(m3, fminf(v[j + 6], v[j + 7]));
        }
        mine = fminf(fminf(m0, m1), fminf(m2, m3)) <= thr;
    } else {
        // plain IP: rank on the raw score with max trees; the key (-score) is formed only when needed
        float m0 = fmaxf(v[0], v[1]), m1 = fmaxf(v[2], v[3]), m2 = fmaxf(v[4], v[5]), m3 = fmaxf(v[6], v[7]);
#pragma unroll
        for (int j = 8; j < 32; j += 8) {
            m0 = fmaxf(m0, fmaxf(v[j], v[j + 1]));
            m1 = fmaxf(m1, fmaxf(v[j + 2], v[j + 3]));
            m2 = fmaxf(m2, fmaxf(v[j + 4], v[j + 5]));
            m3 = fmaxf(m3, fmaxf(v[j + 6], v[j + 7]));
        }
        mine = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) >= -thr;
    }
    if (__any_sync(0xffffffffu, mine)) {
        if ((threadIdx.x & 31) == 0) B200_LIST_STAT(0, 1);
        if (list.append && __any_sync(0xffffffffu, list.n + 32 > list.cap)) {
            if ((threadIdx.x & 31) == 0) B200_LIST_STAT(1, 1);
            list_compact_if_over(list);   // every lane, in lock-step: room for this chunk and a fresh threshold
            thr = fminf(list.thr_key, ext_bound);
        }
        uint32_t mask = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            const float key = use_side ? v[j] : -v[j];
            scratch[j * EPI_THREADS] = key;
            if (key <= thr) mask |= 1u << j;
        }
        if (tail && !use_side) {  // rows past the end of the corpus (zero-filled by TMA) are not candidates
            const int64_t left = n - (int64_t)id0;
            mask = left >= 32 ? mask : left <= 0 ? 0u : (mask & ((1u << left) - 1u));
        }
        if (list.coop) {
            // one (lane, candidate) at a time, the whole warp inserting
            unsigned pending = __ballot_sync(0xffffffffu, mask != 0);
            while (pending) {
                const int src = __ffs(pending) - 1;
                float ckey = 0.f;
                uint32_t cid = 0;
                if ((int)(threadIdx.x & 31) == src) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    ckey = scratch[j * EPI_THREADS];
                    cid = id0 + (uint32_t)j;
                }
                list_insert_coop(list, src, ckey, cid);
                pending = __ballot_sync(0xffffffffu, mask != 0);
            }
        } else {
            while (__any_sync(0xffffffffu, mask != 0)) {
                if (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    list_insert(list, scratch[j * EPI_THREADS], id0 + (uint32_t)j);
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D bf16 tensor map over row-major [rows][d_pad], box = [box_rows][64 elements], 128-byte swizzle
inline bool encode_rows_map(CUtensorMap *map, const void *base, int64_t rows, int d_pad, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)d_pad * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gemm
}  // namespace b200
