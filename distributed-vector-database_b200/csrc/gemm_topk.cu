// gemm_topk.cu -- K2: batched search as a dense contraction on the 5th-gen tensor cores, fused
// with threshold filtering of the scores, + K2s (select between levels) and K4w (exact fp32
// re-rank of the keys that can still matter + coverage certificate).
//
// Replaces hnswlib.Index.knn_query for batches (reference call site src/datanode/handler.py:364;
// hnswlib accepts [nq, dim], the reference only ever passes one row).
//
// K2 (gemm_filter_kernel): persistent CTA PAIRS (cluster of 2, tcgen05 cta_group::2), 384 threads:
//   warp 0   TMA producer: per k-block (128 bytes of every row) one 2-D tensor copy of this CTA's
//            128 query rows (A) and of its half (128 rows) of the 256-row shard tile (B),
//            SWIZZLE_128B, into a 4-stage shared-memory ring; completion bytes of both CTAs are
//            counted on the pair leader's mbarrier
//   warp 1   MMA issuer (leader CTA): tcgen05.mma.cta_group::2, kind::tf32 for fp32 rows /
//            kind::f16 for fp16 rows, M=256 (2 x 128 queries) x N=256, fp32 accumulators in TMEM
//            (two 256-column buffers, ping-pong); tcgen05.commit (multicast to both CTAs) frees
//            smem stages / publishes accumulators
//   warps 2-3 movers: drain the per-query key rings into the per-query candidate buffers in
//            global memory (one atomicAdd reservation per query and round, lane-parallel)
//   warps 4-11 epilogue: two threads per query (one TMEM lane, columns 0-127 / 128-255): tcgen05.ld
//            32 columns at a time, turn the dot product into an approximate distance, keep the rows
//            that beat the query's threshold (a key ring in shared memory; nothing else leaves the SM).
//   The [nq, n_rows] distance matrix never exists in memory.
//
// Thresholds come from a PROBE and LEVELS: the shard's 256-row tiles are visited in bit-reversed
// order, so every prefix of the order is an evenly spread sample.  The probe (the first 16 tiles)
// writes only the best live row of every 32-score chunk; the r-th smallest chunk minimum is the first
// threshold.  Levels then start over at tile 0, each 8x larger than what informed its threshold;
// between levels K2s keeps each query's k' best approximate keys and publishes the next threshold
// (the k'-th best, or a tighter rank while little of the shard has been seen).  Thresholds only
// tighten, and a row that is not in a query's buffer has approximate distance >= its threshold.
//
// K4w reads the last level's buffer as it is: the keys within 2 eps of the k-th smallest approximate
// value are recomputed exactly (same summation order as the scan kernel, so batched and single-query
// searches return bit-identical distances), eps a rigorous bound on the tf32 / fp16 rounding error;
// if the k-th exact distance is < last threshold - eps no row outside the buffer can enter the top-k.
// Queries that fail this certificate (or whose buffer overflowed) are re-searched with the exact scan.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "block_select.cuh"
#include "common.cuh"
#include "gemm_topk.h"
#include "kernels.h"

namespace vdbk {

constexpr int GT_BM = 128;                           // queries per CTA (TMEM lanes); a CTA PAIR covers 256
constexpr int GT_BN = 256;                           // shard rows per tile (TMEM columns)
constexpr int GT_BN_HALF = GT_BN / 2;                // rows of the B tile each CTA of the pair stages
constexpr int GT_KB_BYTES = 128;                     // one 128B swizzle atom per row per k-block
constexpr int GT_A_BYTES = GT_BM * GT_KB_BYTES;      // 16 KB
constexpr int GT_B_BYTES = GT_BN_HALF * GT_KB_BYTES; // 16 KB (this CTA's half of the 256-row tile)
constexpr int GT_STAGE_BYTES = GT_A_BYTES + GT_B_BYTES;
#ifndef VDB_GT_STAGES
#define VDB_GT_STAGES 4       // -D overrides: tuning builds only (tools/level_sweep.sh documents the measured choices)
#endif
#ifndef VDB_GT_RING
#define VDB_GT_RING 40
#endif
#ifndef VDB_GT_STAGES_ARES
#define VDB_GT_STAGES_ARES 4  // query-resident variant: stages of 16 KB (shard rows only)
#endif
#ifndef VDB_GT_RING_ARES
#define VDB_GT_RING_ARES 13
#endif
constexpr int GT_STAGES = VDB_GT_STAGES;
constexpr int GT_THREADS = 384;                      // 12 warps: TMA, MMA, 2 movers, 8 epilogue
constexpr int GT_RING = VDB_GT_RING;                  // keys per ring, streaming variant
constexpr int GT_EPI_THREADS = 256;                  // 2 threads per query (TMEM lane): columns 0-127 / 128-255
constexpr int GT_EPI_WARPS = GT_EPI_THREADS / 32;
constexpr int GT_TMEM_COLS = 512;
constexpr int GT_LEVEL_GROWTH = 8;                   // each level sees 8x the rows seen so far
constexpr int GT_PROBE_TILES = 16;                   // probe level: 16 tiles = 4096 rows, 128 chunk minima per query
constexpr int GT_CTL_BYTES = GT_EPI_THREADS * 8 + 64;  // EpiCtl, checked below

// dynamic shared memory map (base aligned to 1024), two variants of the kernel:
//   streaming (ARES = false): every stage of the ring holds one k-block of the query block (A, 16 KB) and of this
//     CTA's half of the shard tile (B, 16 KB); A is fetched again for every tile (from L2).
//   query-resident (ARES = true, rows of at most 1024 bytes = 8 k-blocks): the CTA's 128 query rows stay in shared
//     memory (128 KB) for as long as the pair works on the same query block, the ring holds shard rows only.  Per
//     tile a CTA then pulls 128 KB instead of 256 KB through the L2 -> SM path, which at the f16 MMA rate was running
//     at ~11 TB/s chip-wide, ~0.9 of what the L2 slices deliver (profiles/README.md); the key rings shrink to pay
//     for it (survivors are ~1 score in 3000 on the levels that matter).
template <bool ARES>
struct GtSmem {
    static constexpr int A_RES_KB = 8;                                        // resident k-blocks (ARES)
    static constexpr int A_RES_BYTES = ARES ? A_RES_KB * GT_A_BYTES : 0;      // 128 KB
    static constexpr int STAGES = ARES ? VDB_GT_STAGES_ARES : GT_STAGES;
    static constexpr int STAGE_BYTES = ARES ? GT_B_BYTES : GT_STAGE_BYTES;
    static constexpr int B_OFF = ARES ? 0 : GT_A_BYTES;                       // B inside a stage
    static constexpr int RING = ARES ? VDB_GT_RING_ARES : GT_RING;           // keys per ring (one ring per epilogue thread)
    static constexpr int RING_STRIDE = RING | 1;                              // odd: same-slot appends of a warp spread over banks
    static constexpr int OFF_STAGE = A_RES_BYTES;
    static constexpr int OFF_RING = OFF_STAGE + STAGES * STAGE_BYTES;
    static constexpr int OFF_NORM = OFF_RING + GT_EPI_THREADS * RING_STRIDE * 8;
    static constexpr int OFF_BAR = OFF_NORM + 2 * GT_BN * 4;                  // + 2048
    static constexpr int BAR_BYTES = 256;                                     // full/empty per stage, 2+2 TMEM, 2 A, TMEM pointer
    static constexpr int OFF_CTL = OFF_BAR + BAR_BYTES;
    static constexpr int BYTES = OFF_CTL + GT_CTL_BYTES + 1024;               // + slack for the 1024-byte alignment
    static_assert((2 * STAGES + 6) * 8 + 4 <= BAR_BYTES, "barrier block");
    static_assert(RING >= 8, "the epilogue reserves room for 4 appends at a time");
    static_assert(BYTES <= 232448, "shared memory budget");
};
struct GemmParams {
    uint32_t n_rows, nq;
    int num_kb;              // k-blocks per row (row bytes / 128)
    int kb_elems;            // elements per k-block (32 fp32 / 64 fp16)
    int MB;                  // 256-query blocks (one per CTA pair)
    int S;                   // slices of this level's position range
    int n_items;             // MB * S
    int n_tiles;             // 256-row tiles of the shard
    int bits;                // tile order = bit reversal over `bits` bits
    int pos_begin, pos_end;  // this level's positions in that order
    int probe;               // probe level: one key per 32-score chunk (its best live row), at fixed positions
    int dbg;                 // experiments only: 2 = epilogue releases TMEM at once, 8 = TMEM reads but no filtering
    const float* sqnorm;     // [n_rows] (L2 only)
    const float* thr;        // [nq] threshold of this level (approximate distance); unused by the probe
    const uint32_t* tomb;    // tombstone bitmap or null: dead rows neither stand for a probe chunk nor pass a level
    uint64_t* buf;           // [nq][cap] candidate keys (approximate distance bits << 32 | row)
    int* cnt;                // [nq] keys in buf
    int cap;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers (tcgen05 / TMA); cta_group::2 (CTA pair)
// ------------------------------------------------------------------------------------------
// cluster helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D tensor copy into THIS CTA's shared memory; completion bytes are counted on the mbarrier at
// shared::cluster address `bar_cluster` (the pair leader's barrier: its MMA consumes both halves)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint32_t bar_cluster,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(x), "r"(y), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives (once the MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (F16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void st_shared_u64(uint32_t addr, uint64_t v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
// start address >> 4 | LBO (ignored for swizzled K-major, 1) | SBO = 8 rows * 128 B | layout 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// instruction descriptor: D fp32, A/B both `fmt` (0 f16, 2 tf32), both K-major, N, M
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
// 32 consecutive floats from shared memory (explicit shared-space loads: a generic pointer into the
// dynamic shared array would compile to generic LD, ordered against every ring store)
__device__ __forceinline__ void lds_f32x32(uint32_t addr, float (&o)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
            : "=f"(o[4 * i]), "=f"(o[4 * i + 1]), "=f"(o[4 * i + 2]), "=f"(o[4 * i + 3])
            : "r"(addr + 16 * i));
}
__host__ __device__ __forceinline__ uint32_t bitrev(uint32_t x, int bits) {
#ifdef __CUDA_ARCH__
    return bits ? (__brev(x) >> (32 - bits)) : 0u;
#else
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

// shared-memory control block between the epilogue threads (producers of candidate keys) and the
// mover warps (consumers)
struct EpiCtl {
    uint32_t head_pub[GT_EPI_THREADS];   // keys appended so far (published by the epilogue thread)
    uint32_t tail[GT_EPI_THREADS];       // keys moved out so far
    uint32_t qbase[GT_EPI_WARPS];        // first query of each epilogue warp's current item
    uint32_t done;                       // epilogue warps that have finished all items
};

static_assert(sizeof(EpiCtl) <= GT_CTL_BYTES, "EpiCtl outgrew its shared-memory slot");

// item -> (slice, query block) and the slice's position range inside the level
struct ItemRange { int mb, p0, p1; };
__device__ __forceinline__ ItemRange item_range(const GemmParams& p, int item) {
    ItemRange r;
    const int slice = item / p.MB;
    r.mb = item - slice * p.MB;
    const long long n = p.pos_end - p.pos_begin;
    r.p0 = p.pos_begin + (int)(n * slice / p.S);
    r.p1 = p.pos_begin + (int)(n * (slice + 1) / p.S);
    return r;
}

// ------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------
template <bool F16, bool L2, bool ARES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GT_THREADS, 1)
gemm_filter_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    pdl_prologue();
    using SM = GtSmem<ARES>;
    constexpr int NST = SM::STAGES, RING = SM::RING, RING_STRIDE = SM::RING_STRIDE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem + SM::OFF_STAGE;                     // ring of operand stages (ARES: after the resident queries)
    uint64_t* ring_all = reinterpret_cast<uint64_t*>(smem + SM::OFF_RING);
    float* norm_s = reinterpret_cast<float*>(smem + SM::OFF_NORM);   // [2][256]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::OFF_BAR);
    uint64_t* empty = full + NST;
    uint64_t* tmem_full = empty + NST;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* a_full = tmem_empty + 2;       // ARES: the pair's query block has landed (leader's barrier counts both CTAs' bytes)
    uint64_t* a_empty = a_full + 1;          // ARES: every MMA that reads the resident block has retired
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_empty + 1);
    EpiCtl* ctl = reinterpret_cast<EpiCtl*>(smem + SM::OFF_CTL);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();     // 0 = pair leader (issues the MMAs), 1 = peer
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);      // leader's arrive.expect_tx; bytes of BOTH CTAs' copies land on the leader's barrier
            mbar_init(&empty[s], 1);     // one multicast tcgen05.commit per use
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 2 * GT_EPI_WARPS);   // 8 epilogue warps x 2 CTAs arrive on the LEADER's barrier
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < GT_EPI_THREADS) {
        ctl->head_pub[threadIdx.x] = 0;
        ctl->tail[threadIdx.x] = 0;
        if (threadIdx.x < GT_EPI_WARPS) ctl->qbase[threadIdx.x] = 0;
        if (threadIdx.x == 0) ctl->done = 0;
    }
    if (warp == 2) tmem_alloc(tmem_ptr, GT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // barriers of both CTAs are initialised before any remote arrive / copy
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own query rows, own half of the shard tile) =================
        if (lane == 0) {
            // shard tiles are re-read from L2 by the pairs that hold the other query blocks of the same
            // slice, so they keep the default policy; the query block is reused by every tile: keep it
            uint64_t pol_stream;
            asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_stream));
            const uint64_t pol_keep = l2_policy_evict_last();
            uint32_t it = 0, a_loads = 0;
            int cur_mb = -1;
            for (int item = pair; item < p.n_items; item += num_pairs) {
                const ItemRange ir = item_range(p, item);
                const int arow = ir.mb * (2 * GT_BM) + (int)cta_rank * GT_BM;
                if (ARES && ir.mb != cur_mb) {
                    // (re)load the resident query block: all k-blocks at once, after the MMAs on the previous block
                    if (a_loads) mbar_wait(a_empty, (a_loads - 1) & 1);
                    if (cta_rank == 0) mbar_arrive_expect_tx(a_full, 2u * (uint32_t)p.num_kb * GT_A_BYTES);
                    const uint32_t abar = map_to_cta(smem_u32(a_full), 0);
                    for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(smem + kb * GT_A_BYTES, &tmA, kb * p.kb_elems, arow, abar, pol_keep);
                    cur_mb = ir.mb;
                    ++a_loads;
                }
                for (int pos = ir.p0; pos < ir.p1; ++pos) {
                    const int tile = (int)bitrev((uint32_t)pos, p.bits);
                    if (tile >= p.n_tiles) continue;
                    const int brow = tile * GT_BN + (int)cta_rank * GT_BN_HALF;
                    for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                        const int s = it % NST;
                        const uint32_t ph = (it / NST) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full[s], 2 * SM::STAGE_BYTES);
                        const uint32_t bar = map_to_cta(smem_u32(&full[s]), 0);
                        uint8_t* sa = stage_base + s * SM::STAGE_BYTES;
                        if (!ARES) tma_load_2d(sa, &tmA, kb * p.kb_elems, arow, bar, pol_keep);
                        tma_load_2d(sa + SM::B_OFF, &tmB, kb * p.kb_elems, brow, bar, pol_stream);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (pair leader only): M = 256 (2 x 128 queries) x N = 256 =================
        // The WHOLE warp walks the loops and waits on the barriers; one elected lane issues.  Everything the
        // tcgen05 instructions take (descriptors, TMEM address, barrier address) is then warp-uniform by
        // construction and lives in uniform registers.  With a single lane inside `if (lane == 0)` the compiler
        // cannot know that and wraps every MMA in an elect / R2UR loop: ~106 dependent instructions per k-block,
        // ~156 cycles per MMA against the 128 the tensor pipe needs -- the issuer, not the data, paced the
        // kernel (ncu source page: no samples on the full / tmem_empty waits, all on the issue loop).
        if (cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc(F16 ? 0u : 2u, 2 * GT_BM, GT_BN);
            const uint32_t stage0 = smem_u32(stage_base), a_res0 = smem_u32(smem);
            const uint64_t desc_hi = make_smem_desc(0);              // everything but the 14-bit start address
            uint32_t it = 0, tcount = 0, a_loads = 0;
            int cur_mb = -1;
            for (int item = pair; item < p.n_items; item += num_pairs) {
                const ItemRange ir = item_range(p, item);
                if (ARES && ir.mb != cur_mb) {
                    if (a_loads && elect_one()) umma_commit_pair(a_empty);   // both producers may overwrite the block once these MMAs retire
                    __syncwarp();
                    mbar_wait(a_full, a_loads & 1);
                    tc_fence_after();
                    cur_mb = ir.mb;
                    ++a_loads;
                }
                for (int pos = ir.p0; pos < ir.p1; ++pos) {
                    if ((int)bitrev((uint32_t)pos, p.bits) >= p.n_tiles) continue;
                    const uint32_t acc = tcount & 1;
                    mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * GT_BN;
                    for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                        const uint32_t s = it % NST;
                        const uint32_t ph = (it / NST) & 1;
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t sa = stage0 + s * SM::STAGE_BYTES;
                        const uint64_t adesc = desc_hi | (uint64_t)(((ARES ? a_res0 + (uint32_t)kb * GT_A_BYTES : sa) & 0x3FFFFu) >> 4);
                        const uint64_t bdesc = desc_hi | (uint64_t)(((sa + SM::B_OFF) & 0x3FFFFu) >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)   // 4 x 32 bytes of K per 128-byte swizzle atom
                                umma<F16>(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
                            umma_commit_pair(&empty[s]);           // both CTAs may refill this stage once the MMAs retire
                            if (kb == p.num_kb - 1) umma_commit_pair(&tmem_full[acc]);   // accumulators ready in both CTAs' TMEM
                        }
                        __syncwarp();
                    }
                    ++tcount;
                }
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ================= movers: key rings -> per-query candidate buffers in global memory =================
        // Lane-parallel: each lane serves two queries; one atomicAdd reserves room for everything that is
        // waiting in a ring, so global-memory latency never sits on the TMEM-drain path of the epilogue.
        if (!p.probe) {
            const int base = (warp - 2) * (GT_EPI_THREADS / 2);    // each mover warp serves half of the rings
            volatile uint32_t* v_head = ctl->head_pub;
            volatile uint32_t* v_tail = ctl->tail;
            constexpr int RPL = GT_EPI_THREADS / 64;      // rings per lane
            for (;;) {
                const bool fin = *reinterpret_cast<volatile uint32_t*>(&ctl->done) == GT_EPI_WARPS;
                __threadfence_block();
                // round: look at all rings of this lane, reserve room for all of them at once (the atomics'
                // round trips overlap), then copy
                uint32_t t[RPL], n[RPL], q[RPL];
                int slot[RPL];
                bool moved = false;
#pragma unroll
                for (int h = 0; h < RPL; ++h) {
                    const int ri = base + h * 32 + lane;          // ring == epilogue thread
                    t[h] = v_tail[ri];
                    n[h] = v_head[ri] - t[h];
                    moved |= n[h] != 0;
                }
                __threadfence_block();
#pragma unroll
                for (int h = 0; h < RPL; ++h) {
                    const int ri = base + h * 32 + lane;
                    slot[h] = 0;
                    if (n[h]) {
                        q[h] = *reinterpret_cast<volatile uint32_t*>(&ctl->qbase[ri >> 5]) + (ri & 31);
                        slot[h] = atomicAdd(p.cnt + q[h], (int)n[h]);
                    }
                }
#pragma unroll
                for (int h = 0; h < RPL; ++h) {
                    if (n[h]) {
                        const int ri = base + h * 32 + lane;
                        const uint64_t* ring = ring_all + (size_t)ri * RING_STRIDE;
                        uint64_t* dst = p.buf + (size_t)q[h] * p.cap;
                        for (uint32_t i = 0; i < n[h]; ++i) {
                            const uint64_t key = *reinterpret_cast<const volatile uint64_t*>(ring + ((t[h] + i) % RING));
                            if (slot[h] + (int)i < p.cap) dst[slot[h] + i] = key;     // beyond cap: counted, flagged by K2s
                        }
                    }
                }
                __threadfence_block();
#pragma unroll
                for (int h = 0; h < RPL; ++h)
                    if (n[h]) v_tail[base + h * 32 + lane] = t[h] + n[h];
                if (!__any_sync(0xffffffffu, moved)) {
                    if (fin) break;
                    __nanosleep(100);
                }
            }
        }
    } else {
        // ================= epilogue: threshold filter (each CTA: its own 128 queries) =================
        // Two threads per query: warps 4-7 take columns 0-127 of the accumulator, warps 8-11 columns 128-255
        // (a warp reaches the TMEM lane quadrant warp % 4).  Half the columns per thread halves the time an
        // accumulator stays busy, which is what paces the MMA at the f16 rate.
        const int et = threadIdx.x - 128;                  // 0..255: ring / epilogue thread
        const int ew = et >> 5;                            // 0..7: epilogue warp
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may read
        const int half = ew >> 2;                          // which 128 columns
        const int ql = quad * 32 + lane;                   // TMEM lane == query within this CTA's block
        const int col0 = half * (GT_BN / 2);
        const uint32_t my_ring_s = smem_u32(ring_all + (size_t)et * RING_STRIDE);
        volatile uint32_t* my_tail = &ctl->tail[et];
        uint32_t head = 0, slot = 0;      // keys appended so far; slot == head % RING
        uint32_t tail_seen = 0;           // last value read from the movers' tail counter
        bool dirty = false;               // keys appended since head was last published
        uint32_t tcount = 0;
        for (int item = pair; item < p.n_items; item += num_pairs) {
            const ItemRange ir = item_range(p, item);
            const uint32_t qbase = (uint32_t)ir.mb * (2 * GT_BM) + cta_rank * GT_BM;
            const uint32_t q = qbase + ql;
            const bool q_ok = q < p.nq;
            float thr = __int_as_float(0xff800000);        // -inf: nothing passes (padding queries)
            if (!p.probe) {
                // the ring still holds keys of the previous item's query until the movers have drained it
                while (*my_tail != head) __nanosleep(64);
                tail_seen = head;
                __syncwarp();
                if (lane == 0) ctl->qbase[ew] = qbase + quad * 32;
                __threadfence_block();
                if (q_ok) thr = p.thr[q];
            }
            const float nthr = -thr;
            uint64_t* probe_dst = p.buf + (size_t)(q_ok ? q : 0) * p.cap;
            float n_next = 0.0f;
            bool first_tile = true;

            for (int pos = ir.p0; pos < ir.p1; ++pos) {
                const int tile = (int)bitrev((uint32_t)pos, p.bits);
                if (tile >= p.n_tiles) continue;
                const uint32_t acc = tcount & 1;
                const uint32_t row0 = (uint32_t)tile * GT_BN;
                if constexpr (L2) {
                    // ||d||^2 of this tile's rows: prefetched into registers one tile ahead (first tile of an
                    // item: loaded here), published to the buffer the previous-but-one tile used
                    float* nsw = norm_s + acc * GT_BN;
                    if (first_tile) {
                        const uint32_t r_a = row0 + et;
                        n_next = r_a < p.n_rows ? p.sqnorm[r_a] : 0.0f;
                    }
                    nsw[et] = n_next;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    // next valid tile of this item
                    int np = pos + 1;
                    while (np < ir.p1 && (int)bitrev((uint32_t)np, p.bits) >= p.n_tiles) ++np;
                    if (np < ir.p1) {
                        const uint32_t r_a = bitrev((uint32_t)np, p.bits) * GT_BN + et;
                        n_next = r_a < p.n_rows ? p.sqnorm[r_a] : 0.0f;
                    }
                }
                first_tile = false;
                mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * GT_BN + col0;
                const uint32_t ns_s = smem_u32(norm_s + acc * GT_BN + col0);
                constexpr int NCH = GT_BN / 2 / 32;            // 4 chunks of 32 columns per thread
                if (p.probe) {
                    // probe level: per 32-score chunk only its best live row is written (fixed position: 8 keys
                    // per tile and query).  The r-th smallest of those chunk minima is the value of a real row with
                    // at least r rows at or below it: a valid threshold for the levels that follow, which visit
                    // these tiles again.
                    uint64_t* dd = probe_dst + (size_t)(pos - p.pos_begin) * 8 + half * NCH;
#pragma unroll 1
                    for (int c = 0; c < NCH; ++c) {
                        uint32_t v[32];
                        tmem_ld32(taddr + c * 32, v);
                        tmem_ld_wait();
                        if (q_ok) {
                            const uint32_t r0 = row0 + col0 + c * 32;        // multiple of 32: one tombstone word
                            uint32_t dead = p.tomb ? p.tomb[r0 >> 5] : 0u;
                            if (r0 + 32 > p.n_rows) dead |= r0 >= p.n_rows ? 0xFFFFFFFFu : ~((1u << (p.n_rows - r0)) - 1u);
                            float s[32];
                            if constexpr (L2) {
                                float nv[32];
                                lds_f32x32(ns_s + c * 128, nv);
#pragma unroll
                                for (int j = 0; j < 32; ++j) s[j] = fmaf(2.0f, __uint_as_float(v[j]), -nv[j]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) s[j] = __uint_as_float(v[j]);
                            }
                            float m = __int_as_float(0xff800000);
                            int jm = -1;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const bool better = !((dead >> j) & 1u) && s[j] > m;     // NaN never wins
                                m = better ? s[j] : m;
                                jm = better ? j : jm;
                            }
                            dd[c] = jm >= 0 ? make_key(-m, r0 + jm) : KEY_SENTINEL;
                        }
                    }
                    // all TMEM reads of this accumulator are done: one arrive per warp on the LEADER's barrier
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));
                } else if (p.dbg == 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));
                } else {
                    // TMEM reads are double-buffered: chunk c+1 is in flight while chunk c is filtered.
                    // Survivors are rare (about 1 score in 3000), so the common case of a chunk of 32 scores is a
                    // min/max tree and ONE branch.  A fair share of the warp-chunks hold a survivor in some lane:
                    // that path walks the 8 group extrema of the tree and tests single scores only inside a
                    // group that has one.
                    auto filter_chunk = [&](const uint32_t (&v)[32], int c) {
                        if (p.dbg == 8) {
                            if (v[0] == 0x12345678u && v[31] == 0x9abcdef0u) ++head;
                            return;
                        }
                        // s[j] = -(approximate distance): a row passes iff s[j] > nthr  (exactly d < thr)
                        float s[32], g[8];
                        if constexpr (L2) {
                            float nv[32];
                            lds_f32x32(ns_s + c * 128, nv);
#pragma unroll
                            for (int j = 0; j < 32; ++j) s[j] = fmaf(2.0f, __uint_as_float(v[j]), -nv[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) s[j] = __uint_as_float(v[j]);
                        }
#pragma unroll
                        for (int gi = 0; gi < 8; ++gi)
                            g[gi] = fmaxf(fmaxf(s[4 * gi], s[4 * gi + 1]), fmaxf(s[4 * gi + 2], s[4 * gi + 3]));
                        const float m = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])),
                                              fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
                        if (m > nthr) {
                            // deleted rows stop here (p.tomb is null while the shard has none): a shard that is
                            // mostly tombstones would otherwise fill the level buffers with dead survivors.
                            // The chunk's 32 rows share one bitmap word.
                            const uint32_t dead = p.tomb ? __ldg(p.tomb + ((row0 + col0 + c * 32) >> 5)) : 0u;
#pragma unroll
                            for (int gi = 0; gi < 8; ++gi) {
                                if (g[gi] > nthr) {
                                    // up to 4 appends: make room first (tail_seen is a stale copy: the ring can only
                                    // be emptier than it says); publish what is pending before waiting for the movers
                                    if (head - tail_seen > (uint32_t)(RING - 4)) {
                                        if (dirty) { __threadfence_block(); ctl->head_pub[et] = head; dirty = false; }
                                        while (head - (tail_seen = *my_tail) > (uint32_t)(RING - 4)) __nanosleep(32);
                                    }
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        if (s[4 * gi + e] > nthr && !((dead >> (4 * gi + e)) & 1u)) {
                                            st_shared_u64(my_ring_s + slot * 8,
                                                          make_key(-s[4 * gi + e], row0 + col0 + c * 32 + 4 * gi + e));
                                            ++head;
                                            if (++slot == RING) slot = 0;
                                        }
                                    }
                                    dirty = true;
                                }
                            }
                        }
                    };
                    // single-buffered: the other epilogue warp of this sub-partition covers the TMEM latency
#pragma unroll 1
                    for (int c = 0; c < NCH; ++c) {
                        uint32_t v[32];
                        tmem_ld32(taddr + c * 32, v);
                        tmem_ld_wait();
                        if (c == NCH - 1) {
                            // this thread's share of the accumulator is in registers: hand it back before
                            // filtering the last chunk
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));
                        }
                        filter_chunk(v, c);
                    }
                    if (dirty) {    // publish this tile's keys to the movers (once per tile, not per chunk)
                        __threadfence_block();
                        ctl->head_pub[et] = head;
                        dirty = false;
                    }
                }
                ++tcount;
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            atomicAdd(&ctl->done, 1u);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // the peer may still be arriving on / copying against this CTA's barriers
    if (warp == 2) tmem_dealloc(tmem_base, GT_TMEM_COLS);
}

// error model of the approximate values (approximate-value units: ||d||^2 - 2 q.d for L2, -q.d otherwise):
// |approximate + offset - exact| <= eps for every row, offset = ||q||^2 (L2) / 1 (ip, cosine)
struct EpsModel {
    float eps_rel;            // bound on |approx dot - exact dot| / (||q|| ||d||)
    float eps_abs;            // + eps_abs * (||q|| + ||d||max): fp16 subnormal rounding of single elements
    int metric;               // 0 = L2, 1 = ip / cosine
    float eps_dd;             // L2 in direct form over rounded rows (candidates of the shadow-plane scan): + eps_dd * ||d||^2max
    float eps_sum;            //   ... + eps_sum * (||q||^2 + ||d||^2max): fp32 accumulation of the dim squared differences
};
// `at` = the value (offset included) whose magnitude scales the fp32 slack
__device__ __forceinline__ float approx_eps(const EpsModel& m, float qn2, float dmax2, float at) {
    const float eb = m.eps_rel * sqrtf(qn2) * sqrtf(dmax2) + m.eps_abs * (sqrtf(qn2) + sqrtf(dmax2));
    return m.metric == 0 ? 2.0f * eb + m.eps_dd * dmax2 + (m.eps_sum + 4e-7f) * (qn2 + dmax2) + 4e-7f * fabsf(at)
                         : eb + 4e-7f * (1.0f + fabsf(at));
}

// ------------------------------------------------------------------------------------------
// K2s: per query, keep the KP best approximate keys of what the level collected; their KP-th value
// is the next level's threshold.  Padding rows of the last tile and tombstoned rows are dropped here.
// ------------------------------------------------------------------------------------------
constexpr int SELECT_COUNT_MAX = 512;     // up to this many keys a level select ranks by counting (select_kernel)
struct SelectParams {
    uint64_t* buf; int* cnt; int cap; int kp;
    float* thr; int* overflow;
    const uint32_t* tomb; uint32_t n_rows;
    int probe_cnt;            // > 0: the probe level wrote this many keys per query at fixed positions; they only
                              // yield the first threshold (the levels that follow visit the probed tiles again)
    int thr_rank;             // the next level's threshold = the thr_rank-th best kept value (<= kp)
    int np2;                  // keys the shared array sk[] holds (>= cap); kept[kp] lies behind it
    int min_rank;             // probe: with fewer live chunk minima than thr_rank, the largest of them serves if
                              // there are at least this many
    // Tight thresholds: the published threshold is min(thr_rank-th best, a_k + margin * eps), a_k = the k-th best
    // approximate value seen so far.  Any value is a VALID threshold (rows are only ever dropped at or above it);
    // this one is the smallest that still lets the certificate of the re-rank pass: the k-th exact distance is
    // <= a_k + offset + eps, the certificate asks for < threshold + offset - eps.  margin = 0: rank rule only.
    int k; float margin;
    const float* qn2; const unsigned int* max_sqnorm_bits;
    EpsModel em;
};

// Radix select on the 32 distance bits (4 passes of 8 bits): O(n) instead of a full sort.  Keys whose
// distance bits equal the KP-th value are taken in arrival order until KP are kept.  Output unordered.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) select_kernel(const SelectParams p) {
    pdl_prologue();
    extern __shared__ uint64_t sk[];
    __shared__ int hist[256];
    __shared__ uint32_t s_prefix, s_rank, s_min, s_max;
    __shared__ int s_valid, s_c1, s_c2;
    __shared__ uint32_t sh[8];
    __shared__ uint32_t s_tr, s_ak;
    const size_t q = blockIdx.x;
    uint64_t* b = p.buf + q * (size_t)p.cap;
    // threshold of the tight rule from the bits of a_k
    auto tight_thr = [&](uint32_t ak_bits) -> float {
        const float a = ordered_to_float(ak_bits);
        const float off = p.em.metric == 0 ? p.qn2[q] : 1.0f;
        const float t = a + p.margin * approx_eps(p.em, p.qn2[q], __uint_as_float(*p.max_sqnorm_bits), a + off);
        return t == t ? t : __int_as_float(0x7f800000);          // NaN: no tightening
    };
    int n = p.probe_cnt > 0 ? p.probe_cnt : p.cnt[q];
    if (n > p.cap) {
        if (threadIdx.x == 0) p.overflow[q] = 1;
        n = p.cap;
    }
    if (threadIdx.x == 0) { s_valid = 0; s_c1 = 0; s_c2 = 0; s_prefix = 0; s_min = 0xFFFFFFFFu; s_max = 0u; }
    __syncthreads();
    int my_valid = 0;
    uint32_t vlo = 0xFFFFFFFFu, vhi = 0u;
    for (int i = threadIdx.x; i < n; i += THREADS) {
        uint64_t key = b[i];
        if (key != KEY_SENTINEL) {
            const uint32_t row = (uint32_t)key;
            if (row >= p.n_rows || (p.tomb && ((p.tomb[row >> 5] >> (row & 31)) & 1u))) key = KEY_SENTINEL;
        }
        sk[i] = key;
        if (key != KEY_SENTINEL) {
            ++my_valid;
            vlo = min(vlo, (uint32_t)(key >> 32)); vhi = max(vhi, (uint32_t)(key >> 32));
        }
    }
    my_valid = warp_sum_int(my_valid);
    vlo = __reduce_min_sync(0xffffffffu, vlo);
    vhi = __reduce_max_sync(0xffffffffu, vhi);
    if ((threadIdx.x & 31) == 0 && my_valid) { atomicAdd(&s_valid, my_valid); atomicMin(&s_min, vlo); atomicMax(&s_max, vhi); }
    __syncthreads();
    const int n_valid = s_valid;
    __syncthreads();
    const bool probe = p.probe_cnt > 0;
    // 1-based rank of the key we are looking for
    const int want = !probe ? p.kp : (n_valid >= p.thr_rank ? p.thr_rank : n_valid);
    if (!probe && n_valid > p.kp && n <= SELECT_COUNT_MAX) {
        // Few keys (the usual case for k' <= 64: one to three hundred survivors per level): rank every key by counting
        // the smaller ones -- keys are distinct, sentinels are the largest value -- and write the k' best straight to
        // their rank.  One pass and two barriers instead of up to four histogram passes with three barriers each.
        const bool tight = p.margin > 0.0f && p.k < p.kp;
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const uint64_t key = sk[i];
            if (key == KEY_SENTINEL) continue;
            int rank = 0;
#pragma unroll 8
            for (int j = 0; j < n; ++j) rank += sk[j] < key;
            if (rank < p.kp) b[rank] = key;
            if (rank == p.thr_rank - 1) s_tr = (uint32_t)(key >> 32);
            if (rank == p.k - 1) s_ak = (uint32_t)(key >> 32);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float thr = ordered_to_float(s_tr);
            if (tight) thr = fminf(thr, tight_thr(s_ak));
            p.thr[q] = thr;
            p.cnt[q] = p.kp;
        }
        return;
    }
    if (probe && n_valid < p.min_rank) {
        if (threadIdx.x == 0) { p.cnt[q] = 0; p.thr[q] = __int_as_float(0x7f800000); }   // too few live rows probed
        return;
    }
    if (probe && n <= SELECT_COUNT_MAX) {
        // the probe's chunk minima (128 per query with 16 probe tiles): ranks by counting, as for the levels below
        const bool tight = p.margin > 0.0f && n_valid >= p.k && p.k < want;
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const uint64_t key = sk[i];
            if (key == KEY_SENTINEL) continue;
            int rank = 0;
#pragma unroll 8
            for (int j = 0; j < n; ++j) rank += sk[j] < key;
            if (rank == want - 1) s_tr = (uint32_t)(key >> 32);
            if (rank == p.k - 1) s_ak = (uint32_t)(key >> 32);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float thr = ordered_to_float(s_tr);
            if (tight) thr = fminf(thr, tight_thr(s_ak));
            p.cnt[q] = 0;
            p.thr[q] = thr;
        }
        return;
    }
    if (!probe && n_valid <= p.kp) {
        // everything valid is kept
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const uint64_t key = sk[i];
            if (key != KEY_SENTINEL) b[atomicAdd(&s_c1, 1)] = key;
        }
        __syncthreads();
        for (int i = n_valid + threadIdx.x; i < p.kp; i += THREADS) b[i] = KEY_SENTINEL;
        // rows dropped so far are >= the current threshold and nothing tighter is known: it stays
        if (threadIdx.x == 0) p.cnt[q] = p.kp;
        return;
    }
    // The passes start at the highest byte in which the values differ: the bytes above it (sign, exponent) are
    // common to all keys and would send every atomicAdd of a pass to ONE histogram bin.
    const uint32_t diff = s_min ^ s_max;
    const int first_pass = diff ? 3 - ((31 - __clz(diff)) >> 3) : 4;     // all values equal: no pass
    uint32_t mask = first_pass == 0 ? 0u : (first_pass == 4 ? 0xFFFFFFFFu : ~((1u << (32 - 8 * first_pass)) - 1u));
    if (threadIdx.x == 0) {
        s_rank = (uint32_t)want;
        s_prefix = s_min & mask;
    }
    __syncthreads();
    for (int pass = first_pass; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += THREADS) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        for (int i = threadIdx.x; i < n; i += THREADS) {
            const uint64_t key = sk[i];
            if (key != KEY_SENTINEL) {
                const uint32_t hi = (uint32_t)(key >> 32);
                if ((hi & mask) == prefix) atomicAdd(&hist[(hi >> shift) & 255], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // 8 bins per lane, warp prefix sum, find the bin that holds rank s_rank
            int local[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { local[i] = hist[threadIdx.x * 8 + i]; sum += local[i]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)threadIdx.x >= o) incl += t;
            }
            const int excl = incl - sum;
            const int rank = (int)s_rank;
            if (rank > excl && rank <= incl) {
                int run = excl;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (rank > run && rank <= run + local[i]) {
                        s_prefix = prefix | ((uint32_t)(threadIdx.x * 8 + i) << shift);
                        s_rank = (uint32_t)(rank - run);
                    }
                    run += local[i];
                }
            }
        }
        mask |= 0xFFu << shift;
        __syncthreads();
    }
    const uint32_t T = s_prefix;                  // distance bits of the want-th best key
    if (probe) {
        float thr = ordered_to_float(T);
        if (p.margin > 0.0f && n_valid >= p.k && p.k < want) {
            // the k-th smallest chunk minimum bounds the k-th best probed row from above (sentinels sort last)
            __syncthreads();
            thr = fminf(thr, tight_thr(block_kth_bits(sk, n, p.k, hist, sh)));
        }
        if (threadIdx.x == 0) { p.cnt[q] = 0; p.thr[q] = thr; }
        return;
    }
    const int need_eq = (int)s_rank;              // how many keys with exactly these bits to keep
    const int n_less = p.kp - need_eq;
    uint64_t* kept = sk + p.np2;                  // [kp] behind the key array
    const bool tight = p.margin > 0.0f && p.k < p.kp;
    const bool tighter = p.thr_rank < p.kp || tight;
    for (int i = threadIdx.x; i < n; i += THREADS) {
        const uint64_t key = sk[i];
        if (key == KEY_SENTINEL) continue;
        const uint32_t hi = (uint32_t)(key >> 32);
        int dst = -1;
        if (hi < T) dst = atomicAdd(&s_c1, 1);
        else if (hi == T) {
            const int s = atomicAdd(&s_c2, 1);
            if (s < need_eq) dst = n_less + s;
        }
        if (dst >= 0) {
            b[dst] = key;
            if (tighter) kept[dst] = key;
        }
    }
    if (threadIdx.x == 0) p.cnt[q] = p.kp;
    if (!tighter) {
        if (threadIdx.x == 0) p.thr[q] = ordered_to_float(T);
        return;
    }
    // The kp best stay candidates; the threshold is the thr_rank-th best of them.  While only a small part of the
    // shard has been seen the kp-th best of the sample is far looser than the certificate needs, and every row
    // that passes it costs a trip through the key rings (rank by counting: keys are distinct).
    __syncthreads();
    for (int t = threadIdx.x; t < p.kp; t += THREADS) {
        const uint64_t key = kept[t];
        int rank = 0;
        for (int j = 0; j < p.kp; ++j) rank += kept[j] < key;
        if (rank == p.thr_rank - 1) s_tr = (uint32_t)(key >> 32);
        if (rank == p.k - 1) s_ak = (uint32_t)(key >> 32);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float thr = ordered_to_float(s_tr);
        if (tight) thr = fminf(thr, tight_thr(s_ak));
        p.thr[q] = thr;
    }
}

// ------------------------------------------------------------------------------------------
// K4w: exact re-rank of the candidates that can still matter + coverage certificate
// ------------------------------------------------------------------------------------------
struct RerankParams {
    const uint64_t* approx;   // [nq][stride] level buffers: approximate keys (distance-like value, row), unordered
    size_t stride;
    const int* overflow;      // [nq] 1 = the candidate buffer overflowed: certificate void
    const float* tau;         // [nq] last level's threshold: every row outside the buffer is >= it (+inf: there is none)
    const void* rows; uint32_t row_bytes; uint32_t ld;
    const uint32_t* labels;
    const float* q;           // prepared queries [nq][ld] fp32
    const float* qn2;         // [nq]
    const unsigned int* max_sqnorm_bits;
    int k, metric;            // metric 0 = L2 (approx value = ||d||^2 - 2 q.d), 1 = ip/cos (approx value = -q.d)
    EpsModel em;              // error model of the approximate values
    int f16_range;            // operands were rounded to fp16: norms beyond its range void the certificate
    int64_t* out_ids; float* out_dist; int* out_counts;
    int* flags;               // [nq] 1 = certificate failed: exact_fallback_kernel re-searches the query
    int* n_flagged;           // flagged queries of this batch (zeroed with the batch's other flags)
    unsigned long long* n_fallback;   // cumulative count of such queries (statistics; never reset)
    const int* cnt;           // [nq] keys in the buffer (may exceed cap: overflow)
    int cap;
    const uint32_t* tomb; uint32_t n_rows;
    int approx_is_dist;       // the keys and tau carry approximate DISTANCES (offset included), not contraction values
};

// Exact fp32 distances of (query, row) in the scan kernel's summation order -> sortable keys; whole warp, NR rows
// at once: the loads of all rows (4 lane chunks each) are issued before the first FMA,
// so a warp keeps NR x 4 x 16 bytes in flight per lane instead of 16 (the window re-rank reads a handful of rows
// per query: its time is memory latency, not bandwidth).
template <typename T, int NR>
__device__ __forceinline__ void exact_keys(const RerankParams& p, const float* qs, const uint32_t (&row)[NR], int lane,
                                           uint64_t (&out)[NR]) {
    const int nld16 = p.row_bytes / 512;
    constexpr int PER16 = 16 / sizeof(T);
    constexpr int U = 4;
    const uint8_t* rp[NR];
    float acc[NR];
    uint32_t label[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        rp[r] = reinterpret_cast<const uint8_t*>(p.rows) + (size_t)row[r] * p.row_bytes + (size_t)lane * 16;
        acc[r] = 0.0f;
        label[r] = __ldg(p.labels + row[r]);
    }
    for (int ch0 = 0; ch0 < nld16; ch0 += U) {
        uint4 raw[NR][U];
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ch0 + u < nld16) raw[r][u] = __ldg(reinterpret_cast<const uint4*>(rp[r] + (size_t)(ch0 + u) * 512));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ch0 + u < nld16) {
                float qq[PER16];
                const float* qp = qs + (size_t)((ch0 + u) * 32 + lane) * PER16;
#pragma unroll
                for (int e = 0; e < PER16; e += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(qp + e);
                    qq[e] = t.x; qq[e + 1] = t.y; qq[e + 2] = t.z; qq[e + 3] = t.w;
                }
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    float dv[PER16];
                    if constexpr (sizeof(T) == 4) {
                        dv[0] = __uint_as_float(raw[r][u].x); dv[1] = __uint_as_float(raw[r][u].y);
                        dv[2] = __uint_as_float(raw[r][u].z); dv[3] = __uint_as_float(raw[r][u].w);
                    } else {
                        const __half2* h = reinterpret_cast<const __half2*>(&raw[r][u]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 f = __half22float2(h[i]);
                            dv[2 * i] = f.x; dv[2 * i + 1] = f.y;
                        }
                    }
                    if (p.metric == 0) {
#pragma unroll
                        for (int e = 0; e < PER16; ++e) {
                            const float t = dv[e] - qq[e];
                            acc[r] = fmaf(t, t, acc[r]);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < PER16; ++e) acc[r] = fmaf(dv[e], qq[e], acc[r]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const float a = warp_sum_butterfly(acc[r]);
        out[r] = make_key(p.metric == 0 ? a : 1.0f - a, label[r]);
    }
}

// ------------------------------------------------------------------------------------------
// K4w: the re-rank straight from the level buffers -- no select after the last level, and only the candidates that can
// still reach the exact top-k are re-ranked.
//   Buffer of query q after the last level: the k' keys the previous select carried over + every row of the last
//   level whose approximate value beat that level's threshold thr.  Every row that is NOT in the buffer failed a
//   threshold >= thr (thresholds only tighten) or was dropped by a select at a value >= thr: approximate >= thr.
//   With |approximate + offset - exact| <= eps for every row:
//   (1) window: let a_k = k-th smallest approximate value in the buffer.  k rows have exact <= a_k + offset + eps,
//       so a row with approximate > a_k + 2 eps (exact > a_k + offset + eps) is strictly worse than k rows: only
//       the keys at or below a_k + 2 eps (a handful beyond k) are read back from the fp32 rows;
//   (2) certificate: k-th exact distance < thr + offset - eps  =>  no row outside the buffer can enter the top-k.
// ------------------------------------------------------------------------------------------
// One block of 128 threads per query; six short stages, a barrier between them:
//   1. a' = k-th smallest of the k' carried keys (rank by counting): a' >= a_k
//   2. list = buffer keys at or below a' + 2 eps (expected ~8k of them: a' is the k-th best of 1/8 of the rows)
//   3. a_k = k-th smallest of the list (radix select over the bits in which the list's values differ)
//   4. window = list keys at or below a_k + 2 eps
//   5. exact distances, two rows in flight per warp   6. results written at their rank (by counting), certificate
constexpr int RW_MAX_THREADS = 256;            // 128 threads per query for k' <= 64, 256 above

// ------------------------------------------------------------------------------------------
// Exact re-search of ONE query by its block (the certificate failed or a level buffer overflowed: floods of
// near-duplicates, values beyond the fp16 range, adversarial insertion orders).  Every live row of the shard is
// scored in fp32 with the scan kernel's summation order; candidates that beat the current k-th key collect in a
// shared-memory buffer that is cut back to the k best whenever it fills.  A block streams the shard at a fraction of
// the HBM rate -- this is the slow, always-correct road; what matters is that it needs no host round trip: the
// search stays one enqueue on the caller's stream whatever the data looks like.
//   buf / scratch: [buf_n] keys each in shared memory, buf_n >= k + rows per round
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_keep_k_smallest(uint64_t* buf, int n, int k, uint64_t* scratch, int* hist, uint32_t* sh,
                                                     int* s_a, int* s_b) {
    // -> number kept (min(n, k)); kept keys end up in buf[0..kept), unordered.  Keys are distinct.
    const int tid = threadIdx.x, NT = blockDim.x;
    __syncthreads();
    if (n <= k) return n;
    const uint32_t T = block_kth_bits(buf, n, k, hist, sh);       // value bits of the k-th smallest
    if (tid == 0) { *s_a = 0; *s_b = 0; }
    __syncthreads();
    int less = 0;
    for (int i = tid; i < n; i += NT) less += (uint32_t)(buf[i] >> 32) < T;
    less = warp_sum_int(less);
    if ((tid & 31) == 0 && less) atomicAdd(s_a, less);
    __syncthreads();
    const int n_less = *s_a, need_eq = k - n_less;               // keys with bits == T to keep: the need_eq smallest
    __syncthreads();
    if (tid == 0) *s_a = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const uint64_t key = buf[i];
        const uint32_t v = (uint32_t)(key >> 32);
        bool keep = v < T;
        if (v == T) {
            int r = 0;
            for (int j = 0; j < n; ++j) { const uint64_t o = buf[j]; r += ((uint32_t)(o >> 32) == T) && o < key; }
            keep = r < need_eq;
        }
        if (keep) scratch[atomicAdd(s_a, 1)] = key;
    }
    __syncthreads();
    const int kept = *s_a;
    for (int i = tid; i < kept; i += NT) buf[i] = scratch[i];
    __syncthreads();
    return kept;
}

template <typename T>
__device__ void exact_scan_block(const RerankParams& p, int q, uint64_t* buf, uint64_t* scratch, int buf_n, int* hist, uint32_t* sh) {
    __shared__ int s_cnt, s_a, s_b;
    __shared__ unsigned long long s_thr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5, k = p.k;
    constexpr int NR = 4;
    const int per_round = nwarps * NR;
    if (tid == 0) { s_cnt = 0; s_thr = KEY_SENTINEL; }
    __syncthreads();
    const float* qv = p.q + (size_t)q * p.ld;
    for (uint32_t r0 = 0; r0 < p.n_rows; r0 += per_round) {
        uint32_t rows[NR];
        bool live[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const uint32_t row = r0 + warp * NR + i;
            live[i] = row < p.n_rows && !(p.tomb && ((p.tomb[row >> 5] >> (row & 31)) & 1u));
            rows[i] = row < p.n_rows ? row : p.n_rows - 1;
        }
        uint64_t keys[NR];
        exact_keys<T, NR>(p, qv, rows, lane, keys);
        if (lane == 0) {
            const unsigned long long thr = s_thr;
#pragma unroll
            for (int i = 0; i < NR; ++i)
                if (live[i] && keys[i] < thr) buf[atomicAdd(&s_cnt, 1)] = keys[i];
        }
        __syncthreads();
        if (s_cnt + per_round > buf_n) {                          // block-uniform: cut back to the k best
            const int kept = block_keep_k_smallest(buf, s_cnt, k, scratch, hist, sh, &s_a, &s_b);
            if (tid == 0) {
                s_cnt = kept;
                if (kept == k) {                                  // the k-th best so far bounds what can still matter
                    unsigned long long mx = 0;
                    for (int i = 0; i < kept; ++i) mx = buf[i] > mx ? buf[i] : mx;
                    s_thr = mx;
                }
            }
            __syncthreads();
        }
    }
    const int m = block_keep_k_smallest(buf, s_cnt, k, scratch, hist, sh, &s_a, &s_b);
    for (int i = tid; i < m; i += blockDim.x) {
        const uint64_t key = buf[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) rank += buf[j] < key;
        p.out_ids[(size_t)q * k + rank] = (int64_t)key_label(key);
        p.out_dist[(size_t)q * k + rank] = key_dist(key);
    }
    for (int i = m + tid; i < k; i += blockDim.x) {
        p.out_ids[(size_t)q * k + i] = -1;
        p.out_dist[(size_t)q * k + i] = __int_as_float(0x7f800000);
    }
    if (tid == 0 && p.out_counts) p.out_counts[q] = m;
}

template <typename T, int NRW>
__global__ void __launch_bounds__(RW_MAX_THREADS) rerank_window_kernel(const RerankParams p, const int kp) {
    pdl_prologue();
    extern __shared__ uint64_t wsm[];
    uint64_t* sk = wsm;                // [cap] buffer keys; from stage 4 the window keys
    uint64_t* lk = wsm + p.cap;        // [cap] list; from stage 5 the exact keys
    __shared__ int hist[256];
    __shared__ uint32_t sh[8];
    __shared__ uint32_t s_a1;
    __shared__ int s_m1, s_m;
    __shared__ float s_kth, s_eps0;
    __shared__ int s_have_kth;
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, RW_THREADS = blockDim.x;
    const int k = p.k;
    const uint64_t* b = p.approx + (size_t)q * p.stride;
    int n = p.cnt[q];
    bool ok = !(n > p.cap || p.overflow[q] != 0);
    n = min(n, p.cap);
    const float qn2 = p.qn2[q];
    const float dmax2 = __uint_as_float(*p.max_sqnorm_bits);
    const float off = p.approx_is_dist ? 0.0f : (p.metric == 0 ? qn2 : 1.0f);
    const float INF = __int_as_float(0x7f800000);
    // value bits of a_x + 2 eps (with a margin for the rounding of the sum); non-finite: keep everything.
    // eps(at) = s_eps0 + 4e-7 |at|: the part with the square roots is computed once per query, not per thread
    auto window_bits = [&](uint32_t a_bits) -> uint32_t {
        const float a = ordered_to_float(a_bits);
        const float wv = a + 2.25f * (s_eps0 + 4e-7f * fabsf(a + off));
        return (wv == wv && fabsf(wv) < INF) ? float_to_ordered(wv) : 0xFFFFFFFFu;
    };
    if (tid == 0) {
        s_a1 = 0xFFFFFFFFu; s_m1 = 0; s_m = 0; s_kth = INF; s_have_kth = 0;
        s_eps0 = approx_eps(p.em, qn2, dmax2, 0.0f);
    }
    // ---- 0. the buffer; padding rows of the last tile and tombstoned rows drop out
    for (int i = tid; i < n; i += RW_THREADS) {
        uint64_t key = b[i];
        if (key != KEY_SENTINEL) {
            const uint32_t row = (uint32_t)key;
            if (row >= p.n_rows || (p.tomb && ((p.tomb[row >> 5] >> (row & 31)) & 1u))) key = KEY_SENTINEL;
        }
        sk[i] = key;
    }
    __syncthreads();
    // ---- 1. bound from the carried keys (sentinels sort last: fewer than k real ones -> keep everything)
    const int nc = min(n, kp);
    if (nc > 64) {
        // k' of 128 / 256: radix select (sentinels carry the largest value bits: a sentinel at rank k reads as
        // "keep everything", like the initial value)
        if (nc >= k) {
            const uint32_t a1 = block_kth_bits(sk, nc, k, hist, sh);
            if (tid == 0) s_a1 = a1;
        }
    } else {
        for (int t = tid; t < nc; t += RW_THREADS) {
            const uint64_t key = sk[t];
            int rank = 0;
            for (int j = 0; j < nc; ++j) rank += sk[j] < key;
            if (rank == k - 1) s_a1 = (uint32_t)(key >> 32);  // keys are distinct (rows are), sentinels are not:
        }                                                      // a sentinel of rank k-1 writes 0xFFFFFFFF as well
    }
    __syncthreads();
    // ---- 2. the keys that can still matter
    const uint32_t w1 = s_a1 == 0xFFFFFFFFu ? 0xFFFFFFFFu : window_bits(s_a1);
    for (int i0 = 0; i0 < n; i0 += RW_THREADS) {
        const int i = i0 + tid;
        const uint64_t key = i < n ? sk[i] : KEY_SENTINEL;
        const bool pass = key != KEY_SENTINEL && (uint32_t)(key >> 32) <= w1;
        const uint32_t bal = __ballot_sync(0xffffffffu, pass);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_m1, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) lk[base + __popc(bal & ((1u << lane) - 1u))] = key;
    }
    __syncthreads();
    const int m1 = s_m1;
    // ---- 3. a_k and the final window
    uint32_t w2 = w1;
    if (m1 > k) {
        if (m1 <= 512) {                       // a few dozen keys as a rule: rank by counting (one barrier) instead of
            for (int t = tid; t < m1; t += RW_THREADS) {                       // a radix select (up to twelve)
                const uint64_t key = lk[t];
                int rank = 0;
                for (int j = 0; j < m1; ++j) rank += lk[j] < key;
                if (rank == k - 1) sh[5] = (uint32_t)(key >> 32);
            }
            __syncthreads();
            w2 = min(w1, window_bits(sh[5]));
        } else {
            w2 = min(w1, window_bits(block_kth_bits(lk, m1, k, hist, sh)));
        }
    }
    // ---- 4. window keys -> sk (stage 2 has finished reading it: barriers above)
    for (int i0 = 0; i0 < m1; i0 += RW_THREADS) {
        const int i = i0 + tid;
        const uint64_t key = i < m1 ? lk[i] : KEY_SENTINEL;
        const bool pass = i < m1 && (uint32_t)(key >> 32) <= w2;
        const uint32_t bal = __ballot_sync(0xffffffffu, pass);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_m, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) sk[base + __popc(bal & ((1u << lane) - 1u))] = key;
    }
    __syncthreads();
    int m = s_m;
    // ---- 5. exact distances -> lk
    const float* qv = p.q + (size_t)q * p.ld;
    // rows in flight per warp.  Measured (B200): 4 rows cost 104 registers -> half the resident blocks of 2 rows;
    // with k' <= 64 (a dozen window rows per query, 8192 queries) 2 rows win (b8192 x 125k: 0.977 vs 1.029 ms/step),
    // with k' >= 128 (~130 window rows per query) 4 rows win (config-3 shard: 5.08 vs 5.14 ms/step): the window is a handful of rows per query,
    for (int c = NRW * warp; c < m; c += NRW * (RW_THREADS / 32)) {   // its time is memory latency (a ragged tail
        uint32_t rows[NRW];                                          // repeats the last row, result discarded)
#pragma unroll
        for (int i = 0; i < NRW; ++i) rows[i] = (uint32_t)sk[min(c + i, m - 1)];
        uint64_t out[NRW];
        exact_keys<T, NRW>(p, qv, rows, lane, out);
#pragma unroll
        for (int i = 0; i < NRW; ++i)
            if (lane == i && c + i < m) lk[c + i] = out[i];
    }
    __syncthreads();
    // ---- 6. results at their rank (keys are distinct: labels are), certificate
    if (m > 2048) { ok = false; m = 0; }            // degenerate (thousands of near-ties): leave it to the scan
    for (int i = tid; i < m; i += RW_THREADS) {
        const uint64_t key = lk[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) rank += lk[j] < key;
        if (rank < k) {
            p.out_ids[(size_t)q * k + rank] = (int64_t)key_label(key);
            p.out_dist[(size_t)q * k + rank] = key_dist(key);
            if (rank == k - 1) { s_kth = key_dist(key); s_have_kth = 1; }
        }
    }
    for (int i = m + tid; i < k; i += RW_THREADS) {
        p.out_ids[(size_t)q * k + i] = -1;
        p.out_dist[(size_t)q * k + i] = INF;
    }
    __syncthreads();
    if (tid == 0) {
        if (p.out_counts) p.out_counts[q] = min(m, k);
        const float a_tau = p.tau[q];
        if (a_tau < INF) {                          // rows outside the buffer exist: prove they cannot matter
            const float tau = a_tau + off;
            ok = ok && s_have_kth && s_kth < tau - approx_eps(p.em, qn2, dmax2, tau);
        }
        // an element above the fp16 range became +-inf in the operand plane (|x_i| <= ||x||): no bound holds
        if (p.f16_range && (qn2 >= 4.0e9f || dmax2 >= 4.0e9f)) ok = false;
        p.flags[q] = ok ? 0 : 1;
        if (!ok) { atomicAdd(p.n_flagged, 1); atomicAdd(p.n_fallback, 1ull); }   // exact_fallback_kernel takes it from here
    }
}

// K4x: exact re-search of the queries K4w flagged, launched after it on every batch.  Almost always there are none:
// every block reads one counter and leaves (the launch overlaps K4w's tail through programmatic dependent launch).
// Otherwise the flagged queries are dealt round-robin to the blocks of the grid.
constexpr int XF_THREADS = 256;
constexpr int XF_BUF = 1024;
template <typename T>
__global__ void __launch_bounds__(XF_THREADS) exact_fallback_kernel(const RerankParams p, const int nq) {
    pdl_prologue();
    if (*reinterpret_cast<const volatile int*>(p.n_flagged) == 0) return;
    __shared__ uint64_t buf[XF_BUF], scratch[XF_BUF];
    __shared__ int hist[256];
    __shared__ uint32_t sh[8];
    int ord = 0;
    for (int q = 0; q < nq; ++q) {
        if (p.flags[q] == 0) continue;                            // block-uniform
        if (ord % (int)gridDim.x == (int)blockIdx.x) {
            exact_scan_block<T>(p, q, buf, scratch, XF_BUF, hist, sh);
            __syncthreads();
        }
        ++ord;
    }
}

// prepared fp32 queries [nq][ld] -> fp16 [nq][ld16] (ld16 <= ld; padding columns are zero)
__global__ void f32_to_f16_kernel(const float* __restrict__ in, int ld, __half* __restrict__ out, int ld16, size_t nq) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= nq * (size_t)ld16) return;
    const size_t r = i / ld16, c = i - r * ld16;
    out[i] = __float2half_rn(in[r * ld + c]);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct GemmWsImpl {
    uint64_t* buf = nullptr; size_t buf_cap = 0;
    int* cnt = nullptr; size_t cnt_cap = 0;
    float* thr = nullptr; size_t thr_cap = 0;
    int* overflow = nullptr; size_t ovf_cap = 0;
    __half* q16 = nullptr; size_t q16_cap = 0;
    int* flags = nullptr; size_t flags_cap = 0;
    int* n_flagged = nullptr;
    unsigned long long* n_fallback = nullptr;     // cumulative, read by gemm_workspace_fallbacks
};
struct GemmPlanImpl {
    long fallbacks = 0;
};

static PFN_cuTensorMapEncodeTiled get_encode() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(sym);
    });
    return fn;
}

// rows x (ld elements) row-major, box = {128 bytes of a row, box_rows}, SWIZZLE_128B
static bool make_tmap(CUtensorMap* tm, const void* base, bool f16, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    auto enc = get_encode();
    if (!enc) return false;
    const uint32_t esz = f16 ? 2 : 4;
    cuuint64_t gdim[2] = {ld, rows};
    cuuint64_t gstride[1] = {ld * esz};
    cuuint32_t box[2] = {128u / esz, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static int kp_for_k(int k) {
    int kp = 32;
    while (kp < 2 * k) kp <<= 1;
    return kp;
}
static int env_int(const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; }
// probe tiles: 8 chunk minima per tile and query; with 4x as many chunks as the threshold rank the rank-th chunk
// minimum is close to the rank-th best probed row (VDB_PROBE_TILES / VDB_GROWTH: sweeps, tools/level_sweep.sh)
static int probe_tiles(int rank) {
    static const int v = env_int("VDB_PROBE_TILES", 0);
    return std::max(v > 0 ? v : GT_PROBE_TILES, (rank + 1) / 2);
}
// each level covers `growth` x the rows that informed its threshold.  Survivors per level ~ threshold rank x growth:
// for large k (k' = 256, ~1300 survivors per query and level at growth 8) the epilogue's slow path and the selects
// dominate, and a level more with half the survivors wins (measured on 1.25M x 512, L2, k = 100, batch 4096:
// growth 8 5.61 ms/step, growth 4 5.15; k = 10 on 1M rows: growth 8 is the best that keeps the buffers half empty)
static int level_growth(int k) {
    static const int v = env_int("VDB_GROWTH", 0);
    return v >= 2 ? v : (k >= 64 ? 4 : GT_LEVEL_GROWTH);
}
// keys a query's level buffer holds: 16 k' (the levels are planned for <= 55 % of that); a shard of at most 32 k'
// rows gets a buffer that holds every row, because its probe may see fewer live chunks than the threshold rank
// and then publishes no threshold (the single level that follows keeps everything)
static int cap_for(int kp, size_t n_rows) {
    static const int mult = std::max(8, env_int("VDB_CAP_MULT", 16));
    int cap = mult * kp;
    if (n_rows <= (size_t)32 * kp)
        while ((size_t)cap < n_rows) cap <<= 1;
    return cap;
}

bool gemm_topk_supported(int dim, int ld, bool f16, int k, size_t n_rows) {
    (void)dim;
    if (k < 1 || k > 128) return false;
    if (n_rows < 1) return false;
    const size_t row_bytes = (size_t)ld * (f16 ? 2 : 4);
    return row_bytes % 128 == 0;
}

template <typename T>
static cudaError_t grow_dev(T*& ptr, size_t& cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    cudaError_t e = cudaMalloc((void**)&ptr, need * sizeof(T));
    if (e == cudaSuccess) cap = need;
    return e;
}

// cudaFuncSetAttribute is per device: a process that holds shards on several GPUs (one handler per GPU under a
// LocalCoordinator) must configure every kernel once on each of them
constexpr int MAX_DEVICES = 64;
static int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev & (MAX_DEVICES - 1);
}

template <bool F16, bool L2, bool ARES>
static cudaError_t launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& gp, int grid, cudaStream_t st) {
    static bool configured[MAX_DEVICES] = {};
    const int slot = current_device_slot();
    constexpr int smem_bytes = GtSmem<ARES>::BYTES;
    if (!configured[slot]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_filter_kernel<F16, L2, ARES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return e;
        configured[slot] = true;
    }
    cudaError_t e = launch_pdl(gemm_filter_kernel<F16, L2, ARES>, dim3(grid), dim3(GT_THREADS), smem_bytes, st, tmA, tmB, gp);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <typename T>
static cudaError_t launch_rerank_window(int kp, const RerankParams& rp, size_t nq, cudaStream_t st) {
    const size_t smem = 2 * (size_t)rp.cap * sizeof(uint64_t);
    static size_t configured[MAX_DEVICES] = {};
    const int slot = current_device_slot();
    if (smem > configured[slot]) {
        cudaError_t e = cudaFuncSetAttribute(rerank_window_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(rerank_window_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[slot] = smem;
    }
    // (64 / 96 / 128 threads per query measured alike at 8192 queries)
    cudaError_t le = kp <= 64 ? launch_pdl(rerank_window_kernel<T, 2>, dim3((unsigned)nq), dim3(128), smem, st, rp, kp)
                              : launch_pdl(rerank_window_kernel<T, 4>, dim3((unsigned)nq), dim3(RW_MAX_THREADS), smem, st, rp, kp);
    count_launch();
    return le != cudaSuccess ? le : cudaGetLastError();
}

// slices of a level's position range: fill the pairs in whole waves, prefer few slices
static int choose_slices(int MB, int n_pos, int num_pairs) {
    int best = 1;
    double best_cost = 1e30;
    const int smax = std::min(n_pos, 128);
    for (int S = 1; S <= smax; ++S) {
        const long items = (long)MB * S;
        const long waves = (items + num_pairs - 1) / num_pairs;
        const double tiles_per_item = (double)n_pos / S;
        const double cost = waves * (tiles_per_item + 1.0);   // ~1 tile-time of pipeline fill per item
        if (cost < best_cost * 0.999) { best_cost = cost; best = S; }
    }
    return best;
}

static cudaError_t ensure_ws(GemmWorkspace& ws) {
    if (ws.impl) return cudaSuccess;
    auto* w = new GemmWsImpl();
    cudaError_t e = cudaMalloc((void**)&w->n_flagged, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&w->n_fallback, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(w->n_fallback, 0, sizeof(unsigned long long));
    if (e != cudaSuccess) { delete w; return e; }
    ws.impl = w;
    return cudaSuccess;
}

cudaError_t gemm_topk_prep_targets(GemmWorkspace& ws, const GemmSearchArgs& a, GemmPrepTargets* out) {
    cudaError_t e = ensure_ws(ws);
    if (e != cudaSuccess) return e;
    auto* w = static_cast<GemmWsImpl*>(ws.impl);
    const bool g16 = a.f16 || a.shadow != nullptr;
    const int gld = a.f16 ? a.ld : (a.shadow ? a.ld16 : a.ld);
    if ((e = grow_dev(w->overflow, w->ovf_cap, a.nq)) != cudaSuccess) return e;
    if (g16 && (e = grow_dev(w->q16, w->q16_cap, a.nq * (size_t)gld)) != cudaSuccess) return e;
    out->q16 = g16 ? w->q16 : nullptr;
    out->gld = gld;
    out->overflow = w->overflow;
    out->n_flagged = w->n_flagged;
    return cudaSuccess;
}

// The launch sequence of one search as data (pure host arithmetic: tests/test_abi.py checks its invariants on CPU
// over a grid of shapes through vdb_debug_level_plan).
//   Probe: the first P positions of the bit-reversed tile order, chunk minima only -> first threshold.  Then levels
//   from position 0 on, each covering `growth` x the positions that informed its threshold; the last one takes what
//   is left if that is at most 1.5 x the growth and the expected survivors stay under 55 % of the buffer.  A select
//   follows every level but the last (the window re-rank reads the last level's buffer as it is).
//   Threshold rank: while at most a quarter of the shard has been seen, max(k'/2, 1.6 k) instead of k' -- the
//   sample's k'-th best is then far looser than the certificate needs, and every row that passes costs a trip
//   through the key rings.  The k' best stay candidates either way.
//   Small batches (one or two query blocks): the survivors of a level are few whatever the threshold, what costs
//   is every launch's ramp and every select between launches -- a larger probe, then levels that grow 64x (as far
//   as an 8192-key buffer allows), i.e. probe + ONE level for a million rows.  Measured on 1M x 512, batch 5..256
//   (tools/small_batch.py): ~290 -> ~265 us.
LevelPlan gemm_topk_level_plan(size_t nq, size_t n_rows, int k) {
    LevelPlan lp;
    lp.kp = kp_for_k(k);
    lp.kq = std::max(lp.kp / 2, std::min(lp.kp, (16 * k + 9) / 10));
    lp.query_blocks = (int)((nq + 2 * GT_BM - 1) / (2 * GT_BM));       // 256-query blocks, one per CTA pair
    lp.small_batch = lp.query_blocks <= 2 && env_int("VDB_SMALL_PLAN", 1) != 0;
    lp.growth = lp.small_batch ? std::max(level_growth(k), std::min(64, 8192 / (3 * lp.kq))) : level_growth(k);
    lp.cap = cap_for(lp.kp, n_rows);
    if (lp.small_batch)
        while (lp.cap < 3 * lp.kq * lp.growth && lp.cap < 8192) lp.cap <<= 1;
    lp.n_tiles = (int)((n_rows + GT_BN - 1) / GT_BN);
    lp.bits = 0;
    while ((1 << lp.bits) < lp.n_tiles) ++lp.bits;
    lp.n_pos = 1 << lp.bits;                 // positions in bit-reversed order; those mapping past n_tiles are skipped
    const int n_pos = lp.n_pos;
    auto rank_after = [&](int seen) { return 4L * seen <= n_pos ? lp.kq : lp.kp; };
    lp.probe_tiles = std::min(n_pos, std::max(probe_tiles(lp.kq), lp.small_batch ? 64 : 0));
    lp.probe_tiles = std::max(1, std::min(lp.probe_tiles, lp.cap / 8));      // the probe writes 8 keys per tile into the buffer
    lp.probe_rank = rank_after(lp.probe_tiles);
    int rank = lp.probe_rank, seen = lp.probe_tiles, pos = 0;
    while (pos < n_pos) {
        // positions scale with tiles by n_pos / n_tiles (< 2): use positions directly
        const long want = (long)seen * lp.growth;
        int next = (int)std::min<long>(n_pos, pos + want);
        const long left = n_pos - pos;
        if (left <= want + want / 2 && (long)rank * left / seen + lp.kp <= (long)lp.cap * 11 / 20) next = n_pos;
        LevelPlan::Level lv{pos, next, 0};
        if (next < n_pos) lv.rank_after = rank = rank_after(next);
        lp.levels.push_back(lv);
        seen = pos = next;
    }
    return lp;
}

cudaError_t gemm_topk_search(GemmPlan& plan, GemmWorkspace& ws, const GemmSearchArgs& a, cudaStream_t st, std::string& err) {
    if (!plan.impl) plan.impl = new GemmPlanImpl();
    cudaError_t e0 = ensure_ws(ws);
    if (e0 != cudaSuccess) return e0;
    auto* w = static_cast<GemmWsImpl*>(ws.impl);
    const LevelPlan lp = gemm_topk_level_plan(a.nq, a.n_rows, a.k);
    const int kp = lp.kp, kq = lp.kq, MB = lp.query_blocks, cap = lp.cap;
    const int n_tiles = (int)((a.n_rows + GT_BN - 1) / GT_BN);
    const int num_pairs_max = a.num_sms / 2;
    const size_t esz = a.f16 ? 2 : 4;
    cudaError_t e;
    if ((e = grow_dev(w->buf, w->buf_cap, a.nq * (size_t)cap)) != cudaSuccess) return e;
    if ((e = grow_dev(w->cnt, w->cnt_cap, a.nq)) != cudaSuccess) return e;
    if ((e = grow_dev(w->thr, w->thr_cap, a.nq)) != cudaSuccess) return e;
    if ((e = grow_dev(w->overflow, w->ovf_cap, a.nq)) != cudaSuccess) return e;
    if ((e = grow_dev(w->flags, w->flags_cap, a.nq)) != cudaSuccess) return e;
    if (!a.prepped && (e = cudaMemsetAsync(w->overflow, 0, a.nq * sizeof(int), st)) != cudaSuccess) return e;
    // operand planes of the contraction: fp16 rows (fp16 shard, or the fp16 shadow of an fp32 shard) against
    // fp16-rounded queries -> kind::f16; plain fp32 rows against fp32 queries -> kind::tf32
    const bool g16 = a.f16 || a.shadow != nullptr;
    const void* grows = a.f16 ? a.rows : (a.shadow ? a.shadow : a.rows);
    const int gld = a.f16 ? a.ld : (a.shadow ? a.ld16 : a.ld);
    const size_t gesz = g16 ? 2 : 4;
    const void* qa = a.q;
    if (g16) {
        if ((e = grow_dev(w->q16, w->q16_cap, a.nq * (size_t)gld)) != cudaSuccess) return e;
        if (!a.prepped) {
            const size_t n = a.nq * (size_t)gld;
            f32_to_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.q, a.ld, w->q16, gld, a.nq);
            count_launch();
        }
        qa = w->q16;
    }
    CUtensorMap tmA, tmB;
    if (!make_tmap(&tmA, qa, g16, a.nq, (uint64_t)gld, GT_BM) || !make_tmap(&tmB, grows, g16, a.n_rows, (uint64_t)gld, GT_BN_HALF)) {
        err = "cuTensorMapEncodeTiled failed";
        return cudaErrorUnknown;
    }
    GemmParams gp{};
    gp.n_rows = a.n_rows; gp.nq = (uint32_t)a.nq;
    gp.num_kb = (int)((size_t)gld * gesz / GT_KB_BYTES);
    gp.kb_elems = (int)(GT_KB_BYTES / gesz);
    { const char* d = getenv("VDB_GEMM_DBG"); gp.dbg = d ? atoi(d) : 0; }
    gp.MB = MB; gp.n_tiles = n_tiles;
    gp.bits = lp.bits;
    gp.sqnorm = a.sqnorm; gp.thr = w->thr; gp.buf = w->buf; gp.cnt = w->cnt; gp.cap = cap; gp.tomb = a.tomb;
    const bool l2 = a.metric == 0;

    // |approx dot - exact dot| <= eps_rel * ||q|| * ||d||  (+ eps_abs_unit * sqrt(dim) * (||q|| + ||d||) for fp16
    // subnormal rounding).  fp16 shard: only the query is rounded (u = 2^-11 = 4.9e-4); shadow plane: query and
    // row are rounded (2u + u^2 = 9.8e-4); tf32 operands are truncated to 10 mantissa bits (2 * 2^-10 = 1.95e-3).
    // Accumulation term: the tensor core adds `dim` exact products in fp32; with truncating adds the partial sums
    // drift by at most dim * 2^-23 * sum|q_i d_i| <= dim * 2^-23 * ||q|| ||d|| (6.1e-5 at dim 512, 1.7e-4 at 1408).
    EpsModel em{};
    {
        const float eps_round = a.f16 ? 4.9e-4f : (a.shadow ? 9.8e-4f : 1.96e-3f);
        em.eps_rel = eps_round + (float)a.dim * 1.1920929e-7f + 1.0e-4f;
        em.eps_abs = g16 ? 3.0e-8f * sqrtf((float)a.dim) : 0.0f;     // 2^-25 per element, Cauchy-Schwarz over dim
        em.metric = a.metric;
    }
    SelectParams sp{};
    sp.k = a.k; sp.qn2 = a.qn2; sp.max_sqnorm_bits = a.d_max_sqnorm_bits; sp.em = em;
    { static const float m = [] { const char* e = getenv("VDB_TIGHT_MARGIN"); return e && *e ? (float)atof(e) : 2.5f; }(); sp.margin = m; }
    sp.buf = w->buf; sp.cnt = w->cnt; sp.cap = cap; sp.kp = kp; sp.thr = w->thr; sp.overflow = w->overflow;
    sp.tomb = a.tomb; sp.n_rows = a.n_rows;
    static bool sel_configured[MAX_DEVICES] = {};
    const int dev_slot = current_device_slot();
    if (!sel_configured[dev_slot]) {
        if ((e = cudaFuncSetAttribute(select_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(select_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024)) != cudaSuccess) return e;
        sel_configured[dev_slot] = true;
    }
    int sel_np = 2;
    while (sel_np < cap) sel_np <<= 1;
    sp.np2 = sel_np;
    const size_t sel_smem = ((size_t)sel_np + kp) * 8;

    // query-resident kernel variant for rows of at most 8 k-blocks (VDB_ARES: bit 0 = probe, bit 1 = levels; A/B runs)
    static const int ares_mode = env_int("VDB_ARES", 3);
    auto run_level = [&](int p0, int p1, bool probe) -> cudaError_t {
        gp.pos_begin = p0; gp.pos_end = p1; gp.probe = probe ? 1 : 0;
        gp.S = choose_slices(MB, p1 - p0, num_pairs_max);
        gp.n_items = MB * gp.S;
        const int grid = 2 * std::min(gp.n_items, num_pairs_max);   // whole CTA pairs
        cudaError_t le;
        if (a.prof_begin) a.prof_begin(a.prof_ctx, st);
        const bool ares = gp.num_kb <= GtSmem<true>::A_RES_KB && (probe ? (ares_mode & 1) : (ares_mode & 2)) != 0;
        if (g16) {
            if (ares) le = l2 ? launch_gemm<true, true, true>(tmA, tmB, gp, grid, st) : launch_gemm<true, false, true>(tmA, tmB, gp, grid, st);
            else      le = l2 ? launch_gemm<true, true, false>(tmA, tmB, gp, grid, st) : launch_gemm<true, false, false>(tmA, tmB, gp, grid, st);
        } else {
            if (ares) le = l2 ? launch_gemm<false, true, true>(tmA, tmB, gp, grid, st) : launch_gemm<false, false, true>(tmA, tmB, gp, grid, st);
            else      le = l2 ? launch_gemm<false, true, false>(tmA, tmB, gp, grid, st) : launch_gemm<false, false, false>(tmA, tmB, gp, grid, st);
        }
        if (a.prof_end) a.prof_end(a.prof_ctx, st);
        return le;
    };
    auto run_select = [&](int probe_cnt, int thr_rank) -> cudaError_t {
        sp.probe_cnt = probe_cnt; sp.thr_rank = thr_rank; sp.min_rank = std::min(kq, thr_rank);
        // one block per query; thousands of queries with a few hundred keys each: small blocks, so that more
        // of them are resident and the barrier chain of a block is short
        cudaError_t le;
        if (a.nq >= 4096 && cap <= 1024) le = launch_pdl(select_kernel<64>, dim3((unsigned)a.nq), dim3(64), sel_smem, st, sp);
        else le = launch_pdl(select_kernel<256>, dim3((unsigned)a.nq), dim3(256), sel_smem, st, sp);
        count_launch();
        return le != cudaSuccess ? le : cudaGetLastError();
    };
    {
        const int P = lp.probe_tiles;
        // the probe writes fixed positions; positions whose tile is past the end must read as sentinel: fill the
        // buffer only when there is such a position (never for shards of a few thousand rows or more)
        bool all_valid = true;
        for (int pp = 0; pp < P; ++pp) all_valid = all_valid && (int)bitrev((uint32_t)pp, gp.bits) < n_tiles;
        if (!all_valid && (e = cudaMemsetAsync(w->buf, 0xFF, a.nq * (size_t)cap * sizeof(uint64_t), st)) != cudaSuccess) return e;
        if ((e = run_level(0, P, true)) != cudaSuccess) return e;
        if ((e = run_select(P * 8, lp.probe_rank)) != cudaSuccess) return e;
        for (const LevelPlan::Level& lv : lp.levels) {
            if ((e = run_level(lv.p0, lv.p1, false)) != cudaSuccess) return e;
            if (lv.rank_after && (e = run_select(0, lv.rank_after)) != cudaSuccess) return e;
        }
    }

    if (!a.prepped && (e = cudaMemsetAsync(w->n_flagged, 0, sizeof(int), st)) != cudaSuccess) return e;
    RerankParams rp{};
    rp.approx = w->buf; rp.stride = (size_t)cap; rp.overflow = w->overflow; rp.tau = w->thr;
    rp.rows = a.rows; rp.row_bytes = (uint32_t)((size_t)a.ld * esz); rp.ld = a.ld;
    rp.labels = a.labels; rp.q = a.q; rp.qn2 = a.qn2; rp.max_sqnorm_bits = a.d_max_sqnorm_bits;
    rp.k = a.k; rp.metric = a.metric;
    rp.em = em;
    rp.f16_range = g16 ? 1 : 0;
    rp.out_ids = a.out_ids; rp.out_dist = a.out_dist; rp.out_counts = a.out_counts;
    rp.flags = w->flags; rp.n_flagged = w->n_flagged; rp.n_fallback = w->n_fallback;
    rp.cnt = w->cnt; rp.cap = cap; rp.tomb = a.tomb; rp.n_rows = a.n_rows;
    e = a.f16 ? launch_rerank_window<__half>(kp, rp, a.nq, st) : launch_rerank_window<float>(kp, rp, a.nq, st);
    if (e != cudaSuccess) return e;
    // K4x: queries whose certificate failed are re-searched exactly ON THE DEVICE -- no host round trip, the whole
    // search is one enqueue.  Without flagged queries (the normal case) its blocks read one counter and leave.
    const unsigned xgrid = (unsigned)std::min<size_t>(a.nq, (size_t)a.num_sms * 2);
    e = a.f16 ? launch_pdl(exact_fallback_kernel<__half>, dim3(xgrid), dim3(XF_THREADS), 0, st, rp, (int)a.nq)
              : launch_pdl(exact_fallback_kernel<float>, dim3(xgrid), dim3(XF_THREADS), 0, st, rp, (int)a.nq);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

int gemm_topk_candidate_kp(int k) { return k >= 1 && k <= 32 ? kp_for_k(k) : 0; }

cudaError_t gemm_topk_candidate_buffers(GemmWorkspace& ws, size_t nq, int kp, CandidateBuffers* out) {
    cudaError_t e = ensure_ws(ws);
    if (e != cudaSuccess) return e;
    auto* w = static_cast<GemmWsImpl*>(ws.impl);
    if ((e = grow_dev(w->buf, w->buf_cap, nq * (size_t)kp)) != cudaSuccess) return e;
    if ((e = grow_dev(w->cnt, w->cnt_cap, nq)) != cudaSuccess) return e;
    if ((e = grow_dev(w->thr, w->thr_cap, nq)) != cudaSuccess) return e;
    out->keys = w->buf; out->stride = (size_t)kp; out->cnt = w->cnt; out->tau = w->thr;
    return cudaSuccess;
}

cudaError_t gemm_topk_rerank_candidates(GemmWorkspace& ws, const GemmSearchArgs& a, int kp, cudaStream_t st) {
    cudaError_t e = ensure_ws(ws);
    if (e != cudaSuccess) return e;
    auto* w = static_cast<GemmWsImpl*>(ws.impl);
    if ((e = grow_dev(w->flags, w->flags_cap, a.nq)) != cudaSuccess) return e;
    // the candidates' distances were computed in fp32 from fp16-ROUNDED rows and the fp32 query: one operand rounded
    // (u = 2^-11), fp32 accumulation of `dim` terms.  L2 was taken in direct form sum (q - d16)^2, whose error also has
    // the term 2 d.(d16 - d) <= 2u ||d||^2 (eps_dd, with u^2 and slack)
    EpsModel em{};
    em.eps_rel = 4.9e-4f + (float)a.dim * 1.1920929e-7f + 1.0e-4f;
    em.eps_abs = 3.0e-8f * sqrtf((float)a.dim);
    em.metric = a.metric;
    em.eps_dd = a.metric == 0 ? 1.1e-3f : 0.0f;
    em.eps_sum = a.metric == 0 ? 2.0f * (float)a.dim * 1.1920929e-7f : 0.0f;       // sum <= 2 (||q||^2 + ||d||^2), dim * 2^-23 of it
    RerankParams rp{};
    rp.approx = w->buf; rp.stride = (size_t)kp; rp.overflow = w->overflow; rp.tau = w->thr;
    rp.rows = a.rows; rp.row_bytes = (uint32_t)((size_t)a.ld * 4); rp.ld = a.ld;
    rp.labels = a.labels; rp.q = a.q; rp.qn2 = a.qn2; rp.max_sqnorm_bits = a.d_max_sqnorm_bits;
    rp.k = a.k; rp.metric = a.metric;
    rp.em = em;
    rp.f16_range = 1;
    rp.out_ids = a.out_ids; rp.out_dist = a.out_dist; rp.out_counts = a.out_counts;
    rp.flags = w->flags; rp.n_flagged = w->n_flagged; rp.n_fallback = w->n_fallback;
    rp.cnt = w->cnt; rp.cap = kp; rp.tomb = a.tomb; rp.n_rows = a.n_rows;
    rp.approx_is_dist = 1;
    if ((e = launch_rerank_window<float>(kp, rp, a.nq, st)) != cudaSuccess) return e;
    const unsigned xgrid = (unsigned)std::min<size_t>(a.nq, (size_t)a.num_sms * 2);
    e = launch_pdl(exact_fallback_kernel<float>, dim3(xgrid), dim3(XF_THREADS), 0, st, rp, (int)a.nq);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// Queries this workspace has re-searched exactly so far (synchronises the device: statistics only).
long gemm_workspace_fallbacks(const GemmWorkspace& ws) {
    auto* w = static_cast<GemmWsImpl*>(ws.impl);
    if (!w || !w->n_fallback) return 0;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, w->n_fallback, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return 0; }
    return (long)v;
}

void gemm_plan_free(GemmPlan& plan) {
    delete static_cast<GemmPlanImpl*>(plan.impl);
    plan.impl = nullptr;
}
void gemm_workspace_free(GemmWorkspace& ws) {
    auto* w = static_cast<GemmWsImpl*>(ws.impl);
    if (!w) return;
    if (w->buf) cudaFree(w->buf);
    if (w->cnt) cudaFree(w->cnt);
    if (w->thr) cudaFree(w->thr);
    if (w->overflow) cudaFree(w->overflow);
    if (w->q16) cudaFree(w->q16);
    if (w->flags) cudaFree(w->flags);
    if (w->n_flagged) cudaFree(w->n_flagged);
    if (w->n_fallback) cudaFree(w->n_fallback);
    delete w;
    ws.impl = nullptr;
}
}  // namespace vdbk
