// gemm_topk.cu -- K2: batched search as a dense contraction on the 5th-gen tensor cores, fused
// with a streaming top-k' select, + K4: exact fp32 re-rank with a coverage certificate.
//
// Replaces hnswlib.Index.knn_query for nq > 8 (reference call site src/datanode/handler.py:364;
// hnswlib accepts [nq, dim], the reference only ever passes one row).
//
// K2 (gemm_topk_kernel), one persistent CTA per SM, 256 threads, warp-specialised:
//   warp 0   TMA producer: per k-block (128 bytes of every row) one 2-D tensor copy of the
//            128-query tile (A) and one of the 256-row shard tile (B), SWIZZLE_128B, into a
//            ring of shared-memory stages (cp.async.bulk.tensor, mbarrier complete_tx)
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma (kind::tf32 for fp32 rows,
//            kind::f16 for fp16 rows), M=128 x N=256, fp32 accumulators in TMEM (two 256-column
//            buffers, ping-pong), tcgen05.commit frees smem stages / publishes accumulators
//   warp 2   TMEM allocator
//   warps 4-7 epilogue: each thread owns one query (one TMEM lane): tcgen05.ld 32 columns at a
//            time, turn the dot product into an approximate distance, compare with the query's
//            running threshold (k'-th best so far), append survivors to a per-query pending
//            buffer in shared memory; full buffers are merged warp-cooperatively (bitonic
//            network in registers) into the query's sorted k'-list in global memory.
//   The [nq, n_rows] distance matrix never exists in memory.
//
// Work = items (slice of the shard x 128-query block); a CTA walks its items.  Every CTA that
// works on a query merges into that query's ONE shared sorted list of the k' best APPROXIMATE
// (tf32 / fp16-rounded-query) candidates (global memory, L2-resident, per-query spin lock), so
// the pruning threshold is that of the union of all rows seen so far.  K4 recomputes those k' candidates exactly (same summation order as the scan
// kernel, so batched and single-query searches return bit-identical distances) and proves the
// result: every row that is not a candidate has approximate distance >= tau (the k'-th
// approximate distance), hence exact distance >= tau - eps with eps a rigorous bound on the
// reduced-precision error; if the k-th exact distance is < tau - eps the exact top-k is inside
// the candidate set.  Queries that fail the certificate are re-searched with the exact scan.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

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
constexpr int GT_STAGES = 4;
constexpr int GT_THREADS = 256;
constexpr int GT_PEND_CAP = 64;                      // pending keys per query (flush at >= 32)
constexpr int GT_PEND_STRIDE = 65;                   // padded row: same-slot appends of a warp spread over banks
constexpr int GT_EPI_THREADS = 128;
constexpr int GT_TMEM_COLS = 512;

// dynamic shared memory map (base aligned to 1024)
constexpr int GT_OFF_PEND = GT_STAGES * GT_STAGE_BYTES;                       // 131072
constexpr int GT_OFF_NORM = GT_OFF_PEND + GT_EPI_THREADS * GT_PEND_STRIDE * 8; // + 66560
constexpr int GT_OFF_BAR = GT_OFF_NORM + 2 * GT_BN * 4;                       // + 2048
constexpr int GT_OFF_CTL = GT_OFF_BAR + 128;                                  // EpiCtl
constexpr int GT_SMEM_BYTES = GT_OFF_CTL + 1600 + 1024;                       // control block + align slack

struct GemmParams {
    uint32_t n_rows, nq;
    int num_kb;              // k-blocks per row (row bytes / 128)
    int kb_elems;            // elements per k-block (32 fp32 / 64 fp16)
    int MB, S, n_tiles, n_items;
    unsigned long long* stats;   // optional debug counters (null = off)
    int dbg;                 // experiments only: 1 = epilogue reads TMEM but selects nothing, 2 = releases at once
    const float* sqnorm;     // [n_rows] (L2 only)
    const uint32_t* tomb;    // bitmap or null
    uint64_t* cand;          // [nq][KP]  ONE sorted candidate list per query, shared by every CTA (L2-resident)
    uint64_t* gpend;         // [nq][KP]  keys accepted since the last merge (unsorted)
    int* gcnt;               // [nq]      how many
    int* locks;              // [nq] spin lock guarding a query's shared state
    uint32_t* thr_g;         // [nq] ordered bits of the list's k'-th approximate value (0xFFFFFFFF until full)
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
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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
// warp-level sorting networks on 64-bit keys
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t shfl_xor64(uint64_t v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a < b ? b : a; }

// one key per lane, full sort, ascending by lane
__device__ __forceinline__ uint64_t warp_sort32(uint64_t key, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const uint64_t other = shfl_xor64(key, stride);
            const bool asc = (lane & size) == 0 || size == 32;
            const bool lower = (lane & stride) == 0;
            key = (lower == asc) ? umin64(key, other) : umax64(key, other);
        }
    }
    return key;
}
// bitonic merge across lanes (strides 16..1): bitonic input -> ascending (or descending) by lane
template <bool ASC>
__device__ __forceinline__ uint64_t warp_bitonic_merge32(uint64_t key, int lane) {
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
        const uint64_t other = shfl_xor64(key, stride);
        const bool lower = (lane & stride) == 0;
        key = (lower == ASC) ? umin64(key, other) : umax64(key, other);
    }
    return key;
}

// Full bitonic sort of KP = R*32 keys held R per lane, element index e = r*32 + lane, ascending in e.
template <int R>
__device__ __forceinline__ void warp_sort_striped(uint64_t (&v)[R], int lane) {
    constexpr int KP = R * 32;
#pragma unroll
    for (int size = 2; size <= KP; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int rs = stride >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rs) == 0) {
                        const bool asc = (((r * 32) & size) == 0) || size == KP;
                        const uint64_t lo = umin64(v[r], v[r + rs]), hi = umax64(v[r], v[r + rs]);
                        v[r] = asc ? lo : hi;
                        v[r + rs] = asc ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint64_t other = shfl_xor64(v[r], stride);
                    const bool asc = ((((r * 32) | lane) & size) == 0) || size == KP;
                    const bool lower = (lane & stride) == 0;
                    v[r] = (lower == asc) ? umin64(v[r], other) : umax64(v[r], other);
                }
            }
        }
    }
}
// Bitonic merge (ascending) of a bitonic sequence of KP = R*32 keys, striped layout.
template <int R>
__device__ __forceinline__ void warp_merge_striped(uint64_t (&v)[R], int lane) {
#pragma unroll
    for (int rs = R / 2; rs > 0; rs >>= 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if ((r & rs) == 0) {
                const uint64_t lo = umin64(v[r], v[r + rs]), hi = umax64(v[r], v[r + rs]);
                v[r] = lo;
                v[r + rs] = hi;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = warp_bitonic_merge32<true>(v[r], lane);
}

// Per-query shared candidate state in global memory (L2-resident), guarded by locks[q]:
//   cand[q][KP]   sorted ascending: the KP best approximate keys merged so far
//   gpend[q][KP]  unsorted keys accepted since the last merge, gcnt[q] of them
//   thr_g[q]      ordered bits of cand[q][KP-1] (0xFFFFFFFF until the list is full): the pruning threshold
// Appending is cheap; the sort + merge runs once per KP accepted keys.
template <int KP>
__device__ __forceinline__ uint64_t heavy_merge(uint64_t* list, const uint64_t* gpend, int n_new, int lane) {
    constexpr int R = KP / 32;
    uint64_t pn[R], m[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int e = r * 32 + lane;
        pn[r] = e < n_new ? __ldcg(reinterpret_cast<const unsigned long long*>(gpend) + e) : KEY_SENTINEL;
        m[r] = __ldcg(reinterpret_cast<const unsigned long long*>(list) + e);
    }
    warp_sort_striped<R>(pn, lane);
    // half-cleaner against the reversed new keys: keeps the KP smallest of the union as a bitonic sequence
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint64_t rev = __shfl_sync(0xffffffffu, pn[R - 1 - r], 31 - lane);
        m[r] = umin64(m[r], rev);
    }
    warp_merge_striped<R>(m, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) __stcg(reinterpret_cast<unsigned long long*>(list) + r * 32 + lane, m[r]);
    return __shfl_sync(0xffffffffu, m[R - 1], 31);
}

struct QueryShared {
    uint64_t* cand; uint64_t* gpend; int* gcnt; int* locks; uint32_t* thr_g;
};

// Hand `key` (one per lane, `valid` lanes only) of query q to the shared state.  Whole converged warp.
// Returns the query's current threshold bits.
template <int KP>
__device__ __forceinline__ uint32_t share_keys(const QueryShared& g, uint32_t q, uint64_t key, bool valid, int lane) {
    const uint32_t mask = __ballot_sync(0xffffffffu, valid);
    const int ns = __popc(mask);
    if (ns == 0) return __ldcg(g.thr_g + q);
    int* lock = g.locks + q;
    if (lane == 0) {
        while (atomicCAS(lock, 0, 1) != 0) __nanosleep(40);
        __threadfence();
    }
    __syncwarp();
    uint64_t* gp = g.gpend + (size_t)q * KP;
    uint64_t* list = g.cand + (size_t)q * KP;
    int gc = __ldcg(g.gcnt + q);
    const int idx = __popc(mask & ((1u << lane) - 1u));
    const int room = KP - gc;
    if (valid && idx < room) __stcg(reinterpret_cast<unsigned long long*>(gp) + gc + idx, key);
    uint32_t thr_bits = 0xFFFFFFFFu;
    if (ns >= room) {   // buffer full: sort it and merge it into the sorted list
        __threadfence();
        __syncwarp();
        const uint64_t last = heavy_merge<KP>(list, gp, KP, lane);
        if (last != KEY_SENTINEL) thr_bits = (uint32_t)(last >> 32);
        __syncwarp();
        if (valid && idx >= room) __stcg(reinterpret_cast<unsigned long long*>(gp) + (idx - room), key);
        gc = ns - room;
        if (lane == 0 && thr_bits != 0xFFFFFFFFu) atomicMin(g.thr_g + q, thr_bits);
    } else {
        gc += ns;
        thr_bits = __ldcg(g.thr_g + q);
    }
    __syncwarp();
    if (lane == 0) {
        __stcg(g.gcnt + q, gc);
        __threadfence();
        atomicExch(lock, 0);
    }
    return thr_bits;
}

// shared-memory control block between the epilogue threads (producers of candidate keys) and the
// merger warps (consumers): one ring of GT_PEND_CAP keys per query
struct EpiCtl {
    uint32_t head_pub[GT_EPI_THREADS];   // keys appended so far (published by the epilogue thread)
    uint32_t tail[GT_EPI_THREADS];       // keys consumed so far (merger)
    float thr_s[GT_EPI_THREADS];         // latest threshold the merger has seen for the query
    uint32_t qbase[4];                   // first query of each epilogue warp's current item
    uint32_t done;                       // epilogue warps that have finished all items
};

// ------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------
template <bool F16, int KP, bool L2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GT_THREADS, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* pend_all = reinterpret_cast<uint64_t*>(smem + GT_OFF_PEND);
    float* norm_s = reinterpret_cast<float*>(smem + GT_OFF_NORM);   // [2][256]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GT_OFF_BAR);
    uint64_t* empty = full + GT_STAGES;
    uint64_t* tmem_full = empty + GT_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    EpiCtl* ctl = reinterpret_cast<EpiCtl*>(smem + GT_OFF_CTL);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();     // 0 = pair leader (issues the MMAs), 1 = peer
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GT_STAGES; ++s) {
            mbar_init(&full[s], 1);      // leader's arrive.expect_tx; bytes of BOTH CTAs' copies land on the leader's barrier
            mbar_init(&empty[s], 1);     // one multicast tcgen05.commit per use
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);   // 4 epilogue warps x 2 CTAs arrive on the LEADER's barrier
        }
        fence_barrier_init();
    }
    if (threadIdx.x < GT_EPI_THREADS) {
        ctl->head_pub[threadIdx.x] = 0;
        ctl->tail[threadIdx.x] = 0;
        ctl->thr_s[threadIdx.x] = __int_as_float(0x7f800000);
        if (threadIdx.x < 4) ctl->qbase[threadIdx.x] = 0;
        if (threadIdx.x == 0) ctl->done = 0;
    }
    if (warp == 2) tmem_alloc(tmem_ptr, GT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // barriers of both CTAs are initialised before any remote arrive / copy
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const QueryShared gsh{p.cand, p.gpend, p.gcnt, p.locks, p.thr_g};

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own query rows, own half of the shard tile) =================
        if (lane == 0) {
            // shard tiles are re-read from L2 by the pairs that hold the other query blocks of the same
            // slice, so they keep the default policy; the query block is reused by every tile: keep it
            uint64_t pol_stream;
            asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_stream));
            const uint64_t pol_keep = l2_policy_evict_last();
            uint32_t it = 0;
            for (int item = pair; item < p.n_items; item += num_pairs) {
                const int slice = item / p.MB, mb = item - slice * p.MB;
                const int t0 = (int)((long long)slice * p.n_tiles / p.S), t1 = (int)((long long)(slice + 1) * p.n_tiles / p.S);
                const int arow = mb * (2 * GT_BM) + (int)cta_rank * GT_BM;
                for (int tile = t0; tile < t1; ++tile) {
                    const int brow = tile * GT_BN + (int)cta_rank * GT_BN_HALF;
                    for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                        const int s = it % GT_STAGES;
                        const uint32_t ph = (it / GT_STAGES) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full[s], 2 * GT_STAGE_BYTES);
                        const uint32_t bar = map_to_cta(smem_u32(&full[s]), 0);
                        uint8_t* sa = smem + s * GT_STAGE_BYTES;
                        tma_load_2d(sa, &tmA, kb * p.kb_elems, arow, bar, pol_keep);
                        tma_load_2d(sa + GT_A_BYTES, &tmB, kb * p.kb_elems, brow, bar, pol_stream);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (pair leader only): M = 256 (2 x 128 queries) x N = 256 =================
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc(F16 ? 0u : 2u, 2 * GT_BM, GT_BN);
            uint32_t it = 0, tcount = 0;
            for (int item = pair; item < p.n_items; item += num_pairs) {
                const int slice = item / p.MB;
                const int t0 = (int)((long long)slice * p.n_tiles / p.S), t1 = (int)((long long)(slice + 1) * p.n_tiles / p.S);
                for (int tile = t0; tile < t1; ++tile, ++tcount) {
                    const uint32_t acc = tcount & 1;
                    mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * GT_BN;
                    for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                        const int s = it % GT_STAGES;
                        const uint32_t ph = (it / GT_STAGES) & 1;
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + s * GT_STAGE_BYTES);
                        const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + GT_A_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)   // 4 x 32 bytes of K per 128-byte swizzle atom
                            umma<F16>(d_tmem, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
                        umma_commit_pair(&empty[s]);           // both CTAs may refill this stage once the MMAs retire
                    }
                    umma_commit_pair(&tmem_full[acc]);         // accumulators ready in both CTAs' TMEM
                }
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ================= mergers: drain the per-query rings into the shared lists =================
        // Global-memory latency (lock, list) stays off the TMEM-drain path of the epilogue warps.
        const int base = (warp - 2) * 64;    // warp 2 serves epilogue warps 4,5; warp 3 serves 6,7
        volatile uint32_t* v_head = ctl->head_pub;
        volatile uint32_t* v_tail = ctl->tail;
        volatile float* v_thr = ctl->thr_s;
        long long mg_busy = 0, mg_hand = 0, mg_kept = 0, mg_t0 = clock64();
        for (;;) {
            const bool fin = *reinterpret_cast<volatile uint32_t*>(&ctl->done) == 4;
            __threadfence_block();
            const uint32_t av0 = v_head[base + lane] - v_tail[base + lane];
            const uint32_t av1 = v_head[base + 32 + lane] - v_tail[base + 32 + lane];
            const uint32_t big0 = __ballot_sync(0xffffffffu, av0 >= 16), big1 = __ballot_sync(0xffffffffu, av1 >= 16);
            const uint32_t any0 = __ballot_sync(0xffffffffu, av0 > 0), any1 = __ballot_sync(0xffffffffu, av1 > 0);
            if ((any0 | any1) == 0) {
                if (fin) {
                    if (p.stats && lane == 0) {
                        atomicAdd(p.stats + 5, (unsigned long long)mg_busy);   // merger warp-cycles in hand-offs
                        atomicAdd(p.stats + 6, (unsigned long long)mg_hand);   // hand-offs
                        atomicAdd(p.stats + 7, (unsigned long long)mg_kept);   // keys surviving the pre-filter
                        atomicAdd(p.stats + 8, (unsigned long long)(clock64() - mg_t0));
                    }
                    break;
                }
                __nanosleep(200);
                continue;
            }
            const bool take_big = (big0 | big1) != 0;
            for (int half = 0; half < 2; ++half) {
                uint32_t m = half == 0 ? (take_big ? big0 : any0) : (take_big ? big1 : any1);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const int qi = base + half * 32 + src;
                    const long long c0 = clock64();
                    const uint32_t t = v_tail[qi];
                    const uint32_t n = min(v_head[qi] - t, 32u);
                    __threadfence_block();
                    const uint64_t* ring = pend_all + (size_t)qi * GT_PEND_STRIDE;
                    uint64_t key = KEY_SENTINEL;
                    if ((uint32_t)lane < n) key = *reinterpret_cast<const volatile uint64_t*>(ring + ((t + lane) & (GT_PEND_CAP - 1)));
                    const uint32_t q = *reinterpret_cast<volatile uint32_t*>(&ctl->qbase[qi >> 5]) + (qi & 31);
                    // pre-filter: keys accepted under a threshold that has since tightened, padding rows of the
                    // last tile, tombstoned rows
                    const uint32_t tb = __ldcg(p.thr_g + q);
                    bool valid = key != KEY_SENTINEL;
                    if (valid) {
                        const uint32_t row = (uint32_t)key;
                        valid = (uint32_t)(key >> 32) < tb && row < p.n_rows &&
                                !(p.tomb && ((p.tomb[row >> 5] >> (row & 31)) & 1u));
                    }
                    mg_kept += __popc(__ballot_sync(0xffffffffu, valid));
                    const uint32_t nb = share_keys<KP>(gsh, q, key, valid, lane);
                    if (lane == 0) {
                        if (nb != 0xFFFFFFFFu) v_thr[qi] = fminf(v_thr[qi], ordered_to_float(nb));
                        __threadfence_block();
                        v_tail[qi] = t + n;
                    }
                    __syncwarp();
                    mg_busy += clock64() - c0;
                    ++mg_hand;
                }
            }
        }
    } else {
        // ================= epilogue: streaming top-k' (each CTA: its own 128 queries) =================
        const int et = threadIdx.x - 128;                  // 0..127 == TMEM lane == query within this CTA's block
        const int ew = et >> 5;                            // == warp % 4 : TMEM lane quadrant
        const uint32_t my_ring_s = smem_u32(pend_all + (size_t)et * GT_PEND_STRIDE);
        volatile uint32_t* my_tail = &ctl->tail[et];
        volatile float* my_thr_s = &ctl->thr_s[et];
        uint32_t head = 0, pub = 0;
        uint32_t tcount = 0;
        long long st_wait_full = 0, st_wait_ring = 0, st_wait_item = 0, st_busy = 0, st_t0 = clock64();
        for (int item = pair; item < p.n_items; item += num_pairs) {
            const int slice = item / p.MB, mb = item - slice * p.MB;
            const int t0 = (int)((long long)slice * p.n_tiles / p.S), t1 = (int)((long long)(slice + 1) * p.n_tiles / p.S);
            const uint32_t qbase = (uint32_t)mb * (2 * GT_BM) + cta_rank * GT_BM;
            const uint32_t q = qbase + et;
            const bool q_ok = q < p.nq;
            // the ring still holds keys of the previous item's query until the merger has drained it
            { const long long c0 = clock64(); while (*my_tail != head) __nanosleep(64); st_wait_item += clock64() - c0; }
            __syncwarp();
            *my_thr_s = q_ok ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
            if (lane == 0) ctl->qbase[ew] = qbase + ew * 32;
            __threadfence_block();
            float thr = q_ok ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);   // +inf / -inf (never passes)
            float n_next0 = 0.0f, n_next1 = 0.0f;

            for (int tile = t0; tile < t1; ++tile, ++tcount) {
                const uint32_t acc = tcount & 1;
                const uint32_t row0 = (uint32_t)tile * GT_BN;
                // the query's shared list is tightened by every CTA working on it: a row that is not below its
                // current k'-th approximate value cannot be in the global top-k'
                uint32_t tg = 0xFFFFFFFFu;
                if (q_ok) tg = __ldcg(p.thr_g + q);
                if constexpr (L2) {
                    // ||d||^2 of this tile's rows: prefetched into registers one tile ahead (first tile of an
                    // item: loaded here), published to the buffer the previous-but-one tile used
                    float* ns = norm_s + acc * GT_BN;
                    if (tile == t0) {
                        const uint32_t r_a = row0 + et, r_b = row0 + 128 + et;
                        n_next0 = r_a < p.n_rows ? p.sqnorm[r_a] : 0.0f;
                        n_next1 = r_b < p.n_rows ? p.sqnorm[r_b] : 0.0f;
                    }
                    ns[et] = n_next0;
                    ns[et + 128] = n_next1;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (tile + 1 < t1) {
                        const uint32_t r_a = row0 + GT_BN + et, r_b = row0 + GT_BN + 128 + et;
                        n_next0 = r_a < p.n_rows ? p.sqnorm[r_a] : 0.0f;
                        n_next1 = r_b < p.n_rows ? p.sqnorm[r_b] : 0.0f;
                    }
                }
                if (tg != 0xFFFFFFFFu) thr = fminf(thr, ordered_to_float(tg));
                { const long long c0 = clock64(); mbar_wait(&tmem_full[acc], (tcount >> 1) & 1); st_wait_full += clock64() - c0; }
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * GT_BN;
                const float* ns = norm_s + acc * GT_BN;
#pragma unroll 1
                for (int c = 0; c < GT_BN / 32; ++c) {
                    if (p.dbg == 2) break;
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 32, v);
                    thr = fminf(thr, *my_thr_s);
                    // a chunk may append up to 32 keys: wait (rare) until the merger has left that much room
                    if (head - *my_tail > (uint32_t)(GT_PEND_CAP - 32)) { const long long c0 = clock64(); while (head - *my_tail > (uint32_t)(GT_PEND_CAP - 32)) __nanosleep(32); st_wait_ring += clock64() - c0; }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float dot = __uint_as_float(v[j]);
                        float a;
                        if constexpr (L2) a = fmaf(-2.0f, dot, ns[c * 32 + j]);
                        else a = -dot;
                        if (a < thr && p.dbg == 0) {
                            st_shared_u64(my_ring_s + (head & (GT_PEND_CAP - 1)) * 8, make_key(a, row0 + c * 32 + j));
                            ++head;
                        }
                    }
                    if (head != pub) {
                        __threadfence_block();
                        ctl->head_pub[et] = head;
                        pub = head;
                    }
                    __syncwarp();
                }
                // all TMEM reads of this accumulator are done: one arrive per warp on the LEADER's barrier
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));
            }
        }
        __syncwarp();
        if (p.stats) {
            st_busy = clock64() - st_t0;
            atomicAdd(p.stats + 0, (unsigned long long)head);            // keys appended
            if (lane == 0) {
                atomicAdd(p.stats + 1, (unsigned long long)st_wait_full);   // epilogue warp-cycles waiting for MMA
                atomicAdd(p.stats + 4, (unsigned long long)st_busy);        // epilogue warp-cycles total
            }
            atomicAdd(p.stats + 2, (unsigned long long)st_wait_ring);    // thread-cycles waiting for ring space
            atomicAdd(p.stats + 3, (unsigned long long)st_wait_item);    // thread-cycles waiting for drain at item switch
        }
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

// Merges what is still waiting in gpend[q] into cand[q]: one warp per query.
template <int KP>
__global__ void finalize_lists_kernel(uint64_t* cand, const uint64_t* gpend, const int* gcnt, size_t nq) {
    const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nq) return;
    const int n = gcnt[w];
    if (n > 0) heavy_merge<KP>(cand + w * KP, gpend + w * KP, n, lane);
}

// ------------------------------------------------------------------------------------------
// K4: exact re-rank of the k' approximate candidates + coverage certificate
// ------------------------------------------------------------------------------------------
struct RerankParams {
    const uint64_t* approx;   // [nq][KP] ascending approximate keys (distance-like value, row)
    const void* rows; uint32_t row_bytes; uint32_t ld;
    const uint32_t* labels;
    const float* q;           // prepared queries [nq][ld] fp32
    const float* qn2;         // [nq]
    const unsigned int* max_sqnorm_bits;
    int k, metric;            // metric 0 = L2 (approx value = ||d||^2 - 2 q.d), 1 = ip/cos (approx value = -q.d)
    float eps_rel;            // bound on |approx dot - exact dot| / (||q|| ||d||)
    int64_t* out_ids; float* out_dist; int* out_counts;
    int* flags;               // [nq] 1 = certificate failed
    int* n_flagged;
};

template <typename T, int KP>
__global__ void __launch_bounds__(256) rerank_kernel(const RerankParams p) {
    __shared__ uint64_t ek[KP];
    const int q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t* ap = p.approx + (size_t)q * KP;
    const float* qv = p.q + (size_t)q * p.ld;
    const int nld16 = p.row_bytes / 512;
    constexpr int PER16 = 16 / sizeof(T);
    for (int c = warp; c < KP; c += 8) {
        const uint64_t key = ap[c];
        uint64_t out = KEY_SENTINEL;
        if (key != KEY_SENTINEL) {
            const uint32_t row = (uint32_t)key;
            const uint8_t* rp = reinterpret_cast<const uint8_t*>(p.rows) + (size_t)row * p.row_bytes;
            float acc = 0.0f;
            for (int ch = 0; ch < nld16; ++ch) {
                float dv[PER16], qq[PER16];
                if constexpr (sizeof(T) == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(rp + (size_t)(ch * 32 + lane) * 16);
                    dv[0] = t.x; dv[1] = t.y; dv[2] = t.z; dv[3] = t.w;
                } else {
                    const uint4 t = *reinterpret_cast<const uint4*>(rp + (size_t)(ch * 32 + lane) * 16);
                    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = __half22float2(h[i]);
                        dv[2 * i] = f.x; dv[2 * i + 1] = f.y;
                    }
                }
                const float* qp = qv + (size_t)(ch * 32 + lane) * PER16;
#pragma unroll
                for (int e = 0; e < PER16; e += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(qp + e);
                    qq[e] = t.x; qq[e + 1] = t.y; qq[e + 2] = t.z; qq[e + 3] = t.w;
                }
                if (p.metric == 0) {
#pragma unroll
                    for (int e = 0; e < PER16; ++e) {
                        const float t = dv[e] - qq[e];
                        acc = fmaf(t, t, acc);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < PER16; ++e) acc = fmaf(dv[e], qq[e], acc);
                }
            }
            acc = warp_sum_butterfly(acc);
            const float dist = p.metric == 0 ? acc : 1.0f - acc;
            out = make_key(dist, p.labels[row]);
        }
        if (lane == 0) ek[c] = out;
    }
    __syncthreads();
    // bitonic sort of KP keys by 256 threads
    for (int size = 2; size <= KP; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < KP / 2; t += 256) {
                const int i = 2 * t - (t & (stride - 1)), j = i + stride;
                const bool asc = (i & size) == 0;
                const uint64_t a = ek[i], b = ek[j];
                if ((a > b) == asc) { ek[i] = b; ek[j] = a; }
            }
            __syncthreads();
        }
    }
    const int k = p.k;
    for (int i = threadIdx.x; i < k; i += 256) {
        const uint64_t key = ek[i];
        const bool real = key != KEY_SENTINEL;
        p.out_ids[(size_t)q * k + i] = real ? (int64_t)key_label(key) : -1;
        p.out_dist[(size_t)q * k + i] = real ? key_dist(key) : __int_as_float(0x7f800000);
    }
    if (threadIdx.x == 0) {
        int cnt = 0;
        for (int i = 0; i < k; ++i) cnt += ek[i] != KEY_SENTINEL;
        if (p.out_counts) p.out_counts[q] = cnt;
        bool ok = true;
        const uint64_t last = ap[KP - 1];
        if (last != KEY_SENTINEL) {   // candidate list full: rows outside it exist, prove they cannot matter
            const float a_tau = key_dist(last);
            const float qn2 = p.qn2[q];
            const float dmax2 = __uint_as_float(*p.max_sqnorm_bits);
            const float eb = p.eps_rel * sqrtf(qn2) * sqrtf(dmax2);
            float tau, eps;
            if (p.metric == 0) { tau = a_tau + qn2; eps = 2.0f * eb + 4e-7f * (qn2 + dmax2 + fabsf(tau)); }
            else               { tau = 1.0f + a_tau; eps = eb + 4e-7f * (1.0f + fabsf(tau)); }
            const uint64_t kth = ek[k - 1];
            ok = kth != KEY_SENTINEL && key_dist(kth) < tau - eps;
        }
        p.flags[q] = ok ? 0 : 1;
        if (!ok) atomicAdd(p.n_flagged, 1);
    }
}

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2half_rn(in[i]);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct GemmWsImpl {
    uint64_t* cand = nullptr; size_t cand_cap = 0;
    int* locks = nullptr; size_t locks_cap = 0;
    uint64_t* gpend = nullptr; size_t gpend_cap = 0;
    int* gcnt = nullptr; size_t gcnt_cap = 0;
    __half* q16 = nullptr; size_t q16_cap = 0;
    int* flags = nullptr; size_t flags_cap = 0;
    uint32_t* thr_g = nullptr; size_t thr_cap = 0;
    int* n_flagged = nullptr;
    int* h_n_flagged = nullptr;   // pinned
    unsigned long long* stats = nullptr;
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

template <bool F16, int KP, bool L2>
static cudaError_t launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& gp, int grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_topk_kernel<F16, KP, L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    gemm_topk_kernel<F16, KP, L2><<<grid, GT_THREADS, GT_SMEM_BYTES, st>>>(tmA, tmB, gp);
    count_launch();
    return cudaGetLastError();
}

template <bool F16, bool L2>
static cudaError_t launch_gemm_kp(int kp, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& gp, int grid, cudaStream_t st) {
    switch (kp) {
        case 32: return launch_gemm<F16, 32, L2>(tmA, tmB, gp, grid, st);
        case 64: return launch_gemm<F16, 64, L2>(tmA, tmB, gp, grid, st);
        case 128: return launch_gemm<F16, 128, L2>(tmA, tmB, gp, grid, st);
        case 256: return launch_gemm<F16, 256, L2>(tmA, tmB, gp, grid, st);
    }
    return cudaErrorInvalidValue;
}

template <typename T>
static cudaError_t launch_rerank(int kp, const RerankParams& rp, size_t nq, cudaStream_t st) {
    switch (kp) {
        case 32: rerank_kernel<T, 32><<<(unsigned)nq, 256, 0, st>>>(rp); break;
        case 64: rerank_kernel<T, 64><<<(unsigned)nq, 256, 0, st>>>(rp); break;
        case 128: rerank_kernel<T, 128><<<(unsigned)nq, 256, 0, st>>>(rp); break;
        case 256: rerank_kernel<T, 256><<<(unsigned)nq, 256, 0, st>>>(rp); break;
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// picks the number of shard slices: fill the SMs in whole waves, prefer few slices
static int choose_slices(int MB, int n_tiles, int num_sms) {
    int best = 1;
    double best_cost = 1e30;
    const int smax = std::min(n_tiles, 96);
    for (int S = 1; S <= smax; ++S) {
        const long items = (long)MB * S;
        const long waves = (items + num_sms - 1) / num_sms;
        const double tiles_per_item = (double)n_tiles / S;
        // time ~ waves * (tiles per item + warm-up of the per-item lists, ~4 tiles worth)
        const double cost = waves * (tiles_per_item + 4.0);
        if (cost < best_cost * 0.999) { best_cost = cost; best = S; }
    }
    return best;
}

cudaError_t gemm_topk_search(GemmPlan& plan, GemmWorkspace& ws, const GemmSearchArgs& a, cudaStream_t st, std::string& err) {
    if (!plan.impl) plan.impl = new GemmPlanImpl();
    if (!ws.impl) {
        auto* w = new GemmWsImpl();
        cudaError_t e = cudaMalloc((void**)&w->n_flagged, sizeof(int));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&w->h_n_flagged, sizeof(int));
        if (e != cudaSuccess) { delete w; return e; }
        ws.impl = w;
    }
    auto* w = static_cast<GemmWsImpl*>(ws.impl);
    const int kp = kp_for_k(a.k);
    const int MB = (int)((a.nq + 2 * GT_BM - 1) / (2 * GT_BM));   // 256-query blocks, one per CTA pair
    const int n_tiles = (int)((a.n_rows + GT_BN - 1) / GT_BN);
    const int num_pairs_max = a.num_sms / 2;
    const int S = choose_slices(MB, n_tiles, num_pairs_max);
    const size_t esz = a.f16 ? 2 : 4;
    cudaError_t e;
    if ((e = grow_dev(w->cand, w->cand_cap, a.nq * (size_t)kp)) != cudaSuccess) return e;
    if ((e = grow_dev(w->locks, w->locks_cap, a.nq)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w->cand, 0xFF, a.nq * (size_t)kp * sizeof(uint64_t), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w->locks, 0, a.nq * sizeof(int), st)) != cudaSuccess) return e;
    if ((e = grow_dev(w->gpend, w->gpend_cap, a.nq * (size_t)kp)) != cudaSuccess) return e;
    if ((e = grow_dev(w->gcnt, w->gcnt_cap, a.nq)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w->gcnt, 0, a.nq * sizeof(int), st)) != cudaSuccess) return e;
    if ((e = grow_dev(w->flags, w->flags_cap, a.nq)) != cudaSuccess) return e;
    if ((e = grow_dev(w->thr_g, w->thr_cap, a.nq)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w->thr_g, 0xFF, a.nq * sizeof(uint32_t), st)) != cudaSuccess) return e;
    const void* qa = a.q;
    if (a.f16) {
        if ((e = grow_dev(w->q16, w->q16_cap, a.nq * (size_t)a.ld)) != cudaSuccess) return e;
        const size_t n = a.nq * (size_t)a.ld;
        f32_to_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.q, w->q16, n);
        count_launch();
        qa = w->q16;
    }
    CUtensorMap tmA, tmB;
    if (!make_tmap(&tmA, qa, a.f16, a.nq, (uint64_t)a.ld, GT_BM) || !make_tmap(&tmB, a.rows, a.f16, a.n_rows, (uint64_t)a.ld, GT_BN_HALF)) {
        err = "cuTensorMapEncodeTiled failed";
        return cudaErrorUnknown;
    }
    GemmParams gp{};
    gp.n_rows = a.n_rows; gp.nq = (uint32_t)a.nq;
    gp.num_kb = (int)((size_t)a.ld * esz / GT_KB_BYTES);
    gp.kb_elems = (int)(GT_KB_BYTES / esz);
    { const char* d = getenv("VDB_GEMM_DBG"); gp.dbg = d ? atoi(d) : 0; }
    if (getenv("VDB_GEMM_STATS")) {
        if (!w->stats) cudaMalloc((void**)&w->stats, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(w->stats, 0, 16 * sizeof(unsigned long long), st);
        gp.stats = w->stats;
    }
    gp.MB = MB; gp.S = S; gp.n_tiles = n_tiles; gp.n_items = MB * S;
    gp.sqnorm = a.sqnorm; gp.tomb = a.tomb; gp.cand = w->cand; gp.gpend = w->gpend; gp.gcnt = w->gcnt; gp.locks = w->locks; gp.thr_g = w->thr_g;
    const int grid = 2 * std::min(gp.n_items, num_pairs_max);   // whole CTA pairs
    const bool l2 = a.metric == 0;
    if (a.f16) e = l2 ? launch_gemm_kp<true, true>(kp, tmA, tmB, gp, grid, st) : launch_gemm_kp<true, false>(kp, tmA, tmB, gp, grid, st);
    else       e = l2 ? launch_gemm_kp<false, true>(kp, tmA, tmB, gp, grid, st) : launch_gemm_kp<false, false>(kp, tmA, tmB, gp, grid, st);
    if (e != cudaSuccess) return e;

    {   // keys still waiting in the per-query pending buffers
        const unsigned blocks = (unsigned)((a.nq * 32 + 255) / 256);
        switch (kp) {
            case 32: finalize_lists_kernel<32><<<blocks, 256, 0, st>>>(w->cand, w->gpend, w->gcnt, a.nq); break;
            case 64: finalize_lists_kernel<64><<<blocks, 256, 0, st>>>(w->cand, w->gpend, w->gcnt, a.nq); break;
            case 128: finalize_lists_kernel<128><<<blocks, 256, 0, st>>>(w->cand, w->gpend, w->gcnt, a.nq); break;
            default: finalize_lists_kernel<256><<<blocks, 256, 0, st>>>(w->cand, w->gpend, w->gcnt, a.nq); break;
        }
        count_launch();
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (gp.stats) {
        unsigned long long hs[16];
        cudaMemcpyAsync(hs, w->stats, sizeof(hs), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "[gemm stats] S=%d items=%d grid=%d | appended=%llu kept=%llu handoffs=%llu | epi wait_full=%.1f%% ring=%.2f%% item=%.2f%% | merger busy=%.1f%%\n",
                S, gp.n_items, grid, hs[0], hs[7], hs[6], 100.0 * hs[1] / (double)(hs[4] + 1),
                100.0 * hs[2] / 32.0 / (double)(hs[4] + 1), 100.0 * hs[3] / 32.0 / (double)(hs[4] + 1), 100.0 * hs[5] / (double)(hs[8] + 1));
    }
    if ((e = cudaMemsetAsync(w->n_flagged, 0, sizeof(int), st)) != cudaSuccess) return e;
    RerankParams rp{};
    rp.approx = w->cand; rp.rows = a.rows; rp.row_bytes = (uint32_t)((size_t)a.ld * esz); rp.ld = a.ld;
    rp.labels = a.labels; rp.q = a.q; rp.qn2 = a.qn2; rp.max_sqnorm_bits = a.d_max_sqnorm_bits;
    rp.k = a.k; rp.metric = a.metric;
    rp.eps_rel = a.f16 ? 6.5e-4f : 2.5e-3f;
    rp.out_ids = a.out_ids; rp.out_dist = a.out_dist; rp.out_counts = a.out_counts;
    rp.flags = w->flags; rp.n_flagged = w->n_flagged;
    e = a.f16 ? launch_rerank<__half>(kp, rp, a.nq, st) : launch_rerank<float>(kp, rp, a.nq, st);
    if (e != cudaSuccess) return e;
    return cudaSuccess;
}

// After gemm_topk_search: how many queries failed the certificate (synchronises `st`), and which.
cudaError_t gemm_topk_flagged(GemmWorkspace& ws, size_t nq, std::vector<int>& flagged, cudaStream_t st) {
    auto* w = static_cast<GemmWsImpl*>(ws.impl);
    flagged.clear();
    cudaError_t e = cudaMemcpyAsync(w->h_n_flagged, w->n_flagged, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    if (*w->h_n_flagged == 0) return cudaSuccess;
    std::vector<int> f(nq);
    e = cudaMemcpyAsync(f.data(), w->flags, nq * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    for (size_t i = 0; i < nq; ++i)
        if (f[i]) flagged.push_back((int)i);
    return cudaSuccess;
}

void gemm_plan_note_fallbacks(GemmPlan& plan, long n) {
    if (!plan.impl) plan.impl = new GemmPlanImpl();
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    static_cast<GemmPlanImpl*>(plan.impl)->fallbacks += n;
}
void gemm_plan_free(GemmPlan& plan) {
    delete static_cast<GemmPlanImpl*>(plan.impl);
    plan.impl = nullptr;
}
void gemm_workspace_free(GemmWorkspace& ws) {
    auto* w = static_cast<GemmWsImpl*>(ws.impl);
    if (!w) return;
    if (w->cand) cudaFree(w->cand);
    if (w->locks) cudaFree(w->locks);
    if (w->gpend) cudaFree(w->gpend);
    if (w->gcnt) cudaFree(w->gcnt);
    if (w->q16) cudaFree(w->q16);
    if (w->flags) cudaFree(w->flags);
    if (w->thr_g) cudaFree(w->thr_g);
    if (w->n_flagged) cudaFree(w->n_flagged);
    if (w->h_n_flagged) cudaFreeHost(w->h_n_flagged);
    delete w;
    ws.impl = nullptr;
}
long gemm_plan_fallbacks(const GemmPlan& plan) {
    return plan.impl ? static_cast<GemmPlanImpl*>(plan.impl)->fallbacks : 0;
}

}  // namespace vdbk
