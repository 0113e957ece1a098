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
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "gemm_topk.h"
#include "kernels.h"

namespace vdbk {

constexpr int GT_BM = 128;
constexpr int GT_BN = 256;
constexpr int GT_KB_BYTES = 128;                     // one 128B swizzle atom per row per k-block
constexpr int GT_A_BYTES = GT_BM * GT_KB_BYTES;      // 16 KB
constexpr int GT_B_BYTES = GT_BN * GT_KB_BYTES;      // 32 KB
constexpr int GT_STAGE_BYTES = GT_A_BYTES + GT_B_BYTES;
constexpr int GT_STAGES = 3;
constexpr int GT_THREADS = 256;
constexpr int GT_PEND_CAP = 64;                      // pending keys per query (flush at >= 32)
constexpr int GT_PEND_STRIDE = 65;                   // padded row: same-slot appends of a warp spread over banks
constexpr int GT_EPI_THREADS = 128;
constexpr int GT_TMEM_COLS = 512;

// dynamic shared memory map (base aligned to 1024)
constexpr int GT_OFF_PEND = GT_STAGES * GT_STAGE_BYTES;                       // 147456
constexpr int GT_OFF_NORM = GT_OFF_PEND + GT_EPI_THREADS * GT_PEND_STRIDE * 8; // + 66560
constexpr int GT_OFF_BAR = GT_OFF_NORM + 2 * GT_BN * 4;                       // + 2048
constexpr int GT_SMEM_BYTES = GT_OFF_BAR + 128 + 1024;                        // barriers + align slack

struct GemmParams {
    uint32_t n_rows, nq;
    int num_kb;              // k-blocks per row (row bytes / 128)
    int kb_elems;            // elements per k-block (32 fp32 / 64 fp16)
    int MB, S, n_tiles, n_items;
    int dbg;                 // experiments only: 1 = epilogue reads TMEM but selects nothing, 2 = releases at once
    const float* sqnorm;     // [n_rows] (L2 only)
    const uint32_t* tomb;    // bitmap or null
    uint64_t* cand;          // [nq][KP]  ONE sorted candidate list per query, shared by every CTA (L2-resident)
    int* locks;              // [nq] spin lock guarding a query's list during a merge
    uint32_t* thr_g;         // [nq] ordered bits of the list's k'-th approximate value (0xFFFFFFFF until full)
};

// ------------------------------------------------------------------------------------------
// PTX wrappers (tcgen05 / TMA); cta_group::1
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(x), "r"(y), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (F16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
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

// Merge `n_pend` pending keys of one query into the query's shared sorted k'-list (global memory, L2).
// Executed by a converged warp under the query's spin lock: every CTA that scans a slice of the shard
// for this query merges into the same list, so the running threshold is the k'-th best of the UNION of
// all rows seen so far by anyone.  m[r] holds list element r*32 + lane.  Returns the list's new last key.
template <int KP>
__device__ __forceinline__ uint64_t flush_query(uint64_t* list, int* lock, uint32_t* thr_slot, const uint64_t* pend,
                                                int n_pend, const uint32_t* __restrict__ tomb, uint32_t n_rows, int lane) {
    constexpr int R = KP / 32;
    uint64_t m[R];
    if (lane == 0) {
        while (atomicCAS(lock, 0, 1) != 0) __nanosleep(64);
    }
    __syncwarp();
    __threadfence();
#pragma unroll
    for (int r = 0; r < R; ++r) m[r] = __ldcg(reinterpret_cast<const unsigned long long*>(list) + r * 32 + lane);
    for (int base = 0; base < n_pend; base += 32) {
        uint64_t p = (base + lane < n_pend) ? pend[base + lane] : KEY_SENTINEL;
        if (p != KEY_SENTINEL) {   // drop padding rows of the last tile and tombstoned rows here (rare path)
            const uint32_t row = (uint32_t)p;
            if (row >= n_rows || (tomb && ((tomb[row >> 5] >> (row & 31)) & 1u))) p = KEY_SENTINEL;
        }
        p = warp_sort32(p, lane);
        // half-cleaner against the list's top 32: keeps the 32 smallest of (top 32 U pending), bitonic
        const uint64_t prev = __shfl_sync(0xffffffffu, p, 31 - lane);
        m[R - 1] = umin64(m[R - 1], prev);
        if constexpr (R == 1) {
            m[0] = warp_bitonic_merge32<true>(m[0], lane);
        } else {
            // sort the top block DESCENDING so that m[0..R-2] (ascending) ++ m[R-1] is bitonic
            m[R - 1] = warp_bitonic_merge32<false>(m[R - 1], lane);
            // bitonic merge of KP elements: register strides first, then lane strides
#pragma unroll
            for (int rs = R / 2; rs > 0; rs >>= 1) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rs) == 0) {
                        const uint64_t lo = umin64(m[r], m[r + rs]), hi = umax64(m[r], m[r + rs]);
                        m[r] = lo;
                        m[r + rs] = hi;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) m[r] = warp_bitonic_merge32<true>(m[r], lane);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) __stcg(reinterpret_cast<unsigned long long*>(list) + r * 32 + lane, m[r]);
    const uint64_t last = __shfl_sync(0xffffffffu, m[R - 1], 31);
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        if (last != KEY_SENTINEL) atomicMin(thr_slot, (uint32_t)(last >> 32));
        __threadfence();
        atomicExch(lock, 0);
    }
    return last;
}

__device__ __forceinline__ float thr_from_key(uint64_t last) {
    return last == KEY_SENTINEL ? __int_as_float(0x7f800000) : key_dist(last);
}

// ------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------
template <bool F16, int KP, bool L2>
__global__ void __launch_bounds__(GT_THREADS, 1)
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GT_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], GT_EPI_THREADS);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, GT_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            // shard tiles are re-read from L2 by the CTAs that hold the other query blocks of the same
            // slice, so they keep the default policy; the query block is reused by every tile: keep it
            uint64_t pol_stream;
            asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_stream));
            const uint64_t pol_keep = l2_policy_evict_last();
            uint32_t it = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int slice = item / p.MB, mb = item - slice * p.MB;
                const int t0 = (int)((long long)slice * p.n_tiles / p.S), t1 = (int)((long long)(slice + 1) * p.n_tiles / p.S);
                for (int tile = t0; tile < t1; ++tile) {
                    for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                        const int s = it % GT_STAGES;
                        const uint32_t ph = (it / GT_STAGES) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full[s], GT_STAGE_BYTES);
                        uint8_t* sa = smem + s * GT_STAGE_BYTES;
                        tma_load_2d(sa, &tmA, kb * p.kb_elems, mb * GT_BM, &full[s], pol_keep);
                        tma_load_2d(sa + GT_A_BYTES, &tmB, kb * p.kb_elems, tile * GT_BN, &full[s], pol_stream);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(F16 ? 0u : 2u, GT_BM, GT_BN);
            uint32_t it = 0, tcount = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
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
                        umma_commit(&empty[s]);            // smem stage reusable once these MMAs retire
                    }
                    umma_commit(&tmem_full[acc]);          // accumulator ready for the epilogue
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: streaming top-k' =================
        const int et = threadIdx.x - 128;                  // 0..127 == TMEM lane == query within block
        const int ew = et >> 5;                            // == warp % 4 : TMEM lane quadrant
        const uint32_t my_pend_s = smem_u32(pend_all + (size_t)et * GT_PEND_STRIDE);
        uint32_t tcount = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int slice = item / p.MB, mb = item - slice * p.MB;
            const int t0 = (int)((long long)slice * p.n_tiles / p.S), t1 = (int)((long long)(slice + 1) * p.n_tiles / p.S);
            const uint32_t q = (uint32_t)mb * GT_BM + et;
            const bool q_ok = q < p.nq;
            float thr = q_ok ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);   // +inf / -inf (never passes)
            int cnt = 0;

            for (int tile = t0; tile < t1; ++tile, ++tcount) {
                const uint32_t acc = tcount & 1;
                const uint32_t row0 = (uint32_t)tile * GT_BN;
                // the query's shared list is tightened by every CTA working on it: a row that is not below its
                // current k'-th approximate value cannot be in the global top-k'
                uint32_t tg = 0xFFFFFFFFu;
                if (q_ok) tg = __ldcg(p.thr_g + q);
                if constexpr (L2) {
                    // stage ||d||^2 of this tile's rows (buffer `acc` was last read two tiles ago, and every
                    // epilogue thread has passed the previous tile's barrier since)
                    float* ns = norm_s + acc * GT_BN;
                    const uint32_t r_a = row0 + et, r_b = row0 + 128 + et;
                    ns[et] = r_a < p.n_rows ? p.sqnorm[r_a] : 0.0f;
                    ns[et + 128] = r_b < p.n_rows ? p.sqnorm[r_b] : 0.0f;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                if (tg != 0xFFFFFFFFu) thr = fminf(thr, ordered_to_float(tg));
                mbar_wait(&tmem_full[acc], (tcount >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * GT_BN;
                const float* ns = norm_s + acc * GT_BN;
#pragma unroll 1
                for (int c = 0; c < GT_BN / 32; ++c) {
                    if (p.dbg == 2) break;
                    uint32_t v[32];
                    tmem_ld32(taddr + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float dot = __uint_as_float(v[j]);
                        float a;
                        if constexpr (L2) a = fmaf(-2.0f, dot, ns[c * 32 + j]);
                        else a = -dot;
                        if (a < thr && p.dbg == 0) {
                            st_shared_u64(my_pend_s + cnt * 8, make_key(a, row0 + c * 32 + j));
                            ++cnt;
                        }
                    }
                    __syncwarp();
                    uint32_t fl = __ballot_sync(0xffffffffu, cnt >= 32);
                    while (fl) {
                        const int src = __ffs(fl) - 1;
                        fl &= fl - 1;
                        const int n_p = __shfl_sync(0xffffffffu, cnt, src);
                        const uint32_t qs = (uint32_t)mb * GT_BM + ew * 32 + src;
                        const uint64_t last = flush_query<KP>(p.cand + (size_t)qs * KP, p.locks + qs, p.thr_g + qs,
                                                              pend_all + (size_t)(ew * 32 + src) * GT_PEND_STRIDE, n_p,
                                                              p.tomb, p.n_rows, lane);
                        if (lane == src) {
                            if (last != KEY_SENTINEL) thr = fminf(thr, key_dist(last));
                            cnt = 0;
                        }
                        __syncwarp();
                    }
                }
                // all TMEM reads of this accumulator are done
                tc_fence_before();
                mbar_arrive(&tmem_empty[acc]);
            }
            // drain what is still pending
            __syncwarp();
            uint32_t fl = __ballot_sync(0xffffffffu, cnt > 0);
            while (fl) {
                const int src = __ffs(fl) - 1;
                fl &= fl - 1;
                const int n_p = __shfl_sync(0xffffffffu, cnt, src);
                const uint32_t qs = (uint32_t)mb * GT_BM + ew * 32 + src;
                flush_query<KP>(p.cand + (size_t)qs * KP, p.locks + qs, p.thr_g + qs,
                                pend_all + (size_t)(ew * 32 + src) * GT_PEND_STRIDE, n_p, p.tomb, p.n_rows, lane);
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, GT_TMEM_COLS);
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
    __half* q16 = nullptr; size_t q16_cap = 0;
    int* flags = nullptr; size_t flags_cap = 0;
    uint32_t* thr_g = nullptr; size_t thr_cap = 0;
    int* n_flagged = nullptr;
    int* h_n_flagged = nullptr;   // pinned
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
    const int MB = (int)((a.nq + GT_BM - 1) / GT_BM);
    const int n_tiles = (int)((a.n_rows + GT_BN - 1) / GT_BN);
    const int S = choose_slices(MB, n_tiles, a.num_sms);
    const size_t esz = a.f16 ? 2 : 4;
    cudaError_t e;
    if ((e = grow_dev(w->cand, w->cand_cap, a.nq * (size_t)kp)) != cudaSuccess) return e;
    if ((e = grow_dev(w->locks, w->locks_cap, a.nq)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w->cand, 0xFF, a.nq * (size_t)kp * sizeof(uint64_t), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w->locks, 0, a.nq * sizeof(int), st)) != cudaSuccess) return e;
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
    if (!make_tmap(&tmA, qa, a.f16, a.nq, (uint64_t)a.ld, GT_BM) || !make_tmap(&tmB, a.rows, a.f16, a.n_rows, (uint64_t)a.ld, GT_BN)) {
        err = "cuTensorMapEncodeTiled failed";
        return cudaErrorUnknown;
    }
    GemmParams gp{};
    gp.n_rows = a.n_rows; gp.nq = (uint32_t)a.nq;
    gp.num_kb = (int)((size_t)a.ld * esz / GT_KB_BYTES);
    gp.kb_elems = (int)(GT_KB_BYTES / esz);
    { const char* d = getenv("VDB_GEMM_DBG"); gp.dbg = d ? atoi(d) : 0; }
    gp.MB = MB; gp.S = S; gp.n_tiles = n_tiles; gp.n_items = MB * S;
    gp.sqnorm = a.sqnorm; gp.tomb = a.tomb; gp.cand = w->cand; gp.locks = w->locks; gp.thr_g = w->thr_g;
    const int grid = std::min(gp.n_items, a.num_sms);
    const bool l2 = a.metric == 0;
    if (a.f16) e = l2 ? launch_gemm_kp<true, true>(kp, tmA, tmB, gp, grid, st) : launch_gemm_kp<true, false>(kp, tmA, tmB, gp, grid, st);
    else       e = l2 ? launch_gemm_kp<false, true>(kp, tmA, tmB, gp, grid, st) : launch_gemm_kp<false, false>(kp, tmA, tmB, gp, grid, st);
    if (e != cudaSuccess) return e;

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
