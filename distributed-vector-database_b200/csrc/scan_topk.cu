// scan_topk.cu -- K1: HBM-streaming exact scan fused with a per-CTA top-k.
//
// Replaces hnswlib.Index.knn_query for 1..8 queries per pass (reference call site
// src/datanode/handler.py:364).  Bandwidth-bound: every live row of the shard is read exactly
// once per pass (algorithmic bytes = n_rows * row_bytes); the distance vector never exists in
// memory.  Persistent CTAs, two per SM when two rings fit (scan_plan), 10 warps each:
//   warp 8        producer: one 1-D bulk async copy (UBLKCP) of 16 contiguous rows per stage
//                 into a ring of shared-memory stages, completion on an mbarrier
//   warps 0..7    consumers: 2 rows each per stage, 128-bit LDS, fp32 FMA against the queries
//                 (held in shared memory).  Per-lane partial sums of 32 (row, query) values are kept in
//                 registers across stages and reduced together (canonical butterfly pairing, 31 shuffles),
//                 so each lane ends with one distance: one threshold test + one ballot per 32 values.
//                 Survivors are pushed (warp-aggregated atomicAdd) into a shared-memory candidate ring.
//   warp 9        selector: the only writer of the CTA's sorted top-k lists; drains the ring 32 slots at a
//                 time, pre-filters against the current k-th key, inserts.  Consumers never take a lock or
//                 wait for an insert (measured: inserts under a lock on the consumer path cost 50 us of a
//                 335 us scan; with the selector the scan runs at the speed of the bare copy ring).
// Each CTA leaves a sorted list of k keys per query; for one or two queries the kernel reduces the lists itself (last
// CTA of a group, then last group), wider passes leave them to merge_topk.cu.
// Candidate mode (ScanParams::cand_keys): the scan ran over an approximate plane of the rows (the fp16 shadow of an
// fp32 shard: half the bytes); the k best (distance, ROW) keys and the k-th distance go to the exact re-rank (K4w,
// gemm_topk.cu) instead of the caller -- vdb_api.cu shadow_scan.
#include <algorithm>
#include <cstdlib>

#include "block_select.cuh"
#include "common.cuh"
#include "kernels.h"

namespace vdbk {

constexpr int SCAN_WARPS = 8;
constexpr int SCAN_R = 2;                                  // rows per consumer warp per stage
constexpr int SCAN_STAGE_ROWS = SCAN_WARPS * SCAN_R;       // 16
constexpr int SCAN_THREADS = (SCAN_WARPS + 2) * 32;        // + producer warp + selector warp
constexpr int SCAN_RING_SMALL = 512;                       // candidate ring slots (power of two), k <= 32
constexpr int SCAN_RING_LARGE = 4096;                      // k > 32: an insert costs O(k/32) and the first rounds of a
                                                           // scan pass everything, so the ring must absorb a longer burst

template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int PER16 = 4;  // elements per 16-byte lane load
    __device__ static __forceinline__ void load(const void* p, float (&v)[4]) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <> struct Elem<__half> {
    static constexpr int PER16 = 8;
    __device__ static __forceinline__ void load(const void* p, float (&v)[8]) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __half22float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
};

// Sorted insert of `key` into list[0..k) (ascending) by the selector warp (the only writer).
// Consumers concurrently read the high word of list[k-1] as their threshold: during the shift it only
// ever moves from the old tail to a smaller-or-equal value, so a racing read is a valid (stale) bound.
__device__ __forceinline__ void list_insert(volatile uint64_t* list, int k, uint64_t key) {
    const int lane = lane_id();
    if (key >= list[k - 1]) return;
    int cnt = 0;
    for (int i = lane; i < k; i += 32) cnt += (list[i] < key) ? 1 : 0;
    const int pos = warp_sum_int(cnt);
    for (int hi = k - 2; hi >= pos; hi -= 32) {
        const int i = hi - lane;
        uint64_t v = 0;
        if (i >= pos) v = list[i];
        __syncwarp();
        if (i >= pos) list[i + 1] = v;
        __syncwarp();
    }
    if (lane == 0) list[pos] = key;
    __syncwarp();
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct ScanCtl {            // shared-memory control block of the candidate ring
    uint32_t tail;          // slots reserved by consumers (monotonic)
    uint32_t head;          // slots retired by the selector (monotonic)
    uint32_t done;          // consumer warps that have finished
    uint32_t pad;
};

// METRIC (0 = squared L2 in direct form, 1 = 1 - dot) and QREG (the query in registers, see below) are template
// parameters: as run-time branches inside the 16-stage unrolled accumulation loop they quadrupled the code (13k SASS
// instructions) and a fifth of the warp samples sat in instruction-cache misses (stall_no_inst).
template <typename T, int NQ, int METRIC, bool QREG>
__global__ void __launch_bounds__(SCAN_THREADS, 2) scan_topk_kernel(const ScanParams p) {
    pdl_prologue();
    constexpr int PER16 = Elem<T>::PER16;
    constexpr int VS = SCAN_R * NQ;   // (row, query) values a warp produces per stage
    // (row, query) values reduced together.  32 amortises the threshold test and the ballot best, but unrolls the
    // accumulation over 32 / VS stages; the register-query variant (short rows: the code per stage is the cost) takes
    // 8: the unrolled loop is a quarter as long and stays in the instruction cache, at 9 shuffles per 8 values
    // instead of 31 per 32.
    constexpr int NV = QREG ? 8 : 32;
    constexpr int G = NV / VS;        // stages whose partial sums are reduced together
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t row_bytes = p.row_bytes;
    const uint32_t stage_bytes = row_bytes * SCAN_STAGE_ROWS;
    const int stages = p.stages;
    uint8_t* stage_base = smem;
    float* qs = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);          // [NQ][ld]
    uint64_t* lists = reinterpret_cast<uint64_t*>(qs + (size_t)NQ * p.ld);                // [NQ][k]
    const uint32_t ring_n = (uint32_t)p.ring, ring_mask = ring_n - 1;
    uint64_t* ring = lists + (size_t)NQ * p.k;                                            // [ring_n]
    uint64_t* full = ring + ring_n;
    uint64_t* empty = full + stages;
    ScanCtl* ctl = reinterpret_cast<ScanCtl*>(empty + stages);
    uint8_t* ring_q = reinterpret_cast<uint8_t*>(ctl + 1);                                // [ring_n]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = p.k;

    for (int i = threadIdx.x; i < NQ * k; i += SCAN_THREADS) lists[i] = KEY_SENTINEL;
    for (int i = threadIdx.x; i < (int)ring_n; i += SCAN_THREADS) ring[i] = KEY_SENTINEL;
    if (threadIdx.x == 0) {
        ctl->tail = 0; ctl->head = 0; ctl->done = 0;
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SCAN_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const uint32_t nchunks = (p.n_rows + SCAN_STAGE_ROWS - 1) / SCAN_STAGE_ROWS;

    if (warp == SCAN_WARPS) {
        // ---------------- producer: starts streaming while the consumers fetch the queries ----------------
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_first();
            int s = 0;
            uint32_t ph = 0;
            for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
                mbar_wait(&empty[s], ph ^ 1);
                const uint32_t r0 = chunk * SCAN_STAGE_ROWS;
                const uint32_t nr = min((uint32_t)SCAN_STAGE_ROWS, p.n_rows - r0);
                const uint32_t bytes = nr * row_bytes;
                mbar_arrive_expect_tx(&full[s], bytes);
                if (p.dbg & 2)
                    bulk_copy_g2s(stage_base + (size_t)s * stage_bytes,
                                  reinterpret_cast<const uint8_t*>(p.rows) + (size_t)r0 * row_bytes, bytes, &full[s]);
                else
                    bulk_copy_g2s_stream(stage_base + (size_t)s * stage_bytes,
                                         reinterpret_cast<const uint8_t*>(p.rows) + (size_t)r0 * row_bytes, bytes,
                                         &full[s], policy);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == SCAN_WARPS + 1) {
        // ---------------- selector: drains the candidate ring into the sorted lists ----------------
        volatile uint64_t* vring = ring;
        volatile uint64_t* vlists = lists;
        volatile ScanCtl* vctl = ctl;
        uint32_t head = 0;
        for (;;) {
            const uint32_t slot = (head + lane) & ring_mask;
            const uint64_t key = vring[slot];
            const uint32_t ready = __ballot_sync(0xffffffffu, key != KEY_SENTINEL);
            const int n = (ready == 0xffffffffu) ? 32 : __ffs(~ready) - 1;   // leading run of written slots
            if (n == 0) {
                // all pushes of a finished warp are visible before its `done` increment
                if (vctl->done == SCAN_WARPS && head == vctl->tail) break;
                __nanosleep(100);
                continue;
            }
            __threadfence_block();
            const int q = (NQ > 1 && lane < n) ? reinterpret_cast<volatile uint8_t*>(ring_q)[slot] : 0;
            if (lane < n) vring[slot] = KEY_SENTINEL;
            __syncwarp();
            head += n;
            if (lane == 0) {
                __threadfence_block();
                vctl->head = head;
            }
            const bool cand = lane < n && key < vlists[(size_t)q * k + (k - 1)];
            uint32_t cm = __ballot_sync(0xffffffffu, cand);
            while (cm) {
                const int src = __ffs(cm) - 1;
                cm &= cm - 1;
                const uint64_t kk = __shfl_sync(0xffffffffu, key, src);
                const int qq = __shfl_sync(0xffffffffu, q, src);
                list_insert(vlists + (size_t)qq * k, k, kk);
            }
        }
    } else {
        // ---------------- consumers ----------------
        for (int i = threadIdx.x; i < NQ * (int)p.ld; i += SCAN_WARPS * 32) {
            const int qi = i / (int)p.ld, c = i - qi * (int)p.ld;
            float v = 0.0f;
            if (qi < p.nq) {
                if (p.q_raw) v = c < p.dim ? p.q_raw[(size_t)qi * p.dim + c] : 0.0f;
                else v = p.q[(size_t)qi * p.ld + c];
            }
            qs[i] = v;
        }
        if (p.q_raw && p.normalize) {
            // same arithmetic as prepare_queries_kernel (insert.cu), so both paths see identical queries
            named_barrier_sync(1, SCAN_WARPS * 32);
            for (int qi = warp; qi < p.nq; qi += SCAN_WARPS) {
                float* qv = qs + (size_t)qi * p.ld;
                const float ssq = warp_sumsq_f32(qv, p.dim);
                const float scale = 1.0f / (sqrtf(ssq) + 1e-30f);
                for (int c = lane; c < p.dim; c += 32) qv[c] = qv[c] * scale;
            }
        }
        named_barrier_sync(1, SCAN_WARPS * 32);

        // Per-lane partial sums of G consecutive stages (32 (row, query) values) stay in registers and are
        // reduced together: 31 shuffles + one threshold test + one ballot per 32 values instead of per stage.
        const int nld16 = row_bytes / 512;  // 16-byte lane loads per row
        // One query against 1 KB rows (512 fp16 -- the shadow plane of a 512-d fp32 shard -- or 256 fp32): a lane meets
        // the same 2 x PER16 query elements in every row, so they live in registers.  Read from shared memory they cost
        // as many bytes as the rows themselves (fp32 query against fp16 rows, shared by SCAN_R = 2 rows): bulk-copy
        // writes + row reads + query reads then run the shared-memory pipe at ~0.75 of its bandwidth and the scan at
        // 4.9 TB/s instead of the HBM rate.
        static_assert(!QREG || NQ == 1, "register-resident query: one query");     // launched only when nld16 == 2
        float qreg[QREG ? 2 : 1][PER16];
        if constexpr (QREG) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int e = 0; e < PER16; ++e) qreg[c][e] = qs[(size_t)(c * 32 + lane) * PER16 + e];
        }
        const uint32_t my_iters = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        volatile ScanCtl* vctl = ctl;
        int s = 0;
        uint32_t ph = 0;
        for (uint32_t it0 = 0; it0 < my_iters; it0 += G) {
            float acc[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (it0 + g < my_iters) {   // warp-uniform
                    mbar_wait(&full[s], ph);
                    const uint8_t* sbase = stage_base + (size_t)s * stage_bytes + (size_t)(warp * SCAN_R) * row_bytes;
                    if constexpr (QREG) {
                        {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                float dv[SCAN_R][PER16];
#pragma unroll
                                for (int r = 0; r < SCAN_R; ++r)
                                    Elem<T>::load(sbase + (size_t)r * row_bytes + (size_t)(c * 32 + lane) * 16, dv[r]);
#pragma unroll
                                for (int r = 0; r < SCAN_R; ++r) {
                                    float a = acc[g * VS + r];
                                    if constexpr (METRIC == 0) {
#pragma unroll
                                        for (int e = 0; e < PER16; ++e) {
                                            const float t = dv[r][e] - qreg[c][e];
                                            a = fmaf(t, t, a);
                                        }
                                    } else {
#pragma unroll
                                        for (int e = 0; e < PER16; ++e) a = fmaf(dv[r][e], qreg[c][e], a);
                                    }
                                    acc[g * VS + r] = a;
                                }
                            }
                        }
                    } else if (!(p.dbg & 1)) {   // (dbg & 1: no accumulation, experiments)
#pragma unroll 2
                        for (int c = 0; c < nld16; ++c) {
                            float dv[SCAN_R][PER16];
#pragma unroll
                            for (int r = 0; r < SCAN_R; ++r)
                                Elem<T>::load(sbase + (size_t)r * row_bytes + (size_t)(c * 32 + lane) * 16, dv[r]);
#pragma unroll
                            for (int qi = 0; qi < NQ; ++qi) {
                                float qv[PER16];
                                const float* qp = qs + (size_t)qi * p.ld + (size_t)(c * 32 + lane) * PER16;
#pragma unroll
                                for (int e = 0; e < PER16; e += 4) {
                                    float4 t = *reinterpret_cast<const float4*>(qp + e);
                                    qv[e] = t.x; qv[e + 1] = t.y; qv[e + 2] = t.z; qv[e + 3] = t.w;
                                }
#pragma unroll
                                for (int r = 0; r < SCAN_R; ++r) {
                                    float a = acc[g * VS + r * NQ + qi];
                                    if constexpr (METRIC == 0) {
#pragma unroll
                                        for (int e = 0; e < PER16; ++e) {
                                            const float t = dv[r][e] - qv[e];
                                            a = fmaf(t, t, a);
                                        }
                                    } else {
#pragma unroll
                                        for (int e = 0; e < PER16; ++e) a = fmaf(dv[r][e], qv[e], a);
                                    }
                                    acc[g * VS + r * NQ + qi] = a;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);   // data is in registers: free the slot
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
            if (p.dbg & 4) {   // experiment: no reduction / selection
                if (acc[0] == 123.456f) lists[0] = 0;
                continue;
            }
            warp_sum_multi<NV>(acc);        // lane L now holds the total of value index vi (NV < 32: 32 / NV lanes do)
            const int vi = value_index_of_lane<NV>(lane);
            const int g = vi / VS, rem = vi - g * VS;
            const int r = rem / NQ, qi = rem - r * NQ;
            const uint32_t it = it0 + g;
            const uint32_t row = (blockIdx.x + it * gridDim.x) * SCAN_STAGE_ROWS + warp * SCAN_R + r;
            const float dist = (METRIC == 0) ? acc[0] : 1.0f - acc[0];
            const bool valid = it < my_iters && row < p.n_rows && qi < p.nq && (NV == 32 || lane == lane_of_value_index<NV>(vi));
            // high word (distance bits) of the list tail = current threshold of query qi
            const uint32_t tail_hi = reinterpret_cast<volatile uint32_t*>(lists + (size_t)qi * k + (k - 1))[1];
            bool pass = valid && float_to_ordered(dist) <= tail_hi;
            uint32_t label = p.label_base + row;
            if (pass) {
                if (p.tomb && ((p.tomb[row >> 5] >> (row & 31)) & 1u)) pass = false;
                else if (p.labels) label = p.labels[row];
            }
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                // warp-aggregated push into the candidate ring; the selector warp owns the lists
                const int n = __popc(m);
                const int leader = __ffs(m) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&ctl->tail, (uint32_t)n);
                base = __shfl_sync(0xffffffffu, base, leader);
                while ((int)(base + n - vctl->head) > (int)ring_n) {
                }
                if (pass) {
                    const uint32_t slot = (base + __popc(m & ((1u << lane) - 1))) & ring_mask;
                    if constexpr (NQ > 1) {     // which query the key belongs to, visible before the key itself
                        ring_q[slot] = (uint8_t)qi;
                        __threadfence_block();
                    }
                    reinterpret_cast<volatile uint64_t*>(ring)[slot] = make_key(dist, label);
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            atomicAdd(&ctl->done, 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.nq * k; i += SCAN_THREADS) {
        const int qi = i / k, j = i - qi * k;
        p.out_keys[((size_t)qi * gridDim.x + blockIdx.x) * k + j] = lists[i];
    }
    if (p.merge_group <= 0) return;

    // ---------------- in-kernel merge: last CTA of a group, then last group ----------------
    // The copy ring (>= 64 KB, idle now) holds the keys of one merge step and its scratch; the candidate ring
    // (>= 4 KB) holds the histogram.  Lists written by other CTAs are read with L1-bypassing loads after the
    // counter handshake (fence + atomic on both sides).
    __shared__ int s_last;
    uint64_t* m_in = reinterpret_cast<uint64_t*>(stage_base);
    uint64_t* m_tmp = m_in + SCAN_MERGE_KEYS;
    int* hist = reinterpret_cast<int*>(ring);
    uint32_t* sh = reinterpret_cast<uint32_t*>(hist + 256);
    uint64_t* m_out = reinterpret_cast<uint64_t*>(sh + 8);          // [k] behind the histogram (k <= 384 with the 4 KB ring)
    const int grid = (int)gridDim.x, groups = (grid + p.merge_group - 1) / p.merge_group;
    const int group = (int)blockIdx.x / p.merge_group;
    const int g0 = group * p.merge_group, g1 = min(grid, g0 + p.merge_group);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&p.merge_ctr[1 + group], 1u) == (unsigned)(g1 - g0 - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    auto emit = [&](int qi, int cnt) {      // m_out[0..k) -> final outputs of query qi
        if (p.cand_keys) {
            for (int j = threadIdx.x; j < k; j += SCAN_THREADS) p.cand_keys[(size_t)qi * p.cand_stride + j] = m_out[j];
            if (threadIdx.x == 0) {
                p.cand_cnt[qi] = cnt;
                p.cand_tau[qi] = cnt == k ? key_dist(m_out[k - 1]) : __int_as_float(0x7f800000);
            }
            return;
        }
        for (int j = threadIdx.x; j < k; j += SCAN_THREADS) {
            const uint64_t key = m_out[j];
            const bool real = key != KEY_SENTINEL;
            p.out_ids[(size_t)qi * k + j] = real ? (int64_t)key_label(key) : -1;
            p.out_dist[(size_t)qi * k + j] = real ? key_dist(key) : __int_as_float(0x7f800000);
        }
        if (threadIdx.x == 0 && p.out_counts) p.out_counts[qi] = cnt;
    };
    for (int qi = 0; qi < p.nq; ++qi) {
        const uint64_t* src = p.out_keys + ((size_t)qi * grid + g0) * k;      // the group's lists are contiguous
        const int n = (g1 - g0) * k;
        for (int i = threadIdx.x; i < n; i += SCAN_THREADS) m_in[i] = __ldcg(src + i);
        __syncthreads();
        const int cnt = block_topk_sorted(m_in, n, k, m_tmp, m_out, hist, sh);
        if (groups == 1) emit(qi, cnt);
        else for (int j = threadIdx.x; j < k; j += SCAN_THREADS) p.group_keys[((size_t)qi * groups + group) * k + j] = m_out[j];
        __syncthreads();
    }
    if (groups == 1) {
        if (threadIdx.x == 0) p.merge_ctr[1] = 0;                  // leave the counters zero for the next launch
        return;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&p.merge_ctr[0], 1u) == (unsigned)(groups - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int qi = 0; qi < p.nq; ++qi) {
        const uint64_t* src = p.group_keys + (size_t)qi * groups * k;
        const int n = groups * k;
        for (int i = threadIdx.x; i < n; i += SCAN_THREADS) m_in[i] = __ldcg(src + i);
        __syncthreads();
        const int cnt = block_topk_sorted(m_in, n, k, m_tmp, m_out, hist, sh);
        emit(qi, cnt);
        __syncthreads();
    }
    for (int i = threadIdx.x; i <= groups; i += SCAN_THREADS) p.merge_ctr[i] = 0;
}

static int scan_ring_slots(int k) { return k > 32 ? SCAN_RING_LARGE : SCAN_RING_SMALL; }
static size_t scan_fixed_smem(int nq_t, int ld, int k, int stages) {
    return (size_t)nq_t * ld * 4 + (size_t)nq_t * k * 8 + (size_t)scan_ring_slots(k) * 9 + (size_t)stages * 16 + 64 + 128;
}

template <typename T, int NQ, int METRIC, bool QREG>
static cudaError_t launch_v(const ScanParams& p, int grid, size_t smem, cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaFuncAttributes fa{};
        cudaError_t e = cudaFuncGetAttributes(&fa, scan_topk_kernel<T, NQ, METRIC, QREG>);
        if (e != cudaSuccess) return e;
        // static + dynamic shared memory of a block may not exceed 227 KB
        e = cudaFuncSetAttribute(scan_topk_kernel<T, NQ, METRIC, QREG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024 - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    cudaError_t e = launch_pdl(scan_topk_kernel<T, NQ, METRIC, QREG>, dim3(grid), dim3(SCAN_THREADS), smem, st, p);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}
template <typename T, int NQ>
static cudaError_t launch_t(const ScanParams& p, int grid, size_t smem, cudaStream_t st) {
    if constexpr (NQ == 1) {
        static const bool qreg_on = [] { const char* e = getenv("VDB_SCAN_QREG"); return !(e && e[0] == '0'); }();
        if (qreg_on && p.row_bytes == 1024 && !(p.dbg & 1))
            return p.metric == 0 ? launch_v<T, 1, 0, true>(p, grid, smem, st) : launch_v<T, 1, 1, true>(p, grid, smem, st);
    }
    return p.metric == 0 ? launch_v<T, NQ, 0, false>(p, grid, smem, st) : launch_v<T, NQ, 1, false>(p, grid, smem, st);
}

int scan_max_k(int nq_t, int ld, uint32_t row_bytes) {
    const size_t budget = 227 * 1024 - 256;      // 256 bytes: the kernel's static shared memory
    const size_t min_stage = (size_t)2 * row_bytes * SCAN_STAGE_ROWS;
    // conservative: counted with the large candidate ring whatever k turns out to be
    const size_t fixed0 = scan_fixed_smem(nq_t, ld, 0, 8) + (size_t)(SCAN_RING_LARGE - SCAN_RING_SMALL) * 9;
    if (min_stage + fixed0 >= budget) return 0;
    return (int)((budget - min_stage - fixed0) / ((size_t)nq_t * 8));
}

// Launch shape.  Two CTAs per SM whenever two rings fit (measured on B200, 1M x 512 fp32, one query:
// 1 CTA x 6 stages 360 us, 2 CTAs x 2 stages 335 us, 3 x 2 339 us): a second CTA's consumers fill the
// bubbles the first one leaves at the reductions.  The rings together keep >= 128 KB of bulk copies in flight
// per SM when shared memory allows (64 KB/SM already reaches ~7.2 TB/s with bare copies).
ScanPlan scan_plan(int nq, uint32_t ld, uint32_t row_bytes, int k, uint32_t n_rows, int num_sms) {
    ScanPlan pl{};
    const int nq_t = nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : 8;
    const size_t stage_bytes = (size_t)row_bytes * SCAN_STAGE_ROWS;
    const size_t sm_budget = 227 * 1024 - 256;   // 256 bytes: the kernel's static shared memory
    int ctas = 2, stages = 0;
    if (const char* e = getenv("VDB_SCAN_CTAS")) ctas = atoi(e);
    if (ctas < 1) ctas = 1;
    for (; ctas >= 1; --ctas) {
        const size_t budget = ctas == 1 ? sm_budget : (228 * 1024) / ctas - 1024;   // 1 KB reserved per CTA
        int want = (int)((128 * 1024 / ctas + stage_bytes - 1) / stage_bytes);
        if (want < 2) want = 2;
        if (want > 8) want = 8;
        if (const char* e = getenv("VDB_SCAN_STAGES")) want = atoi(e);
        stages = want;
        while (stages > 2 && stages * stage_bytes + scan_fixed_smem(nq_t, ld, k, stages) > budget) --stages;
        if (stages * stage_bytes + scan_fixed_smem(nq_t, ld, k, stages) <= budget) break;
        stages = 0;
    }
    if (stages == 0) return pl;     // does not fit at all (pl.grid == 0)
    pl.ctas_per_sm = ctas;
    pl.stages = stages;
    pl.smem = stages * stage_bytes + scan_fixed_smem(nq_t, ld, k, stages);
    const uint32_t nchunks = (n_rows + SCAN_STAGE_ROWS - 1) / SCAN_STAGE_ROWS;
    pl.grid = (int)min((uint32_t)(num_sms * ctas), nchunks);
    if (pl.grid < 1) pl.grid = 1;
    // In-kernel merge for one or two queries (their merges run one after the other on the last CTA; wider passes keep
    // the merge kernel, one block per query): groups of CTAs whose lists fit one merge step, at most two levels.
    pl.merge_group = pl.merge_groups = 0;
    static const bool fuse = [] { const char* e = getenv("VDB_SCAN_FUSED_MERGE"); return !(e && e[0] == '0'); }();
    const size_t merge_bytes = 2 * (size_t)SCAN_MERGE_KEYS * sizeof(uint64_t);
    if (fuse && nq_t <= 2 && k <= SCAN_MERGE_KEYS && (size_t)stages * stage_bytes >= merge_bytes &&
        (size_t)scan_ring_slots(k) * 8 >= 256 * 4 + 32 + (size_t)k * 8) {
        // one level while the lists together are a small merge; otherwise ~sqrt(grid) CTAs per group: both levels then
        // rank a few hundred keys (a merge step costs about as much as it holds keys, and the last one is a serial
        // tail behind the scan: 296 lists of 32 keys as 3 groups of 128 cost ~25 us more than as 18 groups of 17)
        int per_group = std::max(1, SCAN_MERGE_KEYS / k);
        if ((size_t)pl.grid * k > 1024) {
            int r = 1;
            while (r * r < pl.grid) ++r;
            per_group = std::min(per_group, std::max(2, r));
        }
        if (const char* e = getenv("VDB_SCAN_MERGE_GROUP")) per_group = std::max(1, std::min(atoi(e), std::max(1, SCAN_MERGE_KEYS / k)));
        const int groups = (pl.grid + per_group - 1) / per_group;
        if ((size_t)groups * k <= (size_t)SCAN_MERGE_KEYS) { pl.merge_group = per_group; pl.merge_groups = groups; }
    }
    return pl;
}

// Launch K1 with a plan made for >= p.nq queries.  p.nq in 1..8; pl.grid lists of k keys per query come back.
cudaError_t launch_scan_topk(ScanParams p, bool f16, const ScanPlan& pl, cudaStream_t st) {
    const int nq_t = p.nq <= 1 ? 1 : p.nq <= 2 ? 2 : p.nq <= 4 ? 4 : 8;
    if (pl.grid == 0) return cudaErrorInvalidConfiguration;
    if (const char* e = getenv("VDB_SCAN_DBG")) p.dbg = atoi(e);
    p.stages = pl.stages;
    p.ring = scan_ring_slots(p.k);
    const size_t smem = pl.smem;
    const int grid = pl.grid;
#define VDB_SCAN_CASE(NQV)                                                        \
    case NQV:                                                                     \
        return f16 ? launch_t<__half, NQV>(p, grid, smem, st) : launch_t<float, NQV>(p, grid, smem, st);
    switch (nq_t) {
        VDB_SCAN_CASE(1)
        VDB_SCAN_CASE(2)
        VDB_SCAN_CASE(4)
        VDB_SCAN_CASE(8)
    }
#undef VDB_SCAN_CASE
    return cudaErrorInvalidValue;
}

}  // namespace vdbk
