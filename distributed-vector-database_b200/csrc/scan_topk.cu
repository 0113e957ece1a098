// scan_topk.cu -- K1: HBM-streaming exact scan fused with a per-CTA top-k.
//
// Replaces hnswlib.Index.knn_query for 1..8 queries per pass (reference call site
// src/datanode/handler.py:364).  Bandwidth-bound: every live row of the shard is read exactly
// once per pass (algorithmic bytes = n_rows * row_bytes); the distance vector never exists in
// memory.  One persistent CTA per SM:
//   warp 8        producer: one 1-D bulk async copy (UBLKCP) of 16 contiguous rows per stage
//                 into a ring of shared-memory stages, completion on an mbarrier
//   warps 0..7    consumers: 2 rows each per stage, 128-bit LDS, fp32 FMA against the queries
//                 (held in shared memory), canonical butterfly reduction, threshold test
//                 against the CTA's current k-th key, rare warp-cooperative sorted insert.
// Each CTA leaves a sorted list of k keys per query; merge_topk.cu reduces grid lists to one.
#include "common.cuh"
#include "kernels.h"

namespace vdbk {

constexpr int SCAN_WARPS = 8;
constexpr int SCAN_R = 2;                                  // rows per consumer warp per stage
constexpr int SCAN_STAGE_ROWS = SCAN_WARPS * SCAN_R;       // 16
constexpr int SCAN_THREADS = (SCAN_WARPS + 1) * 32;

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

// Sorted insert of `key` into list[0..k) (ascending), executed by a whole converged warp.
__device__ __forceinline__ void list_insert(volatile uint64_t* list, int k, uint64_t key, int* lock) {
    const int lane = lane_id();
    if (lane == 0) {
        while (atomicCAS(lock, 0, 1) != 0) {
        }
    }
    __syncwarp();
    __threadfence_block();
    if (key < list[k - 1]) {
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
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        atomicExch(lock, 0);
    }
}

template <typename T, int NQ>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_topk_kernel(const ScanParams p) {
    constexpr int PER16 = Elem<T>::PER16;
    constexpr int V = SCAN_R * NQ;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t row_bytes = p.row_bytes;
    const uint32_t stage_bytes = row_bytes * SCAN_STAGE_ROWS;
    const int stages = p.stages;
    uint8_t* stage_base = smem;
    float* qs = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);          // [NQ][ld]
    uint64_t* lists = reinterpret_cast<uint64_t*>(qs + (size_t)NQ * p.ld);                // [NQ][k]
    uint64_t* full = lists + (size_t)NQ * p.k;
    uint64_t* empty = full + stages;
    int* locks = reinterpret_cast<int*>(empty + stages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = p.k;

    for (int i = threadIdx.x; i < NQ * (int)p.ld; i += SCAN_THREADS) {
        const int qi = i / (int)p.ld, c = i - qi * (int)p.ld;
        qs[i] = (qi < p.nq) ? p.q[(size_t)qi * p.ld + c] : 0.0f;
    }
    for (int i = threadIdx.x; i < NQ * k; i += SCAN_THREADS) lists[i] = KEY_SENTINEL;
    if (threadIdx.x < NQ) locks[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SCAN_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const uint32_t nchunks = (p.n_rows + SCAN_STAGE_ROWS - 1) / SCAN_STAGE_ROWS;

    if (warp == SCAN_WARPS) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_first();
            uint32_t it = 0;
            for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
                const int s = it % stages;
                const uint32_t ph = (it / stages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const uint32_t r0 = chunk * SCAN_STAGE_ROWS;
                const uint32_t nr = min((uint32_t)SCAN_STAGE_ROWS, p.n_rows - r0);
                const uint32_t bytes = nr * row_bytes;
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_copy_g2s_stream(stage_base + (size_t)s * stage_bytes,
                                     reinterpret_cast<const uint8_t*>(p.rows) + (size_t)r0 * row_bytes, bytes,
                                     &full[s], policy);
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int nld16 = row_bytes / 512;  // 16-byte lane loads per row
        uint32_t it = 0;
        for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, ++it) {
            const int s = it % stages;
            const uint32_t ph = (it / stages) & 1;
            mbar_wait(&full[s], ph);
            const uint32_t row0 = chunk * SCAN_STAGE_ROWS + warp * SCAN_R;
            const uint8_t* sbase = stage_base + (size_t)s * stage_bytes + (size_t)(warp * SCAN_R) * row_bytes;

            float acc[V];
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = 0.0f;

            if (row0 < p.n_rows) {
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
                            float a = acc[r * NQ + qi];
                            if (p.metric == 0) {
#pragma unroll
                                for (int e = 0; e < PER16; ++e) {
                                    const float t = dv[r][e] - qv[e];
                                    a = fmaf(t, t, a);
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < PER16; ++e) a = fmaf(dv[r][e], qv[e], a);
                            }
                            acc[r * NQ + qi] = a;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);   // data is in registers: free the slot

            if (row0 >= p.n_rows) continue;
            warp_sum_multi<V>(acc);
            const int vi = value_index_of_lane<V>(lane);
            const int r = vi / NQ, qi = vi - r * NQ;
            const uint32_t row = row0 + r;
            float dist = (p.metric == 0) ? acc[0] : 1.0f - acc[0];
            const bool holder = (lane == lane_of_value_index<V>(vi)) && row < p.n_rows && qi < p.nq;
            const uint32_t tail_hi = (uint32_t)(((volatile uint64_t*)lists)[(size_t)qi * k + (k - 1)] >> 32);
            const bool pass = holder && float_to_ordered(dist) <= tail_hi;
            uint32_t m = __ballot_sync(0xffffffffu, pass);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float d = __shfl_sync(0xffffffffu, dist, src);
                const uint32_t rw = __shfl_sync(0xffffffffu, row, src);
                const int q = __shfl_sync(0xffffffffu, qi, src);
                if (p.tomb && ((p.tomb[rw >> 5] >> (rw & 31)) & 1u)) continue;
                const uint32_t label = p.labels ? p.labels[rw] : rw;
                list_insert(lists + (size_t)q * k, k, make_key(d, label), &locks[q]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.nq * k; i += SCAN_THREADS) {
        const int qi = i / k, j = i - qi * k;
        p.out_keys[((size_t)qi * gridDim.x + blockIdx.x) * k + j] = lists[i];
    }
}

static size_t scan_fixed_smem(int nq_t, int ld, int k, int stages) {
    return (size_t)nq_t * ld * 4 + (size_t)nq_t * k * 8 + (size_t)stages * 16 + 64 + 128;
}

template <typename T, int NQ>
static cudaError_t launch_t(const ScanParams& p, int grid, size_t smem, cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(scan_topk_kernel<T, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    scan_topk_kernel<T, NQ><<<grid, SCAN_THREADS, smem, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

int scan_max_k(int nq_t, int ld, uint32_t row_bytes) {
    const size_t budget = 227 * 1024;
    const size_t min_stage = (size_t)2 * row_bytes * SCAN_STAGE_ROWS;
    const size_t fixed0 = scan_fixed_smem(nq_t, ld, 0, 8);
    if (min_stage + fixed0 >= budget) return 0;
    return (int)((budget - min_stage - fixed0) / ((size_t)nq_t * 8));
}

// Launch K1.  p.nq in 1..8.  Returns the grid used through *grid_out (lists per query).
cudaError_t launch_scan_topk(ScanParams p, bool f16, int num_sms, int* grid_out, cudaStream_t st) {
    const int nq_t = p.nq <= 1 ? 1 : p.nq <= 2 ? 2 : p.nq <= 4 ? 4 : 8;
    const uint32_t stage_bytes = p.row_bytes * SCAN_STAGE_ROWS;
    const size_t budget = 227 * 1024;
    int stages = 8;
    while (stages > 2 && (size_t)stages * stage_bytes + scan_fixed_smem(nq_t, p.ld, p.k, stages) > budget) --stages;
    const size_t smem = (size_t)stages * stage_bytes + scan_fixed_smem(nq_t, p.ld, p.k, stages);
    if (smem > budget) return cudaErrorInvalidConfiguration;
    p.stages = stages;
    const uint32_t nchunks = (p.n_rows + SCAN_STAGE_ROWS - 1) / SCAN_STAGE_ROWS;
    int grid = (int)min((uint32_t)num_sms, nchunks);
    if (grid < 1) grid = 1;
    *grid_out = grid;
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
