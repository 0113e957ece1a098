// block_select.cuh -- block-wide selection helpers shared by the tensor path's selects / re-rank (gemm_topk.cu) and the
// scan kernel's in-kernel merge (scan_topk.cu).
#pragma once
#include "common.cuh"

namespace vdbk {

// value bits of the `rank`-th smallest (1-based) of the n >= rank keys in sk[].  Radix select
// from the highest byte in which the values differ: the bytes above it are common to all keys and would send every
// atomicAdd of a pass to ONE histogram bin.  Stops as soon as the bin that holds the rank has a single key.
__device__ __forceinline__ uint32_t block_kth_bits(const uint64_t* sk, int n, int rank, int* hist, uint32_t* sh) {
    // sh[0] = min, sh[1] = max, sh[2] = prefix, sh[3] = rank, sh[4] = count in the chosen bin
    const int tid = threadIdx.x, RW_THREADS = blockDim.x;
    if (tid == 0) { sh[0] = 0xFFFFFFFFu; sh[1] = 0u; }
    __syncthreads();
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (int i = tid; i < n; i += RW_THREADS) {
        const uint32_t v = (uint32_t)(sk[i] >> 32);
        lo = min(lo, v); hi = max(hi, v);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((tid & 31) == 0) { atomicMin(&sh[0], lo); atomicMax(&sh[1], hi); }
    __syncthreads();
    lo = sh[0]; hi = sh[1];
    if (lo == hi) return lo;
    int shift = ((31 - __clz(lo ^ hi)) >> 3) << 3;          // byte of the highest differing bit
    uint32_t mask = shift == 24 ? 0u : ~((1u << (shift + 8)) - 1u);
    if (tid == 0) { sh[2] = lo & mask; sh[3] = (uint32_t)rank; sh[4] = 0u; }
    for (; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += RW_THREADS) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = sh[2];
        for (int i = tid; i < n; i += RW_THREADS) {
            const uint32_t v = (uint32_t)(sk[i] >> 32);
            if ((v & mask) == prefix) atomicAdd(&hist[(v >> shift) & 255], 1);
        }
        __syncthreads();
        if (tid < 32) {
            int local[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { local[i] = hist[tid * 8 + i]; sum += local[i]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            const int excl = incl - sum;
            const int r = (int)sh[3];
            if (r > excl && r <= incl) {
                int run = excl;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (r > run && r <= run + local[i]) {
                        sh[2] = prefix | ((uint32_t)(tid * 8 + i) << shift);
                        sh[3] = (uint32_t)(r - run);
                        sh[4] = (uint32_t)local[i];
                    }
                    run += local[i];
                }
            }
        }
        mask |= 0xFFu << shift;
        __syncthreads();
        if (sh[4] == 1u && shift > 0) {            // one key left under this prefix: it is the answer
            const uint32_t prefix1 = sh[2];
            __syncthreads();
            for (int i = tid; i < n; i += RW_THREADS) {
                const uint32_t v = (uint32_t)(sk[i] >> 32);
                if ((v & mask) == prefix1) sh[2] = v;
            }
            __syncthreads();
            break;
        }
    }
    return sh[2];
}

// Top-k of n sortable keys held in shared memory, ascending, for a block of any size.
//   in[n]     keys (KEY_SENTINEL = absent; real keys are distinct); destroyed
//   tmp[n]    scratch
//   out[k]    result, padded with KEY_SENTINEL; returns the number of real keys (<= k)
// Radix select of the k-th value, then the few keys at or below it are ranked by counting.
__device__ __forceinline__ int block_topk_sorted(uint64_t* in, int n, int k, uint64_t* tmp, uint64_t* out, int* hist, uint32_t* sh) {
    const int tid = threadIdx.x, NT = blockDim.x;
    int* cnt = reinterpret_cast<int*>(sh + 6);                       // sh[6], sh[7]: counters
    if (tid == 0) { cnt[0] = 0; cnt[1] = 0; }
    __syncthreads();
    int v = 0;
    for (int i = tid; i < n; i += NT) v += in[i] != KEY_SENTINEL;
    v = warp_sum_int(v);
    if ((tid & 31) == 0 && v) atomicAdd(&cnt[0], v);
    __syncthreads();
    const int n_valid = cnt[0];
    uint32_t T = 0xFFFFFFFEu;                                           // keep every real key
    if (n_valid > k) T = block_kth_bits(in, n, k, hist, sh);            // sentinels sort last
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const uint64_t key = in[i];
        if (key != KEY_SENTINEL && (uint32_t)(key >> 32) <= T) tmp[atomicAdd(&cnt[1], 1)] = key;
    }
    __syncthreads();
    const int m = cnt[1];
    for (int i = tid; i < k; i += NT) out[i] = KEY_SENTINEL;
    __syncthreads();
    for (int i = tid; i < m; i += NT) {
        const uint64_t key = tmp[i];
        int rank = 0;
        for (int j = 0; j < m; ++j) rank += tmp[j] < key;
        if (rank < k) out[rank] = key;
    }
    __syncthreads();
    return min(m, k);
}

}  // namespace vdbk
