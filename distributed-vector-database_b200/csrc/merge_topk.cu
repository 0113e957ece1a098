// merge_topk.cu -- K5: per-query merge of candidate lists into one ascending top-k.
//
// Replaces `sorted(range(len(all_scores)), key=...)[:top_k]` of CoordinatorHandler.search
// (src/coordinator/handler.py:212-216) and is also the cross-CTA reduction after K1/K2.
// Keys are (ordered fp32 distance << 32 | label), so one unsigned 64-bit bitonic sort gives
// ascending (distance, label).  One CTA per query; inputs longer than the 2048-key shared
// buffer are streamed through it, keeping the best k_out after each round.
#include "common.cuh"
#include "kernels.h"

namespace vdbk {

constexpr int MERGE_THREADS = 512;
constexpr int MERGE_BUF = 2048;

__device__ __forceinline__ void bitonic_sort_smem(uint64_t* buf, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += MERGE_THREADS) {
                const int i = 2 * t - (t & (stride - 1));
                const int j = i + stride;
                const bool asc = (i & size) == 0;
                const uint64_t a = buf[i], b = buf[j];
                if ((a > b) == asc) {
                    buf[i] = b;
                    buf[j] = a;
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ uint64_t merge_load(const MergeParams& p, size_t q, int idx) {
    if (p.in_keys) return p.in_keys[q * (size_t)p.n_in + idx];
    const int g = idx / p.k_in, j = idx - g * p.k_in;
    const size_t off = ((size_t)g * p.nq + q) * (size_t)p.k_in + j;
    const int64_t id = p.in_ids[off];
    if (id < 0) return KEY_SENTINEL;
    return make_key(p.in_dist[off], (uint32_t)id);
}

__global__ void __launch_bounds__(MERGE_THREADS) merge_topk_kernel(const MergeParams p) {
    __shared__ uint64_t buf[MERGE_BUF];
    const size_t q = blockIdx.x;
    const int total = p.n_in;
    const int K = p.k_out;
    int n = 2;
    while (n < total && n < MERGE_BUF) n <<= 1;
    while (n < 2 * K && n < MERGE_BUF) n <<= 1;   // room to keep K and still take new keys

    int consumed = min(total, n);
    for (int i = threadIdx.x; i < n; i += MERGE_THREADS) buf[i] = i < consumed ? merge_load(p, q, i) : KEY_SENTINEL;
    __syncthreads();
    bitonic_sort_smem(buf, n);
    while (consumed < total) {
        const int take = min(total - consumed, n - K);
        for (int i = threadIdx.x; i < n - K; i += MERGE_THREADS)
            buf[K + i] = i < take ? merge_load(p, q, consumed + i) : KEY_SENTINEL;
        consumed += take;
        __syncthreads();
        bitonic_sort_smem(buf, n);
    }
    int cnt = 0;
    for (int i = threadIdx.x; i < K; i += MERGE_THREADS) {
        const uint64_t key = i < n ? buf[i] : KEY_SENTINEL;
        const bool real = key != KEY_SENTINEL;
        cnt += real ? 1 : 0;
        if (p.out_keys) p.out_keys[q * K + i] = key;
        if (p.out_ids) p.out_ids[q * K + i] = real ? (int64_t)key_label(key) : -1;
        if (p.out_dist) p.out_dist[q * K + i] = real ? key_dist(key) : __int_as_float(0x7f800000);
    }
    if (p.out_counts) {
        __shared__ int total_cnt;
        if (threadIdx.x == 0) total_cnt = 0;
        __syncthreads();
        cnt = warp_sum_int(cnt);
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&total_cnt, cnt);
        __syncthreads();
        if (threadIdx.x == 0) p.out_counts[q] = total_cnt;
    }
}

cudaError_t launch_merge_topk(const MergeParams& p, cudaStream_t st) {
    if (p.nq == 0) return cudaSuccess;
    if (p.k_out < 1 || p.k_out > MERGE_BUF / 2) return cudaErrorInvalidValue;
    merge_topk_kernel<<<(unsigned)p.nq, MERGE_THREADS, 0, st>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace vdbk
