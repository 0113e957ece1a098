// merge_topk.cu -- K5: per-query merge of candidate lists into one ascending top-k.
//
// Replaces `sorted(range(len(all_scores)), key=...)[:top_k]` of CoordinatorHandler.search
// (src/coordinator/handler.py:212-216) and is also the cross-CTA reduction after K1/K2.
// Keys are (ordered fp32 distance << 32 | label), so one unsigned 64-bit bitonic sort gives
// ascending (distance, label).  One CTA per query; inputs longer than the 2048-key shared
// buffer are streamed through it, keeping the best k_out after each round.
#include <algorithm>

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
    if (p.in_keys && p.G > 0) {   // packed keys laid out [G][nq][k_in] (the exchange buffer): L1-bypassing loads,
        const int g = idx / p.k_in, j = idx - g * p.k_in;          // the lines were written by peer GPUs
        return __ldcv(p.in_keys + ((size_t)g * p.key_stride_g + q * (size_t)p.k_in + j));
    }
    if (p.in_keys) return p.in_keys[q * (size_t)p.n_in + idx];
    const int g = idx / p.k_in, j = idx - g * p.k_in;
    const size_t off = ((size_t)g * p.nq + q) * (size_t)p.k_in + j;
    const int64_t id = p.in_ids[off];
    if (id < 0) return KEY_SENTINEL;
    return make_key(p.in_dist[off], (uint32_t)id);
}

// write K sorted keys from buf[0..) to the outputs of query q
__device__ __forceinline__ void merge_emit(const MergeParams& p, size_t q, const uint64_t* buf, int n_sorted) {
    const int K = p.k_out;
    int cnt = 0;
    for (int i = threadIdx.x; i < K; i += MERGE_THREADS) {
        const uint64_t key = i < n_sorted ? buf[i] : KEY_SENTINEL;
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

// streaming bitonic path: any input length, 2048-key window
__device__ __forceinline__ void merge_stream(const MergeParams& p, size_t q, uint64_t* buf) {
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
    merge_emit(p, q, buf, n);
}

__global__ void __launch_bounds__(MERGE_THREADS) merge_topk_kernel(const MergeParams p) {
    pdl_prologue();
    __shared__ uint64_t buf[MERGE_BUF];
    merge_stream(p, blockIdx.x, buf);
}

// Long inputs (the 148 per-CTA lists of the scan kernel): radix select on the 32 distance bits finds the
// k-th distance in 4 passes over the keys, only keys at or below it are sorted.  Falls back to the streaming
// path when exact distance ties make more than MERGE_BUF keys qualify.
__global__ void __launch_bounds__(MERGE_THREADS) merge_select_kernel(const MergeParams p) {
    pdl_prologue();
    __shared__ uint64_t buf[MERGE_BUF];
    __shared__ int hist[256];
    __shared__ uint32_t s_prefix, s_rank;
    __shared__ int s_valid, s_m;
    const size_t q = blockIdx.x;
    const int n = p.n_in, K = p.k_out;
    if (threadIdx.x == 0) { s_valid = 0; s_m = 0; s_prefix = 0; s_rank = (uint32_t)K; }
    __syncthreads();
    int v = 0;
    for (int i = threadIdx.x; i < n; i += MERGE_THREADS) v += merge_load(p, q, i) != KEY_SENTINEL;
    v = warp_sum_int(v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_valid, v);
    __syncthreads();
    uint32_t T = 0xFFFFFFFFu;          // keep everything valid
    if (s_valid > K) {
        uint32_t mask = 0;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            if (threadIdx.x < 256) hist[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = threadIdx.x; i < n; i += MERGE_THREADS) {
                const uint64_t key = merge_load(p, q, i);
                if (key != KEY_SENTINEL) {
                    const uint32_t hi = (uint32_t)(key >> 32);
                    if ((hi & mask) == prefix) atomicAdd(&hist[(hi >> shift) & 255], 1);
                }
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                int local[8], sum = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) { local[i] = hist[threadIdx.x * 8 + i]; sum += local[i]; }
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)threadIdx.x >= o) incl += t;
                }
                const int excl = incl - sum, rank = (int)s_rank;
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
        T = s_prefix;
    }
    // gather keys whose distance bits are <= T
    for (int i = threadIdx.x; i < n; i += MERGE_THREADS) {
        const uint64_t key = merge_load(p, q, i);
        if (key != KEY_SENTINEL && (uint32_t)(key >> 32) <= T) {
            const int s = atomicAdd(&s_m, 1);
            if (s < MERGE_BUF) buf[s] = key;
        }
    }
    __syncthreads();
    const int m = s_m;
    if (m > MERGE_BUF) {               // a flood of exact ties: take the slow, always-correct road
        __syncthreads();
        merge_stream(p, q, buf);
        return;
    }
    int np = 2;
    while (np < m) np <<= 1;
    for (int i = m + threadIdx.x; i < np; i += MERGE_THREADS) buf[i] = KEY_SENTINEL;
    __syncthreads();
    bitonic_sort_smem(buf, np);
    merge_emit(p, q, buf, np);
}

// ------------------------------------------------------------------------------------------
// K5x: cross-GPU exchange fused with the merge, over NVLink peer memory (no NCCL call).
//
// GPU form of CoordinatorHandler.search's gather + merge (src/coordinator/handler.py:191-216) for G ranks of
// one box.  Every rank has searched its shard for the whole batch; rank r owns the final answer of query slice
// r.  One launch per rank:
//   phase 1  pack (distance, id) into sortable keys and STORE each query's list straight into the owner's
//            receive buffer (peer pointers opened with CUDA IPC; coalesced 8-byte stores over NVLink)
//   phase 2  the last CTA to finish phase 1 publishes (shape, step) in every peer's flag word: one 8-byte store
//            per peer after a system-scope fence
//   phase 3  wait until all G flags carry this step, then merge the G lists of each owned query
// Receive buffers are double-buffered by step parity: a rank can only be two steps ahead of a peer after that
// peer has pushed the step in between, which it does after finishing its merge of the older step.
//
// Liveness.  Every CTA waits in phase 3 while this rank's own flag is only published once ALL its CTAs have
// finished phase 1: a CTA queued behind waiting ones would deadlock two ranks against each other.  The launch
// is therefore COOPERATIVE (the driver places the whole grid at once or not at all, whatever else runs on the
// device) and the grid is sized from the occupancy calculator.  The wait itself is bounded: a peer that never
// arrives (it raised before its launch, or died) or arrives with another (nq, k) makes every CTA give up,
// pad its outputs and record the failure in a host-visible word that vdb_xchg_status / the next merge report.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(MERGE_THREADS, 2) exchange_merge_kernel(const XchgParams x, const MergeParams mp) {
    __shared__ uint64_t buf[MERGE_BUF];
    __shared__ int s_last, s_fail;
    const size_t total = x.nq * (size_t)x.k;
    if (threadIdx.x == 0) s_fail = 0;
    for (size_t i = blockIdx.x * (size_t)MERGE_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * MERGE_THREADS) {
        const size_t q = i / x.k;
        const int j = (int)(i - q * x.k);
        const int dst = (int)(q / x.slice);
        const size_t ql = q - (size_t)dst * x.slice;
        const int64_t id = x.ids[i];
        const uint64_t key = id < 0 ? KEY_SENTINEL : make_key(x.dist[i], (uint32_t)id);
        x.peer_buf[dst][(size_t)x.parity * x.stride_parity + (size_t)x.rank * x.stride_src + ql * x.k + j] = key;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(x.done_counter, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    const uint64_t my_word = ((uint64_t)x.shape << 32) | x.step;
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < x.world)
            *reinterpret_cast<volatile uint64_t*>(x.peer_flag[threadIdx.x] + x.parity * x.world + x.rank) = my_word;
    }
    if (threadIdx.x < x.world) {
        const volatile uint64_t* f = x.local_flag + x.parity * x.world + threadIdx.x;
        const uint64_t t0 = global_timer_ns();
        uint64_t w;
        int fail = 0;
        while ((uint32_t)(w = *f) < x.step) {       // steps only grow; a peer already past this step is a protocol error
            __nanosleep(200);
            if (global_timer_ns() - t0 > x.timeout_ns) { fail = XCHG_ERR_TIMEOUT; break; }
        }
        if (!fail && w != my_word) fail = (uint32_t)w != x.step ? XCHG_ERR_STEP : XCHG_ERR_SHAPE;
        if (fail) {
            s_fail = fail;
            // host-visible record: code, step, peer (first writer wins; all CTAs report the same kind of failure)
            if (atomicCAS(reinterpret_cast<unsigned int*>(x.err), 0u, (unsigned int)fail) == 0u) {
                x.err[1] = x.step;
                x.err[2] = (uint32_t)threadIdx.x;
                __threadfence_system();
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (s_fail) {                                     // no usable input: pad this rank's outputs
        for (size_t i = blockIdx.x * (size_t)MERGE_THREADS + threadIdx.x; i < x.owned * (size_t)mp.k_out;
             i += (size_t)gridDim.x * MERGE_THREADS) {
            if (mp.out_ids) mp.out_ids[i] = -1;
            if (mp.out_dist) mp.out_dist[i] = __int_as_float(0x7f800000);
        }
        return;
    }
    for (size_t ql = blockIdx.x; ql < x.owned; ql += gridDim.x) {   // queries this rank owns (ragged last slice)
        merge_stream(mp, ql, buf);
        __syncthreads();
    }
}

cudaError_t launch_exchange_merge(const XchgParams& x, const MergeParams& mp, int num_sms, cudaStream_t st) {
    if (mp.k_out < 1 || mp.k_out > MERGE_BUF / 2) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(x.done_counter, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    // co-resident grid: what the occupancy calculator says fits, launched cooperatively (see the kernel's header)
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_merge_kernel, MERGE_THREADS, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const size_t resident = (size_t)num_sms * (size_t)std::min(per_sm, 2);
    const size_t useful = std::max(x.slice, (x.nq * x.k + MERGE_THREADS - 1) / MERGE_THREADS);
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min(resident, useful));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(MERGE_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, exchange_merge_kernel, x, mp);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Query all-gather over the copy engines: every rank DMAs its slice of the batch into every peer's query slot
// (cudaMemcpyAsync over NVLink, no SM), then this one-warp kernel publishes "slice of batch `value` has landed" in
// every rank's arrival array; the consumer's one-warp wait kernel sits in front of its search.  Neither kernel needs
// more than a warp slot, and the signal waits for nothing: unlike a collective kernel on a second stream it cannot be
// kept off the SMs by -- or keep off the SMs -- the cooperative exchange kernel.
// ------------------------------------------------------------------------------------------
__global__ void xchg_signal_kernel(const XchgSignal s) {
    __threadfence_system();
    if (threadIdx.x < s.world) *reinterpret_cast<volatile uint32_t*>(s.flag[threadIdx.x]) = s.value;
}

__global__ void xchg_wait_kernel(const XchgWait w) {
    if (threadIdx.x < w.world) {
        const volatile uint32_t* f = w.flag + threadIdx.x;
        const uint64_t t0 = global_timer_ns();
        while (*f < w.value) {
            __nanosleep(200);
            if (global_timer_ns() - t0 > w.timeout_ns) {
                if (atomicCAS(reinterpret_cast<unsigned int*>(w.err), 0u, (unsigned int)XCHG_ERR_TIMEOUT) == 0u) {
                    w.err[1] = w.value;
                    w.err[2] = (uint32_t)threadIdx.x;
                    __threadfence_system();
                }
                break;
            }
        }
    }
    __threadfence_system();
}

cudaError_t launch_xchg_signal(const XchgSignal& s, cudaStream_t st) {
    xchg_signal_kernel<<<1, 32, 0, st>>>(s);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_xchg_wait(const XchgWait& w, cudaStream_t st) {
    xchg_wait_kernel<<<1, 32, 0, st>>>(w);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_merge_topk(const MergeParams& p, cudaStream_t st) {
    if (p.nq == 0) return cudaSuccess;
    if (p.k_out < 1 || p.k_out > MERGE_BUF / 2) return cudaErrorInvalidValue;
    cudaError_t e;
    if (p.n_in >= 512 && p.n_in >= 8 * p.k_out) e = launch_pdl(merge_select_kernel, dim3((unsigned)p.nq), dim3(MERGE_THREADS), 0, st, p);
    else e = launch_pdl(merge_topk_kernel, dim3((unsigned)p.nq), dim3(MERGE_THREADS), 0, st, p);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace vdbk
