// common.cuh -- shared device helpers: sortable (distance,label) keys, warp reductions,
// mbarrier / bulk-copy PTX wrappers for sm_100a.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vdbk {

constexpr uint64_t KEY_SENTINEL = 0xFFFFFFFFFFFFFFFFull;   // sorts after every real key
constexpr uint32_t LABEL_MAX = 0xFFFFFFFEu;

// fp32 -> uint32 whose unsigned order equals the float order (-inf < ... < +inf < NaN(+)).
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}
// key: ascending key order == ascending (distance, label)
__host__ __device__ __forceinline__ uint64_t make_key(float dist, uint32_t label) {
    return (uint64_t(float_to_ordered(dist)) << 32) | label;
}
__host__ __device__ __forceinline__ float key_dist(uint64_t key) { return ordered_to_float(uint32_t(key >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_label(uint64_t key) { return uint32_t(key); }

#ifdef __CUDACC__
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Programmatic dependent launch: every kernel of the search chain starts with this.  launch_dependents lets the
// NEXT kernel of the stream be set up (and its CTAs placed where resources allow) while this one runs; wait
// blocks until the PREVIOUS kernel has completed and its writes are visible.  Placed before the first global
// access, so the chain behaves exactly like plain stream order minus the launch gaps.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Canonical cross-lane sum: butterfly xor 16,8,4,2,1.  Every kernel that reports a distance
// uses this tree on per-lane partials built in the same element order, so the scan kernel and
// the re-rank kernel return bit-identical distances for the same (query,row).
__device__ __forceinline__ float warp_sum_butterfly(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// V values per lane (V = power of two <= 32) reduced with the SAME pairing as the butterfly but
// V-1 + log2(32/V) shuffles instead of 5V.  On return a[0] in lane L is the total of value
// index  value_index_of_lane<V>(L).
template <int V>
__device__ __forceinline__ void warp_sum_multi(float (&a)[V]) {
    const int lane = lane_id();
    int o = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1, o >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            float send = up ? a[i] : a[i + n / 2];
            float keep = up ? a[i + n / 2] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
#pragma unroll
    for (; o > 0; o >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], o);
}
template <int V>
__device__ __forceinline__ int value_index_of_lane(int lane) {
    int idx = 0, o = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1, o >>= 1) idx += (lane & o) ? n / 2 : 0;
    return idx;
}
// lowest lane that holds value index v after warp_sum_multi<V>
template <int V>
__device__ __forceinline__ int lane_of_value_index(int v) {
    int lane = 0, o = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1, o >>= 1) if (v & (n / 2)) lane |= o;
    return lane;
}

// lane-strided sum of squares in the canonical order (16-byte lane chunks, then butterfly)
__device__ __forceinline__ float warp_sumsq_f32(const float* x, int dim) {
    const int lane = lane_id();
    float a = 0.0f;
    for (int c = lane * 4; c < dim; c += 128) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (c + e < dim) a = fmaf(x[c + e], x[c + e], a);
    }
    return warp_sum_butterfly(a);
}

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- shared-memory addressing / mbarrier / bulk async copy (TMA 1-D) -----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy (SASS UBLKCP), completion counted in bytes on `bar`.
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same with an L2 evict-first policy: a one-pass stream must not wash the L2
__device__ __forceinline__ void bulk_copy_g2s_stream(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                                     uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif  // __CUDACC__

#ifdef __CUDACC__
// Host: launch `kern` with programmatic stream serialisation allowed (VDB_PDL=0 falls back to a plain launch).
bool pdl_enabled();
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}
#endif

}  // namespace vdbk
