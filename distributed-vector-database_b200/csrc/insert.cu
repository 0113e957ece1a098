// insert.cu -- K0 query preparation, K3 insert (normalise / convert / norms), synthetic rows,
// tombstone bitmap updates, row gather.
//
// K3 replaces hnswlib.Index.add_items (reference call sites src/datanode/handler.py:112,
// 268-271): for space 'cosine' hnswlib stores x * 1/(sqrtf(sum x^2) + 1e-30f); ||d||^2 of the
// row AS STORED is kept beside it for the L2 expansion used by the tensor-core kernel.
#include "common.cuh"
#include "kernels.h"
#include "synth.cuh"

namespace vdbk {

__global__ void prepare_queries_kernel(const float* __restrict__ q, size_t nq, int dim, int ld, bool normalize,
                                       float* __restrict__ out, float* __restrict__ qn2, __half* __restrict__ out16,
                                       int ld16, int* __restrict__ zero_per_query, int* __restrict__ zero_one) {
    pdl_prologue();
    const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    if (w >= nq) return;
    if (lane == 0) {      // per-batch flags of the tensor path, cleared here instead of by separate memsets
        if (zero_per_query) zero_per_query[w] = 0;
        if (zero_one && w == 0) *zero_one = 0;
    }
    const float* src = q + w * (size_t)dim;
    float scale = 1.0f;
    if (normalize) {
        const float s = warp_sumsq_f32(src, dim);
        scale = 1.0f / (sqrtf(s) + 1e-30f);
    }
    float a = 0.0f;
    // 128-bit path (dim a multiple of 4, 16-byte aligned batch: every CLIP-shaped query block); same arithmetic, same
    // order as the scalar path below
    const bool vec = (dim & 3) == 0 && (ld16 & 3) == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(out) |
                                                         reinterpret_cast<uintptr_t>(out16)) & 15) == 0;
    if (vec) {
        for (int c = lane * 4; c < ld; c += 128) {
            float4 x = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (c < dim) x = *reinterpret_cast<const float4*>(src + c);
            const float v[4] = {x.x * scale, x.y * scale, x.z * scale, x.w * scale};
            *reinterpret_cast<float4*>(out + w * (size_t)ld + c) = make_float4(v[0], v[1], v[2], v[3]);
            if (out16 && c < ld16) {
                __half2 h[2] = {__floats2half2_rn(v[0], v[1]), __floats2half2_rn(v[2], v[3])};
                *reinterpret_cast<uint2*>(out16 + w * (size_t)ld16 + c) = *reinterpret_cast<uint2*>(h);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) a = fmaf(v[e], v[e], a);
        }
    } else {
        for (int c = lane * 4; c < ld; c += 128) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float v = (c + e < dim) ? src[c + e] * scale : 0.0f;
                out[w * (size_t)ld + c + e] = v;
                if (out16 && c + e < ld16) out16[w * (size_t)ld16 + c + e] = __float2half_rn(v);   // fp16 operand copy
                a = fmaf(v, v, a);
            }
        }
    }
    a = warp_sum_butterfly(a);
    if (lane == 0 && qn2) qn2[w] = a;
}

template <typename T>
__global__ void insert_rows_kernel(const float* __restrict__ src, size_t n, int dim, int ld, bool normalize,
                                   T* __restrict__ rows, float* __restrict__ sqnorm, size_t row0,
                                   unsigned int* __restrict__ max_bits) {
    const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    if (w >= n) return;
    const float* x = src + w * (size_t)dim;
    float scale = 1.0f;
    if (normalize) {
        const float s = warp_sumsq_f32(x, dim);
        scale = 1.0f / (sqrtf(s) + 1e-30f);
    }
    T* dst = rows + (row0 + w) * (size_t)ld;
    float a = 0.0f;
    for (int c = lane * 4; c < ld; c += 128) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v = (c + e < dim) ? x[c + e] * scale : 0.0f;
            if constexpr (sizeof(T) == 2) {
                const __half h = __float2half_rn(v);
                dst[c + e] = h;
                v = __half2float(h);
            } else {
                dst[c + e] = v;
            }
            a = fmaf(v, v, a);
        }
    }
    a = warp_sum_butterfly(a);
    if (lane == 0) {
        sqnorm[row0 + w] = a;
        if (max_bits) atomicMax(max_bits, __float_as_uint(a));   // a >= 0: uint order == float order
    }
}

// fp16 shadow of fp32 shard rows [row0, row0+n): the operand plane of the batched tensor-core search
// (kind::f16 runs at twice the tf32 rate; candidates are re-ranked from the fp32 rows, gemm_topk.cu).
__global__ void shadow_rows_kernel(const float* __restrict__ rows, int ld, __half* __restrict__ shadow, int ld16,
                                   size_t row0, size_t n) {
    const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    if (w >= n) return;
    const float* src = rows + (row0 + w) * (size_t)ld;
    __half2* dst = reinterpret_cast<__half2*>(shadow + (row0 + w) * (size_t)ld16);
    for (int c = lane * 2; c < ld16; c += 64) {      // ld >= ld16, padding columns are zero
        const float2 v = *reinterpret_cast<const float2*>(src + c);
        dst[c >> 1] = __floats2half2_rn(v.x, v.y);
    }
}

__global__ void synth_rows_kernel(uint64_t seed, uint64_t row_start, size_t n, int dim, float* __restrict__ out) {
    const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    if (w >= n) return;
    const uint64_t base = synth_row_base(seed, row_start + w);
    long long ss = 0;
    for (int c = lane; c < dim; c += 32) {
        const long long v = synth_int(base, c);
        ss += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);   // exact integer sum
    const double inv = sqrt((double)ss);
    for (int c = lane; c < dim; c += 32) out[w * (size_t)dim + c] = (float)((double)synth_int(base, c) / inv);
}

__global__ void set_bits_kernel(uint32_t* bitmap, const uint32_t* rows, size_t n, bool set) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = rows[i];
    if (set) atomicOr(&bitmap[r >> 5], 1u << (r & 31));
    else atomicAnd(&bitmap[r >> 5], ~(1u << (r & 31)));
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ rows, int ld, int dim, const uint32_t* __restrict__ idx,
                                   size_t n, float* __restrict__ out) {
    const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    if (w >= n) return;
    const T* r = rows + (size_t)idx[w] * ld;
    for (int c = lane; c < dim; c += 32) {
        if constexpr (sizeof(T) == 2) out[w * (size_t)dim + c] = __half2float(r[c]);
        else out[w * (size_t)dim + c] = r[c];
    }
}

__global__ void iota_u32_kernel(uint32_t* out, size_t n, uint32_t start) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = start + (uint32_t)i;
}

static inline unsigned warp_grid(size_t n_warps, int threads) {
    const size_t per = threads / 32;
    return (unsigned)((n_warps + per - 1) / per);
}

cudaError_t launch_prepare_queries(const float* q, size_t nq, int dim, int ld, bool normalize, float* out, float* qn2,
                                   cudaStream_t st, void* out16, int ld16, int* zero_per_query, int* zero_one) {
    if (!nq) return cudaSuccess;
    cudaError_t e = launch_pdl(prepare_queries_kernel, dim3(warp_grid(nq, 256)), dim3(256), 0, st, q, nq, dim, ld, normalize, out,
                               qn2, (__half*)out16, ld16, zero_per_query, zero_one);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_insert_rows(const float* src, size_t n, int dim, int ld, bool normalize, bool f16, void* rows,
                               float* sqnorm, size_t row0, unsigned int* max_sqnorm_bits, cudaStream_t st) {
    if (!n) return cudaSuccess;
    if (f16)
        insert_rows_kernel<__half><<<warp_grid(n, 256), 256, 0, st>>>(src, n, dim, ld, normalize, (__half*)rows,
                                                                      sqnorm, row0, max_sqnorm_bits);
    else
        insert_rows_kernel<float><<<warp_grid(n, 256), 256, 0, st>>>(src, n, dim, ld, normalize, (float*)rows, sqnorm,
                                                                     row0, max_sqnorm_bits);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_shadow_rows(const float* rows, int ld, void* shadow, int ld16, size_t row0, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    shadow_rows_kernel<<<warp_grid(n, 256), 256, 0, st>>>(rows, ld, (__half*)shadow, ld16, row0, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_synth_rows(uint64_t seed, uint64_t row_start, size_t n, int dim, float* out, cudaStream_t st) {
    if (!n) return cudaSuccess;
    synth_rows_kernel<<<warp_grid(n, 256), 256, 0, st>>>(seed, row_start, n, dim, out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_set_bits(uint32_t* bitmap, const uint32_t* rows, size_t n, bool set, cudaStream_t st) {
    if (!n) return cudaSuccess;
    set_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(bitmap, rows, n, set);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_gather_rows(const void* rows, int ld, int dim, bool f16, const uint32_t* idx, size_t n, float* out,
                               cudaStream_t st) {
    if (!n) return cudaSuccess;
    if (f16) gather_rows_kernel<__half><<<warp_grid(n, 256), 256, 0, st>>>((const __half*)rows, ld, dim, idx, n, out);
    else gather_rows_kernel<float><<<warp_grid(n, 256), 256, 0, st>>>((const float*)rows, ld, dim, idx, n, out);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_iota_u32(uint32_t* out, size_t n, uint32_t start, cudaStream_t st) {
    if (!n) return cudaSuccess;
    iota_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, n, start);
    count_launch();
    return cudaGetLastError();
}

}  // namespace vdbk
