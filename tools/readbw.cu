// readbw.cu -- read-only HBM bandwidth micro-benchmark (what can a pure streaming read reach on this B200?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o readbw readbw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ p, size_t n16, unsigned long long* out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// bulk-copy streaming: one producer thread per CTA, STAGES x BYTES ring, consumers only release
template <int STAGES>
__global__ void __launch_bounds__(128) bulk_kernel(const uint8_t* __restrict__ p, size_t bytes, uint32_t chunk) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = bytes / chunk;
    if (threadIdx.x == 0) {
        uint32_t it = 0;
        // issue STAGES ahead, wait in order (no consumer: data is dropped)
        size_t c = blockIdx.x;
        size_t issued = 0, waited = 0;
        size_t mine = (nchunks > blockIdx.x) ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        while (waited < mine) {
            while (issued < mine && issued - waited < STAGES) {
                const int s = issued % STAGES;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(smem + (size_t)s * chunk)),
                             "l"(p + c * (size_t)chunk), "r"(chunk), "r"(smem_u32(&full[s]))
                             : "memory");
                c += gridDim.x;
                ++issued;
            }
            const int s = waited % STAGES;
            const uint32_t ph = (waited / STAGES) & 1;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&full[s])), "r"(ph) : "memory");
            ++waited;
        }
        (void)it;
    }
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    float best = 1e9;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    const size_t bytes = 2048000000ull;
    uint8_t* d; unsigned long long* out;
    cudaMalloc(&d, bytes); cudaMalloc(&out, 8);
    cudaMemset(d, 1, bytes);
    const size_t n16 = bytes / 16;
    for (int grid : {148, 296, 592, 1184, 2368}) {
        float ms = time_ms([&] { ldg_kernel<8><<<grid, 512>>>((const uint4*)d, n16, out); });
        printf("ldg u8  grid %5d x512: %.1f us  %.0f GB/s\n", grid, ms * 1e3, bytes / ms / 1e6);
        ms = time_ms([&] { ldg_kernel<4><<<grid, 512>>>((const uint4*)d, n16, out); });
        printf("ldg u4  grid %5d x512: %.1f us  %.0f GB/s\n", grid, ms * 1e3, bytes / ms / 1e6);
    }
    for (uint32_t chunk : {8192u, 16384u, 32768u, 65536u}) {
        for (int stages : {2, 3, 4, 6}) {
            const size_t smem = (size_t)stages * chunk;
            if (smem > 220 * 1024) continue;
            for (int cps : {1, 2}) {
                if (smem * cps > 220 * 1024) continue;
                float ms = 0;
                auto run = [&](auto kern) {
                    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    ms = time_ms([&] { kern<<<148 * cps, 128, smem>>>(d, bytes, chunk); });
                };
                if (stages == 2) run(bulk_kernel<2>);
                if (stages == 3) run(bulk_kernel<3>);
                if (stages == 4) run(bulk_kernel<4>);
                if (stages == 6) run(bulk_kernel<6>);
                printf("bulk chunk %6u stages %d ctas/sm %d: %.1f us  %.0f GB/s\n", chunk, stages, cps, ms * 1e3, bytes / ms / 1e6);
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
