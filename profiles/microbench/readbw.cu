// readbw.cu -- what read-only HBM bandwidth can one B200 sustain?  (The roofline
// denominator in MEASURED_PEAKS.json is a copy: half reads, half writes.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o readbw readbw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void __launch_bounds__(256) read_sum(const float4 *__restrict__ p, size_t n4, float *out)
{
    float acc = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x * UNROLL;
    for (; i + (UNROLL - 1) * 256 < n4; i += stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(p + i + u * 256);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 1.2345f) out[0] = acc;
}

__global__ void __launch_bounds__(256) copy_k(const float4 *__restrict__ p, float4 *__restrict__ q, size_t n4)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) q[i] = p[i];
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// persistent CTAs, one thread issues cp.async.bulk of CHUNK bytes into a ring of S stages
template <int CHUNK, int S>
__global__ void __launch_bounds__(128) tma_read(const unsigned char *p, size_t bytes, float *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[S];
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = bytes / CHUNK;
    const size_t mine = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](size_t k) {
        const int s = (int)(k % S);
        const size_t c = blockIdx.x + k * gridDim.x;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(CHUNK) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(sm + (size_t)s * CHUNK)), "l"(p + c * CHUNK), "r"(CHUNK), "r"(s32(&full[s])) : "memory");
    };
    if (threadIdx.x == 0)
        for (size_t k = 0; k < S && k < mine; ++k) issue(k);
    float acc = 0.f;
    for (size_t k = 0; k < mine; ++k) {
        const int s = (int)(k % S);
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(s32(&full[s])), "r"((uint32_t)((k / S) & 1)) : "memory");
        acc += reinterpret_cast<const float *>(sm + (size_t)s * CHUNK)[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && k + S < mine) issue(k + S);
    }
    if (acc == 1.2345f) out[0] = acc;
}

template <class F> float time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    const size_t bytes = (size_t)6 << 30;
    unsigned char *p, *q;
    float *out;
    cudaMalloc(&p, bytes); cudaMalloc(&q, bytes); cudaMalloc(&out, 4);
    cudaMemset(p, 1, bytes);
    const size_t n4 = bytes / 16;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float ms;
    ms = time_ms([&] { copy_k<<<sms * 16, 256>>>((const float4 *)p, (float4 *)q, n4); });
    printf("copy  (r+w)            %8.3f ms  %8.1f GB/s\n", ms, 2.0 * bytes / ms / 1e6);
    ms = time_ms([&] { cudaMemcpyAsync(q, p, bytes, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D (r+w)   %8.3f ms  %8.1f GB/s\n", ms, 2.0 * bytes / ms / 1e6);
    ms = time_ms([&] { read_sum<4><<<sms * 8, 256>>>((const float4 *)p, n4, out); });
    printf("read  unroll4 8CTA/SM  %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = time_ms([&] { read_sum<8><<<sms * 8, 256>>>((const float4 *)p, n4, out); });
    printf("read  unroll8 8CTA/SM  %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = time_ms([&] { read_sum<8><<<sms * 4, 256>>>((const float4 *)p, n4, out); });
    printf("read  unroll8 4CTA/SM  %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    ms = time_ms([&] { read_sum<16><<<sms * 4, 256>>>((const float4 *)p, n4, out); });
    printf("read  unroll16 4CTA/SM %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    {
        auto k = tma_read<16384, 8>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
        ms = time_ms([&] { k<<<sms, 128, 16384 * 8>>>(p, bytes, out); });
        printf("TMA 16K x8 1CTA/SM     %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    }
    {
        auto k = tma_read<49152, 4>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 * 4);
        ms = time_ms([&] { k<<<sms, 128, 49152 * 4>>>(p, bytes, out); });
        printf("TMA 48K x4 1CTA/SM     %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    }
    {
        auto k = tma_read<8192, 6>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 6);
        ms = time_ms([&] { k<<<sms * 4, 128, 8192 * 6>>>(p, bytes, out); });
        printf("TMA 8K x6 4CTA/SM      %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    }
    {
        auto k = tma_read<32768, 3>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 3);
        ms = time_ms([&] { k<<<sms * 2, 128, 32768 * 3>>>(p, bytes, out); });
        printf("TMA 32K x3 2CTA/SM     %8.3f ms  %8.1f GB/s\n", ms, bytes / ms / 1e6);
    }
    return 0;
}
