// common.cuh -- error reporting, device staging and small launch helpers shared
// by the translation units of libmdagpu.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mdagpu.h"
#include "geom.cuh"

namespace mdagpu {

// thread-local last-error text, returned by mdagpu_last_error()
char *err_buf();
int set_error(int code, const char *fmt, ...);

#define MDA_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return mdagpu::set_error(MDAGPU_ERR_CUDA, "%s failed at %s:%d: %s", #call,    \
                                     __FILE__, __LINE__, cudaGetErrorString(e_));         \
    } while (0)

#define MDA_CHECK_LAUNCH() MDA_CUDA(cudaGetLastError())

#define MDA_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) return mdagpu::set_error(MDAGPU_ERR_INVALID, __VA_ARGS__);           \
    } while (0)

static inline int num_sms()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// The library's scratch comes from the device's default stream-ordered pool.  By default
// that pool hands everything back to the driver at the next synchronisation, which makes
// every call pay for fresh mappings (tens of ms for the GiB-sized cell lists); keep up to
// MDAGPU_POOL_RETAIN_MB (default: a quarter of the device memory) cached instead.
void pool_init();

// Stream-ordered scratch allocation that is released when the object dies.
struct DevBuf {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t n, cudaStream_t st)
    {
        release();
        pool_init();
        s = st;
        bytes = n;
        if (n == 0) n = 16;
        return cudaMallocAsync(&p, n, st);
    }
    void release()
    {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    template <class T> T *as() const { return (T *)p; }
};

// Transfers of caller-owned host memory (hostcopy.cu): pageable memory goes through a pinned
// ring with the host-side copies on a small thread pool, pinned memory straight to the DMA.
// host_to_device returns once `host` has been read; device_to_host once `host` holds the data.
bool host_is_pinned(const void *p);
cudaError_t host_to_device(void *dev, const void *host, size_t bytes, cudaStream_t st);
cudaError_t device_to_host(void *host, const void *dev, size_t bytes, cudaStream_t st);

// Input that may live on the host (space == MDAGPU_HOST: staged through a
// stream-ordered device buffer) or already on the device (used in place).
struct StagedIn {
    DevBuf buf;
    const void *dev = nullptr;
    cudaError_t init(const void *src, size_t bytes, int space, cudaStream_t st)
    {
        if (space == MDAGPU_DEVICE || src == nullptr) { dev = src; return cudaSuccess; }
        cudaError_t e = buf.alloc(bytes, st);
        if (e != cudaSuccess) return e;
        dev = buf.p;
        if (bytes == 0) return cudaSuccess;
        return host_to_device(buf.p, src, bytes, st);
    }
    template <class T> const T *as() const { return (const T *)dev; }
};

// Output (or in/out) buffer: device pointer used in place, host pointer gets a
// device twin that is copied back by finish().
struct StagedOut {
    DevBuf buf;
    void *dev = nullptr;
    void *host = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;
    cudaError_t init(void *dst, size_t n, int space, cudaStream_t st, bool copy_in = false)
    {
        s = st;
        bytes = n;
        if (space == MDAGPU_DEVICE || dst == nullptr) { dev = dst; return cudaSuccess; }
        host = dst;
        cudaError_t e = buf.alloc(n, st);
        if (e != cudaSuccess) return e;
        dev = buf.p;
        if (copy_in && n) return host_to_device(dev, host, n, st);
        return cudaSuccess;
    }
    cudaError_t finish()
    {
        if (!host || !bytes) return cudaSuccess;
        return device_to_host(host, dev, bytes, s);
    }
    template <class T> T *as() const { return (T *)dev; }
};

// launch counter (mdagpu_launch_count): every kernel launch of this library
// goes through MDA_LAUNCH so bench.py can report gpu_launches truthfully.
extern unsigned long long g_launches;
// Optional per-kernel timing (mdagpu_profile_enable): CUDA events are recorded on
// the launching stream around every launch and summed per name by
// mdagpu_profile_report().  Off by default; costs nothing when off.
extern int g_profile;
void profile_begin(const char *name, cudaStream_t st);
void profile_end(cudaStream_t st);
#define MDA_LAUNCH_AS(name, kernel, grid, block, smem, stream, ...)                       \
    do {                                                                                  \
        if (mdagpu::g_profile) mdagpu::profile_begin(name, stream);                       \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                       \
        if (mdagpu::g_profile) mdagpu::profile_end(stream);                               \
        ++mdagpu::g_launches;                                                             \
    } while (0)
#define MDA_LAUNCH(kernel, grid, block, smem, stream, ...)                                \
    MDA_LAUNCH_AS(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)

// streaming (evict-first) stores for write-once outputs
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream2(double2 *p, double2 v) { __stcs(p, v); }

}  // namespace mdagpu
