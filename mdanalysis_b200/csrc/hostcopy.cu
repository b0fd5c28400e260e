// hostcopy.cu -- transfers between CALLER-OWNED host memory and the device at PCIe rate.
//
// The reference-facing calls receive ordinary NumPy arrays (lib/distances.py hands the backend
// `result` arrays from np.empty, `timeseries()` arrays, ...): pageable memory.  cudaMemcpyAsync
// on pageable memory is staged by the driver through one small internal buffer on the calling
// thread -- measured here at 20 GB/s device-to-host (89 MB distance matrix: 4.3 ms against 1.7 ms
// into pinned memory) and 7 GB/s host-to-device (400 MB of frames: 58 ms against 8 ms).
// Here the staging is done by the library: a ring of pinned chunks per device, the DMA of chunk
// k + 1 in flight while a small pool of host threads copies chunk k between the ring and the
// caller's pages (one core cannot keep up with a x16 Gen5 link; page faults of a fresh
// np.empty() result are taken by the same threads in parallel).
// Pinned / registered host memory is recognised (cudaPointerGetAttributes) and goes straight
// to cudaMemcpyAsync.
#include "common.cuh"

#include <sched.h>
#include <stdlib.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace mdagpu {

namespace {

constexpr size_t HC_CHUNK = 8u << 20;            // bytes per pinned chunk
constexpr int HC_RING = 4;                        // chunks per device
constexpr size_t HC_DIRECT_BELOW = 2u << 20;      // small transfers: the driver's own path

// Persistent host threads for the memcpy between the ring and the caller's memory.
struct CopyPool {
    struct Slice { char *dst; const char *src; size_t n; };
    std::vector<std::thread> threads;
    std::vector<Slice> slices;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    unsigned long long generation = 0;
    int pending = 0;
    int nthreads = 1;
    pid_t pid = 0;

    explicit CopyPool(int n) : nthreads(n), pid(getpid())
    {
        slices.resize(n);
        for (int t = 1; t < n; ++t) threads.emplace_back([this, t] { work(t); });
    }
    void work(int t)
    {
        unsigned long long seen = 0;
        for (;;) {
            Slice s;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return generation != seen; });
                seen = generation;
                s = slices[t];
            }
            if (s.n) memcpy(s.dst, s.src, s.n);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--pending == 0) cv_done.notify_one();
            }
        }
    }
    // dst[0:n] = src[0:n], split over the pool; the caller takes the first slice
    void copy(char *dst, const char *src, size_t n)
    {
        if (nthreads == 1 || n < (256u << 10)) { memcpy(dst, src, n); return; }
        const size_t per = ((n / nthreads) + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(m);
            size_t off = 0;
            for (int t = 0; t < nthreads; ++t) {
                const size_t len = off < n ? (n - off < per || t == nthreads - 1 ? n - off : per) : 0;
                slices[t] = { dst + off, src + off, len };
                off += len;
            }
            pending = nthreads - 1;
            ++generation;
        }
        cv_job.notify_all();
        if (slices[0].n) memcpy(slices[0].dst, slices[0].src, slices[0].n);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

int copy_threads()
{
    const char *env = getenv("MDAGPU_COPY_THREADS");
    if (env && *env) {
        const int n = atoi(env);
        return n < 1 ? 1 : (n > 32 ? 32 : n);
    }
    cpu_set_t set;
    int avail = 1;
    if (sched_getaffinity(0, sizeof set, &set) == 0) avail = CPU_COUNT(&set);
    // one process per GPU (torchrun): the ranks of a node share its cores
    const char *lws = getenv("LOCAL_WORLD_SIZE");
    if (lws && atoi(lws) > 1) avail = avail / atoi(lws) > 1 ? avail / atoi(lws) : 1;
    // measured on the B200 box (16 cores, profiles/microbench/hostcopy_bench.txt): one thread moves
    // 9 GB/s, 8 threads 35-40 GB/s, 12 threads 41 GB/s of the 52 GB/s a pinned transfer reaches
    int n = avail - avail / 4;
    return n < 1 ? 1 : (n > 12 ? 12 : n);
}

CopyPool *pool()
{
    // never destroyed: its threads must not be joined at exit, and after fork() the child gets
    // a fresh pool (the parent's threads do not exist there)
    static CopyPool *p = nullptr;
    static std::mutex m;
    std::lock_guard<std::mutex> lk(m);
    if (!p || p->pid != getpid()) p = new CopyPool(copy_threads());
    return p;
}

struct Ring {
    char *chunk[HC_RING] = { nullptr };
    cudaEvent_t ev[HC_RING] = { nullptr };
    bool busy[HC_RING] = { false };
    std::mutex m;                                    // one staged transfer per device at a time
    bool ready = false;
    cudaError_t init()
    {
        if (ready) return cudaSuccess;
        for (int k = 0; k < HC_RING; ++k) {
            cudaError_t e = cudaHostAlloc((void **)&chunk[k], HC_CHUNK, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        ready = true;
        return cudaSuccess;
    }
    cudaError_t wait(int k)
    {
        if (!busy[k]) return cudaSuccess;
        busy[k] = false;
        return cudaEventSynchronize(ev[k]);
    }
};

Ring *ring_of_current_device()
{
    static Ring rings[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    return &rings[dev];
}

}  // namespace

bool host_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

static bool staged(const void *host, size_t bytes)
{
    static const bool off = getenv("MDAGPU_NO_STAGED_COPY") != nullptr;
    return !off && bytes >= HC_DIRECT_BELOW && !host_is_pinned(host);
}

// dev[0:bytes] = host[0:bytes] in stream order of `st`.  Returns once `host` has been read
// (like cudaMemcpyAsync from pageable memory); the device side completes in stream order.
cudaError_t host_to_device(void *dev, const void *host, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return cudaSuccess;
    Ring *r = staged(host, bytes) ? ring_of_current_device() : nullptr;
    if (!r) return cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st);
    std::lock_guard<std::mutex> lk(r->m);
    cudaError_t e = r->init();
    if (e != cudaSuccess) return e;
    CopyPool *cp = pool();
    int k = 0;
    for (size_t off = 0; off < bytes; off += HC_CHUNK, k = (k + 1) % HC_RING) {
        const size_t n = bytes - off < HC_CHUNK ? bytes - off : HC_CHUNK;
        if ((e = r->wait(k)) != cudaSuccess) return e;              // the DMA that last read this chunk is done
        cp->copy(r->chunk[k], (const char *)host + off, n);
        if ((e = cudaMemcpyAsync((char *)dev + off, r->chunk[k], n, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(r->ev[k], st)) != cudaSuccess) return e;
        r->busy[k] = true;
    }
    return cudaSuccess;
}

// host[0:bytes] = dev[0:bytes] after everything queued on `st`; returns when `host` holds the data.
cudaError_t device_to_host(void *host, const void *dev, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return cudaSuccess;
    Ring *r = staged(host, bytes) ? ring_of_current_device() : nullptr;
    if (!r) {
        cudaError_t e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    }
    std::lock_guard<std::mutex> lk(r->m);
    cudaError_t e = r->init();
    if (e != cudaSuccess) return e;
    CopyPool *cp = pool();
    const size_t nchunk = (bytes + HC_CHUNK - 1) / HC_CHUNK;
    auto issue = [&](size_t c) -> cudaError_t {
        const int k = (int)(c % HC_RING);
        const size_t off = c * HC_CHUNK, n = bytes - off < HC_CHUNK ? bytes - off : HC_CHUNK;
        cudaError_t e2 = r->wait(k);                                 // left over from an earlier upload
        if (e2 == cudaSuccess) e2 = cudaMemcpyAsync(r->chunk[k], (const char *)dev + off, n, cudaMemcpyDeviceToHost, st);
        if (e2 == cudaSuccess) e2 = cudaEventRecord(r->ev[k], st);
        r->busy[k] = true;
        return e2;
    };
    for (size_t c = 0; c < nchunk && c < (size_t)HC_RING; ++c)
        if ((e = issue(c)) != cudaSuccess) return e;
    for (size_t c = 0; c < nchunk; ++c) {
        const int k = (int)(c % HC_RING);
        const size_t off = c * HC_CHUNK, n = bytes - off < HC_CHUNK ? bytes - off : HC_CHUNK;
        if ((e = r->wait(k)) != cudaSuccess) return e;
        cp->copy((char *)host + off, r->chunk[k], n);
        if (c + HC_RING < nchunk && (e = issue(c + HC_RING)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace mdagpu

using namespace mdagpu;

// Public face of the two movers: torch tensors filled from / drained into NumPy arrays at the
// staged rate (FrameBlock.to_device; torch's own .cuda() of a pageable array takes the driver's path).
extern "C" int mdagpu_host_to_device(void *dev, const void *host, uint64_t bytes, mdagpu_stream_t stream)
{
    MDA_REQUIRE(bytes == 0 || (dev && host), "mdagpu_host_to_device: NULL argument");
    MDA_CUDA(host_to_device(dev, host, bytes, (cudaStream_t)stream));
    return MDAGPU_OK;
}

extern "C" int mdagpu_device_to_host(void *host, const void *dev, uint64_t bytes, mdagpu_stream_t stream)
{
    MDA_REQUIRE(bytes == 0 || (dev && host), "mdagpu_device_to_host: NULL argument");
    MDA_CUDA(device_to_host(host, dev, bytes, (cudaStream_t)stream));
    return MDAGPU_OK;
}
