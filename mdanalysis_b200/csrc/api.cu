// api.cu -- library-level entry points of libmdagpu.so: ABI version, error text,
// device selection, launch counter.
#include "common.cuh"

#include <stdlib.h>

#include <map>
#include <string>
#include <vector>

namespace mdagpu {

unsigned long long g_launches = 0;
int g_profile = 0;

struct ProfRec { const char *name; cudaEvent_t a, b; };
static std::vector<ProfRec> g_recs;

void profile_begin(const char *name, cudaStream_t st)
{
    ProfRec r;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
}

void profile_end(cudaStream_t st)
{
    if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, st);
}

void pool_init()
{
    static bool done[64] = { false };
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    done[dev] = true;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) { cudaGetLastError(); return; }
    size_t free_b = 0, total_b = 0;
    unsigned long long keep = 0;
    const char *env = getenv("MDAGPU_POOL_RETAIN_MB");
    if (env && *env) keep = strtoull(env, nullptr, 10) << 20;
    else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) keep = total_b / 4;
    if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) cudaGetLastError();
}

char *err_buf()
{
    static thread_local char buf[512] = "";
    return buf;
}

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    if (code == MDAGPU_ERR_CUDA) cudaGetLastError();   // clear the sticky-free error state
    return code;
}

}  // namespace mdagpu

using namespace mdagpu;

extern "C" int mdagpu_abi_version(void) { return MDAGPU_ABI_VERSION; }

extern "C" const char *mdagpu_last_error(void) { return err_buf(); }

extern "C" int mdagpu_device_count(int *count)
{
    MDA_REQUIRE(count != nullptr, "mdagpu_device_count: NULL argument");
    *count = 0;
    MDA_CUDA(cudaGetDeviceCount(count));
    return MDAGPU_OK;
}

extern "C" int mdagpu_set_device(int device)
{
    MDA_CUDA(cudaSetDevice(device));
    return MDAGPU_OK;
}

extern "C" unsigned long long mdagpu_launch_count(void) { return g_launches; }

extern "C" int mdagpu_profile_enable(int on)
{
    g_profile = on ? 1 : 0;
    return MDAGPU_OK;
}

// Writes "name count total_ms\n" lines for everything launched since the last
// report (synchronises the device) and clears the records.
extern "C" int mdagpu_profile_report(char *buf, uint64_t size)
{
    MDA_REQUIRE(buf && size > 0, "mdagpu_profile_report: NULL buffer");
    MDA_CUDA(cudaDeviceSynchronize());
    std::map<std::string, std::pair<unsigned long long, double>> acc;
    for (auto &r : g_recs) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto &e = acc[r.name];
            e.first += 1;
            e.second += ms;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_recs.clear();
    std::string out;
    char line[256];
    for (auto &kv : acc) {
        snprintf(line, sizeof line, "%s\t%llu\t%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (out.size() + 1 > size) out.resize(size - 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return MDAGPU_OK;
}
