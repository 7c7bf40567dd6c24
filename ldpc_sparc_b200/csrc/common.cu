// Library-wide state: thread-local error message, launch counter, device info.
#include "common.cuh"

#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace lsb {

static thread_local char t_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

static int cached_attr(cudaDeviceAttr attr, int *cache /* [64] */)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (cache[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, attr, dev) != cudaSuccess) return 0;
        cache[dev] = v;
    }
    return cache[dev];
}

int sm_count()
{
    static int cache[64] = {0};
    return cached_attr(cudaDevAttrMultiProcessorCount, cache);
}

int max_smem_optin()
{
    static int cache[64] = {0};
    return cached_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, cache);
}

cudaError_t ensure_dynamic_smem(const void *kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> granted;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t &have = granted[std::make_pair(kernel, dev)];
    if (bytes <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

}  // namespace lsb

extern "C" {

int lsb_version(void) { return 1; }

const char *lsb_last_error(void) { return lsb::t_error; }

int lsb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t lsb_launch_count(void) { return lsb::g_launches.load(); }
void lsb_launch_count_reset(void) { lsb::g_launches.store(0); }

void *lsb_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        lsb::set_error("cudaMallocHost(%llu) failed", (unsigned long long)bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void lsb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

}  // extern "C"
