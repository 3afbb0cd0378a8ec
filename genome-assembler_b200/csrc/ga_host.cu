// ga_host.cu -- host-side pieces of libga_b200.so: error reporting and small utilities.  The raw-byte read parser
// (assemble.py:40-71) is in ga_parse.cu, the contig traversal over the CSR in ga_traverse.cu.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ga_common.cuh"

#include <atomic>
#include <thread>

static thread_local char g_error[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void ga_note_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

extern "C" uint64_t ga_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void ga_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int ga_cuda_fail(cudaError_t e, const char* what) {
    ga_set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return GA_ERR_CUDA;
}

// The stream-ordered allocator (cudaMallocAsync: CSR scratch, note overflow buffer, scan tiles) returns freed
// memory to the driver at every synchronisation unless the pool is told to keep it; re-creating hundreds of MB
// of mappings every step costs milliseconds of host time per call.  Once per device.
void ga_pool_retain() {
    static std::atomic<unsigned long long> done{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    if (done.load(std::memory_order_relaxed) & (1ull << dev)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    done.fetch_or(1ull << dev, std::memory_order_relaxed);
}

extern "C" int ga_version(void) { return 100; }

extern "C" const char* ga_last_error(void) { return g_error; }

extern "C" int ga_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int ga_fill_bytes(void* dev, int value, uint64_t bytes, ga_stream stream) {
    if (bytes == 0) return GA_OK;
    GA_CUDA(cudaMemsetAsync(dev, value, bytes, (cudaStream_t)stream));
    return GA_OK;
}

extern "C" int ga_copy_bytes(void* dst, const void* src, uint64_t bytes, ga_stream stream) {
    if (bytes == 0) return GA_OK;
    if (!dst || !src) {
        ga_set_error("ga_copy_bytes: null pointer");
        return GA_ERR_BAD_ARG;
    }
    GA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return GA_OK;
}

extern "C" void ga_free_host(void* p) { free(p); }

// The hot kernels are random 16-byte probes; a DRAM fetch wider than one 32-byte sector is
// wasted bandwidth for them.  Process-wide hint (cudaLimitMaxL2FetchGranularity).
extern "C" int ga_set_l2_fetch_granularity(int bytes) {
    GA_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
    return GA_OK;
}
