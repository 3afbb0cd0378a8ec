// ga_host.cu -- host-side pieces of libga_b200.so: error reporting, small utilities and the
// contig traversal over the CSR, which stays on the host by design (north star item 4;
// debruijn_graph.py:72-111 unpaired, :222-267 paired; Node.pop_edge = dict.popitem,
// debruijn_node.py:24-26).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ga_common.cuh"

#include <atomic>

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

extern "C" void ga_free_host(void* p) { free(p); }

// The hot kernels are random 16-byte probes; a DRAM fetch wider than one 32-byte sector is
// wasted bandwidth for them.  Process-wide hint (cudaLimitMaxL2FetchGranularity).
extern "C" int ga_set_l2_fetch_granularity(int bytes) {
    GA_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes));
    return GA_OK;
}

// Walk = pop the last remaining edge of the start node, then keep going while the current node
// still has edges and was not branching; every step contributes the successor's last symbol.
extern "C" int ga_traverse_contigs(const int32_t* rowptr, const int32_t* col, const int32_t* indeg,
                                   const uint8_t* branching, const uint8_t* last_char, int64_t n_nodes,
                                   int64_t num_edges_attr, int paired, uint8_t** text_out,
                                   uint64_t** offsets_out, uint64_t* n_contigs, int32_t* left_out) {
    if (!text_out || !offsets_out || !n_contigs || n_nodes < 0 ||
        (n_nodes > 0 && (!rowptr || !indeg || !branching || !last_char))) {
        ga_set_error("ga_traverse_contigs: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    std::vector<int32_t> left((size_t)n_nodes);
    for (int64_t i = 0; i < n_nodes; ++i) left[(size_t)i] = rowptr[i + 1] - rowptr[i];
    std::vector<uint8_t> text;
    std::vector<uint64_t> offs(1, 0);
    int64_t remaining = num_edges_attr;
    auto walk = [&](int64_t start) {
        int64_t j = col[rowptr[start] + --left[(size_t)start]];
        --remaining;
        text.push_back(last_char[j]);
        while (left[(size_t)j] > 0 && !branching[j]) {
            int64_t nx = col[rowptr[j] + --left[(size_t)j]];
            --remaining;
            text.push_back(last_char[nx]);
            j = nx;
        }
        offs.push_back(text.size());
    };
    bool done = false;
    for (int64_t i = 0; i < n_nodes && !done; ++i) {
        while (left[(size_t)i] > 0 && (branching[i] || indeg[i] == 0)) walk(i);
        if (remaining == 0) done = true;   // (:80-81, 231-232)
    }
    if (!done && n_nodes > 0) {
        if (paired) {
            for (int64_t i = 0; i < n_nodes && !done; ++i) {
                while (left[(size_t)i] > 0) walk(i);
                if (remaining == 0) done = true;
            }
        } else {
            // the reference's second loop keeps testing the last node of the first loop
            // (debruijn_graph.py:85-86): only that node's cycle is ever emitted
            int64_t last = n_nodes - 1;
            while (left[(size_t)last] > 0) walk(last);
        }
    }
    uint8_t* t = (uint8_t*)malloc(text.size() ? text.size() : 1);
    uint64_t* o = (uint64_t*)malloc(offs.size() * sizeof(uint64_t));
    if (!t || !o) {
        free(t);
        free(o);
        ga_set_error("ga_traverse_contigs: out of host memory");
        return GA_ERR_BAD_ARG;
    }
    if (left_out && n_nodes > 0) memcpy(left_out, left.data(), (size_t)n_nodes * sizeof(int32_t));
    if (!text.empty()) memcpy(t, text.data(), text.size());
    memcpy(o, offs.data(), offs.size() * sizeof(uint64_t));
    *text_out = t;
    *offsets_out = o;
    *n_contigs = offs.size() - 1;
    return GA_OK;
}
