// ga_peer.cu -- the hash-partition exchange of the multi-GPU build as ONE kernel over NVLink peer memory.
//
// SURVEY 8e / BASELINE north star: "k-mers are routed to their owning GPU by a hash-partition all-to-all over
// NVLink".  On this path the unit that travels is the super-k-mer record (csrc/ga_superkmer.cu): rank g owns a
// contiguous range of bucket ids and counts every rank's records of that range.  Round 1 did it in three steps:
// sort the records into a dense local copy (25 ms per rank at N = 2), then two NCCL all_to_all_single calls over
// the bases and meta arrays (41 ms), none of it overlapped.  Here the owner's receive buffer is mapped into every
// rank (CUDA IPC, ga_peer_alloc / ga_peer_open) and sk_push_records_kernel does sort + send in one pass: it walks
// the rank's bucket-sorted INDEX (4 bytes per record), gathers each 32-byte slot from its L2-sized level-1 bucket
// and stores bases and meta straight into the owner's buffer -- coalesced 512-byte / 256-byte runs per warp on
// the wire, nothing staged in local DRAM, no NCCL call on the data path.  The destination of a position is known
// up front from the all-gathered cut matrix (rows every source sends to every owner), so no remote atomics.
#include <cstdlib>
#include <cstring>

#include "ga_common.cuh"

namespace {

constexpr int PUSH_THREADS = 256;
constexpr int PUSH_PER = 8;
constexpr u32 PUSH_TILE = PUSH_THREADS * PUSH_PER;

struct PushPlan {
    u64 cut[GA_PEER_MAX_RANKS + 1];     // this rank's bucket-sorted positions [cut[g], cut[g+1]) belong to rank g
    u64* dst_bases[GA_PEER_MAX_RANKS];  // where position cut[g] lands in rank g's buffer (this source's segment)
    u64* dst_meta[GA_PEER_MAX_RANKS];
    u32 world;
};

__global__ void __launch_bounds__(PUSH_THREADS)
sk_push_records_kernel(const u64* __restrict__ rec, u64 cap1, const u32* __restrict__ index,
                       const u64* __restrict__ offsets, u32 n_l1, int l2_bits, const PushPlan plan) {
    __shared__ u64 s_cut[GA_PEER_MAX_RANKS + 1];
    __shared__ u64* s_bases[GA_PEER_MAX_RANKS];
    __shared__ u64* s_meta[GA_PEER_MAX_RANKS];
    const u32 tid = threadIdx.x, world = plan.world;
    if (tid <= world) s_cut[tid] = plan.cut[tid];
    if (tid < world) {
        s_bases[tid] = plan.dst_bases[tid];
        s_meta[tid] = plan.dst_meta[tid];
    }
    __syncthreads();
    const u64 tiles_per = (cap1 + PUSH_TILE - 1) / PUSH_TILE;
    const u64 total = (u64)n_l1 * tiles_per;
    for (u64 tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const u32 b1 = (u32)(tile / tiles_per);
        const u64 first = offsets[(u64)b1 << l2_bits], last = offsets[(u64)(b1 + 1u) << l2_bits];
        const u64 lo = first + (tile % tiles_per) * PUSH_TILE;
        if (lo >= last) continue;                            // CTA-uniform
        const u64 slot0 = (u64)b1 * cap1;
        u32 idx[PUSH_PER];
#pragma unroll
        for (int u = 0; u < PUSH_PER; ++u) {
            const u64 p = lo + (u64)u * PUSH_THREADS + tid;
            idx[u] = p < last ? __ldg(index + p) : 0u;
        }
        u64 hi_[PUSH_PER], lo_[PUSH_PER], mt[PUSH_PER];
#pragma unroll
        for (int u = 0; u < PUSH_PER; ++u) {
            const u64 p = lo + (u64)u * PUSH_THREADS + tid;
            if (p < last) {
                u64 pad;
                asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];"
                             : "=l"(hi_[u]), "=l"(lo_[u]), "=l"(mt[u]), "=l"(pad)
                             : "l"(rec + 4u * (slot0 + idx[u])));
            }
        }
#pragma unroll
        for (int u = 0; u < PUSH_PER; ++u) {
            const u64 p = lo + (u64)u * PUSH_THREADS + tid;
            if (p < last) {
                u32 g = 0;
                while (g + 1u < world && p >= s_cut[g + 1u]) ++g;
                const u64 at = p - s_cut[g];
                if (!s_bases[g]) continue;                   // probe runs only (GA_PUSH_SKIP)
                asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(s_bases[g] + 2u * at), "l"(hi_[u]), "l"(lo_[u]) : "memory");
                asm volatile("st.global.u64 [%0], %1;" ::"l"(s_meta[g] + at), "l"(mt[u]) : "memory");
            }
        }
    }
    __threadfence_system();        // the peers read after a stream-ordered barrier that follows this kernel
}

}  // namespace

extern "C" int ga_peer_alloc(uint64_t bytes, void** ptr_out, void* handle_out) {
    if (!ptr_out || !handle_out || bytes == 0) {
        ga_set_error("ga_peer_alloc: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    void* p = nullptr;
    GA_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return ga_cuda_fail(e, "cudaIpcGetMemHandle");
    }
    static_assert(sizeof(h) == GA_PEER_HANDLE_BYTES, "CUDA IPC handle size");
    memcpy(handle_out, &h, sizeof(h));
    *ptr_out = p;
    return GA_OK;
}

extern "C" int ga_peer_open(const void* handle, void** ptr_out) {
    if (!handle || !ptr_out) {
        ga_set_error("ga_peer_open: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    GA_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return GA_OK;
}

extern "C" int ga_peer_close(void* ptr) {
    if (ptr) GA_CUDA(cudaIpcCloseMemHandle(ptr));
    return GA_OK;
}

extern "C" int ga_peer_free(void* ptr) {
    if (ptr) GA_CUDA(cudaFree(ptr));
    return GA_OK;
}

extern "C" int ga_sk_push_records(const void* records_dev, uint64_t l1_capacity, const uint32_t* index_dev,
                                  const uint64_t* offsets_dev, int l1_bits, int l2_bits, uint32_t world,
                                  const uint64_t* cut, void* const* dst_bases, void* const* dst_meta,
                                  ga_stream stream) {
    if (!records_dev || !index_dev || !offsets_dev || !cut || !dst_bases || !dst_meta || world == 0 ||
        world > GA_PEER_MAX_RANKS || l1_capacity == 0 || l1_bits < 0 || l1_bits > 10 || l2_bits < 0 || l2_bits > 10 ||
        ((uintptr_t)records_dev & 31u)) {
        ga_set_error("ga_sk_push_records: bad arguments (1..%d ranks, bucket bits 0..10 each)", GA_PEER_MAX_RANKS);
        return GA_ERR_BAD_ARG;
    }
    PushPlan plan;
    memset(&plan, 0, sizeof(plan));
    plan.world = world;
    for (uint32_t g = 0; g <= world; ++g) plan.cut[g] = cut[g];
    for (uint32_t g = 0; g < world; ++g) {
        if (cut[g + 1] < cut[g] || (cut[g + 1] > cut[g] && (!dst_bases[g] || !dst_meta[g] || ((uintptr_t)dst_bases[g] & 15u)))) {
            ga_set_error("ga_sk_push_records: cut must ascend and every non-empty range needs 16-byte aligned targets");
            return GA_ERR_BAD_ARG;
        }
        plan.dst_bases[g] = (u64*)dst_bases[g];
        plan.dst_meta[g] = (u64*)dst_meta[g];
    }
    if (cut[world] == cut[0]) return GA_OK;
    // probe runs (scripts/gpu_r2_multi.sh): GA_PUSH_SKIP=local|remote drops the stores into this rank's own buffer
    // (the target with the highest address distance is not known here: "local" = the target the caller marked
    // by passing it LAST in the environment variable GA_PUSH_SELF) -- timing only, the result is then wrong
    if (const char* skip = getenv("GA_PUSH_SKIP")) {
        const char* self = getenv("GA_PUSH_SELF");
        const uint32_t me = self ? (uint32_t)atoi(self) : 0u;
        for (uint32_t g = 0; g < world; ++g)
            if ((g == me) == (strcmp(skip, "local") == 0)) plan.dst_bases[g] = nullptr;
    }
    const u32 n_l1 = 1u << l1_bits;
    const u64 total = (u64)n_l1 * ((l1_capacity + PUSH_TILE - 1) / PUSH_TILE);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const u64 most = (u64)sms * 8;
    const unsigned grid = (unsigned)(total < most ? total : most);
    sk_push_records_kernel<<<grid, PUSH_THREADS, 0, (cudaStream_t)stream>>>(
        (const u64*)records_dev, l1_capacity, index_dev, (const u64*)offsets_dev, n_l1, l2_bits, plan);
    GA_LAUNCH_CHECK("sk_push_records");
    return GA_OK;
}
