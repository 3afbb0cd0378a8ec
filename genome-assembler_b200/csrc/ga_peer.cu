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
                       const u64* __restrict__ offsets, u32 n_l1, int l2_bits, u32 ahead, const PushPlan plan) {
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
        // Optional (GA_PUSH_AHEAD=n): the tile that works on records [t, t + TILE) of bucket b1 asks the TMA engine
        // to pull the same slice of the bucket n further on into L2 as whole lines (cp.async.bulk.prefetch.L2,
        // UBLKPF.L2 in SASS) so that the random 32-byte gathers below would hit L2.  Measured at N = 2 on C4
        // (profiles/r02/push_probes.txt): 28.6 ms without, 30.1-35.3 ms with n = 1..4 -- the level-1 buckets of
        // half the input are 29 MB each and the persistent grid is spread over 1.3-2.7 of them, so a useful
        // distance does not fit the L2 next to what is being gathered.  Off by default; what took the time out of
        // this step is sk_push_sorted_kernel below, which does not gather at all.
        if (tid == 0 && b1 + ahead < n_l1) {
            const u32 b2 = b1 + ahead;
            const u64 n2 = offsets[(u64)(b2 + 1u) << l2_bits] - offsets[(u64)b2 << l2_bits];
            const u64 t0 = (tile % tiles_per) * PUSH_TILE;
            if (t0 < n2) {
                const u32 bytes = (u32)min((u64)PUSH_TILE, n2 - t0) * 32u;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rec + 4u * ((u64)b2 * cap1 + t0)), "r"(bytes) : "memory");
            }
        }
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
                [[maybe_unused]] u64 pad;
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

// ------------------------------------------------------------------------------------------------
// Level-2 split + send in ONE pass over the level-1 buckets, no index and no gather.
// The gather above costs 21 of its 28.6 ms at N = 2 even with the remote stores switched off: every 32-byte slot
// is fetched as a lone DRAM sector, in random order, after the index pass has already streamed the same slots
// once (7 ms).  Here a CTA streams a chunk of consecutive slots (coalesced, DRAM speed), counting-sorts the
// chunk by final bucket in shared memory -- the records themselves, not an index -- takes the chunk's place in
// every final bucket with one returned atomic per non-empty bucket on the exact-offset cursors, and stores
// the sorted records straight into the owners' buffers: the ~3.5 records a chunk holds per final bucket leave as
// one 56-byte run of bases and one 28-byte run of meta words.  Slots are read twice (meta word first for the
// histogram, whole slot for the move); the second read hits L2, a chunk is 112 KB.
constexpr int PS_THREADS = 512;
constexpr int PS_PER = 7;
constexpr u32 PS_CHUNK = PS_THREADS * PS_PER;       // 3584 records: 84 KB of staged records, two CTAs per SM
struct PSShared {
    ulonglong2 bases[PS_CHUNK];
    u64 meta[PS_CHUNK];
    u64 gbase[1024];
    u32 hist[1024];
    u32 scan[1024];
    u32 wsum[PS_THREADS / 32];
    u64 cut[GA_PEER_MAX_RANKS + 1];
    u64* dst_bases[GA_PEER_MAX_RANKS];
    u64* dst_meta[GA_PEER_MAX_RANKS];
};

__global__ void __launch_bounds__(PS_THREADS, 2)
sk_push_sorted_kernel(const u64* __restrict__ rec, u64 cap1, const u64* __restrict__ cursors1, u32 cursor_stride,
                      u32 n_l1, int l2_bits, u64* __restrict__ cursors2, const PushPlan plan) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PSShared& sm = *reinterpret_cast<PSShared*>(smem_raw);
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, world = plan.world;
    if (tid <= world) sm.cut[tid] = plan.cut[tid];
    if (tid < world) {
        sm.dst_bases[tid] = plan.dst_bases[tid];
        sm.dst_meta[tid] = plan.dst_meta[tid];
    }
    const u64 chunks_per = (cap1 + PS_CHUNK - 1) / PS_CHUNK;
    const u64 total = (u64)n_l1 * chunks_per;
    for (u64 chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        const u32 b1 = (u32)(chunk / chunks_per);
        const u64 lo = (chunk % chunks_per) * PS_CHUNK;
        const u64 cnt1 = min(cursors1[(u64)b1 * cursor_stride], cap1);
        if (lo >= cnt1) continue;                            // CTA-uniform
        const u64* slots = rec + 4u * ((u64)b1 * cap1 + lo);
        const u32 n = (u32)min(cnt1 - lo, (u64)PS_CHUNK);
        sm.hist[tid] = 0;
        sm.hist[tid + PS_THREADS] = 0;
        __syncthreads();
        u32 br[PS_PER];                                      // final bucket | rank inside the chunk << 16
#pragma unroll
        for (int u = 0; u < PS_PER; ++u) {
            const u32 i = (u32)u * PS_THREADS + tid;
            br[u] = 0;
            if (i < n) {
                const u32 b2 = (u32)(__ldg(slots + 4u * i + 2u) >> 6) & 1023u;
                br[u] = b2 | (atomicAdd(&sm.hist[b2], 1u) << 16);
            }
        }
        __syncthreads();
        {   // exclusive scan of the 1024 counts (two per thread) and the chunk's place in every final bucket
            const u32 a = sm.hist[2u * tid], b = sm.hist[2u * tid + 1u];
            u32 incl = a + b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= (u32)o) incl += t;
            }
            if (lane == 31u) sm.wsum[warp] = incl;
            const u64 first = (u64)b1 << l2_bits;
            const u64 ga = a ? atomicAdd((unsigned long long*)&cursors2[first + 2u * tid], (unsigned long long)a) : 0ull;
            const u64 gb = b ? atomicAdd((unsigned long long*)&cursors2[first + 2u * tid + 1u], (unsigned long long)b) : 0ull;
            __syncthreads();
            u32 before = 0;
            for (u32 q = 0; q < warp; ++q) before += sm.wsum[q];
            const u32 base = before + incl - (a + b);
            sm.scan[2u * tid] = base;
            sm.scan[2u * tid + 1u] = base + a;
            sm.gbase[2u * tid] = ga;
            sm.gbase[2u * tid + 1u] = gb;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PS_PER; ++u) {
            const u32 i = (u32)u * PS_THREADS + tid;
            if (i < n) {
                u64 hi, lw, mt;
                [[maybe_unused]] u64 pad;
                asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];"
                             : "=l"(hi), "=l"(lw), "=l"(mt), "=l"(pad)
                             : "l"(slots + 4u * i));
                const u32 pos = sm.scan[br[u] & 0xFFFFu] + (br[u] >> 16);
                sm.bases[pos] = make_ulonglong2(hi, lw);
                sm.meta[pos] = mt;
            }
        }
        __syncthreads();
        for (u32 t = tid; t < n; t += PS_THREADS) {
            const u64 mt = sm.meta[t];
            const u32 b2 = (u32)(mt >> 6) & 1023u;
            const u64 p = sm.gbase[b2] + (t - sm.scan[b2]);     // this rank's bucket-sorted position of the record
            u32 g = 0;
            while (g + 1u < world && p >= sm.cut[g + 1u]) ++g;
            const u64 at = p - sm.cut[g];
            if (!sm.dst_bases[g]) continue;                      // probe runs only (GA_PUSH_SKIP)
            const ulonglong2 bs = sm.bases[t];
            asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(sm.dst_bases[g] + 2u * at), "l"(bs.x), "l"(bs.y) : "memory");
            asm volatile("st.global.u64 [%0], %1;" ::"l"(sm.dst_meta[g] + at), "l"(mt) : "memory");
        }
        __syncthreads();
    }
    __threadfence_system();        // the peers read after a stream-ordered barrier that follows this kernel
}

}  // namespace

static int ga_fill_plan(PushPlan& plan, uint32_t world, const uint64_t* cut, void* const* dst_bases, void* const* dst_meta,
                        const char* who);

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
    const int rc = ga_fill_plan(plan, world, cut, dst_bases, dst_meta, "ga_sk_push_records");
    if (rc != GA_OK) return rc;
    if (cut[world] == cut[0]) return GA_OK;
    const u32 n_l1 = 1u << l1_bits;
    const u64 total = (u64)n_l1 * ((l1_capacity + PUSH_TILE - 1) / PUSH_TILE);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 8;
    if (const char* e = getenv("GA_PUSH_CTAS")) per_sm = atoi(e) > 0 ? atoi(e) : 8;
    const u64 most = (u64)sms * (u64)per_sm;
    const unsigned grid = (unsigned)(total < most ? total : most);
    // GA_PUSH_AHEAD=n: prefetch n level-1 buckets ahead (off by default -- measured slower, see the kernel)
    u32 ahead = 0;
    if (const char* e = getenv("GA_PUSH_AHEAD")) ahead = (u32)atoi(e);
    if (ahead == 0u) ahead = n_l1;
    sk_push_records_kernel<<<grid, PUSH_THREADS, 0, (cudaStream_t)stream>>>(
        (const u64*)records_dev, l1_capacity, index_dev, (const u64*)offsets_dev, n_l1, l2_bits, ahead, plan);
    GA_LAUNCH_CHECK("sk_push_records");
    return GA_OK;
}

static int ga_fill_plan(PushPlan& plan, uint32_t world, const uint64_t* cut, void* const* dst_bases, void* const* dst_meta,
                        const char* who) {
    memset(&plan, 0, sizeof(plan));
    plan.world = world;
    for (uint32_t g = 0; g <= world; ++g) plan.cut[g] = cut[g];
    for (uint32_t g = 0; g < world; ++g) {
        if (cut[g + 1] < cut[g] || (cut[g + 1] > cut[g] && (!dst_bases[g] || !dst_meta[g] || ((uintptr_t)dst_bases[g] & 15u)))) {
            ga_set_error("%s: cut must ascend and every non-empty range needs 16-byte aligned targets", who);
            return GA_ERR_BAD_ARG;
        }
        plan.dst_bases[g] = (u64*)dst_bases[g];
        plan.dst_meta[g] = (u64*)dst_meta[g];
    }
    // timing probes only (the result is then wrong): GA_PUSH_SKIP=local drops the stores into this rank's own
    // buffer, =remote those into the peers'; GA_PUSH_SELF names this rank (set by ga_multi)
    if (const char* skip = getenv("GA_PUSH_SKIP")) {
        const char* self = getenv("GA_PUSH_SELF");
        const uint32_t me = self ? (uint32_t)atoi(self) : 0u;
        for (uint32_t g = 0; g < world; ++g)
            if ((g == me) == (strcmp(skip, "local") == 0)) plan.dst_bases[g] = nullptr;
    }
    return GA_OK;
}

extern "C" int ga_sk_push_sorted(const void* records_dev, uint64_t l1_capacity, const uint64_t* l1_cursors_dev,
                                 int l1_bits, int l2_bits, uint64_t* cursors_dev, uint32_t world, const uint64_t* cut,
                                 void* const* dst_bases, void* const* dst_meta, ga_stream stream) {
    if (!records_dev || !l1_cursors_dev || !cursors_dev || !cut || !dst_bases || !dst_meta || world == 0 ||
        world > GA_PEER_MAX_RANKS || l1_capacity == 0 || l1_bits < 0 || l1_bits > 10 || l2_bits < 0 || l2_bits > 10 ||
        ((uintptr_t)records_dev & 31u)) {
        ga_set_error("ga_sk_push_sorted: bad arguments (1..%d ranks, bucket bits 0..10 each)", GA_PEER_MAX_RANKS);
        return GA_ERR_BAD_ARG;
    }
    PushPlan plan;
    const int rc = ga_fill_plan(plan, world, cut, dst_bases, dst_meta, "ga_sk_push_sorted");
    if (rc != GA_OK) return rc;
    if (cut[world] == cut[0]) return GA_OK;
    GA_CUDA(cudaFuncSetAttribute(sk_push_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PSShared)));
    const u32 n_l1 = 1u << l1_bits;
    const u64 total = (u64)n_l1 * ((l1_capacity + PS_CHUNK - 1) / PS_CHUNK);
    int dev = 0, sms = 148, per_sm = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    GA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sk_push_sorted_kernel, PS_THREADS, sizeof(PSShared)));
    const u64 most = (u64)sms * (u64)(per_sm > 0 ? per_sm : 1);
    const unsigned grid = (unsigned)(total < most ? total : most);
    sk_push_sorted_kernel<<<grid, PS_THREADS, sizeof(PSShared), (cudaStream_t)stream>>>(
        (const u64*)records_dev, l1_capacity, (const u64*)l1_cursors_dev, (u32)ga_sk_cursor_stride(), n_l1, l2_bits,
        (u64*)cursors_dev, plan);
    GA_LAUNCH_CHECK("sk_push_sorted");
    return GA_OK;
}
