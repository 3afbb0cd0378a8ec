// ga_partition.cu -- hash-range partitioning of the (k-1)-mer occurrence stream.
//
// At the scale of BASELINE config C4 (1.2e10 occurrences) neither the pre-filter sketch (6 GB) nor
// the candidate table fits the 126 MB L2, so counting in read order makes every probe a DRAM miss
// (profiles/r01: 330 + 750 ms of the 1.23 s step).  Both structures are addressed by
// floor(hash * size / 2^64), which is monotone in the hash: if the occurrences are visited in hash
// order, the live part of each structure is a small moving window.  This file materialises the
// packed keys once, bucketed by the top bits of their hash (sequential 8-byte writes), and runs the
// two counting passes of ga_prefilter.cu over that stream.  Results are identical to the read-order
// passes (same cells, same table); only the order of the atomics changes.
//
// Replaces, together with ga_prefilter.cu, the counting loop of debruijn_graph.py:144-152 / 349-367.
#include "ga_common.cuh"

namespace {

constexpr int PART_THREADS = 256;
constexpr int PART_ROUNDS = 16;                       // window steps staged per flush
constexpr int PART_STAGE = PART_THREADS * PART_ROUNDS;  // 4096 keys = 32 KB
constexpr u32 PART_MAX = 512;

// Block-wide flush of the staged keys into their buckets: shared-memory histogram, one global
// cursor reservation per (block, bucket), then the scatter.
__device__ __forceinline__ void flush_stage(const u64* stage, u16* stage_part, u32 staged, u32 n_parts,
                                            u32* hist, u64* gbase, u64* __restrict__ items,
                                            u64 part_capacity, u64* __restrict__ cursors, bool& overflow) {
    for (u32 p = threadIdx.x; p < n_parts; p += PART_THREADS) hist[p] = 0;
    __syncthreads();
    for (u32 i = threadIdx.x; i < staged; i += PART_THREADS) {
        u64 key = stage[i];
        if (key == GA_NONE64) continue;
        u32 part = (u32)__umul64hi(ga_key_hash(key), (u64)n_parts);
        stage_part[i] = (u16)part;
        atomicAdd(&hist[part], 1u);
    }
    __syncthreads();
    for (u32 p = threadIdx.x; p < n_parts; p += PART_THREADS) {
        u32 c = hist[p];
        u64 base = 0;
        if (c) {
            base = atomicAdd(&cursors[p], (u64)c);
            if (base + c > part_capacity) {
                overflow = true;
                base = GA_NONE64;          // drop: the host retries with larger buckets
            }
        }
        gbase[p] = base;
        hist[p] = 0;
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < staged; i += PART_THREADS) {
        u64 key = stage[i];
        if (key == GA_NONE64) continue;
        u32 part = stage_part[i];
        u32 off = atomicAdd(&hist[part], 1u);
        u64 base = gbase[part];
        if (base != GA_NONE64) items[(u64)part * part_capacity + base + off] = key;
    }
    __syncthreads();
}

// One thread per read; the whole block walks window positions in lock step so that it can flush
// its staging buffer together.  64-bit keys only.
template <int SB>
__global__ void __launch_bounds__(PART_THREADS)
partition_kernel(ReadsView rv, int w, u32 n_parts, u64* __restrict__ items, u64 part_capacity,
                 u64* __restrict__ cursors, u32* status) {
    constexpr u32 SPW = 64 / SB;
    constexpr u64 SMASK = (1ull << SB) - 1;
    __shared__ u64 stage[PART_STAGE];
    __shared__ u16 stage_part[PART_STAGE];
    __shared__ u32 hist[PART_MAX];
    __shared__ u64 gbase[PART_MAX];
    __shared__ u32 block_max;
    const u64 mask = ga_key_mask<u64>(w, rv.sym_bits);
    bool overflow = false;
    u32 rounds = 0;   // block-uniform: window steps currently staged
    const u64 n_tiles = (rv.n_reads + PART_THREADS - 1) / PART_THREADS;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u64 r = tile * PART_THREADS + threadIdx.x;
        const bool valid = r < rv.n_reads;
        u32 len = valid ? ga_read_len(rv, rv.paired ? (r & ~1ull) : r) : 0u;
        if (len < (u32)w) len = 0;
        const u64* words = valid ? ga_read_ptr(rv, r) : rv.words;
        if (threadIdx.x == 0) block_max = 0;
        __syncthreads();
        u32 wmax = __reduce_max_sync(0xFFFFFFFFu, len);
        if ((threadIdx.x & 31) == 0) atomicMax(&block_max, wmax);
        __syncthreads();
        const u32 max_len = block_max;
        u64 key = 0;
        for (u32 base = 0; base < max_len; base += SPW) {
            u64 word = base < len ? __ldg(words + base / SPW) : 0ull;
            const u32 lim = max_len - base < SPW ? max_len - base : SPW;
            for (u32 j = 0; j < lim; ++j) {
                key = ((key << rv.sym_bits) | (word & SMASK)) & mask;
                word >>= SB;
                const u32 i = base + j + 1;
                if (i < (u32)w) continue;                       // block-uniform
                stage[rounds * PART_THREADS + threadIdx.x] = i <= len ? key : GA_NONE64;
                if (++rounds == PART_ROUNDS) {
                    __syncthreads();
                    flush_stage(stage, stage_part, PART_STAGE, n_parts, hist, gbase, items, part_capacity, cursors,
                                overflow);
                    rounds = 0;
                }
            }
        }
    }
    if (rounds) {
        __syncthreads();
        flush_stage(stage, stage_part, rounds * PART_THREADS, n_parts, hist, gbase, items, part_capacity, cursors,
                    overflow);
    }
    if (overflow) atomicOr(status, GA_ST_TABLE_FULL);
}

__global__ void __launch_bounds__(256)
prefilter_update_keys_kernel(const u64* __restrict__ keys, u64 n, PrefilterView pf) {
    const u32 cmask = (1u << pf.cell_bits) - 1u;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 wi;
        u32 sh;
        u32 v = ga_prefilter_value(pf, ga_key_hash(keys[i]), wi, sh);
        if (v < pf.limit) {
            u32 old = atomicAdd(pf.words + wi, 1u << sh);
            if (((old >> sh) & cmask) == cmask) atomicOr(pf.words + wi, cmask << sh);
        }
    }
}

__global__ void __launch_bounds__(256)
count_candidates_keys_kernel(const u64* __restrict__ keys, u64 n, PrefilterView pf, Slot<u64>* __restrict__ table,
                             u64 capacity, u32* status) {
    bool full = false;
    // whole warps iterate together and re-converge every step (see ga_for_each_window)
    const u64 step = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (n + step - 1) / step;
    for (u64 it = 0; it < rounds; ++it) {
        const u64 i = it * step + blockIdx.x * (u64)blockDim.x + threadIdx.x;
        __syncwarp();
        if (i >= n) continue;
        const u64 key = keys[i];
        u64 wi;
        u32 sh;
        if (ga_prefilter_value(pf, ga_key_hash(key), wi, sh) < pf.limit) continue;
        u64 s = ga_table_upsert(table, capacity, key);
        if (s == GA_NONE64) full = true;
        else atomicAdd(&table[s].val, 1u);
    }
    if (full) atomicOr(status, GA_ST_TABLE_FULL);
}

}  // namespace

extern "C" int ga_partition_kmers(const ga_reads* reads, int k, uint32_t n_parts, uint64_t* items_dev,
                                  uint64_t part_capacity, uint64_t* cursors_dev, uint32_t* status_dev,
                                  ga_stream stream) {
    if (!reads || !items_dev || !cursors_dev || n_parts == 0 || n_parts > PART_MAX || part_capacity == 0) {
        ga_set_error("ga_partition_kmers: bad arguments (n_parts must be 1..%u)", PART_MAX);
        return GA_ERR_BAD_ARG;
    }
    if (ga_key_words(k, reads->sym_bits) != 1) {
        ga_set_error("ga_partition_kmers: 64-bit keys only");
        return GA_ERR_BAD_ARG;
    }
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    u64 n_tiles = (rv.n_reads + PART_THREADS - 1) / PART_THREADS;
    unsigned grid = (unsigned)(n_tiles < 148ull * 5 ? n_tiles : 148ull * 5);   // persistent: 5 blocks per SM
    cudaStream_t st = (cudaStream_t)stream;
    if (rv.storage_bits == 2)
        partition_kernel<2><<<grid, PART_THREADS, 0, st>>>(rv, k - 1, n_parts, (u64*)items_dev, part_capacity,
                                                          (u64*)cursors_dev, status_dev);
    else
        partition_kernel<8><<<grid, PART_THREADS, 0, st>>>(rv, k - 1, n_parts, (u64*)items_dev, part_capacity,
                                                          (u64*)cursors_dev, status_dev);
    GA_LAUNCH_CHECK("partition");
    return GA_OK;
}

static int check_pf(const char* fn, const ga_prefilter* pf, long long threshold) {
    if (!pf || !pf->words || pf->n_cells == 0 || (pf->cell_bits != 4 && pf->cell_bits != 8) ||
        threshold + 1 > (1ll << pf->cell_bits) - 1) {
        ga_set_error("%s: bad pre-filter", fn);
        return GA_ERR_BAD_ARG;
    }
    return GA_OK;
}

extern "C" int ga_prefilter_update_keys(const uint64_t* keys_dev, uint64_t n, const ga_prefilter* pf,
                                        int64_t threshold, ga_stream stream) {
    int rc = check_pf("ga_prefilter_update_keys", pf, threshold);
    if (rc) return rc;
    if (n == 0) return GA_OK;
    unsigned grid = ga_grid(n, 256);
    if (grid > 148u * 64u) grid = 148u * 64u;
    prefilter_update_keys_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const u64*)keys_dev, n,
                                                                         ga_prefilter_view(pf, threshold));
    GA_LAUNCH_CHECK("prefilter_update_keys");
    return GA_OK;
}

extern "C" int ga_count_candidates_keys(const uint64_t* keys_dev, uint64_t n, const ga_prefilter* pf,
                                        int64_t threshold, void* table_dev, uint64_t capacity,
                                        uint32_t* status_dev, ga_stream stream) {
    int rc = check_pf("ga_count_candidates_keys", pf, threshold);
    if (rc) return rc;
    if (!table_dev || capacity == 0) return GA_ERR_BAD_ARG;
    if (n == 0) return GA_OK;
    unsigned grid = ga_grid(n, 256);
    if (grid > 148u * 64u) grid = 148u * 64u;
    count_candidates_keys_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const u64*)keys_dev, n, ga_prefilter_view(pf, threshold), (Slot<u64>*)table_dev, capacity, status_dev);
    GA_LAUNCH_CHECK("count_candidates_keys");
    return GA_OK;
}
