// ga_count.cu -- exact (k-1)-mer counting in a lock-free open-addressing table, replacing
// DeBruijnGraph._count_kmers / PairedDeBruijnGraph._count_kmers (debruijn_graph.py:144-152,
// 349-367), plus the scans that turn the table into the filter's "solid" set
// (debruijn_graph.py:127-128, 275-278).
#include "ga_common.cuh"

namespace {

template <class K> __global__ void clear_kernel(Slot<K>* table, u64 capacity);
template <> __global__ void clear_kernel<u64>(Slot<u64>* table, u64 capacity) {
    uint4 fill = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0xFFFFFFFFu);
    uint4* p = reinterpret_cast<uint4*>(table);
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < capacity; i += (u64)gridDim.x * blockDim.x)
        p[i] = fill;
}
template <> __global__ void clear_kernel<u128>(Slot<u128>* table, u64 capacity) {
    uint4 k = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    uint4 v = make_uint4(0u, 0xFFFFFFFFu, 0u, 0u);
    uint4* p = reinterpret_cast<uint4*>(table);
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < capacity; i += (u64)gridDim.x * blockDim.x) {
        p[2 * i] = k;
        p[2 * i + 1] = v;
    }
}

// One thread per read (per mate when paired).  Each window does one find-or-insert and one
// 32-bit reduction on the slot's counter.
template <class K, int SB>
__global__ void __launch_bounds__(256) count_kernel(ReadsView rv, int w, Slot<K>* __restrict__ table,
                                                    u64 capacity, u32* status) {
    const K mask = ga_key_mask<K>(w, rv.sym_bits);
    bool full = false;
    GA_FOR_EACH_READ_WARP(rv, r, valid) {
        // windows range over mate 1's length for both mates (debruijn_graph.py:374)
        u32 len = valid ? ga_read_len(rv, rv.paired ? (r & ~1ull) : r) : 0u;
        if (len < (u32)w) len = 0;
        ga_for_each_window<K, SB>(valid ? ga_read_ptr(rv, r) : rv.words, len, w, rv.sym_bits, mask, [&](u32, K key) {
            u64 s = ga_table_upsert(table, capacity, key);
            if (s == GA_NONE64) full = true;
            else atomicAdd(&table[s].val, 1u);
        });
    }
    if (full) atomicOr(status, GA_ST_TABLE_FULL);
}

template <class K>
__global__ void count_keys_kernel(const K* __restrict__ keys, const u32* __restrict__ amounts, u64 n,
                                  Slot<K>* __restrict__ table, u64 capacity, u32* status) {
    bool full = false;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 s = ga_table_upsert(table, capacity, keys[i]);
        if (s == GA_NONE64) full = true;
        else atomicAdd(&table[s].val, amounts ? amounts[i] : 1u);
    }
    if (full) atomicOr(status, GA_ST_TABLE_FULL);
}

template <class K>
__global__ void insert_ids_kernel(const K* __restrict__ keys, u64 n, u32 id_base,
                                  Slot<K>* __restrict__ table, u64 capacity, u32* status) {
    bool full = false;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 s = ga_table_upsert(table, capacity, keys[i]);
        if (s == GA_NONE64) full = true;
        else table[s].val = id_base + (u32)i;
    }
    if (full) atomicOr(status, GA_ST_TABLE_FULL);
}

template <class K> __device__ __forceinline__ K slot_key(const Slot<K>* s);
template <> __device__ __forceinline__ u64 slot_key<u64>(const Slot<u64>* s) { return s->key; }
template <> __device__ __forceinline__ u128 slot_key<u128>(const Slot<u128>* s) {
    return ((u128)s->hi << 64) | s->lo;
}

template <class K>
__global__ void summary_kernel(const Slot<K>* __restrict__ table, u64 capacity, long long threshold,
                               u64* out4) {
    u64 distinct = 0, above = 0, total = 0, mx = 0;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < capacity; i += (u64)gridDim.x * blockDim.x) {
        if (slot_key<K>(table + i) != ga_empty_key<K>()) {
            u32 c = table[i].val;
            ++distinct;
            total += c;
            above += ((long long)c > threshold);
            mx = c > mx ? c : mx;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        distinct += __shfl_down_sync(0xFFFFFFFFu, distinct, off);
        above += __shfl_down_sync(0xFFFFFFFFu, above, off);
        total += __shfl_down_sync(0xFFFFFFFFu, total, off);
        u64 o = __shfl_down_sync(0xFFFFFFFFu, mx, off);
        mx = o > mx ? o : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        if (distinct) atomicAdd(out4 + 0, distinct);
        if (above) atomicAdd(out4 + 1, above);
        if (total) atomicAdd(out4 + 2, total);
        if (mx) atomicMax(out4 + 3, mx);
    }
}

// Compaction of table entries.  `sketch` (rows > 0) switches the test from the exact count to
// the CountMinSketch estimate of the window (countminsketch.py:40-44).
template <class K>
__global__ void export_kernel(const Slot<K>* __restrict__ table, u64 capacity, long long min_exclusive,
                              SketchView sk, int w, int sym_bits, const u8* __restrict__ lut_g,
                              K* __restrict__ keys_out, u32* __restrict__ counts_out, u64* n_out) {
    __shared__ u8 lut[256];
    if (sk.rows > 0) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = lut_g[i];
        __syncthreads();
    }
    // tiles of 4 slots per thread; one append-counter atomic per tile (every thread of the block
    // takes part in the block scan, so the trip count is uniform)
    constexpr u32 ITEMS = 4;
    const u64 tile = (u64)blockDim.x * ITEMS;
    const u64 n_tiles = (capacity + tile - 1) / tile;
    const u64 rounds = (n_tiles + gridDim.x - 1) / gridDim.x;
    for (u64 it = 0; it < rounds; ++it) {
        const u64 t = it * gridDim.x + blockIdx.x;
        K keys[ITEMS];
        u32 cnts[ITEMS];
        u32 n = 0;
        for (u32 j = 0; j < ITEMS; ++j) {
            u64 i = t * tile + (u64)j * blockDim.x + threadIdx.x;
            if (t >= n_tiles || i >= capacity) continue;
            K key = slot_key<K>(table + i);
            if (key == ga_empty_key<K>()) continue;
            u32 c = table[i].val;
            long long score = c;
            if (sk.rows > 0) score = ga_sketch_estimate(sk, ga_murmur_key<K>(key, w, sym_bits, lut));
            if (score > min_exclusive) {
                keys[n] = key;
                cnts[n] = c;
                ++n;
            }
        }
        u64 pos = ga_block_append(n_out, n);
        for (u32 j = 0; j < n; ++j) {
            if (keys_out) keys_out[pos + j] = keys[j];
            if (counts_out) counts_out[pos + j] = cnts[j];
        }
    }
}

template <class K>
__global__ void lookup_kernel(const Slot<K>* __restrict__ table, u64 capacity, const K* __restrict__ keys,
                              u64 n, u32* __restrict__ out) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u32 v = ga_table_find(table, capacity, keys[i]);
        out[i] = v == GA_NONE32 ? 0u : v;
    }
}

// owner rank of a key for the hash-partitioned exchange: independent of the slot hash's top bits
template <class K>
__global__ void key_owner_kernel(const K* __restrict__ keys, u64 n, u32 n_parts, int* __restrict__ owner) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 h = ga_key_hash(keys[i]);
        owner[i] = (int)__umul64hi((h << 32) | (h >> 32), (u64)n_parts);
    }
}

int check_table_args(const char* fn, const void* table, u64 capacity, int key_words) {
    if (!table || capacity == 0 || (key_words != 1 && key_words != 2)) {
        ga_set_error("%s: bad table arguments (capacity=%llu key_words=%d)", fn, capacity, key_words);
        return GA_ERR_BAD_ARG;
    }
    return GA_OK;
}

}  // namespace

extern "C" int ga_key_words(int k, int sym_bits) {
    if (k < 2 || sym_bits < 1 || sym_bits > 8) return 0;
    long bits = (long)(k - 1) * sym_bits;
    if (bits <= 63) return 1;
    if (bits <= 127) return 2;
    return 0;
}

extern "C" int ga_slot_bytes(int key_words) { return key_words == 1 ? 16 : key_words == 2 ? 32 : 0; }

extern "C" int ga_table_clear(void* table_dev, uint64_t capacity, int key_words, ga_stream stream) {
    int rc = check_table_args("ga_table_clear", table_dev, capacity, key_words);
    if (rc) return rc;
    unsigned grid = ga_grid(capacity, 256);
    if (grid > 148u * 16u) grid = 148u * 16u;
    if (key_words == 1) clear_kernel<u64><<<grid, 256, 0, (cudaStream_t)stream>>>((Slot<u64>*)table_dev, capacity);
    else clear_kernel<u128><<<grid, 256, 0, (cudaStream_t)stream>>>((Slot<u128>*)table_dev, capacity);
    GA_LAUNCH_CHECK("clear");
    return GA_OK;
}

extern "C" int ga_count_kmers(const ga_reads* reads, int k, void* table_dev, uint64_t capacity,
                              uint32_t* status_dev, ga_stream stream) {
    if (!reads) return GA_ERR_BAD_ARG;
    int kw = ga_key_words(k, reads->sym_bits);
    int rc = check_table_args("ga_count_kmers", table_dev, capacity, kw);
    if (rc) return rc;
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    const int w = k - 1;
    unsigned grid = ga_grid(rv.n_reads, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define GA_COUNT(K, SB) count_kernel<K, SB><<<grid, 256, 0, st>>>(rv, w, (Slot<K>*)table_dev, capacity, status_dev)
    if (kw == 1 && rv.storage_bits == 2) GA_COUNT(u64, 2);
    else if (kw == 1) GA_COUNT(u64, 8);
    else if (rv.storage_bits == 2) GA_COUNT(u128, 2);
    else GA_COUNT(u128, 8);
#undef GA_COUNT
    GA_LAUNCH_CHECK("count");
    return GA_OK;
}

extern "C" int ga_count_keys(const void* keys_dev, const uint32_t* amounts_dev, uint64_t n, int key_words,
                             void* table_dev, uint64_t capacity, uint32_t* status_dev, ga_stream stream) {
    int rc = check_table_args("ga_count_keys", table_dev, capacity, key_words);
    if (rc) return rc;
    if (n == 0) return GA_OK;
    unsigned grid = ga_grid(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (key_words == 1)
        count_keys_kernel<u64><<<grid, 256, 0, st>>>((const u64*)keys_dev, amounts_dev, n, (Slot<u64>*)table_dev, capacity, status_dev);
    else
        count_keys_kernel<u128><<<grid, 256, 0, st>>>((const u128*)keys_dev, amounts_dev, n, (Slot<u128>*)table_dev, capacity, status_dev);
    GA_LAUNCH_CHECK("count_keys");
    return GA_OK;
}

extern "C" int ga_table_insert_ids(const void* keys_dev, uint64_t n, int key_words, uint32_t id_base,
                                   void* table_dev, uint64_t capacity, uint32_t* status_dev,
                                   ga_stream stream) {
    int rc = check_table_args("ga_table_insert_ids", table_dev, capacity, key_words);
    if (rc) return rc;
    if (n == 0) return GA_OK;
    unsigned grid = ga_grid(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (key_words == 1)
        insert_ids_kernel<u64><<<grid, 256, 0, st>>>((const u64*)keys_dev, n, id_base, (Slot<u64>*)table_dev, capacity, status_dev);
    else
        insert_ids_kernel<u128><<<grid, 256, 0, st>>>((const u128*)keys_dev, n, id_base, (Slot<u128>*)table_dev, capacity, status_dev);
    GA_LAUNCH_CHECK("insert_ids");
    return GA_OK;
}

extern "C" int ga_key_owner(const void* keys_dev, uint64_t n, int key_words, uint32_t n_parts, int32_t* owner_dev,
                            ga_stream stream) {
    if ((key_words != 1 && key_words != 2) || n_parts == 0 || (n && (!keys_dev || !owner_dev))) {
        ga_set_error("ga_key_owner: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    if (n == 0) return GA_OK;
    unsigned grid = ga_grid(n, 256);
    if (key_words == 1) key_owner_kernel<u64><<<grid, 256, 0, (cudaStream_t)stream>>>((const u64*)keys_dev, n, n_parts, owner_dev);
    else key_owner_kernel<u128><<<grid, 256, 0, (cudaStream_t)stream>>>((const u128*)keys_dev, n, n_parts, owner_dev);
    GA_LAUNCH_CHECK("key_owner");
    return GA_OK;
}

extern "C" int ga_table_summary(const void* table_dev, uint64_t capacity, int key_words, int64_t threshold,
                                uint64_t* out4_dev, ga_stream stream) {
    int rc = check_table_args("ga_table_summary", table_dev, capacity, key_words);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GA_CUDA(cudaMemsetAsync(out4_dev, 0, 4 * sizeof(u64), st));
    unsigned grid = ga_grid(capacity, 256);
    if (grid > 148u * 16u) grid = 148u * 16u;
    if (key_words == 1) summary_kernel<u64><<<grid, 256, 0, st>>>((const Slot<u64>*)table_dev, capacity, threshold, (u64*)out4_dev);
    else summary_kernel<u128><<<grid, 256, 0, st>>>((const Slot<u128>*)table_dev, capacity, threshold, (u64*)out4_dev);
    GA_LAUNCH_CHECK("summary");
    return GA_OK;
}

static int launch_export(const void* table_dev, u64 capacity, int key_words, long long min_exclusive,
                         const ga_sketch* sketch, int k, int sym_bits, const u8* lut_dev, void* keys_out,
                         u32* counts_out, u64* n_out, cudaStream_t st) {
    SketchView sk;
    if (sketch) sk = ga_sketch_view(sketch);
    else {
        sk.cells = nullptr;
        sk.rows = 0;
    }
    unsigned grid = ga_grid(capacity, 256);
    if (grid > 148u * 16u) grid = 148u * 16u;
    if (key_words == 1)
        export_kernel<u64><<<grid, 256, 0, st>>>((const Slot<u64>*)table_dev, capacity, min_exclusive, sk,
                                                 k - 1, sym_bits, lut_dev, (u64*)keys_out, counts_out, n_out);
    else
        export_kernel<u128><<<grid, 256, 0, st>>>((const Slot<u128>*)table_dev, capacity, min_exclusive, sk,
                                                  k - 1, sym_bits, lut_dev, (u128*)keys_out, counts_out, n_out);
    GA_LAUNCH_CHECK("export");
    return GA_OK;
}

extern "C" int ga_table_export(const void* table_dev, uint64_t capacity, int key_words, int64_t min_exclusive,
                               void* keys_out_dev, uint32_t* counts_out_dev, uint64_t* n_out_dev,
                               ga_stream stream) {
    int rc = check_table_args("ga_table_export", table_dev, capacity, key_words);
    if (rc) return rc;
    return launch_export(table_dev, capacity, key_words, min_exclusive, nullptr, 2, 2, nullptr, keys_out_dev,
                         counts_out_dev, (u64*)n_out_dev, (cudaStream_t)stream);
}

extern "C" int ga_select_solid(const void* table_dev, uint64_t capacity, int key_words, int k, int sym_bits,
                               int64_t threshold, const ga_sketch* sketch, const uint8_t* lut_dev,
                               void* keys_out_dev, uint32_t* counts_out_dev, uint64_t* n_out_dev,
                               ga_stream stream) {
    int rc = check_table_args("ga_select_solid", table_dev, capacity, key_words);
    if (rc) return rc;
    if (sketch && (sketch->rows < 1 || sketch->rows > GA_MAX_SKETCH_ROWS || !lut_dev)) {
        ga_set_error("ga_select_solid: bad sketch arguments");
        return GA_ERR_BAD_ARG;
    }
    return launch_export(table_dev, capacity, key_words, threshold, sketch, k, sym_bits, lut_dev,
                         keys_out_dev, counts_out_dev, (u64*)n_out_dev, (cudaStream_t)stream);
}

extern "C" int ga_table_lookup(const void* table_dev, uint64_t capacity, int key_words, const void* keys_dev,
                               uint64_t n, uint32_t* counts_out_dev, ga_stream stream) {
    int rc = check_table_args("ga_table_lookup", table_dev, capacity, key_words);
    if (rc) return rc;
    if (n == 0) return GA_OK;
    unsigned grid = ga_grid(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (key_words == 1)
        lookup_kernel<u64><<<grid, 256, 0, st>>>((const Slot<u64>*)table_dev, capacity, (const u64*)keys_dev, n, counts_out_dev);
    else
        lookup_kernel<u128><<<grid, 256, 0, st>>>((const Slot<u128>*)table_dev, capacity, (const u128*)keys_dev, n, counts_out_dev);
    GA_LAUNCH_CHECK("lookup");
    return GA_OK;
}
