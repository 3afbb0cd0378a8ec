// ga_build.cu -- edge / node insertion with first-occurrence stamps, replacing
// DeBruijnGraph._build_graph (debruijn_graph.py:113-142) and
// PairedDeBruijnGraph._build_graph (:269-317).  The reference's dict insertion order is
// reproduced by stamps: occurrence e = read * estride + position,
// node stamp = min(2e | prefix, 2e+1 | suffix), edge stamp = min e.  Ordering itself happens in
// ga_csr.cu; these kernels only fold stamps in with atomicMin.
#include "ga_common.cuh"

namespace {

__device__ __forceinline__ void stamp_min(u64* p, u64 v) {
    if (v < __ldcg(p)) atomicMin(p, v);
}

// Unpaired: one thread per read.
template <class K, int SB>
__global__ void __launch_bounds__(256)
build_unpaired_kernel(ReadsView rv, int w, const Slot<K>* __restrict__ solid, u64 solid_cap,
                      u64* __restrict__ node_stamp, StampSlot* __restrict__ edges, u64 edge_cap,
                      u32* status) {
    const K mask = ga_key_mask<K>(w, rv.sym_bits);
    bool full = false;
    GA_FOR_EACH_READ_WARP(rv, r, valid) {
        u32 len = valid ? ga_read_len(rv, r) : 0u;
        if (len <= (u32)w) len = 0;  // fewer than two windows: no edge occurrence
        const u64 e0 = (rv.first_read + r) * (u64)rv.estride;
        u32 prev = GA_NONE32;
        ga_for_each_window<K, SB>(valid ? ga_read_ptr(rv, r) : rv.words, len, w, rv.sym_bits, mask, [&](u32 pos, K key) {
            u32 id = ga_table_find(solid, solid_cap, key);
            if (pos > 0 && prev != GA_NONE32 && id != GA_NONE32) {
                const u64 e = e0 + (pos - 1);
                stamp_min(node_stamp + prev, 2 * e);
                stamp_min(node_stamp + id, 2 * e + 1);
                if (ga_stamp_upsert(edges, edge_cap, ((u64)prev << 32) | id, e) == GA_NONE64) full = true;
            }
            prev = id;
        });
    }
    if (full) atomicOr(status, GA_ST_STAMP_FULL);
}

// ---- unpaired, alphabets of at most 4 symbols (DNA): no edge table ---------------------------
// aux word of an id-table slot: five 6-bit fields, each the earliest read epoch (0..61, 63 = none)
// that has already folded in the corresponding stamp: field 0 the node stamp, field 1+s the stamp of
// the edge leaving through symbol s.  Epochs partition reads by index, so a stamp offered from a
// later epoch is larger than the one the recorded epoch folded in (or will fold in before the kernel
// ends): the later occurrence can stop at the slot.  That is the common case (coverage >> 1) and it
// keeps the per-occurrence working set to the id table, which fits the L2.
#define GA_EPOCH_NONE 63u
#define GA_EPOCHS 62u
__device__ __forceinline__ u32 epoch_field(u32 aux, u32 f) { return (aux >> (6u * f)) & 63u; }

// Lower the selected fields of *p to `epoch` (true minimum under concurrency: CAS loop).
__device__ __forceinline__ u32 epoch_lower(u32* p, u32 seen, u32 fields, u32 epoch) {
    u32 old = seen;
    for (;;) {
        u32 neu = old;
        for (u32 f = 0; f < 5u; ++f)
            if (((fields >> f) & 1u) && epoch_field(neu, f) > epoch)
                neu = (neu & ~(63u << (6u * f))) | (epoch << (6u * f));
        if (neu == old) return old;
        u32 prev = atomicCAS(p, old, neu);
        if (prev == old) return neu;
        old = prev;
    }
}

template <class K> __device__ __forceinline__ u32* slot_aux(Slot<K>* s) { return &s->aux; }

template <class K, int SB>
__global__ void __launch_bounds__(256)
build_unpaired_dna_kernel(ReadsView rv, int w, Slot<K>* solid, u64 solid_cap, u64* __restrict__ node_stamp,
                          u64* __restrict__ edge_stamp, u64 reads_per_epoch) {
    const K mask = ga_key_mask<K>(w, rv.sym_bits);
    const u32 smask = (1u << rv.sym_bits) - 1u;
    GA_FOR_EACH_READ_WARP(rv, r, valid) {
        u32 len = valid ? ga_read_len(rv, r) : 0u;
        if (len <= (u32)w) len = 0;
        const u64 e0 = (rv.first_read + r) * (u64)rv.estride;
        u64 ep64 = r / reads_per_epoch;
        const u32 epoch = ep64 < GA_EPOCHS ? (u32)ep64 : GA_EPOCHS - 1u;
        u64 prev_slot = GA_NONE64;
        u32 prev_id = 0, prev_aux = 0;
        ga_for_each_window<K, SB>(valid ? ga_read_ptr(rv, r) : rv.words, len, w, rv.sym_bits, mask, [&](u32 pos, K key) {
            u32 id, aux;
            u64 slot = ga_table_find_slot(solid, solid_cap, key, id, aux);
            if (pos > 0 && prev_slot != GA_NONE64 && slot != GA_NONE64) {
                const u64 e = e0 + (pos - 1);
                const u32 sym = (u32)key & smask;
                u32 want = 0;
                if (epoch_field(prev_aux, 0) >= epoch) {
                    atomicMin(node_stamp + prev_id, 2 * e);
                    want |= 1u;
                }
                if (epoch_field(prev_aux, 1u + sym) >= epoch) {
                    atomicMin(edge_stamp + 4ull * prev_id + sym, e);
                    want |= 2u << sym;
                }
                if (want) {
                    u32 neu = epoch_lower(slot_aux(solid + prev_slot), prev_aux, want, epoch);
                    if (slot == prev_slot) aux = neu;   // homopolymer: prefix and suffix share a slot
                }
                if (epoch_field(aux, 0) >= epoch) {
                    atomicMin(node_stamp + id, 2 * e + 1);
                    aux = epoch_lower(slot_aux(solid + slot), aux, 1u, epoch);
                }
            }
            prev_slot = slot;
            prev_id = id;
            prev_aux = aux;
        });
    }
}

// ---- two-phase build: what is still unstamped after a prefix of the reads ---------------------
__device__ __forceinline__ void bloom_bits(u64 h, u64 n_words, u64& word, u32& bits) {
    word = ga_slot_of(h, n_words);
    bits = (1u << (h & 31u)) | (1u << ((h >> 5) & 31u)) | (1u << ((h >> 10) & 31u));
}

template <class K>
__global__ void unstamped_scan_kernel(const Slot<K>* __restrict__ solid, u64 solid_cap, const K* __restrict__ keys,
                                      u64 n_solid, const u64* __restrict__ edge_stamp, int w, int sym_bits,
                                      u8* __restrict__ mask_out, u64* n_open) {
    const K mask = ga_key_mask<K>(w, sym_bits);
    const u32 n_sym = 1u << sym_bits;
    u64 mine = 0;
    for (u64 id = blockIdx.x * (u64)blockDim.x + threadIdx.x; id < n_solid; id += (u64)gridDim.x * blockDim.x) {
        const K key = keys[id];
        u32 open = 0;
        for (u32 c = 0; c < n_sym; ++c) {
            if (edge_stamp[4 * id + c] != GA_NONE64) continue;
            K succ = ((key << sym_bits) | (K)c) & mask;
            if (ga_table_find(solid, solid_cap, succ) != GA_NONE32) open |= 1u << c;
        }
        mask_out[id] = (u8)open;
        mine += open != 0;
    }
    for (int off = 16; off > 0; off >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, off);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_open, mine);
}

template <class K>
__global__ void unstamped_table_kernel(const K* __restrict__ keys, const u8* __restrict__ open_mask, u64 n_solid,
                                       Slot<K>* __restrict__ table, u64 capacity, u32* __restrict__ bloom,
                                       u64 bloom_words, u32* status) {
    bool full = false;
    for (u64 id = blockIdx.x * (u64)blockDim.x + threadIdx.x; id < n_solid; id += (u64)gridDim.x * blockDim.x) {
        u32 open = open_mask[id];
        if (!open) continue;
        const K key = keys[id];
        u64 s = ga_table_upsert(table, capacity, key);
        if (s == GA_NONE64) {
            full = true;
            continue;
        }
        table[s].val = (u32)id;
        table[s].aux = open;
        u64 word;
        u32 bits;
        bloom_bits(ga_key_hash(key), bloom_words, word, bits);
        atomicOr(bloom + word, bits);
    }
    if (full) atomicOr(status, GA_ST_TABLE_FULL);
}

// Remaining reads: an occurrence matters only if its prefix node still has an open edge through the
// next symbol.  One Bloom word per window; hits (about 1 in 100) go to the open-node table.
template <class K, int SB>
__global__ void __launch_bounds__(256)
build_dna_tail_kernel(ReadsView rv, int w, const u32* __restrict__ bloom, u64 bloom_words,
                      const Slot<K>* __restrict__ open_table, u64 open_cap, const Slot<K>* __restrict__ solid,
                      u64 solid_cap, u64* __restrict__ node_stamp, u64* __restrict__ edge_stamp) {
    const K mask = ga_key_mask<K>(w, rv.sym_bits);
    const u32 smask = (1u << rv.sym_bits) - 1u;
    GA_FOR_EACH_READ_WARP(rv, r, valid) {
        u32 len = valid ? ga_read_len(rv, r) : 0u;
        if (len <= (u32)w) len = 0;
        const u64 e0 = (rv.first_read + r) * (u64)rv.estride;
        bool prev_hit = false;
        K prev_key = 0;
        ga_for_each_window<K, SB>(valid ? ga_read_ptr(rv, r) : rv.words, len, w, rv.sym_bits, mask, [&](u32 pos, K key) {
            if (prev_hit) {   // the previous window may be an open node: is its edge through my last symbol open?
                K pk;
                u32 id, open;
                u64 s = ga_slot_of(ga_key_hash(prev_key), open_cap);
                for (u64 probes = 0; probes < open_cap; ++probes) {
                    const Slot<K>* slot = open_table + s;
                    uint4 raw = __ldg(reinterpret_cast<const uint4*>(slot));
                    if (sizeof(K) == 8) {
                        pk = (K)(((u64)raw.y << 32) | raw.x);
                        id = raw.z;
                        open = raw.w;
                    } else {
                        uint4 raw2 = __ldg(reinterpret_cast<const uint4*>(slot) + 1);
                        pk = (K)(((u128)(((u64)raw.w << 32) | raw.z) << 64) | (((u64)raw.y << 32) | raw.x));
                        id = raw2.x;
                        open = raw2.y;
                    }
                    if (pk == prev_key) {
                        const u32 sym = (u32)key & smask;
                        if ((open >> sym) & 1u) {
                            const u64 e = e0 + (pos - 1);
                            u32 succ = ga_table_find(solid, solid_cap, key);   // solid by construction of the mask
                            atomicMin(edge_stamp + 4ull * id + sym, e);
                            atomicMin(node_stamp + id, 2 * e);
                            if (succ != GA_NONE32) atomicMin(node_stamp + succ, 2 * e + 1);
                        }
                        break;
                    }
                    if (pk == ga_empty_key<K>()) break;
                    if (++s == open_cap) s = 0;
                }
            }
            u64 word;
            u32 bits;
            bloom_bits(ga_key_hash(key), bloom_words, word, bits);
            prev_hit = (__ldg(bloom + word) & bits) == bits;
            prev_key = key;
        });
    }
}

// Keep the two smallest values ever offered (all values distinct): m[0] <= m[1].
__device__ __forceinline__ void two_min(u64* m, u64 v) {
    u64 old = atomicMin(m, v);
    u64 loser = old > v ? old : v;
    if (loser != GA_NONE64) atomicMin(m + 1, loser);
}

// Paired: one thread per pair; mates advance in lock step over mate 1's length.
template <class K, int SB>
__global__ void __launch_bounds__(256)
build_paired_kernel(ReadsView rv, int w, const Slot<K>* __restrict__ solid, u64 solid_cap,
                    StampSlot* __restrict__ queries, u64 query_cap, StampSlot* __restrict__ qedges,
                    u64 qedge_cap, u64* __restrict__ dh, u32* status) {
    constexpr u32 SPW = 64 / SB;
    constexpr u64 SMASK = (1ull << SB) - 1;
    const K mask = ga_key_mask<K>(w, rv.sym_bits);
    const u32 smask = (1u << rv.sym_bits) - 1u;
    const u64 n_pairs = rv.n_reads / 2;
    bool full = false;
    const u32 lane = threadIdx.x & 31u;
    for (u64 pb = blockIdx.x * (u64)blockDim.x + (threadIdx.x - lane); pb < n_pairs;
         pb += (u64)gridDim.x * blockDim.x) {
        const u64 p = pb + lane;
        const bool valid = p < n_pairs;
        u32 len = valid ? ga_read_len(rv, 2 * p) : 0u;
        if (len <= (u32)w) len = 0;
        const u64* wa = valid ? ga_read_ptr(rv, 2 * p) : rv.words;
        const u64* wb = valid ? ga_read_ptr(rv, 2 * p + 1) : rv.words;
        const u64 e0 = (rv.first_read + p) * (u64)rv.estride;
        const u32 max_len = __reduce_max_sync(0xFFFFFFFFu, len);   // warp-uniform walk, see ga_for_each_window
        K ka = 0, kb = 0;
        u32 prev_a = GA_NONE32, prev_b = GA_NONE32;
        u64 prev_q = GA_NONE64;  // query slot of the previous position when it was part of an accepted occurrence
        for (u32 base = 0; base < max_len; base += SPW) {
            u64 word_a = base < len ? __ldg(wa + base / SPW) : 0ull;
            u64 word_b = base < len ? __ldg(wb + base / SPW) : 0ull;
            u32 lim = max_len - base < SPW ? max_len - base : SPW;
            for (u32 j = 0; j < lim; ++j) {
                ka = ((ka << rv.sym_bits) | (K)(word_a & SMASK)) & mask;
                kb = ((kb << rv.sym_bits) | (K)(word_b & SMASK)) & mask;
                word_a >>= SB;
                word_b >>= SB;
                u32 i = base + j + 1;
                __syncwarp();
                if (i < (u32)w || i > len) continue;
                u32 pos = i - (u32)w;
                u32 ida = ga_table_find(solid, solid_cap, ka);
                u32 idb = ida == GA_NONE32 ? GA_NONE32 : ga_table_find(solid, solid_cap, kb);
                bool here = ida != GA_NONE32 && idb != GA_NONE32;
                u64 cur_q = GA_NONE64;
                if (pos > 0 && here && prev_a != GA_NONE32 && prev_b != GA_NONE32) {
                    const u64 e = e0 + (pos - 1);
                    // prefix pair: already folded in with a smaller stamp if the previous
                    // occurrence was accepted too (it was that occurrence's suffix pair)
                    u64 qp = prev_q;
                    if (qp == GA_NONE64)
                        qp = ga_stamp_upsert(queries, query_cap, ((u64)prev_a << 32) | prev_b, 2 * e);
                    u64 qs = ga_stamp_upsert(queries, query_cap, ((u64)ida << 32) | idb, 2 * e + 1);
                    if (qp == GA_NONE64 || qs == GA_NONE64) full = true;
                    else {
                        if (ga_stamp_upsert(qedges, qedge_cap, (qp << 32) | qs, e) == GA_NONE64) full = true;
                        if (prev_a == ida && prev_b == idb)  // both mates homopolymer (App. A-9 ii)
                            two_min(dh + 2 * ((u64)((u32)ka & smask) * 256u + ((u32)kb & smask)), e);
                    }
                    cur_q = qs;
                }
                prev_a = ida;
                prev_b = idb;
                prev_q = cur_q;
            }
        }
    }
    if (full) atomicOr(status, GA_ST_STAMP_FULL);
}

}  // namespace

extern "C" int ga_build_unpaired(const ga_reads* reads, int k, const void* solid_dev, uint64_t solid_capacity,
                                 uint64_t* node_stamp_dev, void* edge_table_dev, uint64_t edge_capacity,
                                 uint32_t* status_dev, ga_stream stream) {
    if (!reads || !solid_dev || !node_stamp_dev || !edge_table_dev || solid_capacity == 0 || edge_capacity == 0) {
        ga_set_error("ga_build_unpaired: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    int kw = ga_key_words(k, reads->sym_bits);
    if (!kw || reads->paired) {
        ga_set_error("ga_build_unpaired: unsupported key width or paired reads");
        return GA_ERR_BAD_ARG;
    }
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    unsigned grid = ga_grid(rv.n_reads, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define GA_BUILD(K, SB)                                                                              \
    build_unpaired_kernel<K, SB><<<grid, 256, 0, st>>>(rv, k - 1, (const Slot<K>*)solid_dev, solid_capacity, \
                                                      (u64*)node_stamp_dev, (StampSlot*)edge_table_dev, \
                                                      edge_capacity, status_dev)
    if (kw == 1 && rv.storage_bits == 2) GA_BUILD(u64, 2);
    else if (kw == 1) GA_BUILD(u64, 8);
    else if (rv.storage_bits == 2) GA_BUILD(u128, 2);
    else GA_BUILD(u128, 8);
#undef GA_BUILD
    GA_LAUNCH_CHECK("build_unpaired");
    return GA_OK;
}

extern "C" int ga_build_unpaired_dna(const ga_reads* reads, int k, void* solid_dev, uint64_t solid_capacity,
                                     uint64_t* node_stamp_dev, uint64_t* edge_stamp_dev, uint32_t* status_dev,
                                     ga_stream stream) {
    (void)status_dev;
    if (!reads || !solid_dev || !node_stamp_dev || !edge_stamp_dev || solid_capacity == 0) {
        ga_set_error("ga_build_unpaired_dna: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    int kw = ga_key_words(k, reads->sym_bits);
    if (!kw || reads->paired || reads->sym_bits > 2) {
        ga_set_error("ga_build_unpaired_dna: needs unpaired reads over at most 4 symbols");
        return GA_ERR_BAD_ARG;
    }
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    u64 reads_per_epoch = (rv.n_reads + GA_EPOCHS - 1) / GA_EPOCHS;
    if (reads_per_epoch == 0) reads_per_epoch = 1;
    unsigned grid = ga_grid(rv.n_reads, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define GA_BUILD(K, SB)                                                                                  \
    build_unpaired_dna_kernel<K, SB><<<grid, 256, 0, st>>>(rv, k - 1, (Slot<K>*)solid_dev, solid_capacity, \
                                                          (u64*)node_stamp_dev, (u64*)edge_stamp_dev,     \
                                                          reads_per_epoch)
    if (kw == 1 && rv.storage_bits == 2) GA_BUILD(u64, 2);
    else if (kw == 1) GA_BUILD(u64, 8);
    else if (rv.storage_bits == 2) GA_BUILD(u128, 2);
    else GA_BUILD(u128, 8);
#undef GA_BUILD
    GA_LAUNCH_CHECK("build_unpaired_dna");
    return GA_OK;
}

extern "C" int ga_unstamped_scan(const void* solid_dev, uint64_t solid_capacity, const void* solid_keys_dev,
                                 uint64_t n_solid, int key_words, const uint64_t* edge_stamp_dev, int k,
                                 int sym_bits, uint8_t* mask_out_dev, uint64_t* n_open_dev, ga_stream stream) {
    if (!solid_dev || !solid_keys_dev || !edge_stamp_dev || !mask_out_dev || !n_open_dev || sym_bits > 2 ||
        (key_words != 1 && key_words != 2) || solid_capacity == 0) {
        ga_set_error("ga_unstamped_scan: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    if (n_solid == 0) return GA_OK;
    unsigned grid = ga_grid(n_solid, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (key_words == 1)
        unstamped_scan_kernel<u64><<<grid, 256, 0, st>>>((const Slot<u64>*)solid_dev, solid_capacity, (const u64*)solid_keys_dev,
                                                        n_solid, (const u64*)edge_stamp_dev, k - 1, sym_bits, mask_out_dev, (u64*)n_open_dev);
    else
        unstamped_scan_kernel<u128><<<grid, 256, 0, st>>>((const Slot<u128>*)solid_dev, solid_capacity, (const u128*)solid_keys_dev,
                                                         n_solid, (const u64*)edge_stamp_dev, k - 1, sym_bits, mask_out_dev, (u64*)n_open_dev);
    GA_LAUNCH_CHECK("unstamped_scan");
    return GA_OK;
}

extern "C" int ga_unstamped_table_build(const void* solid_keys_dev, const uint8_t* mask_dev, uint64_t n_solid,
                                        int key_words, void* open_table_dev, uint64_t open_capacity,
                                        uint32_t* bloom_dev, uint64_t bloom_words, uint32_t* status_dev,
                                        ga_stream stream) {
    if (!solid_keys_dev || !mask_dev || !open_table_dev || !bloom_dev || open_capacity == 0 || bloom_words == 0 ||
        (key_words != 1 && key_words != 2)) {
        ga_set_error("ga_unstamped_table_build: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    if (n_solid == 0) return GA_OK;
    unsigned grid = ga_grid(n_solid, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (key_words == 1)
        unstamped_table_kernel<u64><<<grid, 256, 0, st>>>((const u64*)solid_keys_dev, mask_dev, n_solid, (Slot<u64>*)open_table_dev,
                                                         open_capacity, bloom_dev, bloom_words, status_dev);
    else
        unstamped_table_kernel<u128><<<grid, 256, 0, st>>>((const u128*)solid_keys_dev, mask_dev, n_solid, (Slot<u128>*)open_table_dev,
                                                          open_capacity, bloom_dev, bloom_words, status_dev);
    GA_LAUNCH_CHECK("unstamped_table");
    return GA_OK;
}

extern "C" int ga_build_unpaired_dna_tail(const ga_reads* reads, int k, const uint32_t* bloom_dev, uint64_t bloom_words,
                                          const void* open_table_dev, uint64_t open_capacity, const void* solid_dev,
                                          uint64_t solid_capacity, uint64_t* node_stamp_dev, uint64_t* edge_stamp_dev,
                                          ga_stream stream) {
    if (!reads || !bloom_dev || !open_table_dev || !solid_dev || !node_stamp_dev || !edge_stamp_dev ||
        bloom_words == 0 || open_capacity == 0 || solid_capacity == 0) {
        ga_set_error("ga_build_unpaired_dna_tail: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    int kw = ga_key_words(k, reads->sym_bits);
    if (!kw || reads->paired || reads->sym_bits > 2) {
        ga_set_error("ga_build_unpaired_dna_tail: needs unpaired reads over at most 4 symbols");
        return GA_ERR_BAD_ARG;
    }
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    unsigned grid = ga_grid(rv.n_reads, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define GA_TAIL(K, SB)                                                                                       \
    build_dna_tail_kernel<K, SB><<<grid, 256, 0, st>>>(rv, k - 1, bloom_dev, bloom_words,                     \
                                                      (const Slot<K>*)open_table_dev, open_capacity,         \
                                                      (const Slot<K>*)solid_dev, solid_capacity,             \
                                                      (u64*)node_stamp_dev, (u64*)edge_stamp_dev)
    if (kw == 1 && rv.storage_bits == 2) GA_TAIL(u64, 2);
    else if (kw == 1) GA_TAIL(u64, 8);
    else if (rv.storage_bits == 2) GA_TAIL(u128, 2);
    else GA_TAIL(u128, 8);
#undef GA_TAIL
    GA_LAUNCH_CHECK("build_dna_tail");
    return GA_OK;
}

extern "C" int ga_build_paired(const ga_reads* reads, int k, const void* solid_dev, uint64_t solid_capacity,
                               void* query_table_dev, uint64_t query_capacity, void* qedge_table_dev,
                               uint64_t qedge_capacity, uint64_t* dh_dev, uint32_t* status_dev,
                               ga_stream stream) {
    if (!reads || !solid_dev || !query_table_dev || !qedge_table_dev || !dh_dev || solid_capacity == 0 ||
        query_capacity == 0 || qedge_capacity == 0 || query_capacity >= 0xFFFFFFFFull) {
        ga_set_error("ga_build_paired: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    int kw = ga_key_words(k, reads->sym_bits);
    if (!kw || !reads->paired || (reads->n_reads & 1)) {
        ga_set_error("ga_build_paired: unsupported key width or unpaired reads");
        return GA_ERR_BAD_ARG;
    }
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    unsigned grid = ga_grid(rv.n_reads / 2, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define GA_BUILD(K, SB)                                                                            \
    build_paired_kernel<K, SB><<<grid, 256, 0, st>>>(rv, k - 1, (const Slot<K>*)solid_dev, solid_capacity, \
                                                    (StampSlot*)query_table_dev, query_capacity,   \
                                                    (StampSlot*)qedge_table_dev, qedge_capacity,   \
                                                    (u64*)dh_dev, status_dev)
    if (kw == 1 && rv.storage_bits == 2) GA_BUILD(u64, 2);
    else if (kw == 1) GA_BUILD(u64, 8);
    else if (rv.storage_bits == 2) GA_BUILD(u128, 2);
    else GA_BUILD(u128, 8);
#undef GA_BUILD
    GA_LAUNCH_CHECK("build_paired");
    return GA_OK;
}

// ------------------------------------------------------------------------------------------------
// Paired build across GPUs: every rank fills its own query / query-edge tables from its read shard
// (ga_build_paired; stamps carry global read indices, ids come from the replicated solid table).  The
// tables are exported as plain lists -- query-edge keys re-expressed by the two QUERY KEYS they join, because
// slot numbers mean nothing in another rank's table -- gathered, and folded into one pair of tables with
// min(stamp), which is what one GPU walking all reads would have built (min is associative).
namespace {

__global__ void __launch_bounds__(256)
stamp_export_kernel(const StampSlot* __restrict__ table, u64 cap, const StampSlot* __restrict__ queries,
                    u64* __restrict__ key_out, u64* __restrict__ key2_out, u64* __restrict__ stamp_out,
                    u64 out_cap, u64* __restrict__ n_out) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < cap + ((32u - (cap & 31u)) & 31u);
         i += (u64)gridDim.x * blockDim.x) {
        const bool take = i < cap && table[i].key != GA_NONE64;
        const u64 at = ga_warp_append(n_out, take);
        if (!take || at >= out_cap) continue;
        const u64 key = table[i].key;
        if (queries) {       // a query-edge: slots of the two queries -> their keys
            key_out[at] = queries[key >> 32].key;
            key2_out[at] = queries[key & 0xFFFFFFFFull].key;
        } else {
            key_out[at] = key;
        }
        stamp_out[at] = table[i].stamp;
    }
}

__global__ void __launch_bounds__(256)
merge_queries_kernel(const u64* __restrict__ keys, const u64* __restrict__ stamps, u64 n, StampSlot* __restrict__ queries,
                     u64 qcap, u32* status) {
    bool full = false;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        if (ga_stamp_upsert(queries, qcap, keys[i], stamps[i]) == GA_NONE64) full = true;
    if (full) atomicOr(status, GA_ST_STAMP_FULL);
}

__global__ void __launch_bounds__(256)
merge_qedges_kernel(const u64* __restrict__ pkeys, const u64* __restrict__ skeys, const u64* __restrict__ stamps, u64 n,
                    const StampSlot* __restrict__ queries, u64 qcap, StampSlot* __restrict__ qedges, u64 qecap,
                    u32* status) {
    bool full = false;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 qp = ga_stamp_find(queries, qcap, pkeys[i]), qs = ga_stamp_find(queries, qcap, skeys[i]);
        if (qp == GA_NONE64 || qs == GA_NONE64 ||
            ga_stamp_upsert(qedges, qecap, (qp << 32) | qs, stamps[i]) == GA_NONE64)
            full = true;
    }
    if (full) atomicOr(status, GA_ST_STAMP_FULL);
}

}  // namespace

extern "C" int ga_stamp_table_export(const void* table_dev, uint64_t capacity, const void* query_table_dev,
                                     uint64_t* key_out_dev, uint64_t* key2_out_dev, uint64_t* stamp_out_dev,
                                     uint64_t out_capacity, uint64_t* n_out_dev, ga_stream stream) {
    if (!table_dev || capacity == 0 || !key_out_dev || !stamp_out_dev || !n_out_dev ||
        (query_table_dev && !key2_out_dev)) {
        ga_set_error("ga_stamp_table_export: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    unsigned grid = ga_grid(capacity, 256);
    if (grid > 148u * 16u) grid = 148u * 16u;
    stamp_export_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const StampSlot*)table_dev, capacity,
                                                                (const StampSlot*)query_table_dev, (u64*)key_out_dev,
                                                                (u64*)key2_out_dev, (u64*)stamp_out_dev, out_capacity,
                                                                (u64*)n_out_dev);
    GA_LAUNCH_CHECK("stamp_table_export");
    return GA_OK;
}

extern "C" int ga_paired_merge(const uint64_t* query_keys_dev, const uint64_t* query_stamps_dev, uint64_t n_queries,
                               const uint64_t* edge_pkeys_dev, const uint64_t* edge_skeys_dev,
                               const uint64_t* edge_stamps_dev, uint64_t n_edges, void* query_table_dev,
                               uint64_t query_capacity, void* qedge_table_dev, uint64_t qedge_capacity,
                               uint32_t* status_dev, ga_stream stream) {
    if (!query_table_dev || !qedge_table_dev || !status_dev || query_capacity == 0 || qedge_capacity == 0 ||
        query_capacity >= 0xFFFFFFFFull || (n_queries && (!query_keys_dev || !query_stamps_dev)) ||
        (n_edges && (!edge_pkeys_dev || !edge_skeys_dev || !edge_stamps_dev))) {
        ga_set_error("ga_paired_merge: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (n_queries) {
        merge_queries_kernel<<<ga_grid(n_queries, 256), 256, 0, st>>>((const u64*)query_keys_dev, (const u64*)query_stamps_dev,
                                                                    n_queries, (StampSlot*)query_table_dev, query_capacity,
                                                                    status_dev);
        GA_LAUNCH_CHECK("paired_merge_queries");
    }
    if (n_edges) {
        merge_qedges_kernel<<<ga_grid(n_edges, 256), 256, 0, st>>>((const u64*)edge_pkeys_dev, (const u64*)edge_skeys_dev,
                                                                 (const u64*)edge_stamps_dev, n_edges,
                                                                 (const StampSlot*)query_table_dev, query_capacity,
                                                                 (StampSlot*)qedge_table_dev, qedge_capacity, status_dev);
        GA_LAUNCH_CHECK("paired_merge_qedges");
    }
    return GA_OK;
}
