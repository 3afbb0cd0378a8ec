// ga_common.cuh -- shared device/host helpers for the B200 (sm_100a) k-mer counting and
// de Bruijn graph construction kernels.  Included by every .cu file of libga_b200.so.
//
// Vocabulary (follows the reference, tonycheang/genome-assembler):
//   "window"  a (k-1)-mer of a read (debruijn_graph.py:154-157); the reference calls these k-mers
//   "solid"   a window whose count (or sketch estimate) is > the filter threshold (:127-128)
//   "stamp"   first-occurrence ordinal that replaces the reference's dict insertion order
//             (SURVEY App. C): edge occurrence e = read_index * estride + position,
//             node stamp = min(2e as prefix, 2e+1 as suffix), edge stamp = min e.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ga_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;
typedef unsigned __int128 u128;

#define GA_NONE32 0xFFFFFFFFu
#define GA_NONE64 0xFFFFFFFFFFFFFFFFull

// status word bits (device uint32_t status[0])
#define GA_ST_TABLE_FULL 1u
#define GA_ST_BAD_SYMBOL 2u
#define GA_ST_STAMP_FULL 4u
#define GA_ST_U16_OVERFLOW 8u

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
void ga_set_error(const char* fmt, ...);
int ga_cuda_fail(cudaError_t e, const char* what);
#define GA_CUDA(call)                                             \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return ga_cuda_fail(_e, #call);    \
    } while (0)
void ga_pool_retain();          // default memory pool of the current device keeps freed blocks (call before cudaMallocAsync)
void ga_note_launches(int n);   // counts this library's own kernel launches (ga_launch_count)
#define GA_LAUNCH_CHECK(name)        \
    do {                             \
        ga_note_launches(1);         \
        GA_CUDA(cudaGetLastError()); \
    } while (0)

static inline unsigned ga_grid(u64 n, unsigned block) {
    u64 g = (n + block - 1) / block;
    if (g == 0) g = 1;
    if (g > 0x7FFFFFFFull) g = 0x7FFFFFFFull;
    return (unsigned)g;
}

// ---------------------------------------------------------------------------------------------
// hashing
__host__ __device__ __forceinline__ u64 ga_mix64(u64 z) {
    z = (z ^ (z >> 32)) * 0xd6e8feb86659fd93ull;
    z = (z ^ (z >> 32)) * 0xd6e8feb86659fd93ull;
    return z ^ (z >> 32);
}
__host__ __device__ __forceinline__ u64 ga_key_hash(u64 k) { return ga_mix64(k); }
__host__ __device__ __forceinline__ u64 ga_key_hash(u128 k) {
    return ga_mix64((u64)k ^ ga_mix64((u64)(k >> 64) + 0x9e3779b97f4a7c15ull));
}
// slot = floor(hash * capacity / 2^64): any capacity, no modulo
__device__ __forceinline__ u64 ga_slot_of(u64 hash, u64 capacity) { return __umul64hi(hash, capacity); }

// ---------------------------------------------------------------------------------------------
// count / id table slots.  `val` is the occurrence count (count table) or the dense solid id
// (solid table).  An all-ones key marks an empty slot; keys use at most 63 / 127 bits.
template <class K> struct Slot;
template <> struct __align__(16) Slot<u64> {
    u64 key;
    u32 val;
    u32 aux;
};
template <> struct __align__(32) Slot<u128> {
    u64 lo, hi;
    u32 val;
    u32 aux;
    u64 pad;
};

template <class K> __host__ __device__ __forceinline__ K ga_empty_key() { return ~(K)0; }

__device__ __forceinline__ u64 ga_load_key_cg(const Slot<u64>* s) { return __ldcg(&s->key); }
__device__ __forceinline__ u128 ga_load_key_cg(const Slot<u128>* s) {
    ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(s));
    return ((u128)v.y << 64) | v.x;
}
__device__ __forceinline__ u64 ga_cas_key(Slot<u64>* s, u64 desired) {
    return atomicCAS(&s->key, GA_NONE64, desired);
}
__device__ __forceinline__ u128 ga_cas_key(Slot<u128>* s, u128 desired) {
    u64 vlo = (u64)desired, vhi = (u64)(desired >> 64), olo, ohi;
    u64 ones = GA_NONE64;
    asm volatile(
        "{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %2};\n\tmov.b128 v, {%3, %4};\n\t"
        "atom.global.cas.b128 o, [%5], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(olo), "=l"(ohi)
        : "l"(ones), "l"(vlo), "l"(vhi), "l"(s)
        : "memory");
    return ((u128)ohi << 64) | olo;
}
// A 16-byte key is read with one vector load; should the hardware ever split it while a
// 128-bit CAS lands, a half could still show the empty pattern.  Treat any key with an
// all-ones half as "unknown" and let the CAS (atomic on all 16 bytes) decide.
__device__ __forceinline__ bool ga_maybe_empty(u64 k) { return k == GA_NONE64; }
__device__ __forceinline__ bool ga_maybe_empty(u128 k) {
    return (u64)k == GA_NONE64 || (u64)(k >> 64) == GA_NONE64;
}

// Find-or-insert `key`; returns the slot index, or GA_NONE64 if the table is full.
template <class K>
__device__ __forceinline__ u64 ga_table_upsert(Slot<K>* __restrict__ table, u64 capacity, K key) {
    u64 s = ga_slot_of(ga_key_hash(key), capacity);
    for (u64 probes = 0; probes < capacity; ++probes) {
        K cur = ga_load_key_cg(table + s);
        if (cur == key) return s;
        if (ga_maybe_empty(cur)) {
            K old = ga_cas_key(table + s, key);
            if (old == ga_empty_key<K>() || old == key) return s;
        }
        if (++s == capacity) s = 0;
    }
    return GA_NONE64;
}

// Read-only lookup (table no longer being written): slot's val, or GA_NONE32 when absent.
__device__ __forceinline__ void ga_load_slot_ro(const Slot<u64>* s, u64& key, u32& val) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(s));
    key = ((u64)v.y << 32) | v.x;
    val = v.z;
}
__device__ __forceinline__ void ga_load_slot_ro(const Slot<u128>* s, u128& key, u32& val) {
    const uint4* p = reinterpret_cast<const uint4*>(s);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    key = ((u128)(((u64)a.w << 32) | a.z) << 64) | (((u64)a.y << 32) | a.x);
    val = b.x;
}
// Coherent (L2) whole-slot load for tables whose aux word is still being updated.
__device__ __forceinline__ void ga_load_slot_cg(const Slot<u64>* s, u64& key, u32& val, u32& aux) {
    uint4 v = __ldcg(reinterpret_cast<const uint4*>(s));
    key = ((u64)v.y << 32) | v.x;
    val = v.z;
    aux = v.w;
}
__device__ __forceinline__ void ga_load_slot_cg(const Slot<u128>* s, u128& key, u32& val, u32& aux) {
    const uint4* p = reinterpret_cast<const uint4*>(s);
    uint4 a = __ldcg(p), b = __ldcg(p + 1);
    key = ((u128)(((u64)a.w << 32) | a.z) << 64) | (((u64)a.y << 32) | a.x);
    val = b.x;
    aux = b.y;
}
// Slot index of `key` (GA_NONE64 when absent) with its val / aux words.
template <class K>
__device__ __forceinline__ u64 ga_table_find_slot(const Slot<K>* table, u64 capacity, K key, u32& val, u32& aux) {
    u64 s = ga_slot_of(ga_key_hash(key), capacity);
    for (u64 probes = 0; probes < capacity; ++probes) {
        K cur;
        ga_load_slot_cg(table + s, cur, val, aux);
        if (cur == key) return s;
        if (cur == ga_empty_key<K>()) return GA_NONE64;
        if (++s == capacity) s = 0;
    }
    return GA_NONE64;
}

template <class K>
__device__ __forceinline__ u32 ga_table_find(const Slot<K>* __restrict__ table, u64 capacity, K key) {
    u64 s = ga_slot_of(ga_key_hash(key), capacity);
    for (u64 probes = 0; probes < capacity; ++probes) {
        K cur;
        u32 val;
        ga_load_slot_ro(table + s, cur, val);
        if (cur == key) return val;
        if (cur == ga_empty_key<K>()) return GA_NONE32;
        if (++s == capacity) s = 0;
    }
    return GA_NONE32;
}

// ---------------------------------------------------------------------------------------------
// stamp tables: 64-bit key -> 64-bit minimum stamp, both initialised to all-ones.
struct __align__(16) StampSlot {
    u64 key;
    u64 stamp;
};
// Returns the slot index (GA_NONE64 when full) after folding `stamp` in with atomicMin.
// A table that is too small must fail fast (the caller doubles it and runs again): the probe chain is cut at
// GA_STAMP_PROBES slots -- at the <= 50 % load the callers size for, chains stay far below that; walking a full
// table slot by slot for every insert is what made an undersized first attempt take hundreds of milliseconds.
#define GA_STAMP_PROBES 512ull
__device__ __forceinline__ u64 ga_stamp_upsert(StampSlot* __restrict__ table, u64 capacity, u64 key,
                                               u64 stamp) {
    u64 s = ga_slot_of(ga_mix64(key), capacity);
    const u64 limit = capacity < GA_STAMP_PROBES ? capacity : GA_STAMP_PROBES;
    for (u64 probes = 0; probes < limit; ++probes) {
        u64 cur = __ldcg(&table[s].key);
        if (cur == GA_NONE64) {
            u64 old = atomicCAS(&table[s].key, GA_NONE64, key);
            if (old == GA_NONE64) cur = key; else cur = old;
        }
        if (cur == key) {
            if (stamp < __ldcg(&table[s].stamp)) atomicMin(&table[s].stamp, stamp);
            return s;
        }
        if (++s == capacity) s = 0;
    }
    return GA_NONE64;
}
__device__ __forceinline__ u64 ga_stamp_find(const StampSlot* __restrict__ table, u64 capacity, u64 key) {
    u64 s = ga_slot_of(ga_mix64(key), capacity);
    for (u64 probes = 0; probes < capacity; ++probes) {
        u64 cur = table[s].key;
        if (cur == key) return s;
        if (cur == GA_NONE64) return GA_NONE64;
        if (++s == capacity) s = 0;
    }
    return GA_NONE64;
}

// ---------------------------------------------------------------------------------------------
// reads on the device (mirrors ga_reads in include/ga_b200.h)
struct ReadsView {
    const u64* words;
    const u64* offsets;   // per-read first word, or nullptr (uniform: r * stride_words)
    const u32* lengths;   // per-read symbols, or nullptr (uniform_len)
    u64 n_reads;          // reads (unpaired) or mates (paired: 2 * pairs, mate 1 at even index)
    u64 first_read;       // global index of read 0 / pair 0 (for stamps)
    u32 uniform_len, stride_words;
    int storage_bits;     // 2 or 8 bits per stored symbol
    int sym_bits;         // bits per symbol inside a key
    int paired;
    u32 estride;          // stamp stride per read (>= longest read)
};
static inline ReadsView ga_view(const ga_reads* r) {
    ReadsView v;
    v.words = (const u64*)r->words;
    v.offsets = (const u64*)r->offsets;
    v.lengths = (const u32*)r->lengths;
    v.n_reads = r->n_reads;
    v.first_read = r->first_read;
    v.uniform_len = r->uniform_len;
    v.stride_words = r->stride_words;
    v.storage_bits = r->storage_bits;
    v.sym_bits = r->sym_bits;
    v.paired = r->paired;
    v.estride = r->estride;
    return v;
}
__device__ __forceinline__ const u64* ga_read_ptr(const ReadsView& v, u64 r) {
    return v.words + (v.offsets ? v.offsets[r] : r * (u64)v.stride_words);
}
__device__ __forceinline__ u32 ga_read_len(const ReadsView& v, u64 r) {
    return v.lengths ? v.lengths[r] : v.uniform_len;
}

template <class K> __host__ __device__ __forceinline__ K ga_key_mask(int w, int sym_bits) {
    int bits = w * sym_bits;
    return (bits >= (int)(8 * sizeof(K))) ? ~(K)0 : (((K)1 << bits) - 1);
}

// Walk the windows of one read, WARP-UNIFORMLY: all 32 lanes of the warp must call this together
// (a lane without a read passes len = 0).  Every step re-converges the warp with __syncwarp():
// the callbacks contain data-dependent probe loops, and without an explicit convergence point the
// lanes of a warp drift apart for good and each instruction issues for 3-4 lanes instead of 32
// (profiles/r01: "Avg. Threads Executed").  SB = storage bits per symbol (2 or 8).  f(pos, key) is
// called for every window start `pos` (0-based) with the packed key (first symbol most significant).
template <class K, int SB, class F>
__device__ __forceinline__ void ga_for_each_window(const u64* __restrict__ words, u32 len, int w,
                                                   int sym_bits, K mask, F&& f) {
    constexpr u32 SPW = 64 / SB;
    constexpr u64 SMASK = (1ull << SB) - 1;
    const u32 max_len = __reduce_max_sync(0xFFFFFFFFu, len);
    K key = 0;
    for (u32 base = 0; base < max_len; base += SPW) {
        u64 word = base < len ? __ldg(words + base / SPW) : 0ull;
        const u32 lim = max_len - base < SPW ? max_len - base : SPW;
        for (u32 j = 0; j < lim; ++j) {
            u32 c = (u32)(word & SMASK);
            word >>= SB;
            key = ((key << sym_bits) | (K)c) & mask;
            const u32 i = base + j + 1;
            __syncwarp();
            if (i >= (u32)w && i <= len) f(i - (u32)w, key);
        }
    }
}

// Grid-stride loop over reads in which whole warps stay together: `r` is this lane's read and
// `valid` says whether it exists.
#define GA_FOR_EACH_READ_WARP(rv, r, valid)                                                              \
    for (u64 _rb = blockIdx.x * (u64)blockDim.x + (threadIdx.x & ~31u), r = _rb + (threadIdx.x & 31u),   \
             valid = r < (rv).n_reads;                                                                    \
         _rb < (rv).n_reads;                                                                              \
         _rb += (u64)gridDim.x * blockDim.x, r = _rb + (threadIdx.x & 31u), valid = r < (rv).n_reads)

// ---------------------------------------------------------------------------------------------
// MurmurHash3_x86_32, seed 0, over the ASCII bytes of a window (countminsketch.py:46-95).
__host__ __device__ __forceinline__ u32 ga_rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }
__host__ __device__ __forceinline__ u32 ga_mm3_block(u32 h, u32 b) {
    b *= 0xcc9e2d51u;
    b = ga_rotl32(b, 15);
    b *= 0x1b873593u;
    h ^= b;
    h = ga_rotl32(h, 13);
    return h * 5u + 0xe6546b64u;
}
__host__ __device__ __forceinline__ u32 ga_mm3_tail(u32 h, u32 t) {
    t *= 0xcc9e2d51u;
    t = ga_rotl32(t, 15);
    t *= 0x1b873593u;
    return h ^ t;
}
__host__ __device__ __forceinline__ u32 ga_mm3_final(u32 h, u32 n) {
    h ^= n;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    return h ^ (h >> 16);
}
// hash of the window held in `key` (w symbols of sym_bits each, first symbol most significant);
// `lut` maps a symbol code to its byte.
template <class K>
__host__ __device__ __forceinline__ u32 ga_murmur_key(K key, int w, int sym_bits, const u8* lut) {
    u32 h = 0, blk = 0;
    const u32 smask = (1u << sym_bits) - 1u;
    for (int i = 0; i < w; ++i) {
        u32 code = (u32)(key >> ((w - 1 - i) * sym_bits)) & smask;
        blk |= (u32)lut[code] << (8 * (i & 3));
        if ((i & 3) == 3) {
            h = ga_mm3_block(h, blk);
            blk = 0;
        }
    }
    if (w & 3) h = ga_mm3_tail(h, blk);
    return ga_mm3_final(h, (u32)w);
}
__host__ __device__ __forceinline__ u32 ga_murmur_bytes(const u8* p, u32 n) {
    u32 h = 0, blk = 0;
    for (u32 i = 0; i < n; ++i) {
        blk |= (u32)p[i] << (8 * (i & 3));
        if ((i & 3) == 3) {
            h = ga_mm3_block(h, blk);
            blk = 0;
        }
    }
    if (n & 3) h = ga_mm3_tail(h, blk);
    return ga_mm3_final(h, n);
}

// sketch geometry passed by value to kernels (countminsketch.py:15-18, 26-32)
struct SketchView {
    u32* cells;           // all rows back to back, 32-bit working cells
    u32 width[GA_MAX_SKETCH_ROWS];
    u64 row_off[GA_MAX_SKETCH_ROWS];
    int rows;
};
static inline SketchView ga_sketch_view(const ga_sketch* s) {
    SketchView v;
    v.cells = (u32*)s->cells;
    v.rows = s->rows;
    u64 off = 0;
    for (int i = 0; i < GA_MAX_SKETCH_ROWS; ++i) {
        v.width[i] = i < s->rows ? s->width[i] : 1;
        v.row_off[i] = off;
        if (i < s->rows) off += s->width[i];
    }
    return v;
}
__device__ __forceinline__ u32 ga_sketch_estimate(const SketchView& sk, u32 h) {
    u32 est = GA_NONE32;
    for (int r = 0; r < sk.rows; ++r) {
        u32 c = __ldg(sk.cells + sk.row_off[r] + (h % sk.width[r]));
        est = c < est ? c : est;
    }
    return est;
}

// Block-aggregated append for blocks of up to 1024 threads: every thread of the block must call
// it (uniform control flow); returns the first output index reserved for this thread's `n` items.
__device__ __forceinline__ u64 ga_block_append(u64* counter, u32 n) {
    __shared__ u32 warp_sum[32];
    __shared__ u64 block_base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    u32 incl = n;
    for (int off = 1; off < 32; off <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u32 v = lane < nw ? warp_sum[lane] : 0u, inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane < nw) warp_sum[lane] = inc - v;           // exclusive prefix per warp
        u32 total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        if (lane == 0) block_base = total ? atomicAdd(counter, (u64)total) : 0ull;
    }
    __syncthreads();
    u64 pos = block_base + warp_sum[wid] + (incl - n);
    __syncthreads();                                       // shared scratch is reused by the next call
    return pos;
}

// Pre-filter sketch (ga_prefilter.cu): one row of small saturating counters, packed in 32-bit words.
struct PrefilterView {
    u32* words;
    u64 n_cells;
    u32 lg_per;     // log2(cells per word): 3 for 4-bit cells, 2 for 8-bit cells
    u32 cell_bits;  // 4 or 8
    u32 limit;      // threshold + 1: a cell that reached it marks its k-mers as candidates
};
static inline PrefilterView ga_prefilter_view(const ga_prefilter* p, long long threshold) {
    PrefilterView v;
    v.words = (u32*)p->words;
    v.n_cells = p->n_cells;
    v.cell_bits = (u32)p->cell_bits;
    v.lg_per = p->cell_bits == 4 ? 3u : 2u;
    v.limit = (u32)(threshold < 0 ? 0 : threshold + 1);
    return v;
}
__device__ __forceinline__ u32 ga_prefilter_value(const PrefilterView& pf, u64 hash, u64& word_index, u32& shift) {
    u64 cell = ga_slot_of(hash, pf.n_cells);
    word_index = cell >> pf.lg_per;
    shift = (u32)(cell & ((1u << pf.lg_per) - 1u)) * pf.cell_bits;
    return (__ldcg(pf.words + word_index) >> shift) & ((1u << pf.cell_bits) - 1u);
}

// warp-aggregated append: returns this thread's output index (valid only if `take`)
__device__ __forceinline__ u64 ga_warp_append(u64* counter, bool take) {
    unsigned mask = __ballot_sync(0xFFFFFFFFu, take);
    if (mask == 0) return 0;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    u64 base = 0;
    if (lane == leader) base = atomicAdd(counter, (u64)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}
