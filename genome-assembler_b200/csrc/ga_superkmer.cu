// ga_superkmer.cu -- bucketed counting + edge stamping for unpaired DNA reads with 64-bit keys.
//
// Replaces, for that case, both hot loops of the reference: DeBruijnGraph._count_kmers
// (debruijn_graph.py:144-152) and DeBruijnGraph._build_graph (:113-142).
//
// Why: a B200 SM retires one random global (or L2) access per ~2 cycles per lane, so ANY design that
// touches a global hash table or sketch once per (k-1)-mer occurrence is capped near 1.4e11
// accesses/s chip-wide -- several times below the HBM roofline of this path (profiles/r01).  Random
// accesses have to land in shared memory.  So the occurrence stream is cut into buckets small enough
// for an exact shared-memory table:
//
//   1. sk_scatter_reads: a window's bucket is a function of its CONTENT (hash of the smallest m-mer
//      hash inside it), so every occurrence of a window meets in one bucket.  Consecutive windows of
//      a read mostly share their minimizer; a run travels as ONE 24-byte record (up to 64 bases, the
//      ordinal of its first window, the run length) instead of 8 bytes per window.  Records are
//      scattered to <= 1024 level-1 buckets through a shared-memory stage; a global histogram of
//      the full (level-1, level-2) bucket id is kept on the side.
//   2. sk_scatter_buckets: exact offsets from that histogram; each level-1 bucket is split into its
//      <= 1024 level-2 buckets.
//   3. sk_bucket: one CTA per bucket.  Exact counts in a shared-memory open-addressing table; windows
//      with count > threshold are the solid ones; a second walk over the bucket's records takes, for
//      every solid window p and next symbol c, the smallest occurrence ordinal of "p followed by c".
//      Output: solid keys and 4 candidate edge stamps each.
//   4. sk_resolve: edge (p, c) exists iff p[1:]+c is solid too (solidity is a property of the string,
//      so the candidate IS the reference's first-insertion ordinal); node stamps follow from the
//      edge stamps.  tests/superkmer_model.py states this on the CPU.
//
// The result (solid keys, node_stamp, edge_stamp) is what ga_csr_plan_unpaired_dna consumes.
#include "ga_common.cuh"

namespace {

constexpr u32 FULL = 0xFFFFFFFFu;

__host__ __device__ __forceinline__ u32 sk_hash32(u32 x) {
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    return x ^ (x >> 16);
}

// reverse the order of the 32 two-bit symbols of a word (reads store symbol i at bits 2i..2i+1,
// records and keys keep the first symbol most significant)
__device__ __forceinline__ u64 sk_rev2(u64 x) {
    u64 y = __brevll(x);
    return ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
}

// record meta word: ordinal of the first window << 16 | level-2 bucket << 6 | has_next << 5 | (windows - 1)
__device__ __forceinline__ u32 meta_windows(u64 meta) { return (u32)(meta & 31u) + 1u; }
__device__ __forceinline__ bool meta_has_next(u64 meta) { return (meta >> 5) & 1u; }
__device__ __forceinline__ u32 meta_b2(u64 meta) { return (u32)(meta >> 6) & 1023u; }
__device__ __forceinline__ u64 meta_ordinal(u64 meta) { return meta >> 16; }

// ------------------------------------------------------------------------------------------------
// 1. reads -> records in level-1 buckets
constexpr int S1_THREADS = 256;
constexpr int S1_WARPS = S1_THREADS / 32;
constexpr u32 S1_STAGE = 3072;                 // records staged per CTA between flushes
constexpr u32 S1_MARGIN = S1_WARPS * 128;      // most records one round can add
constexpr u32 S1_MAXB = 1024;                  // level-1 buckets

struct S1Shared {
    ulonglong2 bases[S1_STAGE];
    u64 meta[S1_STAGE];
    u32 br[S1_STAGE];              // level-1 bucket << 16 | rank inside this flush
    u64 gbase[S1_MAXB];
    u32 hist[S1_MAXB];
    u64 words[S1_WARPS][8];
    u32 count;
};

__device__ __forceinline__ void s1_flush(S1Shared& sm, u32 n_l1, ulonglong2* __restrict__ out_bases,
                                         u64* __restrict__ out_meta, u64 cap1, u64* __restrict__ cursors,
                                         bool& overflow) {
    __syncthreads();
    const u32 n = sm.count;
    for (u32 p = threadIdx.x; p < n_l1; p += S1_THREADS) {
        const u32 c = sm.hist[p];
        u64 base = 0;
        if (c) {
            base = atomicAdd(&cursors[p], (u64)c);
            if (base + c > cap1) {
                overflow = true;
                base = GA_NONE64;      // dropped: the host retries with larger buckets
            }
        }
        sm.gbase[p] = base;
        sm.hist[p] = 0;
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < n; i += S1_THREADS) {
        const u32 br = sm.br[i];
        const u32 b1 = br >> 16;
        const u64 base = sm.gbase[b1];
        if (base == GA_NONE64) continue;
        const u64 dst = (u64)b1 * cap1 + base + (br & 0xFFFFu);
        out_bases[dst] = sm.bases[i];
        out_meta[dst] = sm.meta[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) sm.count = 0;
    __syncthreads();
}

// One warp per read and round; lanes own window positions lane, lane+32, lane+64, lane+96 of the
// current 128-window chunk.
__global__ void __launch_bounds__(S1_THREADS)
sk_scatter_reads_kernel(ReadsView rv, int w, int m, int l1_bits, int l2_bits, ulonglong2* __restrict__ out_bases,
                        u64* __restrict__ out_meta, u64 cap1, u64* __restrict__ cursors, u32* __restrict__ ghist,
                        u32* status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S1Shared& sm = *reinterpret_cast<S1Shared*>(smem_raw);
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 n_l1 = 1u << l1_bits;
    const int bits = l1_bits + l2_bits;
    const u32 n = (u32)(w - m + 1);                       // m-mers per window, 1..16
    u32 P = 1;
    while (P * 2 <= n) P *= 2;
    const u32 mmask = m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
    bool overflow = false;
    for (u32 p = threadIdx.x; p < n_l1; p += S1_THREADS) sm.hist[p] = 0;
    if (threadIdx.x == 0) sm.count = 0;
    __syncthreads();

    const u64 n_tiles = (rv.n_reads + S1_WARPS - 1) / S1_WARPS;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u64 r = tile * S1_WARPS + warp;
        const bool valid = r < rv.n_reads;
        const u32 len = valid ? ga_read_len(rv, r) : 0u;
        const u32 nwt = len >= (u32)w ? len - (u32)w + 1u : 0u;     // windows of this read
        const u32 nch = (nwt + 127u) / 128u;
        const u64* rp = valid ? ga_read_ptr(rv, r) : rv.words;
        const u32 n_words = (len + 31u) / 32u;
        const u64 e_read = (rv.first_read + r) * (u64)rv.estride;
        for (u32 c = 0;; ++c) {
            const bool more = c < nch;
            if (!__syncthreads_or(more)) break;
            if (more) {
                const u32 base_word = c * 4u;
                if (lane < 8u) {
                    const u32 wi = base_word + lane;
                    sm.words[warp][lane] = wi < n_words ? __ldg(rp + wi) : 0ull;
                }
                __syncwarp();
                const u64* sw = sm.words[warp];
                auto get64 = [&](u32 pos) -> u64 {       // 32 symbols from symbol `pos` on, symbol 0 in the low bits
                    const u32 wi = (pos >> 5) - base_word, sh = (pos & 31u) * 2u;
                    const u64 lo = sw[wi], hi = sw[wi + 1];
                    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
                };
                u32 v[5];
#pragma unroll
                for (int t = 0; t < 5; ++t) {
                    const u32 q = c * 128u + 32u * t + lane;
                    v[t] = q + (u32)m <= len ? sk_hash32((u32)get64(q) & mmask) : 0xFFFFFFFFu;
                }
                auto shifted = [&](u32 a, u32 b, u32 s) -> u32 {
                    const u32 src = (lane + s) & 31u;
                    const u32 x = __shfl_sync(FULL, a, src), y = __shfl_sync(FULL, b, src);
                    return lane + s < 32u ? x : y;
                };
                for (u32 s = 1; s < P; s <<= 1) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) v[t] = min(v[t], shifted(v[t], v[t + 1], s));
                    v[4] = min(v[4], shifted(v[4], 0xFFFFFFFFu, s));
                }
                if (n > P) {
                    const u32 s = n - P;
#pragma unroll
                    for (int t = 0; t < 4; ++t) v[t] = min(v[t], shifted(v[t], v[t + 1], s));
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const u32 p = c * 128u + 32u * t + lane;
                    const bool vw = p < nwt;
                    // the minimum of n hashes crowds near zero: hash it again before taking bucket bits
                    const u32 b = bits ? sk_hash32(v[t] ^ 0x5bd1e995u) >> (32 - bits) : 0u;
                    const u32 prevb = __shfl_up_sync(FULL, b, 1);
                    const bool start = vw && (lane == 0u || b != prevb);
                    const u32 flags = __ballot_sync(FULL, start);
                    const u32 nvalid = __popc(__ballot_sync(FULL, vw));
                    if (flags == 0u) continue;
                    u32 base = 0;
                    if (lane == 0u) base = atomicAdd(&sm.count, (u32)__popc(flags));
                    base = __shfl_sync(FULL, base, 0);
                    if (start) {
                        const u32 higher = flags & ~((2u << lane) - 1u);
                        const u32 nxt = higher ? (u32)__ffs(higher) - 1u : 32u;
                        const u32 nwin = min(nxt, nvalid) - lane;
                        const u32 idx = base + __popc(flags & ((1u << lane) - 1u));
                        const u64 hi = sk_rev2(get64(p)), lo = sk_rev2(get64(p + 32u));
                        const u32 has_next = p + nwin - 1u + (u32)w < len ? 1u : 0u;
                        const u32 b1 = b >> l2_bits, b2 = b & ((1u << l2_bits) - 1u);
                        const u32 rank = atomicAdd(&sm.hist[b1], 1u);
                        atomicAdd(&ghist[b], 1u);
                        sm.bases[idx] = make_ulonglong2(hi, lo);
                        sm.meta[idx] = ((e_read + p) << 16) | ((u64)b2 << 6) | ((u64)has_next << 5) | (u64)(nwin - 1u);
                        sm.br[idx] = (b1 << 16) | rank;
                    }
                }
            }
            __syncthreads();
            if (sm.count > S1_STAGE - S1_MARGIN) s1_flush(sm, n_l1, out_bases, out_meta, cap1, cursors, overflow);
        }
    }
    s1_flush(sm, n_l1, out_bases, out_meta, cap1, cursors, overflow);
    if (overflow) atomicOr(status, GA_ST_TABLE_FULL);
}

// ------------------------------------------------------------------------------------------------
// 2. histogram -> offsets; level-1 buckets -> final buckets
__global__ void __launch_bounds__(1024) sk_offsets_kernel(const u32* __restrict__ hist, u64 n, u64* __restrict__ offsets,
                                                          u64* __restrict__ cursors) {
    __shared__ u64 part[1024];
    const u64 span = (n + 1023) / 1024;
    const u64 lo = min(n, threadIdx.x * span), hi = min(n, lo + span);
    u64 sum = 0;
    for (u64 i = lo; i < hi; ++i) sum += hist[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 run = 0;
        for (int i = 0; i < 1024; ++i) {
            const u64 t = part[i];
            part[i] = run;
            run += t;
        }
        offsets[n] = run;
    }
    __syncthreads();
    u64 run = part[threadIdx.x];
    for (u64 i = lo; i < hi; ++i) {
        offsets[i] = run;
        cursors[i] = run;
        run += hist[i];
    }
}

constexpr int S2_THREADS = 256;
constexpr int S2_PER = 8;
constexpr u32 S2_CHUNK = S2_THREADS * S2_PER;

__global__ void __launch_bounds__(S2_THREADS)
sk_scatter_buckets_kernel(const ulonglong2* __restrict__ in_bases, const u64* __restrict__ in_meta, u64 cap1,
                          const u64* __restrict__ cursors1, u32 n_l1, int l2_bits, u64* __restrict__ cursors2,
                          ulonglong2* __restrict__ out_bases, u64* __restrict__ out_meta) {
    __shared__ u32 hist[1024];
    __shared__ u64 gbase[1024];
    const u32 n_l2 = 1u << l2_bits;
    const u64 chunks_per = (cap1 + S2_CHUNK - 1) / S2_CHUNK;
    const u64 total = (u64)n_l1 * chunks_per;
    for (u64 chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        const u32 b1 = (u32)(chunk / chunks_per);
        const u64 lo = (chunk % chunks_per) * S2_CHUNK;
        const u64 cnt1 = min(cursors1[b1], cap1);
        if (lo >= cnt1) continue;                           // CTA-uniform
        for (u32 p = threadIdx.x; p < n_l2; p += S2_THREADS) hist[p] = 0;
        __syncthreads();
        ulonglong2 bs[S2_PER];
        u64 mt[S2_PER];
        u32 rank[S2_PER];
#pragma unroll
        for (int u = 0; u < S2_PER; ++u) {
            const u64 i = lo + (u64)u * S2_THREADS + threadIdx.x;
            if (i < cnt1) {
                const u64 src = (u64)b1 * cap1 + i;
                bs[u] = in_bases[src];
                mt[u] = in_meta[src];
                rank[u] = atomicAdd(&hist[meta_b2(mt[u])], 1u);
            }
        }
        __syncthreads();
        for (u32 p = threadIdx.x; p < n_l2; p += S2_THREADS) {
            const u32 c = hist[p];
            gbase[p] = c ? atomicAdd(&cursors2[((u64)b1 << l2_bits) + p], (u64)c) : 0ull;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < S2_PER; ++u) {
            const u64 i = lo + (u64)u * S2_THREADS + threadIdx.x;
            if (i < cnt1) {
                const u64 dst = gbase[meta_b2(mt[u])] + rank[u];
                out_bases[dst] = bs[u];
                out_meta[dst] = mt[u];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// 3. one CTA per bucket: exact counts, solid windows, candidate edge stamps
constexpr int SB_THREADS = 1024;
constexpr u32 SB_MAX_SLOTS = 16384;
constexpr u32 SB_MAX_SOLID = 1024;
constexpr u32 SB_PROBE_MAX = 192;

__device__ __forceinline__ u32 sk_slot_hash(u64 key) {
    u64 h = key * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    return (u32)(h >> 36);
}

// counts: SM = true packs two 16-bit counters per word (shared memory is the scarce resource),
// SM = false (spill path, global scratch) uses one 32-bit word per slot
template <bool SM> __device__ __forceinline__ u32 cnt_get(const u32* cnt, u32 s) {
    if (SM) return (((const volatile u32*)cnt)[s >> 1] >> ((s & 1u) * 16u)) & 0xFFFFu;
    return ((const volatile u32*)cnt)[s];
}
template <bool SM> __device__ __forceinline__ void cnt_add(u32* cnt, u32 s) {
    if (SM) atomicAdd(cnt + (s >> 1), 1u << ((s & 1u) * 16u));
    else atomicAdd(cnt + s, 1u);
}

struct BucketMem {
    u64* keys;        // [cap]
    u32* cnt;         // SM: [cap / 2], else [cap]
    u64* solid_keys;  // [max_solid]
    u64* stamps;      // [max_solid * 4]
};

struct BucketCtl {     // shared-memory control block of one CTA
    u32 n_solid;
    u32 overflow;
    u32 bucket;
    u32 pad;
    u64 n_windows;
    u64 out_base;
};

// returns false when the bucket does not fit (cap / max_solid): the caller lists it for the spill path
template <bool SM>
__device__ bool sk_bucket_body(const ulonglong2* __restrict__ bases, const u64* __restrict__ meta, u64 lo, u64 hi,
                               int w, u32 threshold, BucketMem mem, u32 cap_limit, u32 max_solid, BucketCtl& ctl,
                               u64* __restrict__ solid_keys_out, u64* __restrict__ edge_stamp_out, u64 out_capacity,
                               u64* n_solid_global) {
    const u32 tid = threadIdx.x, T = blockDim.x;
    const u64 mask = ga_key_mask<u64>(w, 2);
    // A. windows in the bucket -> table size
    u64 nw = 0;
    for (u64 i = lo + tid; i < hi; i += T) nw += meta_windows(meta[i]);
    for (int off = 16; off > 0; off >>= 1) nw += __shfl_down_sync(FULL, nw, off);
    if (tid == 0) {
        ctl.n_windows = 0;
        ctl.n_solid = 0;
        ctl.overflow = 0;
    }
    __syncthreads();
    if ((tid & 31u) == 0 && nw) atomicAdd((unsigned long long*)&ctl.n_windows, (unsigned long long)nw);
    __syncthreads();
    nw = ctl.n_windows;
    u32 cap = 256;
    while (cap < cap_limit && (u64)cap < 2 * nw) cap <<= 1;
    const u32 cmask = cap - 1u;
    // B. clear
    for (u32 s = tid; s < cap; s += T) mem.keys[s] = GA_NONE64;
    for (u32 s = tid; s < (SM ? cap / 2 : cap); s += T) mem.cnt[s] = 0;
    __syncthreads();
    volatile u64* vkeys = mem.keys;
    volatile u32* vovf = &ctl.overflow;
    // C. count (saturating just above the threshold: only "count > threshold" is asked)
    for (u64 i = lo + tid; i < hi && !*vovf; i += T) {
        const ulonglong2 b = bases[i];
        const u32 nwin = meta_windows(meta[i]);
        u64 key = b.x >> (64 - 2 * w);
        for (u32 j = 0; j < nwin; ++j) {
            u32 s = sk_slot_hash(key) & cmask;
            u32 probes = 0;
            for (;;) {
                const u64 cur = vkeys[s];
                if (cur == key) break;
                if (cur == GA_NONE64) {
                    const u64 old = atomicCAS((unsigned long long*)(mem.keys + s), GA_NONE64, key);
                    if (old == GA_NONE64 || old == key) break;
                }
                if (++probes > SB_PROBE_MAX) {
                    s = GA_NONE32;
                    break;
                }
                s = (s + 1u) & cmask;
            }
            if (s == GA_NONE32) {
                *vovf = 1u;
                break;
            }
            if (cnt_get<SM>(mem.cnt, s) <= threshold) cnt_add<SM>(mem.cnt, s);
            const u32 idx = (u32)w + j;
            const u64 word = idx < 32u ? b.x : b.y;
            key = ((key << 2) | ((word >> (62u - 2u * (idx & 31u))) & 3ull)) & mask;
        }
    }
    __syncthreads();
    if (ctl.overflow) return false;
    // D. solid windows: counter word -> solid index + 1 (0 = not solid)
    if (SM) {
        for (u32 wi = tid; wi < cap / 2; wi += T) {
            const u32 word = mem.cnt[wi];
            const u32 c0 = word & 0xFFFFu, c1 = word >> 16;
            const u32 s0 = c0 > threshold, s1 = c1 > threshold;
            u32 neu = 0;
            if (s0 + s1) {
                const u32 base = atomicAdd(&ctl.n_solid, s0 + s1);
                if (base + s0 + s1 <= max_solid) {
                    if (s0) {
                        mem.solid_keys[base] = mem.keys[2 * wi];
                        neu |= base + 1u;
                    }
                    if (s1) {
                        mem.solid_keys[base + s0] = mem.keys[2 * wi + 1];
                        neu |= (base + s0 + 1u) << 16;
                    }
                }
            }
            mem.cnt[wi] = neu;
        }
    } else {
        for (u32 s = tid; s < cap; s += T) {
            u32 neu = 0;
            if (mem.cnt[s] > threshold) {
                const u32 base = atomicAdd(&ctl.n_solid, 1u);
                if (base < max_solid) {
                    mem.solid_keys[base] = mem.keys[s];
                    neu = base + 1u;
                }
            }
            mem.cnt[s] = neu;
        }
    }
    __syncthreads();
    const u32 n_solid = ctl.n_solid;
    if (n_solid > max_solid) return false;
    if (n_solid == 0) return true;
    // E. candidate edge stamps: smallest ordinal of "solid window followed by symbol c"
    for (u32 s = tid; s < 4 * n_solid; s += T) mem.stamps[s] = GA_NONE64;
    if (tid == 0) ctl.out_base = atomicAdd((unsigned long long*)n_solid_global, (unsigned long long)n_solid);
    __syncthreads();
    for (u64 i = lo + tid; i < hi; i += T) {
        const ulonglong2 b = bases[i];
        const u64 mt = meta[i];
        const u32 nwin = meta_windows(mt);
        const u32 nfollow = meta_has_next(mt) ? nwin : nwin - 1u;     // windows that have a next symbol
        const u64 e0 = meta_ordinal(mt);
        u64 key = b.x >> (64 - 2 * w);
        for (u32 j = 0; j < nfollow; ++j) {
            const u32 idx = (u32)w + j;
            const u64 word = idx < 32u ? b.x : b.y;
            const u32 c = (u32)(word >> (62u - 2u * (idx & 31u))) & 3u;
            u32 s = sk_slot_hash(key) & cmask;
            u32 probes = 0;                                           // present: counted in C
            while (vkeys[s] != key && probes++ <= SB_PROBE_MAX) s = (s + 1u) & cmask;
            const u32 sol = probes <= SB_PROBE_MAX ? cnt_get<SM>(mem.cnt, s) : 0u;
            if (sol) {
                u64* p = mem.stamps + 4u * (sol - 1u) + c;
                const u64 e = e0 + j;
                if (e < *(volatile u64*)p) atomicMin((unsigned long long*)p, (unsigned long long)e);
            }
            key = ((key << 2) | (u64)c) & mask;
        }
    }
    __syncthreads();
    // F. output
    const u64 base = ctl.out_base;
    if (base + n_solid <= out_capacity) {
        for (u32 s = tid; s < n_solid; s += T) solid_keys_out[base + s] = mem.solid_keys[s];
        for (u32 s = tid; s < 4 * n_solid; s += T) edge_stamp_out[4 * base + s] = mem.stamps[s];
    }
    return true;
}

// counters: [0] next bucket, [1] solid windows so far, [2] buckets listed for the spill path
__global__ void __launch_bounds__(SB_THREADS, 1)
sk_bucket_kernel(const ulonglong2* __restrict__ bases, const u64* __restrict__ meta, const u64* __restrict__ offsets,
                 u64 n_buckets, int w, u32 threshold, u32 cap_limit, u32 max_solid, u64* __restrict__ solid_keys_out,
                 u64* __restrict__ edge_stamp_out, u64 out_capacity, u64* counters, u32* __restrict__ spill_list,
                 u64 spill_capacity, u32* status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BucketCtl ctl;
    BucketMem mem;
    mem.keys = reinterpret_cast<u64*>(smem_raw);
    mem.solid_keys = mem.keys + cap_limit;
    mem.stamps = mem.solid_keys + max_solid;
    mem.cnt = reinterpret_cast<u32*>(mem.stamps + 4 * (size_t)max_solid);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) ctl.bucket = (u32)min((u64)atomicAdd((unsigned long long*)&counters[0], 1ull), n_buckets);
        __syncthreads();
        const u64 b = ctl.bucket;
        if (b >= n_buckets) break;
        const u64 lo = offsets[b], hi = offsets[b + 1];
        if (lo == hi) continue;
        const bool ok = sk_bucket_body<true>(bases, meta, lo, hi, w, threshold, mem, cap_limit, max_solid, ctl,
                                             solid_keys_out, edge_stamp_out, out_capacity, counters + 1);
        if (!ok && threadIdx.x == 0) {
            const u64 at = atomicAdd((unsigned long long*)&counters[2], 1ull);
            if (at < spill_capacity) spill_list[at] = (u32)b;
            else atomicOr(status, GA_ST_TABLE_FULL);
        }
    }
}

// spill path: the same body over global scratch (one slice per CTA), for buckets whose distinct or
// solid windows exceed the shared-memory table
__global__ void __launch_bounds__(SB_THREADS, 1)
sk_bucket_spill_kernel(const ulonglong2* __restrict__ bases, const u64* __restrict__ meta,
                       const u64* __restrict__ offsets, const u32* __restrict__ spill_list, u64 n_spill, int w,
                       u32 threshold, u32 cap, unsigned char* __restrict__ scratch, u64 scratch_per_cta,
                       u64* __restrict__ solid_keys_out, u64* __restrict__ edge_stamp_out, u64 out_capacity,
                       u64* counters, u32* status) {
    __shared__ BucketCtl ctl;
    BucketMem mem;
    unsigned char* mine = scratch + (u64)blockIdx.x * scratch_per_cta;
    mem.keys = reinterpret_cast<u64*>(mine);
    mem.solid_keys = mem.keys + cap;
    mem.stamps = mem.solid_keys + cap;
    mem.cnt = reinterpret_cast<u32*>(mem.stamps + 4 * (size_t)cap);
    for (u64 oi = blockIdx.x; oi < n_spill; oi += gridDim.x) {
        __syncthreads();
        const u64 b = spill_list[oi];
        const u64 lo = offsets[b], hi = offsets[b + 1];
        const bool ok = sk_bucket_body<false>(bases, meta, lo, hi, w, threshold, mem, cap, cap, ctl, solid_keys_out,
                                              edge_stamp_out, out_capacity, counters + 1);
        if (!ok && threadIdx.x == 0) atomicOr(status, GA_ST_TABLE_FULL);
        __threadfence();
    }
}

// ------------------------------------------------------------------------------------------------
// 4. candidate stamps -> the reference's edges and node stamps
__global__ void __launch_bounds__(256)
sk_resolve_kernel(const u64* __restrict__ keys, u64 n, int w, const Slot<u64>* __restrict__ solid, u64 solid_cap,
                  u64* __restrict__ edge_stamp, u64* __restrict__ node_stamp) {
    const u64 mask = ga_key_mask<u64>(w, 2);
    for (u64 id = blockIdx.x * (u64)blockDim.x + threadIdx.x; id < n; id += (u64)gridDim.x * blockDim.x) {
        const u64 key = keys[id];
        u64 mine = GA_NONE64;
        for (u32 c = 0; c < 4; ++c) {
            const u64 e = edge_stamp[4 * id + c];
            if (e == GA_NONE64) continue;
            const u32 sid = ga_table_find(solid, solid_cap, ((key << 2) | (u64)c) & mask);
            if (sid == GA_NONE32) {
                edge_stamp[4 * id + c] = GA_NONE64;          // successor filtered out: no edge
            } else {
                mine = min(mine, 2 * e);
                atomicMin((unsigned long long*)(node_stamp + sid), (unsigned long long)(2 * e + 1));
            }
        }
        if (mine != GA_NONE64) atomicMin((unsigned long long*)(node_stamp + id), (unsigned long long)mine);
    }
}

size_t sk_bucket_smem(u32 cap_limit, u32 max_solid) {
    return (size_t)cap_limit * 8 + (size_t)max_solid * 8 + (size_t)max_solid * 32 + (size_t)cap_limit * 2;
}

bool is_pow2(u64 v) { return v && !(v & (v - 1)); }

}  // namespace

extern "C" int ga_sk_minimizer_len(int k) {
    const int w = k - 1;
    if (w < 1) return 0;
    const int lo = w < 11 ? w : 11;
    int m = w - 15;
    if (m > 16) m = 16;
    if (m < lo) m = lo;
    return m;
}

extern "C" int ga_sk_scatter_reads(const ga_reads* reads, int k, int l1_bits, int l2_bits, void* rec_bases_dev,
                                   uint64_t* rec_meta_dev, uint64_t l1_capacity, uint64_t* l1_cursors_dev,
                                   uint32_t* hist_dev, uint32_t* status_dev, ga_stream stream) {
    if (!reads || !rec_bases_dev || !rec_meta_dev || !l1_cursors_dev || !hist_dev || !status_dev || l1_capacity == 0 ||
        l1_bits < 0 || l1_bits > 10 || l2_bits < 0 || l2_bits > 10) {
        ga_set_error("ga_sk_scatter_reads: bad arguments (bucket bits must be 0..10 each)");
        return GA_ERR_BAD_ARG;
    }
    const int w = k - 1;
    if (reads->paired || reads->storage_bits != 2 || reads->sym_bits != 2 || w < 1 || w > 31) {
        ga_set_error("ga_sk_scatter_reads: needs unpaired 2-bit reads and 2 <= k <= 32");
        return GA_ERR_BAD_ARG;
    }
    if (reads->estride == 0 || (reads->first_read + reads->n_reads) > ((1ull << 47) / reads->estride)) {
        ga_set_error("ga_sk_scatter_reads: occurrence ordinals exceed 47 bits");
        return GA_ERR_BAD_ARG;
    }
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    static bool attr_set = false;
    if (!attr_set) {
        GA_CUDA(cudaFuncSetAttribute(sk_scatter_reads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(S1Shared)));
        attr_set = true;
    }
    const u64 n_tiles = (rv.n_reads + S1_WARPS - 1) / S1_WARPS;
    const unsigned grid = (unsigned)(n_tiles < 148ull * 2 ? n_tiles : 148ull * 2);
    sk_scatter_reads_kernel<<<grid, S1_THREADS, sizeof(S1Shared), (cudaStream_t)stream>>>(
        rv, w, ga_sk_minimizer_len(k), l1_bits, l2_bits, (ulonglong2*)rec_bases_dev, (u64*)rec_meta_dev, l1_capacity,
        (u64*)l1_cursors_dev, hist_dev, status_dev);
    GA_LAUNCH_CHECK("sk_scatter_reads");
    return GA_OK;
}

extern "C" int ga_sk_offsets(const uint32_t* hist_dev, uint64_t n_buckets, uint64_t* offsets_dev,
                             uint64_t* cursors_dev, ga_stream stream) {
    if (!hist_dev || !offsets_dev || !cursors_dev || n_buckets == 0) {
        ga_set_error("ga_sk_offsets: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    sk_offsets_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(hist_dev, n_buckets, (u64*)offsets_dev, (u64*)cursors_dev);
    GA_LAUNCH_CHECK("sk_offsets");
    return GA_OK;
}

extern "C" int ga_sk_scatter_buckets(const void* rec_bases_dev, const uint64_t* rec_meta_dev, uint64_t l1_capacity,
                                     const uint64_t* l1_cursors_dev, int l1_bits, int l2_bits, uint64_t* cursors_dev,
                                     void* out_bases_dev, uint64_t* out_meta_dev, ga_stream stream) {
    if (!rec_bases_dev || !rec_meta_dev || !l1_cursors_dev || !cursors_dev || !out_bases_dev || !out_meta_dev ||
        l1_capacity == 0 || l1_bits < 0 || l1_bits > 10 || l2_bits < 0 || l2_bits > 10) {
        ga_set_error("ga_sk_scatter_buckets: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    const u32 n_l1 = 1u << l1_bits;
    const u64 total = (u64)n_l1 * ((l1_capacity + S2_CHUNK - 1) / S2_CHUNK);
    const unsigned grid = (unsigned)(total < 148ull * 8 ? total : 148ull * 8);
    sk_scatter_buckets_kernel<<<grid, S2_THREADS, 0, (cudaStream_t)stream>>>(
        (const ulonglong2*)rec_bases_dev, (const u64*)rec_meta_dev, l1_capacity, (const u64*)l1_cursors_dev, n_l1,
        l2_bits, (u64*)cursors_dev, (ulonglong2*)out_bases_dev, (u64*)out_meta_dev);
    GA_LAUNCH_CHECK("sk_scatter_buckets");
    return GA_OK;
}

extern "C" int ga_sk_count_build(const void* bases_dev, const uint64_t* meta_dev, const uint64_t* offsets_dev,
                                 uint64_t n_buckets, int k, int64_t threshold, uint32_t table_slots,
                                 uint32_t max_solid, uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev,
                                 uint64_t out_capacity, uint64_t* counters_dev, uint32_t* spill_list_dev,
                                 uint64_t spill_capacity, uint32_t* status_dev, ga_stream stream) {
    const int w = k - 1;
    if (!bases_dev || !meta_dev || !offsets_dev || !solid_keys_out_dev || !edge_stamp_out_dev || !counters_dev ||
        !spill_list_dev || !status_dev || n_buckets == 0 || w < 1 || w > 31 || threshold < 0 || threshold > 60000 ||
        !is_pow2(table_slots) || table_slots < 256 || table_slots > SB_MAX_SLOTS || max_solid == 0 ||
        max_solid > SB_MAX_SOLID) {
        ga_set_error("ga_sk_count_build: bad arguments (0 <= threshold <= 60000, table_slots a power of two in "
                     "256..%u, max_solid 1..%u)", SB_MAX_SLOTS, SB_MAX_SOLID);
        return GA_ERR_BAD_ARG;
    }
    const size_t smem = sk_bucket_smem(table_slots, max_solid);
    GA_CUDA(cudaFuncSetAttribute(sk_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)(n_buckets < 148ull ? n_buckets : 148ull);
    sk_bucket_kernel<<<grid, SB_THREADS, smem, (cudaStream_t)stream>>>(
        (const ulonglong2*)bases_dev, (const u64*)meta_dev, (const u64*)offsets_dev, n_buckets, w, (u32)threshold,
        table_slots, max_solid, (u64*)solid_keys_out_dev, (u64*)edge_stamp_out_dev, out_capacity, (u64*)counters_dev,
        spill_list_dev, spill_capacity, status_dev);
    GA_LAUNCH_CHECK("sk_bucket");
    return GA_OK;
}

extern "C" uint64_t ga_sk_spill_scratch_bytes(uint32_t table_slots) {
    // keys + solid keys (8 B each) + 4 stamps (32 B) + one 32-bit counter per slot
    return (uint64_t)table_slots * (8 + 8 + 32 + 4);
}

extern "C" int ga_sk_count_build_spill(const void* bases_dev, const uint64_t* meta_dev, const uint64_t* offsets_dev,
                                       const uint32_t* spill_list_dev, uint64_t n_spill, int k, int64_t threshold,
                                       uint32_t table_slots, void* scratch_dev, uint32_t n_ctas,
                                       uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev,
                                       uint64_t out_capacity, uint64_t* counters_dev, uint32_t* status_dev,
                                       ga_stream stream) {
    const int w = k - 1;
    if (!bases_dev || !meta_dev || !offsets_dev || !spill_list_dev || !scratch_dev || !solid_keys_out_dev ||
        !edge_stamp_out_dev || !counters_dev || !status_dev || w < 1 || w > 31 || threshold < 0 ||
        !is_pow2(table_slots) || table_slots < 256 || n_ctas == 0) {
        ga_set_error("ga_sk_count_build_spill: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    if (n_spill == 0) return GA_OK;
    sk_bucket_spill_kernel<<<n_ctas, SB_THREADS, 0, (cudaStream_t)stream>>>(
        (const ulonglong2*)bases_dev, (const u64*)meta_dev, (const u64*)offsets_dev, spill_list_dev, n_spill, w,
        (u32)(threshold > 0xFFFFFFF0ll ? 0xFFFFFFF0ll : threshold), table_slots, (unsigned char*)scratch_dev,
        ga_sk_spill_scratch_bytes(table_slots), (u64*)solid_keys_out_dev, (u64*)edge_stamp_out_dev, out_capacity,
        (u64*)counters_dev, status_dev);
    GA_LAUNCH_CHECK("sk_bucket_spill");
    return GA_OK;
}

extern "C" int ga_sk_resolve(const uint64_t* solid_keys_dev, uint64_t n_solid, int k, const void* solid_dev,
                             uint64_t solid_capacity, uint64_t* edge_stamp_dev, uint64_t* node_stamp_dev,
                             ga_stream stream) {
    const int w = k - 1;
    if (!solid_keys_dev || !solid_dev || !edge_stamp_dev || !node_stamp_dev || solid_capacity == 0 || w < 1 || w > 31) {
        ga_set_error("ga_sk_resolve: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    if (n_solid == 0) return GA_OK;
    unsigned grid = ga_grid(n_solid, 256);
    sk_resolve_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const u64*)solid_keys_dev, n_solid, w,
                                                             (const Slot<u64>*)solid_dev, solid_capacity,
                                                             (u64*)edge_stamp_dev, (u64*)node_stamp_dev);
    GA_LAUNCH_CHECK("sk_resolve");
    return GA_OK;
}
