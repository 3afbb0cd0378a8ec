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

// hash of one m-mer (only its rank among the <= 16 m-mers of a window matters: one multiplication with an
// xor-fold is enough to break the bias towards poly-A; the bucket bits come from sk_bucket_of the minimum)
__device__ __forceinline__ u32 sk_mmer_hash(u32 x) {
    x *= 0x9E3779B1u;
    return x ^ (x >> 15);
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

// A level-1 bucket holds 32-byte slots {bases hi, bases lo, meta, 0}: ONE full-sector store per record.
// Two arrays (16 bytes of bases here, the meta word there) cost two partial-sector requests per record, and
// those requests -- not instructions, not DRAM bytes -- were what bounded every variant of the scatter
// kernel (scripts/scatter_probe.py: 26.8 ms -> 12.9 ms on a quarter of C4 with nothing else changed).
__device__ __forceinline__ void sk_store_slot(u64* __restrict__ rec, u64 slot, u64 hi, u64 lo, u64 meta) {
    asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(rec + 4u * slot), "l"(hi), "l"(lo), "l"(meta), "l"(0ull)
                 : "memory");
}
__device__ __forceinline__ void sk_load_slot(const u64* __restrict__ rec, u64 slot, ulonglong2& bases, u64& meta) {
    [[maybe_unused]] u64 pad;
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(bases.x), "=l"(bases.y), "=l"(meta), "=l"(pad)
                 : "l"(rec + 4u * slot));
}
// The level-1 cursors sit 128 bytes apart: packed into 8 KB they all live in a handful of L2 slices, whose
// atomic units then bound the kernel (one returned atomic per record: 18.3 ms packed, 5.5 ms spread, same probe).
constexpr u32 SK_CURSOR_STRIDE = 16;

// ------------------------------------------------------------------------------------------------
// 1. reads -> records in level-1 buckets
constexpr int S1_THREADS = 512;
constexpr int S1_WARPS = S1_THREADS / 32;
constexpr u32 S1_STAGE = 3072;                 // records staged per CTA between flushes (>= S1_WARPS * 128)
constexpr u32 S1_MAXB = 1024;                  // level-1 buckets

struct S1Shared {
    ulonglong2 bases[S1_STAGE];
    u64 meta[S1_STAGE];
    u32 br[S1_STAGE];              // level-1 bucket << 16 | rank inside this flush
    u64 gbase[S1_MAXB];
    u32 hist[S1_MAXB];
    u64 words[S1_WARPS][8];
    u32 desc[S1_WARPS][128];       // records of the current chunk: position | (windows-1) << 7 | bucket << 12
    u32 wcount[S1_WARPS];
};

// the minimum of n hashes crowds near zero: mix it once more before taking the bucket bits from the top
// (bits == 0: one bucket)
__device__ __forceinline__ u32 sk_bucket_of(u32 min_hash, int bits) {
    const u32 x = (min_hash ^ (min_hash >> 13)) * 0x85EBCA77u;
    return __funnelshift_rc(x, 0u, 32u - (u32)bits);
}

// stage -> level-1 buckets: one global cursor bump per non-empty bucket, then the records of a bucket land
// next to each other.
template <class SM, int THREADS>
__device__ __forceinline__ void s1_flush(SM& sm, u32 n, u32 n_l1, u64* __restrict__ out_rec, u64 cap1,
                                         u64* __restrict__ cursors, bool& overflow) {
    __syncthreads();
    for (u32 p = threadIdx.x; p < n_l1; p += THREADS) {
        const u32 c = sm.hist[p];
        u64 base = 0;
        if (c) {
            base = atomicAdd((unsigned long long*)&cursors[p * SK_CURSOR_STRIDE], (unsigned long long)c);
            if (base + c > cap1) {
                overflow = true;
                base = GA_NONE64;      // dropped: the host retries with larger buckets
            }
        }
        sm.gbase[p] = base;
        sm.hist[p] = 0;
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < n; i += THREADS) {
        const u32 br = sm.br[i];
        const u32 b1 = br >> 16;
        const u64 base = sm.gbase[b1];
        if (base == GA_NONE64) continue;
        const ulonglong2 bs = sm.bases[i];
        sk_store_slot(out_rec, (u64)b1 * cap1 + base + (br & 0xFFFFu), bs.x, bs.y, sm.meta[i]);
    }
    __syncthreads();
}

// One warp per read and round.  Lane l owns the four consecutive window positions 4l..4l+3 of the
// current 128-window chunk, so the sliding minimum over a window's m-mers needs the lane's own
// values plus whole-lane minima of up to three following lanes and a prefix of one more.
// NMM: m-mers per window when known at compile time (16 for k >= 27, the common case), 0 = runtime.
template <int NMM>
__global__ void __launch_bounds__(S1_THREADS)
sk_scatter_reads_kernel(ReadsView rv, int w, int m, int l1_bits, int l2_bits, u64* __restrict__ out_rec, u64 cap1,
                        u64* __restrict__ cursors, u64* __restrict__ ghist, u32* status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S1Shared& sm = *reinterpret_cast<S1Shared*>(smem_raw);
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 n_l1 = 1u << l1_bits;
    const int bits = l1_bits + l2_bits;
    const u32 n = NMM ? (u32)NMM : (u32)(w - m + 1);      // m-mers per window, 1..16
    const u32 mmask = m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
    const u32 lt_mask = (1u << lane) - 1u;
    bool overflow = false;
    u32 staged = 0;                                       // records in the stage; identical in every thread
    for (u32 p = threadIdx.x; p < n_l1; p += S1_THREADS) sm.hist[p] = 0;
    __syncthreads();

    const u64 n_tiles = (rv.n_reads + S1_WARPS - 1) / S1_WARPS;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const u64 r = tile * S1_WARPS + warp;
        const bool valid = r < rv.n_reads;
        // paired input: both mates are plain reads here (counting only); windows run over mate 1's length
        const u32 len = valid ? ga_read_len(rv, rv.paired ? (r & ~1ull) : r) : 0u;
        const u32 nwt = len >= (u32)w ? len - (u32)w + 1u : 0u;     // windows of this read
        const u32 nch = (nwt + 127u) / 128u;
        const u64* rp = valid ? ga_read_ptr(rv, r) : rv.words;
        const u32 n_words = (len + 31u) / 32u;
        const u64 e_read = (rv.first_read + r) * (u64)rv.estride;
        for (u32 c = 0;; ++c) {
            const bool more = c < nch;
            if (!__syncthreads_or(more)) break;
            u32 mycount = 0;
            const u32 base_word = c * 4u;
            const u64* sw = sm.words[warp];
            auto get64 = [&](u32 pos) -> u64 {           // 32 symbols from symbol `pos` on, symbol 0 in the low bits
                const u32 wi = (pos >> 5) - base_word, sh = (pos & 31u) * 2u;
                const u64 lo = sw[wi], hi = sw[wi + 1];
                return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
            };
            if (more) {
                if (lane < 8u) {
                    const u32 wi = base_word + lane;
                    sm.words[warp][lane] = wi < n_words ? __ldg(rp + wi) : 0ull;
                }
                __syncwarp();
                // m-mer hashes: v[t] at position 4*lane + t, ov[t] at 128 + 4*lane + t (lanes 0..3 matter)
                const u32 q0 = c * 128u + 4u * lane;
                u32 v[4], ov[4];
                {
                    const u64 x = get64(q0);
                    const u64 y = lane < 4u ? get64(q0 + 128u) : 0ull;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        v[t] = q0 + t + (u32)m <= len ? sk_mmer_hash((u32)(x >> (2 * t)) & mmask) : 0xFFFFFFFFu;
                        ov[t] = (lane < 4u && q0 + 128u + t + (u32)m <= len) ? sk_mmer_hash((u32)(y >> (2 * t)) & mmask)
                                                                             : 0xFFFFFFFFu;
                    }
                }
                auto fetch = [&](u32 a, u32 b, u32 d) -> u32 {      // value of lane + d, continuing into the overflow lanes
                    const u32 src = (lane + d) & 31u;
                    const u32 x = __shfl_sync(FULL, a, src), y = __shfl_sync(FULL, b, src);
                    return lane + d < 32u ? x : y;
                };
                const u32 m4 = min(min(v[0], v[1]), min(v[2], v[3]));
                const u32 m4o = min(min(ov[0], ov[1]), min(ov[2], ov[3]));
                const u32 c1 = fetch(m4, m4o, 1), c2 = min(c1, fetch(m4, m4o, 2)), c3 = min(c2, fetch(m4, m4o, 3));
                const u32 p1 = v[0], p2 = min(v[0], v[1]), p3 = min(p2, v[2]);          // prefix minima
                const u32 p1o = ov[0], p2o = min(ov[0], ov[1]), p3o = min(p2o, ov[2]);
                u32 bkt[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const u32 end = (u32)t + n;              // m-mer positions t .. end-1 relative to 4*lane
                    u32 res = v[t];
#pragma unroll
                    for (int u = t + 1; u < 4; ++u)
                        if ((u32)u < end) res = min(res, v[u]);
                    if (end > 4u) {
                        const u32 f = (end - 4u) >> 2, rem = (end - 4u) & 3u;
                        if (f >= 1u) res = min(res, f == 1u ? c1 : (f == 2u ? c2 : c3));
                        if (rem) {
                            const u32 a = rem == 1u ? p1 : (rem == 2u ? p2 : p3);
                            const u32 b = rem == 1u ? p1o : (rem == 2u ? p2o : p3o);
                            res = min(res, fetch(a, b, f + 1u));
                        }
                    }
                    bkt[t] = sk_bucket_of(res, bits);
                }
                // record starts: bucket change, or a multiple of 32 windows (records hold at most 32)
                const u32 nvalid = min(nwt - c * 128u, 128u);
                const u32 prev3 = __shfl_up_sync(FULL, bkt[3], 1);
                bool st[4];
                u32 F[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const bool vw = 4u * lane + t < nvalid;
                    const bool change = t == 0 ? ((lane & 7u) == 0u || bkt[0] != prev3) : bkt[t] != bkt[t - 1];
                    st[t] = vw && change;
                    F[t] = __ballot_sync(FULL, st[t]);
                }
                const u32 any = F[0] | F[1] | F[2] | F[3];
                const u32 higher = any & ~((2u << lane) - 1u);
                u32 next_lane = 128u;                        // first start in a later lane
                if (higher) {
                    const u32 l2 = (u32)__ffs(higher) - 1u;
                    const u32 t2 = (F[0] >> l2) & 1u ? 0u : ((F[1] >> l2) & 1u ? 1u : ((F[2] >> l2) & 1u ? 2u : 3u));
                    next_lane = 4u * l2 + t2;
                }
                u32 off = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (st[t]) {
                        u32 next = next_lane;
#pragma unroll
                        for (int u = 3; u > t; --u)
                            if (st[u]) next = 4u * lane + u;
                        const u32 prel = 4u * lane + t;
                        const u32 nwin = min(next, nvalid) - prel;
                        sm.desc[warp][off + __popc(F[t] & lt_mask)] = prel | ((nwin - 1u) << 7) | (bkt[t] << 12);
                    }
                    off += __popc(F[t]);
                }
                mycount = off;
            }
            if (lane == 0u) sm.wcount[warp] = mycount;
            __syncthreads();
            // deterministic reservation: every thread computes the same prefix over the warps
            u32 mine = lane < (u32)S1_WARPS ? sm.wcount[lane] : 0u;
            u32 incl = mine;
#pragma unroll
            for (int o = 1; o < S1_WARPS; o <<= 1) {
                const u32 t = __shfl_up_sync(FULL, incl, o);
                if (lane >= (u32)o) incl += t;
            }
            const u32 total = __shfl_sync(FULL, incl, S1_WARPS - 1);
            const u32 my_base = __shfl_sync(FULL, incl - mine, warp);
            if (staged + total > S1_STAGE) {
                s1_flush<S1Shared, S1_THREADS>(sm, staged, n_l1, out_rec, cap1, cursors, overflow);
                staged = 0;
            }
            for (u32 l = lane; l < mycount; l += 32u) {
                const u32 d = sm.desc[warp][l];
                const u32 prel = d & 127u, nwin = ((d >> 7) & 31u) + 1u, b = d >> 12;
                const u32 p = c * 128u + prel;
                const u64 hi = sk_rev2(get64(p)), lo = sk_rev2(get64(p + 32u));
                const u32 has_next = p + nwin - 1u + (u32)w < len ? 1u : 0u;
                const u32 b1 = b >> l2_bits, b2 = b & ((1u << l2_bits) - 1u);
                const u32 rank = atomicAdd(&sm.hist[b1], 1u);
                atomicAdd((unsigned long long*)&ghist[b], (1ull << 32) | (unsigned long long)nwin);
                const u32 idx = staged + my_base + l;
                sm.bases[idx] = make_ulonglong2(hi, lo);
                sm.meta[idx] = ((e_read + p) << 16) | ((u64)b2 << 6) | ((u64)has_next << 5) | (u64)(nwin - 1u);
                sm.br[idx] = (b1 << 16) | rank;
            }
            staged += total;
        }
    }
    s1_flush<S1Shared, S1_THREADS>(sm, staged, n_l1, out_rec, cap1, cursors, overflow);
    if (overflow) atomicOr(status, GA_ST_TABLE_FULL);
}

// ------------------------------------------------------------------------------------------------
// 1b. the same cut, one LANE per read (windows of 16 m-mers: 26 <= k-1 <= 31, every real configuration).
// The warp-per-read kernel above pays ~610 warp instructions per 150-base read, most of them shuffles and
// ballots of the cooperative sliding minimum and of the record search, at 24 of 32 lanes.  Here a thread walks
// its own read 32 windows at a time with everything in registers (~200 warp instructions per read):
//   * the chunk's 47 m-mer hashes are taken from three 64-bit words with funnel shifts; the sliding minimum
//     over 16 m-mers is a suffix minimum of one aligned block of 16 combined with a running prefix minimum of
//     the next block (3 min per window, no shuffles, no branches);
//   * a window starts a record when its bucket differs from its predecessor's (or at a multiple of 32 windows):
//     the start mask stays in a register, the bucket of the i-th record goes to a per-thread column in shared
//     memory (predicated store);
//   * each record then takes its slot in its level-1 bucket with ONE returned global atomic and is written
//     straight from registers with one 256-bit store; a thread keeps BATCH atomics in flight before it builds
//     the records.  No shared-memory stage, no barrier: a stage of a few thousand records over 1024 level-1
//     buckets holds 1-2 records per bucket, so its flush coalesces nothing and costs three barriers plus one
//     returned atomic per bucket (a staged variant of this kernel measured the same time; profiles/r01).
// Buckets, record cuts and record contents are bit-identical to the kernel above (a GPU test compares the two).
template <int THREADS, int BATCH>
__global__ void __launch_bounds__(THREADS)
sk_scatter_reads_lane_kernel(ReadsView rv, int w, int m, int l1_bits, int l2_bits, u64* __restrict__ out_rec, u64 cap1,
                               u64* __restrict__ cursors, u64* __restrict__ ghist, u32* status) {
    __shared__ u32 bk[32][THREADS];                      // bucket of the record that starts at window j of this thread's chunk
    const u32 tid = threadIdx.x;
    const int bits = l1_bits + l2_bits;
    const u32 l2_mask = (1u << l2_bits) - 1u;
    const u32 mmask = m >= 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
    bool overflow = false;
    for (u64 r = (u64)blockIdx.x * THREADS + tid; r < rv.n_reads; r += (u64)gridDim.x * THREADS) {
        // paired input: both mates are plain reads here (counting only); windows run over mate 1's length
        const u32 len = ga_read_len(rv, rv.paired ? (r & ~1ull) : r);
        const u32 nwt = len >= (u32)w ? len - (u32)w + 1u : 0u;     // windows of this read
        const u32 nch = (nwt + 31u) / 32u;
        const u64* rp = ga_read_ptr(rv, r);
        const u32 n_words = nwt ? (len + 31u) / 32u : 0u;
        const u64 e_read = (rv.first_read + r) * (u64)rv.estride;
        // words c, c+1, c+2 of the read while chunk c (windows 32c .. 32c+31) is walked
        u64 w0 = 0, w1 = n_words > 0u ? __ldg(rp) : 0ull, w2 = n_words > 1u ? __ldg(rp + 1) : 0ull;
        for (u32 c = 0; c < nch; ++c) {
            w0 = w1;
            w1 = w2;
            w2 = c + 2u < n_words ? __ldg(rp + c + 2u) : 0ull;
            const u32 nv = min(nwt - c * 32u, 32u);
            u32 mask = 0;
            {
                const u32 R[4] = {(u32)w0, (u32)(w0 >> 32), (u32)w1, (u32)(w1 >> 32)};
                auto mm = [&](int j) -> u32 {
                    return sk_mmer_hash(__funnelshift_r(R[j >> 4], R[(j >> 4) + 1], 2u * ((u32)j & 15u)) & mmask);
                };
                u32 prev = 0;
                // the bucket of a record start goes to row j of this thread's column: a fixed address per j (a
                // running pointer bumped under the same predicate made every bump wait for the store before it
                // to read its address register -- 22 % of this kernel's stall samples, profiles/r02)
                auto window = [&](int j, u32 min_hash) {
                    const u32 bkt = sk_bucket_of(min_hash, bits);
                    if (j == 0 || bkt != prev) {
                        bk[j][tid] = bkt;
                        mask |= 1u << j;
                    }
                    prev = bkt;
                };
                u32 S[16], T[16];
#pragma unroll
                for (int o = 0; o < 16; ++o) S[o] = mm(o);
#pragma unroll
                for (int o = 14; o >= 0; --o) S[o] = min(S[o], S[o + 1]);
                u32 P = 0xFFFFFFFFu;
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    window(o, o ? min(S[o], P) : S[0]);
                    T[o] = mm(16 + o);
                    P = min(P, T[o]);
                }
#pragma unroll
                for (int o = 14; o >= 0; --o) T[o] = min(T[o], T[o + 1]);
                P = 0xFFFFFFFFu;
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    window(16 + o, o ? min(T[o], P) : T[0]);
                    if (o < 15) P = min(P, mm(32 + o));
                }
                mask &= nv >= 32u ? FULL : (1u << nv) - 1u;
            }
            const u32 cur = (u32)__popc(mask);
            u32 rest = mask;
            for (u32 done = 0; done < cur; done += BATCH) {
                u64 pos[BATCH];
                u32 what[BATCH];                          // start window | windows << 8 | bucket << 12 (bucket: 20 bits)
#pragma unroll
                for (int q = 0; q < BATCH; ++q) {
                    pos[q] = 0;
                    what[q] = 0;
                    if (done + q < cur) {
                        const u32 j = (u32)__ffs(rest) - 1u;
                        rest &= rest - 1u;
                        const u32 nwin = (rest ? (u32)__ffs(rest) - 1u : nv) - j;
                        const u32 b = bk[j][tid];
                        what[q] = j | ((nwin - 1u) << 5) | (b << 12);
                        pos[q] = atomicAdd((unsigned long long*)&cursors[(b >> l2_bits) * SK_CURSOR_STRIDE], 1ull);
                        atomicAdd((unsigned long long*)&ghist[b], (1ull << 32) | (unsigned long long)nwin);
                    }
                }
#pragma unroll
                for (int q = 0; q < BATCH; ++q) {
                    if (done + q < cur) {
                        const u32 j = what[q] & 31u, nwin = ((what[q] >> 5) & 31u) + 1u, b = what[q] >> 12;
                        const u32 p = c * 32u + j, sh = 2u * j;
                        const u64 hi = sk_rev2(sh ? (w0 >> sh) | (w1 << (64u - sh)) : w0);
                        const u64 lo = sk_rev2(sh ? (w1 >> sh) | (w2 << (64u - sh)) : w1);
                        const u32 has_next = p + nwin - 1u + (u32)w < len ? 1u : 0u;
                        if (pos[q] < cap1) {
                            sk_store_slot(out_rec, (u64)(b >> l2_bits) * cap1 + pos[q], hi, lo,
                                          ((e_read + p) << 16) | ((u64)(b & l2_mask) << 6) | ((u64)has_next << 5) |
                                              (u64)(nwin - 1u));
                        } else {
                            overflow = true;              // dropped: the host retries with larger buckets
                        }
                    }
                }
            }
        }
    }
    if (overflow) atomicOr(status, GA_ST_TABLE_FULL);
}

// ------------------------------------------------------------------------------------------------
// 2. histogram -> offsets; level-1 buckets -> final buckets
// exclusive prefix sum of the per-bucket record counts, three small kernels: per-tile sums, a scan of
// the tile sums (one CTA), per-tile scan + write
constexpr u32 OFF_TILE = 2048;      // buckets per CTA (256 threads x 8)

__global__ void __launch_bounds__(256) sk_offsets_tile_sums_kernel(const u64* __restrict__ hist, u64 n, u64* __restrict__ tile_sum) {
    __shared__ u64 part[8];
    const u64 base = (u64)blockIdx.x * OFF_TILE;
    u64 sum = 0;
    for (u32 j = 0; j < 8; ++j) {
        const u64 i = base + j * 256u + threadIdx.x;
        if (i < n) sum += hist[i] >> 32;
    }
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(FULL, sum, off);
    if ((threadIdx.x & 31u) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int i = 0; i < 8; ++i) t += part[i];
        tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) sk_offsets_scan_tiles_kernel(u64* __restrict__ tile_sum, u64 n_tiles, u64* __restrict__ total_out) {
    __shared__ u64 part[1024];
    const u64 span = (n_tiles + 1023) / 1024;
    const u64 lo = min(n_tiles, threadIdx.x * span), hi = min(n_tiles, lo + span);
    u64 sum = 0;
    for (u64 i = lo; i < hi; ++i) sum += tile_sum[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 run = 0;
        for (int i = 0; i < 1024; ++i) {
            const u64 t = part[i];
            part[i] = run;
            run += t;
        }
        *total_out = run;
    }
    __syncthreads();
    u64 run = part[threadIdx.x];
    for (u64 i = lo; i < hi; ++i) {
        const u64 t = tile_sum[i];
        tile_sum[i] = run;
        run += t;
    }
}

__global__ void __launch_bounds__(256) sk_offsets_write_kernel(const u64* __restrict__ hist, u64 n, const u64* __restrict__ tile_base,
                                                               u64* __restrict__ offsets, u64* __restrict__ cursors) {
    __shared__ u64 warp_base[8];
    const u64 first = (u64)blockIdx.x * OFF_TILE + (u64)threadIdx.x * 8u;     // 8 consecutive buckets per thread
    u64 v[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = first + j < n ? hist[first + j] >> 32 : 0ull;
        sum += v[j];
    }
    const u32 lane = threadIdx.x & 31u;
    u64 incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u64 t = __shfl_up_sync(FULL, incl, off);
        if (lane >= (u32)off) incl += t;
    }
    if (lane == 31u) warp_base[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 run = tile_base[blockIdx.x];
        for (int i = 0; i < 8; ++i) {
            const u64 t = warp_base[i];
            warp_base[i] = run;
            run += t;
        }
    }
    __syncthreads();
    u64 run = warp_base[threadIdx.x >> 5] + incl - sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (first + j < n) {
            offsets[first + j] = run;
            cursors[first + j] = run;
        }
        run += v[j];
    }
}

constexpr int S2_THREADS = 256;
constexpr int S2_PER = 8;
constexpr u32 S2_CHUNK = S2_THREADS * S2_PER;

// INDEX = false: records move to their final bucket.  INDEX = true: only a 32-bit index (position inside the
// level-1 bucket) is written per record; the bucket kernel then gathers the records from the level-1 bucket,
// which is L2-sized (1024 final buckets share it and are processed back to back).  Same result, a quarter of
// the traffic; the dense form is what the multi-GPU exchange needs.
template <bool INDEX>
__global__ void __launch_bounds__(S2_THREADS)
sk_scatter_buckets_kernel(const u64* __restrict__ in_rec, u64 cap1,
                          const u64* __restrict__ cursors1, u32 n_l1, int l2_bits, u64* __restrict__ cursors2,
                          ulonglong2* __restrict__ out_bases, u64* __restrict__ out_meta, u32* __restrict__ out_index) {
    __shared__ u32 hist[1024];
    __shared__ u64 gbase[1024];
    const u32 n_l2 = 1u << l2_bits;
    const u64 chunks_per = (cap1 + S2_CHUNK - 1) / S2_CHUNK;
    const u64 total = (u64)n_l1 * chunks_per;
    for (u64 chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        const u32 b1 = (u32)(chunk / chunks_per);
        const u64 lo = (chunk % chunks_per) * S2_CHUNK;
        const u64 cnt1 = min(cursors1[b1 * SK_CURSOR_STRIDE], cap1);
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
                if (INDEX) mt[u] = __ldg(in_rec + 4u * src + 2u);       // the meta word of the slot
                else sk_load_slot(in_rec, src, bs[u], mt[u]);
                rank[u] = atomicAdd(&hist[meta_b2(mt[u])], 1u);
            }
        }
        __syncthreads();
        {   // one returned atomic per non-empty final bucket, all of a thread's (<= 4) in flight together
            u64 got[1024 / S2_THREADS];
#pragma unroll
            for (u32 u = 0; u < 1024 / S2_THREADS; ++u) {
                const u32 p = threadIdx.x + u * S2_THREADS;
                const u32 c = p < n_l2 ? hist[p] : 0u;
                got[u] = c ? atomicAdd((unsigned long long*)&cursors2[((u64)b1 << l2_bits) + p], (unsigned long long)c) : 0ull;
            }
#pragma unroll
            for (u32 u = 0; u < 1024 / S2_THREADS; ++u) {
                const u32 p = threadIdx.x + u * S2_THREADS;
                if (p < n_l2) gbase[p] = got[u];
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < S2_PER; ++u) {
            const u64 i = lo + (u64)u * S2_THREADS + threadIdx.x;
            if (i < cnt1) {
                const u64 dst = gbase[meta_b2(mt[u])] + rank[u];
                if (INDEX) {
                    out_index[dst] = (u32)i;
                } else {
                    out_bases[dst] = bs[u];
                    out_meta[dst] = mt[u];
                }
            }
        }
        __syncthreads();
    }
}

// Index form with the chunk sorted by final bucket before it is written: a chunk of 8192 slots holds ~8 entries per
// final bucket, which leave as one 32-byte run (lanes next to each other write addresses next to each other) instead
// of eight 4-byte stores into sectors that other CTAs are writing too.
constexpr int S3_THREADS = 512;
constexpr int S3_PER = 16;
constexpr u32 S3_CHUNK = S3_THREADS * S3_PER;
struct S3Shared {
    u64 gbase[1024];
    u32 hist[1024];
    u32 scan[1024];
    u32 sorted_i[S3_CHUNK];
    u16 sorted_b[S3_CHUNK];
    u32 wsum[32];
};
__global__ void __launch_bounds__(S3_THREADS)
sk_index_buckets_sorted_kernel(const u64* __restrict__ in_rec, u64 cap1, const u64* __restrict__ cursors1, u32 n_l1,
                               int l2_bits, u64* __restrict__ cursors2, u32* __restrict__ out_index) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S3Shared& sm = *reinterpret_cast<S3Shared*>(smem_raw);
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u64 chunks_per = (cap1 + S3_CHUNK - 1) / S3_CHUNK;
    const u64 total = (u64)n_l1 * chunks_per;
    for (u64 chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        const u32 b1 = (u32)(chunk / chunks_per);
        const u64 lo = (chunk % chunks_per) * S3_CHUNK;
        const u64 cnt1 = min(cursors1[b1 * SK_CURSOR_STRIDE], cap1);
        if (lo >= cnt1) continue;                           // CTA-uniform
        sm.hist[tid] = 0;
        sm.hist[tid + S3_THREADS] = 0;
        __syncthreads();
        u32 br[S3_PER];                                     // level-2 bucket | rank inside the chunk << 16
#pragma unroll
        for (int u = 0; u < S3_PER; ++u) {
            const u64 i = lo + (u64)u * S3_THREADS + tid;
            br[u] = 0;
            if (i < cnt1) {
                const u32 b2 = meta_b2(__ldg(in_rec + 4u * ((u64)b1 * cap1 + i) + 2u));
                br[u] = b2 | (atomicAdd(&sm.hist[b2], 1u) << 16);
            }
        }
        __syncthreads();
        {   // exclusive scan of the 1024 counts (two per thread) and the buckets' places in the index
            const u32 a = sm.hist[2u * tid], b = sm.hist[2u * tid + 1u];
            u32 incl = a + b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(FULL, incl, o);
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
        for (int u = 0; u < S3_PER; ++u) {
            const u64 i = lo + (u64)u * S3_THREADS + tid;
            if (i < cnt1) {
                const u32 b2 = br[u] & 0xFFFFu;
                const u32 pos = sm.scan[b2] + (br[u] >> 16);
                sm.sorted_i[pos] = (u32)i;
                sm.sorted_b[pos] = (u16)b2;
            }
        }
        __syncthreads();
        const u32 n = (u32)min(cnt1 - lo, (u64)S3_CHUNK);
        for (u32 t = tid; t < n; t += S3_THREADS) {
            const u32 b2 = sm.sorted_b[t];
            out_index[sm.gbase[b2] + (t - sm.scan[b2])] = sm.sorted_i[t];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// 3. one CTA per bucket: exact counts, solid windows, candidate edge stamps -- ONE walk over the records
//
// Work is flattened per warp: each lane loads one record, the warp prefix-sums the window counts and
// then walks the windows of its 32 records 32 at a time, so every lane probes an independent window
// (full lanes, no serial key roll, balanced warps).  Table and stamps live in shared memory and are
// addressed with explicit ld/atom.shared.
//
// Table = two arrays over the same slots: K[slot] the window's 64-bit key (all ones: empty), claimed with one
// 64-bit CAS, and A[slot] a 32-bit state word that only ever moves forward, by 32-bit CAS:
//     0                   nothing recorded yet
//     FIRST  | rec + 1    seen once, in record `rec` of the bucket (0: that occurrence has no next symbol)
//     REPEAT | count      seen `count` times, 2 <= count <= threshold: still below the filter
//     SOLID  | index      count > threshold: `index` names its key + 4 stamp slots (PENDING while they are set up)
// cand[p][c] = min ordinal of "p followed by c" is taken on the spot for every occurrence that finds its window
// SOLID (compare, then a 64-bit CAS only when the stored ordinal is larger).  An occurrence that comes earlier
// cannot know whether its window will pass the filter: it leaves a 4-byte note (slot, record) in a queue -- the
// first occurrence through the FIRST state, written to the queue by the second one.  After the walk the notes
// whose slot ended SOLID (a few hundred per bucket: the first `threshold` occurrences of each solid window)
// re-read their one record and fold in the stamp; all the others (sequencing-error windows seen two or three
// times) are dropped after one look at A.  No record is walked twice: the flagged second walk of the previous
// generation (23 % of the kernel: 35 % of all records on C4, nearly all of them because of error windows that
// never become solid) is gone, together with the per-record flags.
#ifndef GA_SB_THREADS
#define GA_SB_THREADS 512
#endif
#ifndef GA_SB_CTAS
#define GA_SB_CTAS 2
#endif
constexpr int SB_THREADS = GA_SB_THREADS;
constexpr int SB_CTAS_PER_SM = GA_SB_CTAS;
// dynamic shared memory per CTA (table + queue + solid windows): what is left of the SM's 227 KB after the
// per-CTA static control block and the 1 KB the system reserves per CTA
constexpr u32 SB_POOL_BYTES = ((232448u / GA_SB_CTAS - 1024u - 1536u) / 1024u) * 1024u;
constexpr u32 SB_MAX_SLOTS = 8192;             // upper bound of the table_slots argument
constexpr u32 SB_PROBE_MAX = 192;
constexpr u32 SB_SLOT_BYTES = 12;              // key 8 + state 4
constexpr u32 SB_SOLID_BYTES = 40;             // key 8 + 4 stamps
constexpr u32 SB_NOTE_BYTES = 8;
constexpr u32 SB_NOTE_SPILL = 16384;           // notes per CTA in the global overflow buffer (256 KB)
#ifndef GA_SK_TAIL
#define GA_SK_TAIL 16
#endif
#ifndef GA_SK_TAIL_ROUNDS2
#define GA_SK_TAIL_ROUNDS2 2                      // short spans once <= this many half-rounds of records are left
#endif
constexpr u32 SB_MAX_SEG = 16;                // sources a bucket can be gathered from (multi-GPU exchange)

__device__ __forceinline__ u32 sk_slot_hash(u64 key) {
    u32 h = (u32)key * 0x9E3779B1u ^ (u32)(key >> 32) * 0x85EBCA77u;
    h ^= h >> 15;
    return h * 0xC2B2AE3Du;                     // use the TOP bits
}

constexpr u32 SA_FIRST = 1u << 30, SA_REPEAT = 2u << 30, SA_SOLID = 3u << 30, SA_PAYLOAD = (1u << 30) - 1u;
constexpr u32 SA_PENDING = SA_SOLID | SA_PAYLOAD;

// ---- table, queue and solid storage: shared memory (byte addresses in the shared window) or global scratch
struct MemShared {
    u32 keys, state, queue, skeys, stamps;      // shared byte addresses
    __device__ __forceinline__ u64 cas_k(u32 s, u64 cmp, u64 val) const {
        u64 old;
        asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(keys + 8u * s), "l"(cmp), "l"(val) : "memory");
        return old;
    }
    __device__ __forceinline__ u32 ld_a(u32 s) const {
        u32 v;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(state + 4u * s));
        return v;
    }
    __device__ __forceinline__ u32 cas_a(u32 s, u32 cmp, u32 val) const {
        u32 old;
        asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(state + 4u * s), "r"(cmp), "r"(val) : "memory");
        return old;
    }
    __device__ __forceinline__ void st_a(u32 s, u32 v) const {
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(state + 4u * s), "r"(v) : "memory");
    }
    __device__ __forceinline__ void clear(u32 cap, u32 tid, u32 T) const {
        for (u32 s = tid; s < cap; s += T) {
            asm volatile("st.shared.u64 [%0], %1;" ::"r"(keys + 8u * s), "l"(GA_NONE64) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(state + 4u * s), "r"(0u) : "memory");
        }
    }
    __device__ __forceinline__ void ld_k2(u32 s0, u64& k0, u64& k1) const {     // slots s0 (even) and s0 + 1
        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(k0), "=l"(k1) : "r"(keys + 8u * s0));
    }
    // notes, 8 bytes: first << 63 | slot << 49 | payload (c << 47 | ordinal, or what the FIRST state held)
    __device__ __forceinline__ void q_st(u32 i, bool first, u32 slot, u64 payload) const {
        const u64 v = ((u64)first << 63) | ((u64)slot << 49) | payload;
        asm volatile("st.shared.u64 [%0], %1;" ::"r"(queue + 8u * i), "l"(v) : "memory");
    }
    __device__ __forceinline__ void q_ld(u32 i, bool& first, u32& slot, u64& payload) const {
        u64 v;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(queue + 8u * i));
        first = v >> 63;
        slot = (u32)(v >> 49) & 0x3FFFu;
        payload = v & ((1ull << 49) - 1ull);
    }
    __device__ __forceinline__ void skey_st(u32 i, u64 v) const {
        asm volatile("st.shared.u64 [%0], %1;" ::"r"(skeys + 8u * i), "l"(v) : "memory");
    }
    __device__ __forceinline__ u64 skey_ld(u32 i) const {
        u64 v;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(skeys + 8u * i));
        return v;
    }
    __device__ __forceinline__ void stamp_st(u32 i, u64 v) const {
        asm volatile("st.shared.u64 [%0], %1;" ::"r"(stamps + 8u * i), "l"(v) : "memory");
    }
    __device__ __forceinline__ u64 stamp_ld(u32 i) const {
        u64 v;
        asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(stamps + 8u * i));
        return v;
    }
    __device__ __forceinline__ u64 stamp_cas(u32 i, u64 cmp, u64 val) const {
        u64 old;
        asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(stamps + 8u * i), "l"(cmp), "l"(val) : "memory");
        return old;
    }
};

struct MemGlobal {
    u64* keys;
    u32* state;
    u64* queue;
    u64* skeys;
    u64* stamps;
    __device__ __forceinline__ u64 cas_k(u32 s, u64 cmp, u64 val) const {
        return atomicCAS((unsigned long long*)(keys + s), cmp, val);
    }
    __device__ __forceinline__ u32 ld_a(u32 s) const { return ((volatile u32*)state)[s]; }
    __device__ __forceinline__ u32 cas_a(u32 s, u32 cmp, u32 val) const { return atomicCAS(state + s, cmp, val); }
    __device__ __forceinline__ void st_a(u32 s, u32 v) const { ((volatile u32*)state)[s] = v; }
    __device__ __forceinline__ void clear(u32 cap, u32 tid, u32 T) const {
        for (u32 s = tid; s < cap; s += T) {
            keys[s] = GA_NONE64;
            state[s] = 0u;
        }
    }
    __device__ __forceinline__ void ld_k2(u32 s0, u64& k0, u64& k1) const {
        k0 = ((volatile u64*)keys)[s0];
        k1 = ((volatile u64*)keys)[s0 + 1u];
    }
    // (tables beyond 2^14 slots: two words per note)
    __device__ __forceinline__ void q_st(u32 i, bool first, u32 slot, u64 payload) const {
        queue[2u * (size_t)i] = ((u64)first << 63) | slot;
        queue[2u * (size_t)i + 1u] = payload;
    }
    __device__ __forceinline__ void q_ld(u32 i, bool& first, u32& slot, u64& payload) const {
        const u64 v = queue[2u * (size_t)i];
        first = v >> 63;
        slot = (u32)v;
        payload = queue[2u * (size_t)i + 1u];
    }
    __device__ __forceinline__ void skey_st(u32 i, u64 v) const { skeys[i] = v; }
    __device__ __forceinline__ u64 skey_ld(u32 i) const { return skeys[i]; }
    __device__ __forceinline__ void stamp_st(u32 i, u64 v) const { stamps[i] = v; }
    __device__ __forceinline__ u64 stamp_ld(u32 i) const { return ((volatile u64*)stamps)[i]; }
    __device__ __forceinline__ u64 stamp_cas(u32 i, u64 cmp, u64 val) const {
        return atomicCAS((unsigned long long*)(stamps + i), cmp, val);
    }
};

struct BucketCtl {     // shared-memory control block of one CTA
    u32 n_solid;       // solid windows of this pass
    u32 n_q;           // notes in the queue
    u32 overflow;
    u32 bucket;
    u32 n_distinct;
    u64 out_base;
    u32 sp;            // pending (parts << 16 | part) items of the current bucket
    // what a bucket of x windows will hold -- distinct windows, notes, solid windows -- is predicted by a running
    // straight-line fit y = a + b x per quantity over the buckets this CTA has done: sequencing-error windows
    // whose minimizer was hit by the error land in a random bucket, so every bucket gets about the same number of
    // them whatever its size, and a plain ratio over-estimates the large buckets (needless passes) and
    // under-estimates the small ones (failed passes)
    float mx, mxx, my[3], mxy[3];
    u32 n_obs;
    u32 est[3];        // the prediction for the current bucket (all its passes together)
    u32 n_seg;
    u32 next_batch;    // dynamic record hand-out of the walk
    u32 stack[40];
    u64 seg_lo[SB_MAX_SEG];        // the bucket's records: segment s holds [seg_lo[s], +seg_pre[s+1]-seg_pre[s])
    u64 seg_pre[SB_MAX_SEG + 1];
    // sources form (multi-GPU, records stay where they were cut): segment s is source rank s; its index entries
    // sit in seg_idx[s], its slots in seg_rec[s] (= first slot of this bucket's level-1 bucket on that rank)
    const u64* seg_rec[SB_MAX_SEG];
    const u32* seg_idx[SB_MAX_SEG];
};

// Sources form: the level-1 slots and the bucket-sorted index of every source rank, mapped into this process
// over NVLink (ga_peer_open); entry s is rank s's.  n == 0: not used.
struct SkSources {
    const u64* rec[SB_MAX_SEG];
    const u32* index[SB_MAX_SEG];
    u64 cap1[SB_MAX_SEG];
    u64 first_bucket;              // global id of this launch's bucket 0 (the level-1 bucket comes from the global id)
    u64* counter;                  // solid windows written so far, shared by all ranks (the output is one rank's), or null
    u32 n;
};

// where a bucket's records are: index == nullptr: at the positions the offsets give, bases and meta in two
// dense arrays (dense form); otherwise offsets address `index`, `bases` points at the level-1 buckets'
// 32-byte slots and the record is slot base + index[position] (base = first slot of the bucket's level-1
// bucket; `meta` is not used)
struct SkGather {
    const u32* index;
    u64 base;
    bool sources;                  // sources form: index / slots per segment in the control block
};
constexpr u32 SK_ENT_BITS = 26, SK_ENT_MASK = (1u << SK_ENT_BITS) - 1u;    // index entry | segment << 26

__device__ __forceinline__ u64 shfl64(u64 v, u32 src) {
    const u32 lo = __shfl_sync(FULL, (u32)v, src), hi = __shfl_sync(FULL, (u32)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}

// One batch = up to 32 records, one per lane, held by lanes 0..n-1.  The windows of the batch are dealt to
// the lanes 32 at a time (warp prefix sum + ballot/REDUX find the owner record of each window).
// f(top, ord, follows, where) is called once per window: `top` holds the window's symbols from bit 63
// down (key = top >> (64 - 2w), the symbol after it right below), `ord` its occurrence ordinal, `follows`
// whether a next symbol exists, `where` = (record tag + 1) << 5 | window number inside the record (the tag is
// what finds the record again: its index entry, or its number inside the bucket in the dense form).
// A record's tag (its index entry, or its number inside the bucket: below 2^25, the host checks) travels with the
// number of identical records it stands for (warp-level merge below): tag | (copies - 1) << 26.
[[maybe_unused]] constexpr u32 SK_TAG_BITS = 26, SK_TAG_MASK = (1u << SK_TAG_BITS) - 1u;

// Optional (-DGA_SK_MERGE, off): identical records inside one batch of 32 are merged before their windows are
// dealt out.  At the coverage of real read sets most records are exact copies of one another (BASELINE config
// C4: 300x -- every genomic super-k-mer is cut out of ~240 reads, 44 % of the copies untouched by sequencing
// errors, read ends or the 32-window cut), and a bucket holds only ~6 genomic sites, so a random batch of 32
// records carries each frequent record 2-3 times.  Copies have the same windows with the same next symbols;
// their ordinals differ by a constant.  So the lowest lane of a group keeps the record with the group's smallest
// ordinal and a weight (copies), the other lanes drop theirs: counts add the weight, stamps take the smallest
// ordinal -- bit-identical results (GPU suite green with the flag on).  Measured on C4 (profiles/r02): 23.2 % of
// the windows never reach the table (9.30e9 of 12.1e9; a CPU model of random batches predicts 24 %, of batches
// sorted by site 60 %), yet the kernel takes 143.1 ms against 138.9 ms: the windows that disappear are the cheap
// ones (a probe that hits, one state load, one stamp compare), the insertions and state changes of the error
// windows stay, and MATCH x3 + the gap-closing shuffles cost more than the 1.7 dealt rounds they save.
// Only the symbols the record uses take part in the comparison (the 64-symbol field runs on into the read).
[[maybe_unused]] __device__ __forceinline__ void sk_merge_copies(u64& hi, u64& lo, u64& meta, u32& rtag, bool& have, int w) {
    const u32 lane = threadIdx.x & 31u;
    const u32 used = have ? meta_windows(meta) + (u32)w - 1u + (meta_has_next(meta) ? 1u : 0u) : 0u;   // <= 63 symbols
    const u64 mh = used >= 32u ? hi : (used ? hi & ~(~0ull >> (2u * used)) : 0ull);
    const u64 ml = used <= 32u ? 0ull : lo & ~(~0ull >> (2u * (used - 32u)));
    const u32 sig = have ? (u32)(meta & 63u) : 64u + lane;          // windows, has_next; empty lanes stay alone
    const u32 group = __match_any_sync(FULL, mh) & __match_any_sync(FULL, ml) & __match_any_sync(FULL, sig);
    const u32 leader = (u32)__ffs(group) - 1u;
    u32 rest = group & (group - 1u);                                 // the group's lanes above its leader
    u32 most = __reduce_max_sync(FULL, (u32)__popc(rest));
    if (most == 0u) return;                                          // warp-uniform: no two records alike
    const u64 ord0 = meta_ordinal(meta);
    u64 omin = ord0;
    for (; most; --most) {
        const u32 src = rest ? (u32)__ffs(rest) - 1u : lane;
        rest &= rest - 1u;
        omin = min(omin, shfl64(ord0, src));
    }
    if (lane == leader) {
        meta = (omin << 16) | (meta & 0xFFFFull);
        rtag |= ((u32)__popc(group) - 1u) << SK_TAG_BITS;
    } else {
        have = false;
    }
    // the walk deals windows to records held by lanes 0..n-1: close the gaps the dropped copies left
    const u32 keep = __ballot_sync(FULL, have);
    const u32 src = __fns(keep, 0u, (int)lane + 1);                  // lane of the (lane+1)-th survivor, or all ones
    have = src < 32u;
    const u32 from = have ? src : lane;
    hi = shfl64(hi, from);
    lo = shfl64(lo, from);
    meta = shfl64(meta, from);
    rtag = __shfl_sync(FULL, rtag, from);
}

template <class F>
__device__ __forceinline__ void sk_for_each_window(u64 rhi, u64 rlo, u64 meta, u32 rtag, bool have, F&& f) {
    const u32 lane = threadIdx.x & 31u;
    const u32 npiece = have ? meta_windows(meta) : 0u;
    u32 incl = npiece;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const u32 t = __shfl_up_sync(FULL, incl, off);
        if (lane >= (u32)off) incl += t;
    }
    const u32 start = incl - npiece;
    const u32 total = __shfl_sync(FULL, incl, 31);
    const u32 le_mask = 0xFFFFFFFFu >> (31u - lane);
    for (u32 xb = 0; xb < total; xb += 32u) {
        const u32 before = __popc(__ballot_sync(FULL, start < xb));
        const u32 bit = (npiece && start >= xb && start < xb + 32u) ? 1u << (start - xb) : 0u;
        const u32 marks = __reduce_or_sync(FULL, bit);
        const u32 x = xb + lane;
        const bool active = x < total;
        const u32 owner = active ? before + __popc(marks & le_mask) - 1u : 0u;
        const u32 j = x - __shfl_sync(FULL, start, owner);
        const u64 ohi = shfl64(rhi, owner), olo = shfl64(rlo, owner);
        const u64 om = shfl64(meta, owner);
        const u32 otag = __shfl_sync(FULL, rtag, owner);
        if (active) {
            const u32 nwin = meta_windows(om);
            const u64 top = j ? (ohi << (2u * j)) | (olo >> (64u - 2u * j)) : ohi;
#ifdef GA_SK_MERGE
            f(top, meta_ordinal(om) + j, j + 1u < nwin || meta_has_next(om), (((otag & SK_TAG_MASK) + 1u) << 5) | j,
              (otag >> SK_TAG_BITS) + 1u);
#else
            f(top, meta_ordinal(om) + j, j + 1u < nwin || meta_has_next(om), ((otag + 1u) << 5) | j, 1u);
#endif
        }
        __syncwarp();
    }
}

// returns false when the pass does not fit (table, queue or solid area full): the caller splits it or lists it
// for the spill path
template <class Mem, bool DEEP, bool SRC>
__device__ __forceinline__ bool sk_bucket_body(const ulonglong2* __restrict__ bases, const u64* __restrict__ meta,
                                               SkGather gather, int w, u32 threshold, const Mem& mem, u32 cap,
                                               u32 q_cap, u64* __restrict__ q_spill, u32 q_spill_cap, u32 max_solid,
                                               bool count_only, u32 parts, u32 part,
                                               BucketCtl& ctl, u64* __restrict__ solid_keys_out,
                                               u64* __restrict__ edge_stamp_out, u64 out_capacity,
                                               u64* n_solid_global) {
    const u32 tid = threadIdx.x, T = blockDim.x, lane = tid & 31u, warp = tid >> 5, W = T >> 5;
    const u32 kshift = 64u - 2u * (u32)w;
    const u32 n_seg = ctl.n_seg;
    const u64 nrec = ctl.seg_pre[n_seg];
    // position of the bucket's idx-th entry (segments are few: linear scan); in the index form that entry is a
    // 32-bit index and the record sits at gather.base + index
    auto where2 = [&](u64 idx, u32& sg) -> u64 {
        sg = 0;
        while (sg + 1u < n_seg && idx >= ctl.seg_pre[sg + 1]) ++sg;
        return ctl.seg_lo[sg] + (idx - ctl.seg_pre[sg]);
    };
    auto where = [&](u64 idx) -> u64 {
        u32 sg;
        return where2(idx, sg);
    };
    const bool indexed = gather.index != nullptr || SRC;
#ifdef GA_SB_PROFILE
    // probe build: warp-cycles per phase summed into counters[8..13] = n_solid_global[7..12] (clear, walk, wait at
    // the barrier that ends the walk, notes, output, whole body)
    const long long pt0 = clock64();
#define GA_SB_TICK(slot, from) do { if (lane == 0) atomicAdd((unsigned long long*)(n_solid_global + 7 + (slot)), (unsigned long long)(clock64() - (from))); } while (0)
#else
#define GA_SB_TICK(slot, from) do { } while (0)
#endif
    // B. clear
    mem.clear(cap, tid, T);
    if (tid == 0) {
        ctl.n_solid = 0;
        ctl.n_q = 0;
        ctl.overflow = 0;
        ctl.n_distinct = 0;
        ctl.next_batch = W * 32u;      // records handed out so far (the warps' first spans are 32 each)
    }
    __syncthreads();
#ifdef GA_SB_PROFILE
    GA_SB_TICK(0, pt0);
    const long long pt1 = clock64();
#endif
    volatile u32* vovf = &ctl.overflow;
    const u32 pmask = parts - 1u;     // this pass takes the windows whose hash bits 3.. equal `part`
    // smallest ordinal of "solid window `idx` followed by symbol c"
    auto stamp = [&](u32 idx, u32 c, u64 ord) {
        const u32 at = 4u * idx + c;
        u64 cur = mem.stamp_ld(at);
        while (ord < cur) {
            const u64 old = mem.stamp_cas(at, cur, ord);
            if (old == cur) break;
            cur = old;
        }
    };
    // notes beyond the queue's room in the pool go to this CTA's slice of a global buffer (two words each): the
    // queue is sized from a running estimate, and one bucket with many repeated error windows must not cost a pass
    auto note = [&](bool first, u32 slot, u64 payload) {
        const u32 at = atomicAdd(&ctl.n_q, 1u);
        if (at < q_cap) {
            mem.q_st(at, first, slot, payload);
        } else if (at - q_cap < q_spill_cap) {
            q_spill[2u * (size_t)(at - q_cap)] = ((u64)first << 63) | slot;
            q_spill[2u * (size_t)(at - q_cap) + 1u] = payload;
        } else {
            *vovf = 2u;
        }
    };
    // C. the walk.  Hand-out in RECORDS: a warp takes 32 records at a time while more than a round's worth is left
    // and GA_SK_TAIL records at a time after that, so that the last round of a bucket is not left to a third of
    // the warps.  A span is start | (short ? 1 << 31 : 0); ctl.next_batch counts records.
    u32 inserted = 0;
    constexpr u32 TAIL = GA_SK_TAIL;
    const u32 nrec32 = (u32)min(nrec, (u64)0x7FFFFFFFu);
    auto next_span = [&]() -> u32 {
        u32 v = 0;
        if (lane == 0) {
            const u32 cur = *(volatile u32*)&ctl.next_batch;
            const bool shortspan = cur < nrec32 && nrec32 - cur <= W * 16u * GA_SK_TAIL_ROUNDS2;
            v = atomicAdd(&ctl.next_batch, shortspan ? TAIL : 32u) | (shortspan ? 0x80000000u : 0u);
        }
        return __shfl_sync(FULL, v, 0);
    };
    auto span_take = [&](u32 sp) -> u32 { return (sp >> 31) ? TAIL : 32u; };
    // In the index form the record is two dependent loads away (index entry, then the gather): the next
    // span's index entry is requested one span ahead.
    auto entry = [&](u32 sp) -> u32 {
        const u32 idx = (sp & 0x7FFFFFFFu) + lane;
        if (!indexed || lane >= span_take(sp) || idx >= nrec32) return 0u;
        if (!SRC) return __ldg(gather.index + where(idx));
        u32 sg;
        const u64 at = where2(idx, sg);
        return __ldg(ctl.seg_idx[sg] + at) | (sg << SK_ENT_BITS);        // the entry may live on another GPU
    };
    // this lane's record of span `sp` (index entry `e`): nothing for lanes beyond the span or the bucket
    auto fetch = [&](u32 sp, u32 e, ulonglong2& b, u64& mt) {
        const u64 idx = (u64)(sp & 0x7FFFFFFFu) + lane;
        b = make_ulonglong2(0, 0);
        mt = 0;
        if (lane >= span_take(sp) || idx >= nrec) return;
        if (SRC) {
            sk_load_slot(ctl.seg_rec[e >> SK_ENT_BITS], e & SK_ENT_MASK, b, mt);   // local or over NVLink
        } else if (gather.index) {
            sk_load_slot((const u64*)bases, gather.base + e, b, mt);       // a 32-byte slot of the level-1 bucket
        } else {
            const u64 i = where(idx);
            b = bases[i];
            mt = meta[i];
        }
    };
    // DEEP (sources form, GA_SK_DEEP=1): records that sit on another GPU are 2-3 us away, twice over (index entry,
    // then the slot).  The walk then runs two spans ahead: index entries are requested two spans before their
    // records are dealt, the records themselves one span before.  Measured on C4 at N = 2 (profiles/r02): bucket
    // kernel 90.0 / 107.3 ms on the two ranks against 91.1 / 101.1 ms on demand -- what the remote gather costs
    // (the same kernel takes 67.9 ms over local records) is not latency but the rate at which NVLink serves lone
    // 32-byte reads (0.46e9 of them per rank: ~5e9/s), so the depth buys nothing and costs balance at the end
    // of a bucket.  Off by default; on one GPU the same depth was +4.5 % (round 1).
    u32 ent = entry(warp * 32u);
    u32 sp_ahead = 0, ent_ahead = 0;
    ulonglong2 b = make_ulonglong2(0, 0);
    u64 mt = 0;
    if (DEEP) {
        sp_ahead = next_span();
        ent_ahead = entry(sp_ahead);
        fetch(warp * 32u, ent, b, mt);
    }
    for (u32 sp = warp * 32u, sp_next = 0; (sp & 0x7FFFFFFFu) < nrec32 && !*vovf; sp = sp_next) {
        const u32 bt = sp & 0x7FFFFFFFu;               // first record of the span
        const u64 idx = (u64)bt + lane;
        const bool have = lane < span_take(sp) && idx < nrec;
        ulonglong2 b_ahead = make_ulonglong2(0, 0);
        u64 mt_ahead = 0;
        u32 ent_next;
        if (DEEP) {
            const u32 sp_far = next_span();
            const u32 ent_far = entry(sp_far);
            fetch(sp_ahead, ent_ahead, b_ahead, mt_ahead);
            sp_next = sp_ahead;
            ent_next = ent_ahead;
            sp_ahead = sp_far;
            ent_ahead = ent_far;
        } else {
            sp_next = next_span();
            ent_next = entry(sp_next);
            fetch(sp, ent, b, mt);
        }
        const u32 ent_cur = ent;
        ent = ent_next;
        u32 rtag = gather.index ? ent_cur : bt + lane;     // what finds the record again (sources form: its number in the bucket)
        bool mine_left = have;
#ifdef GA_SK_MERGE
        sk_merge_copies(b.x, b.y, mt, rtag, mine_left, w);
#endif
#ifdef GA_SB_PROFILE
        {   // windows that reach the table after the merge -> counters[15]
            const u32 left = __reduce_add_sync(FULL, mine_left ? meta_windows(mt) : 0u);
            if (lane == 0) atomicAdd((unsigned long long*)(n_solid_global + 14), (unsigned long long)left);
        }
#endif
        sk_for_each_window(b.x, b.y, mt, rtag, mine_left,
                           [&](u64 top, u64 ord, bool follows, u32 where_j, u32 copies) {
            const u64 key = top >> kshift;
            const u32 h = sk_slot_hash(key);
            if (((h >> 3) & pmask) != part) return;
            follows = follows && !count_only;
            // probe two neighbouring slots at a time (one 128-bit load): half the iterations, and the warp waits
            // for its slowest lane
            u32 s0 = __umulhi(h, cap >> 1) << 1, s;
            for (u32 probes = 0;; ++probes) {
                if (probes > SB_PROBE_MAX) {
                    *vovf = 1u;                            // probe limit: table full
                    return;
                }
                u64 K0, K1;
                mem.ld_k2(s0, K0, K1);
                if (K0 == key) {
                    s = s0;
                    break;
                }
                if (K1 == key) {
                    s = s0 + 1u;
                    break;
                }
                if (K0 == GA_NONE64 || K1 == GA_NONE64) {
                    s = K0 == GA_NONE64 ? s0 : s0 + 1u;
                    const u64 old = mem.cas_k(s, GA_NONE64, key);
                    if (old == GA_NONE64) {
                        ++inserted;
                        break;
                    }
                    if (old == key) break;
                    continue;                              // somebody else's key landed there: look at the pair again
                }
                s0 = s0 + 2u >= cap ? 0u : s0 + 2u;
            }
            const u32 c = (u32)(top >> (kshift - 2u)) & 3u;
            const u64 mine = ((u64)c << 47) | ord;         // what a note about this occurrence holds
            u32 a = mem.ld_a(s);
            for (;;) {
                if (a >= SA_SOLID) {                       // SOLID: the stamp, or a note while the slots are set up
                    if (follows) {
                        if (a != SA_PENDING) stamp(a & SA_PAYLOAD, c, ord);
                        else note(false, s, mine);
                    }
                    return;
                }
                const u32 cnt = a == 0u ? 0u : (a < SA_REPEAT ? 1u : a & SA_PAYLOAD);
                if (cnt + copies > threshold) {            // this occurrence (with its copies) takes the window above the threshold
                    const u32 old = mem.cas_a(s, a, SA_PENDING);
                    if (old != a) {
                        a = old;
                        continue;
                    }
                    const u32 at = atomicAdd(&ctl.n_solid, 1u);
                    if (at >= max_solid) {
                        *vovf = 4u;
                        return;
                    }
                    mem.skey_st(at, key);
#pragma unroll
                    for (u32 q = 0; q < 4u; ++q) mem.stamp_st(4u * at + q, (follows && q == c) ? ord : GA_NONE64);
                    __threadfence_block();
                    mem.st_a(s, SA_SOLID | at);
                    if (a >= SA_FIRST && a < SA_REPEAT && (a & SA_PAYLOAD)) note(true, s, a & SA_PAYLOAD);
                    return;
                }
                if (a == 0u && copies == 1u) {             // first occurrence: where to find it again goes in the slot
                    const u32 old = mem.cas_a(s, 0u, SA_FIRST | (follows ? where_j : 0u));
                    if (old == 0u) return;
                    a = old;
                    continue;
                }
                // second .. threshold-th occurrence: count, leave a note (and move the first one to the queue)
                const u32 old = mem.cas_a(s, a, SA_REPEAT | (cnt + copies));
                if (old != a) {
                    a = old;
                    continue;
                }
                if (a < SA_REPEAT && (a & SA_PAYLOAD)) note(true, s, a & SA_PAYLOAD);
                if (follows) note(false, s, mine);
                return;
            }
        });
        if (DEEP) {
            b = b_ahead;
            mt = mt_ahead;
        }
    }
    for (int off = 16; off > 0; off >>= 1) inserted += __shfl_down_sync(FULL, inserted, off);
    if (lane == 0 && inserted) atomicAdd(&ctl.n_distinct, inserted);
#ifdef GA_SB_PROFILE
    GA_SB_TICK(1, pt1);
    const long long pt2 = clock64();
#endif
    __syncthreads();
#ifdef GA_SB_PROFILE
    GA_SB_TICK(2, pt2);
    const long long pt3 = clock64();
#endif
    if (ctl.overflow) return false;
    const u32 n_solid = ctl.n_solid;
    if (n_solid == 0) return true;
    if (tid == 0) ctl.out_base = atomicAdd((unsigned long long*)n_solid_global, (unsigned long long)n_solid);
    // D. the notes whose window ended solid.  A note of the occurrence itself carries (c, ordinal); the note that
    //    moved a FIRST state to the queue names the record and the window number: one 32-byte load.
    const u32 n_q = ctl.n_q;
    for (u32 i = tid; i < n_q; i += T) {
        bool first;
        u32 s;
        u64 payload;
        if (i < q_cap) {
            mem.q_ld(i, first, s, payload);
        } else {
            const u64 v = q_spill[2u * (size_t)(i - q_cap)];
            first = v >> 63;
            s = (u32)v;
            payload = q_spill[2u * (size_t)(i - q_cap) + 1u];
        }
        const u32 a = mem.ld_a(s);
        if (a < SA_SOLID) continue;
        const u32 idx = a & SA_PAYLOAD;
        if (!first) {
            stamp(idx, (u32)(payload >> 47) & 3u, payload & ((1ull << 47) - 1ull));
            continue;
        }
        const u32 tag = (u32)(payload >> 5) - 1u, j = (u32)payload & 31u;
        ulonglong2 b;
        u64 mt;
        if (SRC) {
            u32 sg;
            const u64 at = where2(tag, sg);
            sk_load_slot(ctl.seg_rec[sg], __ldg(ctl.seg_idx[sg] + at), b, mt);
        } else if (gather.index) {
            sk_load_slot((const u64*)bases, gather.base + tag, b, mt);
        } else {
            const u64 at = where(tag);
            b = bases[at];
            mt = meta[at];
        }
        const u64 top = j ? (b.x << (2u * j)) | (b.y >> (64u - 2u * j)) : b.x;
        stamp(idx, (u32)(top >> (kshift - 2u)) & 3u, meta_ordinal(mt) + j);
    }
    __syncthreads();
#ifdef GA_SB_PROFILE
    GA_SB_TICK(3, pt3);
    const long long pt4 = clock64();
#endif
    // E. output
    const u64 base = ctl.out_base;
    if (base + n_solid <= out_capacity) {
        for (u32 s = tid; s < n_solid; s += T) solid_keys_out[base + s] = mem.skey_ld(s);
        if (edge_stamp_out)
            for (u32 s = tid; s < 4 * n_solid; s += T) edge_stamp_out[4 * base + s] = mem.stamp_ld(s);
    }
#ifdef GA_SB_PROFILE
    GA_SB_TICK(4, pt4);
    GA_SB_TICK(5, pt0);
#endif
    return true;
}

// counters: [0] next bucket, [1] solid windows so far, [2] passes listed for the spill path,
// [3] passes run | failed passes << 32, [4] distinct windows and [5] notes over the passes that fitted (statistics)
// hist: per bucket, records << 32 | windows
//
// A bucket is done in `parts` passes (a power of two), pass `part` taking the windows whose hash
// bits select it, so that the distinct windows, the notes and the solid windows of one pass fit the CTA's pool.
// parts comes from running estimates of the three per window of the bucket (the first buckets of a CTA start
// pessimistic); a pass that still does not fit is split in two; only passes that would need more than 32 parts
// (or buckets of more than 65536 records) go to the spill list (entry = bucket | parts << 32 | part << 48).
template <bool DEEP, bool SRC>
__global__ void __launch_bounds__(SB_THREADS, SB_CTAS_PER_SM)
sk_bucket_kernel(const ulonglong2* __restrict__ bases, const u64* __restrict__ meta, const u64* __restrict__ offsets,
                 u32 n_seg, const u64* __restrict__ hist, u64 n_buckets, int w, u32 threshold, u32 cap_limit,
                 u32 solid_limit,
                 u64* __restrict__ solid_keys_out, u64* __restrict__ edge_stamp_out, u64 out_capacity, u64* counters,
                 u64* __restrict__ spill_list, u64 spill_capacity, u32* status, const u32* __restrict__ index,
                 u64 l1_capacity, int l2_bits, u64* __restrict__ note_spill, u32 note_spill_cap,
                 const __grid_constant__ SkSources src) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BucketCtl ctl;
    // keep the pool's shared address in a register: left to itself the compiler re-derives it from the
    // CTA's shared window (S2UR + ULEA) inside the probe loops
    u32 pool;
    {
        const u32 raw = (u32)__cvta_generic_to_shared(smem_raw);
        asm volatile("mov.u32 %0, %1;" : "=r"(pool) : "r"(raw));
    }
    if (threadIdx.x == 0) ctl.n_obs = 0;
#ifdef GA_SB_PROFILE
    const long long pk0 = clock64();
#endif
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) ctl.bucket = (u32)min((u64)atomicAdd((unsigned long long*)&counters[0], 1ull), n_buckets);
        __syncthreads();
        const u64 b = ctl.bucket;
        if (b >= n_buckets) {
#ifdef GA_SB_PROFILE
            if ((threadIdx.x & 31u) == 0) atomicAdd((unsigned long long*)&counters[14], (unsigned long long)(clock64() - pk0));
#endif
            break;
        }
        const u64 nw = hist[b] & 0xFFFFFFFFull;
        if (nw == 0) continue;
        if (threadIdx.x == 0) {       // segment s of the bucket: offsets[s][b] .. offsets[s][b+1]
            u64 run = 0;
            for (u32 sg = 0; sg < n_seg; ++sg) {
                const u64 lo = offsets[sg * (n_buckets + 1) + b], hi = offsets[sg * (n_buckets + 1) + b + 1];
                ctl.seg_lo[sg] = lo;
                ctl.seg_pre[sg] = run;
                run += hi - lo;
            }
            ctl.seg_pre[n_seg] = run;
            ctl.n_seg = n_seg;
            if (SRC)
                for (u32 sg = 0; sg < src.n; ++sg) {
                    ctl.seg_rec[sg] = src.rec[sg] + 4u * (((src.first_bucket + b) >> l2_bits) * src.cap1[sg]);
                    ctl.seg_idx[sg] = src.index[sg];
                }
            // prediction for this bucket: pessimistic ratios until the fit has seen a few buckets
            const float x = (float)nw;
            const float first_guess[3] = {0.32f, 0.2f, 0.03f};
            for (int q = 0; q < 3; ++q) {
                float y = first_guess[q] * x + 256.f;
                if (ctl.n_obs >= 4u) {
                    const float var = ctl.mxx - ctl.mx * ctl.mx;
                    float bq = var > 1e-4f * ctl.mx * ctl.mx ? (ctl.mxy[q] - ctl.mx * ctl.my[q]) / var : 0.f;
                    float aq = ctl.my[q] - bq * ctl.mx;
                    if (bq < 0.f || aq < 0.f) {               // degenerate fit: fall back to the plain ratio
                        bq = ctl.my[q] / fmaxf(ctl.mx, 1.f);
                        aq = 0.f;
                    }
                    y = aq + bq * x;
                }
                ctl.est[q] = (u32)fminf(fmaxf(y, 0.f), 4.0e9f);
            }
            // passes: the expected solid windows (x 2), notes (x 1.25: what does not fit goes to the global note
            // buffer) and distinct windows (table at a load of at most 0.62) of one pass must fit the pool (a pass
            // that overflows anyway is split below)
            const u64 est_d = (u64)ctl.est[0] + 48u, est_q = (u64)ctl.est[1] + 32u, est_s = (u64)ctl.est[2] + 16u;
            u32 parts = 1;
            while (parts < 32u) {
                const u64 d = est_d / parts * 13u / 8u, q = est_q / parts * 5u / 4u + 64u, so = est_s / parts * 2u + 32u;
                if (d <= (u64)cap_limit && est_s / parts <= (u64)solid_limit &&
                    d * SB_SLOT_BYTES + q * SB_NOTE_BYTES + min(so, (u64)solid_limit) * SB_SOLID_BYTES + 32u <= (u64)SB_POOL_BYTES)
                    break;
                parts <<= 1;
            }
            for (u32 q = 0; q < parts; ++q) ctl.stack[q] = (parts << 16) | (parts - 1u - q);
            ctl.sp = parts;
        }
        for (;;) {
            __syncthreads();
            const u32 sp = ctl.sp;
            if (sp == 0) break;
            const u32 item = ctl.stack[sp - 1];
            const u64 ed = (u64)ctl.est[0] + 48u, eq = (u64)ctl.est[1] + 32u, es = (u64)ctl.est[2] + 16u;
            const u32 parts = item >> 16, part = item & 0xFFFFu;
            __syncthreads();           // everybody has read the item before thread 0 replaces it below (dropping this
                                       // barrier where the body's own would do measured 139.1 against 139.3 ms: noise)
            bool ok = false;
            if (parts <= 32u) {
                // the pool of this pass: solid windows and notes for the expected numbers + a margin, the table gets
                // the rest up to 2.5 x the expected distinct windows (short probe chains; more only costs clearing)
                const u32 max_solid = (u32)min(min(es / parts * 2u + 64u, (u64)solid_limit),
                                               (u64)(SB_POOL_BYTES / 4u / SB_SOLID_BYTES));
                const u32 q_cap = (u32)min(eq / parts * 5u / 4u + 96u, (u64)(SB_POOL_BYTES / 4u / SB_NOTE_BYTES));
                const u64 want_d = ed / parts * 5u / 2u + 128u;
                u32 cap = (u32)min(min(want_d, (u64)cap_limit),
                                   (u64)((SB_POOL_BYTES - max_solid * SB_SOLID_BYTES - q_cap * SB_NOTE_BYTES - 32u) / SB_SLOT_BYTES));
                cap &= ~1u;                                    // slots are probed in pairs
                if (cap < 64u) cap = 64u;
                MemShared mem;
                mem.keys = pool;
                mem.skeys = pool + 8u * cap;
                mem.stamps = mem.skeys + 8u * max_solid;
                mem.queue = mem.stamps + 32u * max_solid;
                mem.state = mem.queue + SB_NOTE_BYTES * q_cap;
                const SkGather gather{index, (b >> l2_bits) * l1_capacity, SRC};
                ok = sk_bucket_body<MemShared, DEEP, SRC>(bases, meta, gather, w, threshold, mem, cap, q_cap,
                                    note_spill + 2u * (size_t)blockIdx.x * note_spill_cap, note_spill_cap, max_solid,
                                    edge_stamp_out == nullptr, parts, part, ctl,
                                    solid_keys_out, edge_stamp_out, out_capacity, src.counter ? src.counter : counters + 1);
            }
            if (threadIdx.x == 0) {
                u32 top = sp - 1u;
                atomicAdd((unsigned long long*)&counters[3], ok ? 1ull : (1ull << 32) + 1ull);   // passes, failed passes << 32
                if (ok) {
                    atomicAdd((unsigned long long*)&counters[4], (unsigned long long)ctl.n_distinct);   // statistics
                    atomicAdd((unsigned long long*)&counters[5], (unsigned long long)ctl.n_q);
                    if (parts == 1u) {                         // a whole bucket in one pass: a clean observation for the fit
                        const float x = (float)nw, y[3] = {(float)ctl.n_distinct, (float)ctl.n_q, (float)ctl.n_solid};
                        const float wgt = ctl.n_obs < 8u ? 1.f / (float)(ctl.n_obs + 1u) : 0.125f;
                        if (ctl.n_obs == 0u) ctl.mx = ctl.mxx = 0.f;
                        ctl.mx += (x - ctl.mx) * wgt;
                        ctl.mxx += (x * x - ctl.mxx) * wgt;
                        for (int q = 0; q < 3; ++q) {
                            if (ctl.n_obs == 0u) ctl.my[q] = ctl.mxy[q] = 0.f;
                            ctl.my[q] += (y[q] - ctl.my[q]) * wgt;
                            ctl.mxy[q] += (x * y[q] - ctl.mxy[q]) * wgt;
                        }
                        ++ctl.n_obs;
                    }
                } else if (parts < 32u && top + 2u <= 40u) {
                    atomicAdd((unsigned long long*)&counters[ctl.overflow == 2u ? 7 : 6], ctl.overflow == 4u ? 1ull << 32 : 1ull);
                    ctl.stack[top++] = ((parts * 2u) << 16) | (part + parts);
                    ctl.stack[top++] = ((parts * 2u) << 16) | part;
                    // this bucket holds more than predicted: size its remaining passes for half as much again
                    for (int q = 0; q < 3; ++q) ctl.est[q] += (ctl.est[q] >> 1) + 64u;
                } else {
                    const u64 at = atomicAdd((unsigned long long*)&counters[2], 1ull);
                    const u32 lp = parts > 32u ? 1u : parts, lq = parts > 32u ? 0u : part;
                    if (at < spill_capacity) spill_list[at] = b | ((u64)lp << 32) | ((u64)lq << 48);
                    else atomicOr(status, GA_ST_TABLE_FULL);
                }
                ctl.sp = top;
            }
        }
    }
}

// spill path: the same body over global scratch (one slice per CTA), for passes that do not fit the
// shared-memory pool even after splitting
template <bool SRC>
__global__ void __launch_bounds__(SB_THREADS, 1)
sk_bucket_spill_kernel(const ulonglong2* __restrict__ bases, const u64* __restrict__ meta,
                       const u64* __restrict__ offsets, u32 n_seg, u64 n_buckets,
                       const u64* __restrict__ spill_list, u64 n_spill, int w,
                       u32 threshold, u32 cap, unsigned char* __restrict__ scratch, u64 scratch_per_cta,
                       u64* __restrict__ solid_keys_out, u64* __restrict__ edge_stamp_out, u64 out_capacity,
                       u64* counters, u32* status, const u32* __restrict__ index, u64 l1_capacity, int l2_bits,
                       const __grid_constant__ SkSources src) {
    __shared__ BucketCtl ctl;
    MemGlobal mem;
    unsigned char* mine = scratch + (u64)blockIdx.x * scratch_per_cta;
    // per slot: key 8, solid key 8 + 4 stamps 32, two 16-byte notes, state 4
    mem.keys = reinterpret_cast<u64*>(mine);
    mem.skeys = mem.keys + cap;
    mem.stamps = mem.skeys + cap;
    mem.queue = mem.stamps + 4 * (size_t)cap;
    mem.state = reinterpret_cast<u32*>(mem.queue + 4 * (size_t)cap);
    for (u64 oi = blockIdx.x; oi < n_spill; oi += gridDim.x) {
        __syncthreads();
        const u64 entry = spill_list[oi];
        const u64 b = entry & 0xFFFFFFFFull;
        const u32 parts = (u32)(entry >> 32) & 0xFFFFu, part = (u32)(entry >> 48);
        if (threadIdx.x == 0) {
            u64 run = 0;
            for (u32 sg = 0; sg < n_seg; ++sg) {
                const u64 lo = offsets[sg * (n_buckets + 1) + b], hi = offsets[sg * (n_buckets + 1) + b + 1];
                ctl.seg_lo[sg] = lo;
                ctl.seg_pre[sg] = run;
                run += hi - lo;
            }
            ctl.seg_pre[n_seg] = run;
            ctl.n_seg = n_seg;
            if (SRC)
                for (u32 sg = 0; sg < src.n; ++sg) {
                    ctl.seg_rec[sg] = src.rec[sg] + 4u * (((src.first_bucket + b) >> l2_bits) * src.cap1[sg]);
                    ctl.seg_idx[sg] = src.index[sg];
                }
        }
        __syncthreads();
        const SkGather gather{index, (b >> l2_bits) * l1_capacity, SRC};
        const bool ok = sk_bucket_body<MemGlobal, false, SRC>(bases, meta, gather, w, threshold, mem, cap, 2u * cap, nullptr, 0u, cap,
                                       edge_stamp_out == nullptr, parts, part, ctl, solid_keys_out, edge_stamp_out,
                                       out_capacity, src.counter ? src.counter : counters + 1);
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

bool is_pow2(u64 v) { return v && !(v & (v - 1)); }

// SMs of the current device (persistent grids are sized in multiples of it)
int ga_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        n <= 0)
        n = 148;
    return n;
}

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

extern "C" int ga_sk_cursor_stride(void) { return (int)SK_CURSOR_STRIDE; }

extern "C" int ga_sk_scatter_reads(const ga_reads* reads, int k, int l1_bits, int l2_bits, void* records_dev,
                                   uint64_t l1_capacity, uint64_t* l1_cursors_dev, uint64_t* hist_dev,
                                   uint32_t* status_dev, ga_stream stream) {
    if (!reads || !records_dev || !l1_cursors_dev || !hist_dev || !status_dev || l1_capacity == 0 ||
        l1_bits < 0 || l1_bits > 10 || l2_bits < 0 || l2_bits > 10 || ((uintptr_t)records_dev & 31u)) {
        ga_set_error("ga_sk_scatter_reads: bad arguments (bucket bits must be 0..10 each, records 32-byte aligned)");
        return GA_ERR_BAD_ARG;
    }
    const int w = k - 1;
    if (reads->storage_bits != 2 || reads->sym_bits != 2 || w < 1 || w > 31) {
        ga_set_error("ga_sk_scatter_reads: needs 2-bit reads and 2 <= k <= 32");
        return GA_ERR_BAD_ARG;
    }
    if (reads->estride == 0 || (reads->first_read + reads->n_reads) > ((1ull << 47) / reads->estride)) {
        ga_set_error("ga_sk_scatter_reads: occurrence ordinals exceed 47 bits");
        return GA_ERR_BAD_ARG;
    }
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    const int m = ga_sk_minimizer_len(k);
    // which kernel: windows of 16 m-mers take the lane-per-read kernel when there are enough level-1 buckets for
    // its one-atomic-per-record cursors (with 8 buckets every record of the GPU bumps one of 8 words: C2's scatter
    // went from 0.7 to 1.9 ms); GA_SK_SCATTER=warp|lane|lane128 picks one by hand (A/B runs and the test that
    // compares their records)
    const int variant = [&] {
        const char* e = getenv("GA_SK_SCATTER");
        return !e ? (l1_bits >= 5 ? 1 : 0) : (!strcmp(e, "warp") ? 0 : (!strcmp(e, "lane128") ? 2 : (!strcmp(e, "lane8") ? 3 : 1)));
    }();
#define GA_SK_ARGS \
    rv, w, m, l1_bits, l2_bits, (u64*)records_dev, l1_capacity, (u64*)l1_cursors_dev, (u64*)hist_dev, status_dev
    if (w - m + 1 == 16 && variant != 0) {
        // persistent grid: as many CTAs as fit (registers and the bucket columns in shared memory decide)
        int per_sm = 0;
        if (variant == 1) GA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sk_scatter_reads_lane_kernel<256, 4>, 256, 0));
        else if (variant == 3) GA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sk_scatter_reads_lane_kernel<256, 8>, 256, 0));
        else GA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sk_scatter_reads_lane_kernel<128, 4>, 128, 0));
        const unsigned threads = variant == 2 ? 128 : 256;
        const u64 n_ctas = (rv.n_reads + threads - 1) / threads, most = (u64)ga_sm_count() * (u64)(per_sm > 0 ? per_sm : 1);
        const unsigned grid = (unsigned)(n_ctas < most ? n_ctas : most);
        if (variant == 1) sk_scatter_reads_lane_kernel<256, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(GA_SK_ARGS);
        else if (variant == 3) sk_scatter_reads_lane_kernel<256, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(GA_SK_ARGS);
        else sk_scatter_reads_lane_kernel<128, 4><<<grid, 128, 0, (cudaStream_t)stream>>>(GA_SK_ARGS);
    } else {
        GA_CUDA(cudaFuncSetAttribute(sk_scatter_reads_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(S1Shared)));
        GA_CUDA(cudaFuncSetAttribute(sk_scatter_reads_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(S1Shared)));
        const u64 n_tiles = (rv.n_reads + S1_WARPS - 1) / S1_WARPS;
        const u64 most = (u64)ga_sm_count() * 2;
        const unsigned grid = (unsigned)(n_tiles < most ? n_tiles : most);
        if (w - m + 1 == 16) sk_scatter_reads_kernel<16><<<grid, S1_THREADS, sizeof(S1Shared), (cudaStream_t)stream>>>(GA_SK_ARGS);
        else sk_scatter_reads_kernel<0><<<grid, S1_THREADS, sizeof(S1Shared), (cudaStream_t)stream>>>(GA_SK_ARGS);
    }
#undef GA_SK_ARGS
    GA_LAUNCH_CHECK("sk_scatter_reads");
    return GA_OK;
}

extern "C" int ga_sk_offsets(const uint64_t* hist_dev, uint64_t n_buckets, uint64_t* offsets_dev,
                             uint64_t* cursors_dev, ga_stream stream) {
    if (!hist_dev || !offsets_dev || !cursors_dev || n_buckets == 0) {
        ga_set_error("ga_sk_offsets: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    const u64 n_tiles = (n_buckets + OFF_TILE - 1) / OFF_TILE;
    cudaStream_t st = (cudaStream_t)stream;
    u64* tiles = nullptr;
    ga_pool_retain();
    GA_CUDA(cudaMallocAsync((void**)&tiles, n_tiles * sizeof(u64), st));
    sk_offsets_tile_sums_kernel<<<(unsigned)n_tiles, 256, 0, st>>>((const u64*)hist_dev, n_buckets, tiles);
    sk_offsets_scan_tiles_kernel<<<1, 1024, 0, st>>>(tiles, n_tiles, (u64*)offsets_dev + n_buckets);
    sk_offsets_write_kernel<<<(unsigned)n_tiles, 256, 0, st>>>((const u64*)hist_dev, n_buckets, tiles, (u64*)offsets_dev,
                                                               (u64*)cursors_dev);
    ga_note_launches(2);
    GA_CUDA(cudaFreeAsync(tiles, st));
    GA_LAUNCH_CHECK("sk_offsets");
    return GA_OK;
}

extern "C" int ga_sk_scatter_buckets(const void* records_dev, uint64_t l1_capacity, const uint64_t* l1_cursors_dev,
                                     int l1_bits, int l2_bits, uint64_t* cursors_dev, void* out_bases_dev,
                                     uint64_t* out_meta_dev, uint32_t* out_index_dev, ga_stream stream) {
    const bool dense = out_bases_dev && out_meta_dev && !out_index_dev;
    const bool index = out_index_dev && !out_bases_dev && !out_meta_dev;
    if (!records_dev || !l1_cursors_dev || !cursors_dev || (!dense && !index) || ((uintptr_t)records_dev & 31u) ||
        l1_capacity == 0 || l1_capacity > 0xFFFFFFFFull || l1_bits < 0 || l1_bits > 10 || l2_bits < 0 || l2_bits > 10) {
        ga_set_error("ga_sk_scatter_buckets: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    const u32 n_l1 = 1u << l1_bits;
    const u64 total = (u64)n_l1 * ((l1_capacity + S2_CHUNK - 1) / S2_CHUNK);
    const u64 most = (u64)ga_sm_count() * 8;
    const unsigned grid = (unsigned)(total < most ? total : most);
    if (index && !getenv("GA_SK_INDEX_STAGED")) {
        GA_CUDA(cudaFuncSetAttribute(sk_index_buckets_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(S3Shared)));
        const u64 total3 = (u64)n_l1 * ((l1_capacity + S3_CHUNK - 1) / S3_CHUNK);
        int per_sm3 = 0;
        GA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm3, sk_index_buckets_sorted_kernel, S3_THREADS,
                                                              sizeof(S3Shared)));
        const u64 most3 = (u64)ga_sm_count() * (u64)(per_sm3 > 0 ? per_sm3 : 1);
        const unsigned grid3 = (unsigned)(total3 < most3 ? total3 : most3);
        sk_index_buckets_sorted_kernel<<<grid3, S3_THREADS, sizeof(S3Shared), (cudaStream_t)stream>>>(
            (const u64*)records_dev, l1_capacity, (const u64*)l1_cursors_dev, n_l1, l2_bits, (u64*)cursors_dev,
            out_index_dev);
    } else if (index)
        sk_scatter_buckets_kernel<true><<<grid, S2_THREADS, 0, (cudaStream_t)stream>>>(
            (const u64*)records_dev, l1_capacity, (const u64*)l1_cursors_dev, n_l1, l2_bits, (u64*)cursors_dev, nullptr,
            nullptr, out_index_dev);
    else
        sk_scatter_buckets_kernel<false><<<grid, S2_THREADS, 0, (cudaStream_t)stream>>>(
            (const u64*)records_dev, l1_capacity, (const u64*)l1_cursors_dev, n_l1, l2_bits, (u64*)cursors_dev,
            (ulonglong2*)out_bases_dev, (u64*)out_meta_dev, nullptr);
    GA_LAUNCH_CHECK("sk_scatter_buckets");
    return GA_OK;
}

// host description of the sources form -> kernel parameter; false: bad description
static bool sk_sources(const ga_sk_sources* from, uint32_t n_segments, SkSources& src) {
    memset(&src, 0, sizeof(src));
    if (!from) return true;
    if (from->n_sources == 0 || from->n_sources > SB_MAX_SEG || from->n_sources != n_segments) return false;
    for (uint32_t g = 0; g < from->n_sources; ++g) {
        if (!from->records[g] || !from->index[g] || ((uintptr_t)from->records[g] & 31u) || from->l1_capacity[g] == 0 ||
            from->l1_capacity[g] >= (1ull << 25))
            return false;
        src.rec[g] = (const u64*)from->records[g];
        src.index[g] = from->index[g];
        src.cap1[g] = from->l1_capacity[g];
    }
    src.first_bucket = from->first_bucket;
    src.counter = (u64*)from->solid_counter;
    src.n = from->n_sources;
    return true;
}

static int sk_count_build(const void* bases_dev, const uint64_t* meta_dev, const uint64_t* offsets_dev,
                          uint32_t n_segments, const uint64_t* hist_dev, uint64_t n_buckets, int k,
                          int64_t threshold,
                          uint32_t table_slots, uint32_t max_solid, uint64_t* solid_keys_out_dev,
                          uint64_t* edge_stamp_out_dev, uint64_t out_capacity, uint64_t* counters_dev,
                          uint64_t* spill_list_dev, uint64_t spill_capacity, uint32_t* status_dev,
                          const uint32_t* index_dev, uint64_t l1_capacity, int l2_bits, const ga_sk_sources* sources,
                          ga_stream stream) {
    const int w = k - 1;
    SkSources src;
    if (!sk_sources(sources, n_segments, src)) {
        ga_set_error("ga_sk_count_build_from: bad sources (1..%u, one per segment, 32-byte aligned slots, capacities below 2^25)",
                     SB_MAX_SEG);
        return GA_ERR_BAD_ARG;
    }
    if (sources) bases_dev = sources->records[0];        // the classic arguments are not used in the sources form
    if ((!bases_dev) || (!meta_dev && !index_dev && !sources) || (index_dev && ((uintptr_t)bases_dev & 31u)) || !offsets_dev ||
        !hist_dev || !solid_keys_out_dev || !counters_dev || !spill_list_dev || !status_dev || n_buckets == 0 || w < 1 || w > 31 || threshold < 0 ||
        threshold > 60000 || !is_pow2(table_slots) || table_slots < 256 || table_slots > SB_MAX_SLOTS ||
        max_solid == 0 || n_segments == 0 || n_segments > SB_MAX_SEG) {
        ga_set_error("ga_sk_count_build: bad arguments (0 <= threshold <= 60000, table_slots a power of two in "
                     "256..%u, max_solid >= 1)", SB_MAX_SLOTS);
        return GA_ERR_BAD_ARG;
    }
    // the attribute is per device: set it on every call (cheap) instead of caching a process-wide flag
    const bool deep = sources && getenv("GA_SK_DEEP");           // opt-in: loads two spans ahead of the walk (measured: +-0)
    GA_CUDA(cudaFuncSetAttribute(sk_bucket_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_POOL_BYTES));
    GA_CUDA(cudaFuncSetAttribute(sk_bucket_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_POOL_BYTES));
    GA_CUDA(cudaFuncSetAttribute(sk_bucket_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_POOL_BYTES));
    if (max_solid > 16000) max_solid = 16000;
    const u64 most = (u64)ga_sm_count() * SB_CTAS_PER_SM;
    const unsigned grid = (unsigned)(n_buckets < most ? n_buckets : most);
    // notes that do not fit a CTA's queue in shared memory: SB_NOTE_SPILL two-word notes per CTA, stream-ordered
    u64* note_spill = nullptr;
    ga_pool_retain();
    GA_CUDA(cudaMallocAsync((void**)&note_spill, (size_t)grid * SB_NOTE_SPILL * 16u, (cudaStream_t)stream));
#define GA_SB_LAUNCH(DEEP, SRC)                                                                                     \
    sk_bucket_kernel<DEEP, SRC><<<grid, SB_THREADS, SB_POOL_BYTES, (cudaStream_t)stream>>>(                          \
        (const ulonglong2*)bases_dev, (const u64*)meta_dev, (const u64*)offsets_dev, n_segments,                    \
        (const u64*)hist_dev, n_buckets, w,                                                                         \
        (u32)threshold, table_slots, max_solid, (u64*)solid_keys_out_dev, (u64*)edge_stamp_out_dev, out_capacity,   \
        (u64*)counters_dev, (u64*)spill_list_dev, spill_capacity, status_dev, sources ? nullptr : index_dev,        \
        l1_capacity, l2_bits, note_spill, SB_NOTE_SPILL, src)
    if (deep) GA_SB_LAUNCH(true, true);
    else if (sources) GA_SB_LAUNCH(false, true);
    else GA_SB_LAUNCH(false, false);
#undef GA_SB_LAUNCH
    ga_note_launches(1);
    const cudaError_t launched = cudaGetLastError();
    GA_CUDA(cudaFreeAsync(note_spill, (cudaStream_t)stream));
    GA_CUDA(launched);
    return GA_OK;
}

extern "C" int ga_sk_count_build(const void* bases_dev, const uint64_t* meta_dev, const uint64_t* offsets_dev,
                                 uint32_t n_segments, const uint64_t* hist_dev, uint64_t n_buckets, int k,
                                 int64_t threshold,
                                 uint32_t table_slots, uint32_t max_solid, uint64_t* solid_keys_out_dev,
                                 uint64_t* edge_stamp_out_dev, uint64_t out_capacity, uint64_t* counters_dev,
                                 uint64_t* spill_list_dev, uint64_t spill_capacity, uint32_t* status_dev,
                                 const uint32_t* index_dev, uint64_t l1_capacity, int l2_bits, ga_stream stream) {
    return sk_count_build(bases_dev, meta_dev, offsets_dev, n_segments, hist_dev, n_buckets, k, threshold, table_slots,
                          max_solid, solid_keys_out_dev, edge_stamp_out_dev, out_capacity, counters_dev, spill_list_dev,
                          spill_capacity, status_dev, index_dev, l1_capacity, l2_bits, nullptr, stream);
}

extern "C" int ga_sk_count_build_from(const ga_sk_sources* sources, const uint64_t* offsets_dev, const uint64_t* hist_dev,
                                      uint64_t n_buckets, int k, int64_t threshold, uint32_t table_slots,
                                      uint32_t max_solid, uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev,
                                      uint64_t out_capacity, uint64_t* counters_dev, uint64_t* spill_list_dev,
                                      uint64_t spill_capacity, uint32_t* status_dev, int l2_bits, ga_stream stream) {
    if (!sources) {
        ga_set_error("ga_sk_count_build_from: no sources");
        return GA_ERR_BAD_ARG;
    }
    return sk_count_build(nullptr, nullptr, offsets_dev, sources->n_sources, hist_dev, n_buckets, k, threshold,
                          table_slots, max_solid, solid_keys_out_dev, edge_stamp_out_dev, out_capacity, counters_dev,
                          spill_list_dev, spill_capacity, status_dev, nullptr, 0, l2_bits, sources, stream);
}

extern "C" uint64_t ga_sk_spill_scratch_bytes(uint32_t table_slots) {
    // per slot: key 8 B + state 4 B, a solid key 8 B + 4 stamps 32 B, two 16-byte notes
    return (uint64_t)table_slots * (8 + 4 + 8 + 32 + 32);
}

static int sk_count_build_spill(const void* bases_dev, const uint64_t* meta_dev, const uint64_t* offsets_dev,
                                uint32_t n_segments, uint64_t n_buckets, const uint64_t* spill_list_dev, uint64_t n_spill, int k, int64_t threshold,
                                uint32_t table_slots, void* scratch_dev, uint32_t n_ctas,
                                uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev,
                                uint64_t out_capacity, uint64_t* counters_dev, uint32_t* status_dev,
                                const uint32_t* index_dev, uint64_t l1_capacity, int l2_bits, const ga_sk_sources* sources,
                                ga_stream stream) {
    const int w = k - 1;
    SkSources src;
    if (!sk_sources(sources, n_segments, src)) {
        ga_set_error("ga_sk_count_build_spill_from: bad sources");
        return GA_ERR_BAD_ARG;
    }
    if (sources) bases_dev = sources->records[0];
    if (!bases_dev || (!meta_dev && !index_dev && !sources) || !offsets_dev || !spill_list_dev || !scratch_dev || !solid_keys_out_dev ||
        !counters_dev || !status_dev || w < 1 || w > 31 || threshold < 0 ||
        !is_pow2(table_slots) || table_slots < 256 || n_ctas == 0 || n_segments == 0 || n_segments > SB_MAX_SEG ||
        n_buckets == 0) {
        ga_set_error("ga_sk_count_build_spill: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    if (n_spill == 0) return GA_OK;
    auto* spill = sources ? sk_bucket_spill_kernel<true> : sk_bucket_spill_kernel<false>;
    spill<<<n_ctas, SB_THREADS, 0, (cudaStream_t)stream>>>(
        (const ulonglong2*)bases_dev, (const u64*)meta_dev, (const u64*)offsets_dev, n_segments, n_buckets,
        (const u64*)spill_list_dev, n_spill, w,
        (u32)(threshold > 0xFFFFFFF0ll ? 0xFFFFFFF0ll : threshold), table_slots, (unsigned char*)scratch_dev,
        ga_sk_spill_scratch_bytes(table_slots), (u64*)solid_keys_out_dev, (u64*)edge_stamp_out_dev, out_capacity,
        (u64*)counters_dev, status_dev, sources ? nullptr : index_dev, l1_capacity, l2_bits, src);
    GA_LAUNCH_CHECK("sk_bucket_spill");
    return GA_OK;
}

extern "C" int ga_sk_count_build_spill(const void* bases_dev, const uint64_t* meta_dev, const uint64_t* offsets_dev,
                                       uint32_t n_segments, uint64_t n_buckets, const uint64_t* spill_list_dev, uint64_t n_spill, int k, int64_t threshold,
                                       uint32_t table_slots, void* scratch_dev, uint32_t n_ctas,
                                       uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev,
                                       uint64_t out_capacity, uint64_t* counters_dev, uint32_t* status_dev,
                                       const uint32_t* index_dev, uint64_t l1_capacity, int l2_bits,
                                       ga_stream stream) {
    return sk_count_build_spill(bases_dev, meta_dev, offsets_dev, n_segments, n_buckets, spill_list_dev, n_spill, k,
                                threshold, table_slots, scratch_dev, n_ctas, solid_keys_out_dev, edge_stamp_out_dev,
                                out_capacity, counters_dev, status_dev, index_dev, l1_capacity, l2_bits, nullptr, stream);
}

extern "C" int ga_sk_count_build_spill_from(const ga_sk_sources* sources, const uint64_t* offsets_dev, uint64_t n_buckets,
                                            const uint64_t* spill_list_dev, uint64_t n_spill, int k, int64_t threshold,
                                            uint32_t table_slots, void* scratch_dev, uint32_t n_ctas,
                                            uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev,
                                            uint64_t out_capacity, uint64_t* counters_dev, uint32_t* status_dev,
                                            int l2_bits, ga_stream stream) {
    if (!sources) {
        ga_set_error("ga_sk_count_build_spill_from: no sources");
        return GA_ERR_BAD_ARG;
    }
    return sk_count_build_spill(nullptr, nullptr, offsets_dev, sources->n_sources, n_buckets, spill_list_dev, n_spill, k,
                                threshold, table_slots, scratch_dev, n_ctas, solid_keys_out_dev, edge_stamp_out_dev,
                                out_capacity, counters_dev, status_dev, nullptr, 0, l2_bits, sources, stream);
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
