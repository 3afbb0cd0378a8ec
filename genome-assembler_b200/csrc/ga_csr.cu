// ga_csr.cu -- turn the stamp tables of ga_build.cu into the ordered CSR the host traversal
// consumes: nodes in `self.nodes` iteration order, each row in edge
// insertion order, in-degrees, was_branching (debruijn_graph.py:139-142, 313-317), last
// symbol and packed node keys.  The paired variant first resolves the fuzzy second key
// (_find_matching_node, debruijn_graph.py:319-347) per mate-1 group, in stamp order.
#include <cub/cub.cuh>

#include <vector>

#include "ga_common.cuh"

struct ga_csr_plan {
    int paired = 0, key_words = 1, sym_bits = 2;
    long long n_nodes = 0, n_edges = 0;
    const void* solid_keys = nullptr;
    u32* node_a = nullptr;        // [n_nodes] solid id of each node's (mate-1) key, node order
    u32* node_b = nullptr;        // [n_nodes] solid id of the mate-2 key (paired)
    const StampSlot* etable = nullptr;  // final edge table: key = src << 32 | dst
    u64 ecap = 0;
    u32* rank = nullptr;          // unpaired: solid id -> node index (edge keys hold solid ids)
    int* extra_in = nullptr;      // paired: in-edges held by orphaned prefix nodes
    // "dna4" mode (unpaired, <= 4 symbols): edges live in edge_stamp4[4 * solid id + symbol]
    const u64* edge_stamp4 = nullptr;
    const void* solid = nullptr;  // id table, to turn a successor key into its id
    u64 solid_cap = 0;
    int w = 0;
    cudaStream_t stream = nullptr;
    std::vector<void*> owned;
};

namespace {

struct Scratch {
    cudaStream_t st;
    std::vector<void*>* keep;   // allocations that outlive the call (owned by the plan)
    std::vector<void*> temp;    // freed when the call ends
    cudaError_t err = cudaSuccess;
    template <class T> T* get(u64 n, bool persistent = false) {
        ga_pool_retain();
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, (n ? n : 1) * sizeof(T), st);
        if (e != cudaSuccess) {
            err = e;
            return nullptr;
        }
        (persistent ? *keep : temp).push_back(p);
        return (T*)p;
    }
    void release() {
        for (void* p : temp) cudaFreeAsync(p, st);
        temp.clear();
    }
};


// (stamp, id) of every node that was touched by an accepted occurrence
// counter[0] = nodes taken; counter[2] = largest stamp taken (the radix sort then skips the all-zero high digits)
__global__ void compact_nodes_kernel(const u64* __restrict__ node_stamp, u64 n, u64* __restrict__ stamps,
                                     u32* __restrict__ ids, u64* counter) {
    u64 step = (u64)gridDim.x * blockDim.x;
    u64 rounds = (n + step - 1) / step;
    u64 top = 0;
    for (u64 it = 0; it < rounds; ++it) {
        u64 i = it * step + blockIdx.x * (u64)blockDim.x + threadIdx.x;
        u64 s = i < n ? node_stamp[i] : GA_NONE64;
        bool take = s != GA_NONE64;
        u64 pos = ga_warp_append(counter, take);
        if (take) {
            stamps[pos] = s;
            ids[pos] = (u32)i;
            top = max(top, s);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const u32 lo = __shfl_down_sync(0xFFFFFFFFu, (u32)top, off), hi = __shfl_down_sync(0xFFFFFFFFu, (u32)(top >> 32), off);
        top = max(top, ((u64)hi << 32) | lo);
    }
    if ((threadIdx.x & 31) == 0 && top) atomicMax((unsigned long long*)(counter + 2), (unsigned long long)top);
}

// bits a radix sort has to look at for keys <= top (whole digits of 8 bits)
int key_bits(u64 top) {
    int bits = 8;
    while (bits < 64 && (top >> bits)) bits += 8;
    return bits;
}

__global__ void count_slots_kernel(const StampSlot* __restrict__ t, u64 cap, u64* counter) {
    u64 n = 0;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < cap; i += (u64)gridDim.x * blockDim.x)
        n += t[i].key != GA_NONE64;
    for (int off = 16; off > 0; off >>= 1) n += __shfl_down_sync(0xFFFFFFFFu, n, off);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(counter, n);
}

__global__ void scatter_rank_kernel(const u32* __restrict__ ids_sorted, u64 n, u32* __restrict__ rank) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        rank[ids_sorted[i]] = (u32)i;
}

// ---- CSR emission from a final edge table ----------------------------------------------------
__global__ void degree_kernel(const StampSlot* __restrict__ t, u64 cap, const u32* __restrict__ rank,
                              int* __restrict__ outdeg, int* __restrict__ indeg) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < cap; i += (u64)gridDim.x * blockDim.x) {
        u64 key = t[i].key;
        if (key == GA_NONE64) continue;
        u32 s = (u32)(key >> 32), d = (u32)key;
        if (rank) {
            s = rank[s];
            d = rank[d];
        }
        atomicAdd(outdeg + s, 1);
        atomicAdd(indeg + d, 1);
    }
}

__global__ void place_edges_kernel(const StampSlot* __restrict__ t, u64 cap, const u32* __restrict__ rank,
                                   const int* __restrict__ rowptr, int* __restrict__ cursor,
                                   int* __restrict__ col, u64* __restrict__ estamp) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < cap; i += (u64)gridDim.x * blockDim.x) {
        u64 key = t[i].key;
        if (key == GA_NONE64) continue;
        u32 s = (u32)(key >> 32), d = (u32)key;
        if (rank) {
            s = rank[s];
            d = rank[d];
        }
        int pos = rowptr[s] + atomicAdd(cursor + s, 1);
        col[pos] = (int)d;
        estamp[pos] = t[i].stamp;
    }
}

template <class K>
__global__ void finish_nodes_kernel(long long n_nodes, const int* __restrict__ rowptr, int* __restrict__ col,
                                    u64* __restrict__ estamp, const int* __restrict__ indeg,
                                    const K* __restrict__ solid_keys, const u32* __restrict__ node_a,
                                    const u32* __restrict__ node_b, u32 smask, u8* __restrict__ branching,
                                    u8* __restrict__ last_sym, K* __restrict__ keys_a, K* __restrict__ keys_b) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_nodes;
         i += (long long)gridDim.x * blockDim.x) {
        int beg = rowptr[i], end = rowptr[i + 1];
        for (int a = beg + 1; a < end; ++a) {   // rows are tiny: insertion sort by edge stamp
            u64 st = estamp[a];
            int c = col[a];
            int b = a - 1;
            while (b >= beg && estamp[b] > st) {
                estamp[b + 1] = estamp[b];
                col[b + 1] = col[b];
                --b;
            }
            estamp[b + 1] = st;
            col[b + 1] = c;
        }
        branching[i] = (u8)((end - beg > 1) || indeg[i] > 1);
        K ka = solid_keys[node_a[i]];
        last_sym[i] = (u8)((u32)ka & smask);
        if (keys_a) keys_a[i] = ka;
        if (keys_b && node_b) keys_b[i] = solid_keys[node_b[i]];
    }
}

// ---- dna4 mode: rows straight from the per-node edge stamps ------------------------------------
__global__ void count_stamps_kernel(const u64* __restrict__ stamps, u64 n, u64* counter) {
    u64 c = 0;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        c += stamps[i] != GA_NONE64;
    for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(counter, c);
}

__global__ void dna4_outdeg_kernel(const u64* __restrict__ edge_stamp4, const u32* __restrict__ node_a,
                                   u64 n_nodes, int* __restrict__ outdeg) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n_nodes; i += (u64)gridDim.x * blockDim.x) {
        const u64* st = edge_stamp4 + 4ull * node_a[i];
        outdeg[i] = (st[0] != GA_NONE64) + (st[1] != GA_NONE64) + (st[2] != GA_NONE64) + (st[3] != GA_NONE64);
    }
}

template <class K>
__global__ void dna4_rows_kernel(const u64* __restrict__ edge_stamp4, const u32* __restrict__ node_a,
                                 u64 n_nodes, const K* __restrict__ solid_keys, const Slot<K>* __restrict__ solid,
                                 u64 solid_cap, const u32* __restrict__ rank, int w, int sym_bits,
                                 const int* __restrict__ rowptr, int* __restrict__ col, int* __restrict__ indeg) {
    const K mask = ga_key_mask<K>(w, sym_bits);
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n_nodes; i += (u64)gridDim.x * blockDim.x) {
        const u32 id = node_a[i];
        const K key = solid_keys[id];
        u64 st[4];
        u32 sy[4];
        int n = 0;
        for (u32 s = 0; s < 4u; ++s) {
            u64 v = edge_stamp4[4ull * id + s];
            if (v == GA_NONE64) continue;
            int j = n++;
            while (j > 0 && st[j - 1] > v) {   // insertion sort by first occurrence
                st[j] = st[j - 1];
                sy[j] = sy[j - 1];
                --j;
            }
            st[j] = v;
            sy[j] = s;
        }
        int base = rowptr[i];
        for (int j = 0; j < n; ++j) {
            K succ = ((key << sym_bits) | (K)sy[j]) & mask;
            u32 succ_rank = rank[ga_table_find(solid, solid_cap, succ)];
            col[base + j] = (int)succ_rank;
            atomicAdd(indeg + succ_rank, 1);
        }
    }
}

template <class K>
__global__ void dna4_finish_kernel(u64 n_nodes, const int* __restrict__ rowptr, const int* __restrict__ indeg,
                                   const K* __restrict__ solid_keys, const u32* __restrict__ node_a, u32 smask,
                                   u8* __restrict__ branching, u8* __restrict__ last_sym, K* __restrict__ keys_a) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n_nodes; i += (u64)gridDim.x * blockDim.x) {
        branching[i] = (u8)((rowptr[i + 1] - rowptr[i] > 1) || indeg[i] > 1);
        K ka = solid_keys[node_a[i]];
        last_sym[i] = (u8)((u32)ka & smask);
        if (keys_a) keys_a[i] = ka;
    }
}

// ---- paired: query compaction, per-group greedy, edge remap ------------------------------------
__global__ void compact_queries_kernel(const StampSlot* __restrict__ q, u64 cap, u64* __restrict__ stamps,
                                       u32* __restrict__ slots, u64* counter) {
    u64 step = (u64)gridDim.x * blockDim.x;
    u64 rounds = (cap + step - 1) / step;
    for (u64 it = 0; it < rounds; ++it) {
        u64 i = it * step + blockIdx.x * (u64)blockDim.x + threadIdx.x;
        bool take = i < cap && q[i].key != GA_NONE64;
        u64 pos = ga_warp_append(counter, take);
        if (take) {
            stamps[pos] = q[i].stamp;
            slots[pos] = (u32)i;
        }
    }
}

__global__ void gather_a_kernel(const StampSlot* __restrict__ q, const u32* __restrict__ slots_sorted, u64 n,
                                u32* __restrict__ a_out) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        a_out[i] = (u32)(q[slots_sorted[i]].key >> 32);
}

// suffix of x (>= 3 symbols) equals prefix of y (debruijn_graph.py:336-347 with text=x, pattern=y)
template <class K> __device__ __forceinline__ bool suffix_prefix(K x, K y, int w, int b) {
    for (int s = 0; s + 3 <= w; ++s) {
        int keep = (w - s) * b;
        K xm = x & ((((K)1) << keep) - 1);   // keep < bits(K): w*b <= 63 / 127
        if (xm == (y >> (s * b))) return true;
    }
    return false;
}

// One thread per group head walks its group (members sorted by stamp) and chooses, for every
// queried (A,B), the first earlier key whose B overlaps (SURVEY App. C.2 + A-9 i).
template <class K>
__global__ void resolve_groups_kernel(const StampSlot* __restrict__ q, const u32* __restrict__ slots, u64 n,
                                      const K* __restrict__ solid_keys, int w, int b, u32* __restrict__ rep,
                                      int* __restrict__ iskey, u64* __restrict__ gstamp) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u32 a = (u32)(q[slots[i]].key >> 32);
        if (i > 0 && (u32)(q[slots[i - 1]].key >> 32) == a) continue;   // not a group head
        u64 first_stamp = q[slots[i]].stamp;
        for (u64 m = i; m < n; ++m) {
            StampSlot qm = q[slots[m]];
            if ((u32)(qm.key >> 32) != a) break;
            K bm = solid_keys[(u32)qm.key];
            u32 found = GA_NONE32;
            for (u64 j = i; j < m; ++j) {
                if (!iskey[j]) continue;
                StampSlot qj = q[slots[j]];
                if ((qm.stamp & 1ull) && qj.stamp == qm.stamp - 1) continue;  // same occurrence: unseen
                K bj = solid_keys[(u32)qj.key];
                if (suffix_prefix<K>(bm, bj, w, b) || suffix_prefix<K>(bj, bm, w, b)) {
                    found = (u32)j;
                    break;
                }
            }
            iskey[m] = found == GA_NONE32;
            rep[m] = found == GA_NONE32 ? (u32)m : found;
            gstamp[m] = first_stamp;
        }
    }
}

__global__ void compact_keys_kernel(const int* __restrict__ iskey, const int* __restrict__ kpos,
                                    const u64* __restrict__ gstamp, u64 n, u64* __restrict__ kg,
                                    u32* __restrict__ kidx) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        if (iskey[i]) {
            kg[kpos[i]] = gstamp[i];
            kidx[kpos[i]] = (u32)i;
        }
}

// node order known: fill node -> (A id, B id) and sorted-member -> node rank for keys
__global__ void place_nodes_kernel(const StampSlot* __restrict__ q, const u32* __restrict__ slots,
                                   const u32* __restrict__ node_member, u64 n_nodes, u32* __restrict__ node_a,
                                   u32* __restrict__ node_b, u32* __restrict__ rank_of_member) {
    for (u64 r = blockIdx.x * (u64)blockDim.x + threadIdx.x; r < n_nodes; r += (u64)gridDim.x * blockDim.x) {
        u32 m = node_member[r];
        u64 key = q[slots[m]].key;
        node_a[r] = (u32)(key >> 32);
        node_b[r] = (u32)key;
        rank_of_member[m] = (u32)r;
    }
}

__global__ void map_slots_kernel(const u32* __restrict__ slots, const u32* __restrict__ rep,
                                 const u32* __restrict__ rank_of_member, u64 n, u32* __restrict__ node_of_slot) {
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        node_of_slot[slots[i]] = rank_of_member[rep[i]];
}

template <class K> __device__ __forceinline__ K make_key(u64 lo, u64 hi);
template <> __device__ __forceinline__ u64 make_key<u64>(u64 lo, u64) { return lo; }
template <> __device__ __forceinline__ u128 make_key<u128>(u64 lo, u64 hi) { return ((u128)hi << 64) | lo; }

// Orphaned self-loops (SURVEY App. A-9 ii): the first "prefix pair == suffix pair" occurrence
// also created the node, so the reference hangs that edge off an unreachable duplicate.
template <class K>
__global__ void find_orphans_kernel(const u64* __restrict__ dh, const Slot<K>* __restrict__ solid, u64 solid_cap,
                                    u64 unit_lo, u64 unit_hi, const StampSlot* __restrict__ q, u64 qcap,
                                    const u32* __restrict__ node_of_slot, const u32* __restrict__ node_a,
                                    const u32* __restrict__ node_b, u32* __restrict__ orphan_slot,
                                    u64* __restrict__ orphan_second, u64* counters, int* __restrict__ extra_in) {
    const K unit = make_key<K>(unit_lo, unit_hi);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < 65536u; i += gridDim.x * blockDim.x) {
        u64 first = dh[2 * i];
        if (first == GA_NONE64) continue;
        u32 ida = ga_table_find(solid, solid_cap, (K)(i >> 8) * unit);
        u32 idb = ga_table_find(solid, solid_cap, (K)(i & 255u) * unit);
        if (ida == GA_NONE32 || idb == GA_NONE32) continue;
        u64 s = ga_stamp_find(q, qcap, ((u64)ida << 32) | idb);
        if (s == GA_NONE64) continue;
        u32 node = node_of_slot[s];
        if (node_a[node] != ida || node_b[node] != idb) continue;   // merged into another key
        if (q[s].stamp != 2 * first) continue;                        // the node existed before
        u64 pos = atomicAdd(counters + 0, 1ull);
        orphan_slot[pos] = (u32)s;
        orphan_second[pos] = dh[2 * i + 1];
        atomicAdd(extra_in + node, 1);
    }
}

__global__ void remap_edges_kernel(const StampSlot* __restrict__ qe, u64 qecap, const u32* __restrict__ node_of_slot,
                                   const u32* __restrict__ orphan_slot, const u64* __restrict__ orphan_second,
                                   const u64* __restrict__ counters, StampSlot* __restrict__ etable, u64 ecap,
                                   u32* status) {
    const u64 n_orphans = counters[0];
    bool full = false;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < qecap; i += (u64)gridDim.x * blockDim.x) {
        u64 key = qe[i].key;
        if (key == GA_NONE64) continue;
        u32 qp = (u32)(key >> 32), qs = (u32)key;
        u64 stamp = qe[i].stamp;
        if (qp == qs)
            for (u64 o = 0; o < n_orphans; ++o)
                if (orphan_slot[o] == qp) stamp = orphan_second[o];
        if (stamp == GA_NONE64) continue;   // the only occurrence belonged to the orphan
        u64 ek = ((u64)node_of_slot[qp] << 32) | node_of_slot[qs];
        if (ga_stamp_upsert(etable, ecap, ek, stamp) == GA_NONE64) full = true;
    }
    if (full) atomicOr(status, GA_ST_STAMP_FULL);
}

template <class KeyT, class ValT>
cudaError_t sort_pairs(Scratch& sc, const KeyT* kin, KeyT* kout, const ValT* vin, ValT* vout, u64 n,
                       int end_bit = (int)(8 * sizeof(KeyT))) {
    if (n == 0) return cudaSuccess;
    size_t bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, (long long)n, 0, end_bit, sc.st);
    if (e != cudaSuccess) return e;
    void* tmp = sc.get<u8>(bytes);
    if (!tmp) return sc.err;
    return cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, (long long)n, 0, end_bit, sc.st);
}

cudaError_t exclusive_sum(Scratch& sc, const int* in, int* out, u64 n) {
    if (n == 0) return cudaSuccess;
    size_t bytes = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (long long)n, sc.st);
    if (e != cudaSuccess) return e;
    void* tmp = sc.get<u8>(bytes);
    if (!tmp) return sc.err;
    return cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (long long)n, sc.st);
}

unsigned scan_grid(u64 n) {
    unsigned g = ga_grid(n, 256);
    return g > 148u * 16u ? 148u * 16u : g;
}

#define GA_TRY(expr)                                   \
    do {                                               \
        cudaError_t _e = (expr);                       \
        if (_e != cudaSuccess) {                       \
            sc.release();                              \
            ga_csr_plan_free(plan);                    \
            return ga_cuda_fail(_e, #expr);            \
        }                                              \
    } while (0)
#define GA_NEED(ptr)                                   \
    do {                                               \
        if (!(ptr)) {                                  \
            sc.release();                              \
            ga_csr_plan_free(plan);                    \
            return ga_cuda_fail(sc.err, "cudaMallocAsync"); \
        }                                              \
    } while (0)

template <class K>
int plan_paired_impl(ga_csr_plan* plan, Scratch& sc, const void* solid_dev, u64 solid_cap, int w, int b,
                     const StampSlot* q, u64 qcap, const StampSlot* qe, u64 qecap, const u64* dh,
                     int64_t* n_nodes_out, int64_t* n_edges_out, int64_t* attr_out) {
    cudaStream_t st = sc.st;
    const K* solid_keys = (const K*)plan->solid_keys;
    // counters: [0] queries, [1] query edges, [2] orphans, [3] final edges
    u64* counters = sc.get<u64>(8);
    u32* status = sc.get<u32>(1);
    GA_NEED(counters);
    GA_NEED(status);
    GA_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(u64), st));
    GA_TRY(cudaMemsetAsync(status, 0, sizeof(u32), st));
    count_slots_kernel<<<scan_grid(qcap), 256, 0, st>>>(q, qcap, counters + 0);
    ga_note_launches(1);
    count_slots_kernel<<<scan_grid(qecap), 256, 0, st>>>(qe, qecap, counters + 1);
    ga_note_launches(1);
    u64 host[8];
    GA_TRY(cudaMemcpyAsync(host, counters, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    GA_TRY(cudaStreamSynchronize(st));
    const u64 nq = host[0], nqe = host[1];

    // 1. queries sorted by (A, stamp): sort by stamp, then stable sort by A
    u64* st0 = sc.get<u64>(nq);
    u64* st1 = sc.get<u64>(nq);
    u32* sl0 = sc.get<u32>(nq);
    u32* sl1 = sc.get<u32>(nq);
    u32* a0 = sc.get<u32>(nq);
    u32* a1 = sc.get<u32>(nq);
    u32* slots = sc.get<u32>(nq);
    GA_NEED(st0); GA_NEED(st1); GA_NEED(sl0); GA_NEED(sl1); GA_NEED(a0); GA_NEED(a1); GA_NEED(slots);
    GA_TRY(cudaMemsetAsync(counters + 4, 0, sizeof(u64), st));
    compact_queries_kernel<<<scan_grid(qcap), 256, 0, st>>>(q, qcap, st0, sl0, counters + 4);
    ga_note_launches(1);
    GA_TRY((sort_pairs<u64, u32>(sc, st0, st1, sl0, sl1, nq)));
    if (nq) gather_a_kernel<<<scan_grid(nq), 256, 0, st>>>(q, sl1, nq, a0);
    ga_note_launches(1);
    GA_TRY((sort_pairs<u32, u32>(sc, a0, a1, sl1, slots, nq)));

    // 2. per-group greedy choice of representative keys
    u32* rep = sc.get<u32>(nq);
    int* iskey = sc.get<int>(nq + 1);
    int* kpos = sc.get<int>(nq + 1);
    u64* gstamp = sc.get<u64>(nq);
    GA_NEED(rep); GA_NEED(iskey); GA_NEED(kpos); GA_NEED(gstamp);
    GA_TRY(cudaMemsetAsync(iskey, 0, (nq + 1) * sizeof(int), st));
    if (nq) resolve_groups_kernel<K><<<ga_grid(nq, 128), 128, 0, st>>>(q, slots, nq, solid_keys, w, b, rep, iskey, gstamp);
    ga_note_launches(1);
    GA_TRY(exclusive_sum(sc, iskey, kpos, nq + 1));
    int n_keys = 0;
    GA_TRY(cudaMemcpyAsync(&n_keys, kpos + nq, sizeof(int), cudaMemcpyDeviceToHost, st));
    GA_TRY(cudaStreamSynchronize(st));
    const u64 n_nodes = (u64)n_keys;

    // 3. node order: groups by the stamp of their first key, keys by stamp inside a group
    u64* kg0 = sc.get<u64>(n_nodes);
    u64* kg1 = sc.get<u64>(n_nodes);
    u32* ki0 = sc.get<u32>(n_nodes);
    u32* node_member = sc.get<u32>(n_nodes);
    u32* rank_of_member = sc.get<u32>(nq);
    u32* node_of_slot = sc.get<u32>(qcap);
    plan->node_a = sc.get<u32>(n_nodes, true);
    plan->node_b = sc.get<u32>(n_nodes, true);
    plan->extra_in = sc.get<int>(n_nodes, true);
    GA_NEED(kg0); GA_NEED(kg1); GA_NEED(ki0); GA_NEED(node_member); GA_NEED(rank_of_member);
    GA_NEED(node_of_slot); GA_NEED(plan->node_a); GA_NEED(plan->node_b); GA_NEED(plan->extra_in);
    GA_TRY(cudaMemsetAsync(plan->extra_in, 0, (n_nodes ? n_nodes : 1) * sizeof(int), st));
    if (nq) compact_keys_kernel<<<scan_grid(nq), 256, 0, st>>>(iskey, kpos, gstamp, nq, kg0, ki0);
    ga_note_launches(1);
    GA_TRY((sort_pairs<u64, u32>(sc, kg0, kg1, ki0, node_member, n_nodes)));
    if (n_nodes) {
        place_nodes_kernel<<<scan_grid(n_nodes), 256, 0, st>>>(q, slots, node_member, n_nodes, plan->node_a,
                                                              plan->node_b, rank_of_member);
    ga_note_launches(1);
        map_slots_kernel<<<scan_grid(nq), 256, 0, st>>>(slots, rep, rank_of_member, nq, node_of_slot);
    ga_note_launches(1);
    }

    // 4. orphaned self-loops, then distinct (rep(P), rep(S)) edges with their first occurrence
    u32* orphan_slot = sc.get<u32>(65536);
    u64* orphan_second = sc.get<u64>(65536);
    GA_NEED(orphan_slot); GA_NEED(orphan_second);
    K unit = 0;
    for (int i = 0; i < w; ++i) unit |= (K)1 << (i * b);
    u64 unit_lo = (u64)unit, unit_hi = sizeof(K) > 8 ? (u64)((u128)unit >> 64) : 0;
    if (n_nodes)
        find_orphans_kernel<K><<<64, 256, 0, st>>>(dh, (const Slot<K>*)solid_dev, solid_cap, unit_lo, unit_hi, q, qcap,
                                                   node_of_slot, plan->node_a, plan->node_b, orphan_slot,
                                                   orphan_second, counters + 2, plan->extra_in);
    ga_note_launches(1);
    u64 ecap = 2 * nqe + 64;
    StampSlot* etable = sc.get<StampSlot>(ecap, true);
    GA_NEED(etable);
    GA_TRY(cudaMemsetAsync(etable, 0xFF, ecap * sizeof(StampSlot), st));
    if (nqe)
        remap_edges_kernel<<<scan_grid(qecap), 256, 0, st>>>(qe, qecap, node_of_slot, orphan_slot, orphan_second,
                                                            counters + 2, etable, ecap, status);
    ga_note_launches(1);
    count_slots_kernel<<<scan_grid(ecap), 256, 0, st>>>(etable, ecap, counters + 3);
    ga_note_launches(1);
    GA_TRY(cudaGetLastError());
    GA_TRY(cudaMemcpyAsync(host, counters, 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    GA_TRY(cudaStreamSynchronize(st));
    plan->etable = etable;
    plan->ecap = ecap;
    plan->n_nodes = (long long)n_nodes;
    plan->n_edges = (long long)host[3];
    *n_nodes_out = plan->n_nodes;
    *n_edges_out = plan->n_edges;
    *attr_out = plan->n_edges + (long long)host[2];
    sc.release();
    return GA_OK;
}

}  // namespace

extern "C" void ga_csr_plan_free(ga_csr_plan* plan) {
    if (!plan) return;
    for (void* p : plan->owned) cudaFreeAsync(p, plan->stream);
    delete plan;
}

extern "C" int ga_csr_plan_unpaired(const uint64_t* node_stamp_dev, uint64_t n_solid, const void* solid_keys_dev,
                                    int key_words, int sym_bits, const void* edge_table_dev,
                                    uint64_t edge_capacity, ga_stream stream, ga_csr_plan** plan_out,
                                    int64_t* n_nodes, int64_t* n_edges) {
    if (!plan_out || !n_nodes || !n_edges || !edge_table_dev || edge_capacity == 0 ||
        (key_words != 1 && key_words != 2) || (n_solid && (!node_stamp_dev || !solid_keys_dev))) {
        ga_set_error("ga_csr_plan_unpaired: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    ga_csr_plan* plan = new ga_csr_plan();
    plan->key_words = key_words;
    plan->sym_bits = sym_bits;
    plan->solid_keys = solid_keys_dev;
    plan->stream = (cudaStream_t)stream;
    plan->etable = (const StampSlot*)edge_table_dev;
    plan->ecap = edge_capacity;
    Scratch sc{plan->stream, &plan->owned};
    cudaStream_t st = plan->stream;

    u64* counters = sc.get<u64>(3);
    u64* stamps0 = sc.get<u64>(n_solid);
    u32* ids0 = sc.get<u32>(n_solid);
    GA_NEED(counters); GA_NEED(stamps0); GA_NEED(ids0);
    GA_TRY(cudaMemsetAsync(counters, 0, 3 * sizeof(u64), st));
    if (n_solid)
        compact_nodes_kernel<<<scan_grid(n_solid), 256, 0, st>>>((const u64*)node_stamp_dev, n_solid, stamps0, ids0, counters);
    ga_note_launches(1);
    count_slots_kernel<<<scan_grid(edge_capacity), 256, 0, st>>>(plan->etable, edge_capacity, counters + 1);
    ga_note_launches(1);
    GA_TRY(cudaGetLastError());
    u64 host[3];
    GA_TRY(cudaMemcpyAsync(host, counters, sizeof(host), cudaMemcpyDeviceToHost, st));
    GA_TRY(cudaStreamSynchronize(st));
    const u64 nn = host[0];
    u64* stamps1 = sc.get<u64>(nn);
    plan->node_a = sc.get<u32>(nn, true);
    plan->rank = sc.get<u32>(n_solid, true);
    GA_NEED(stamps1); GA_NEED(plan->node_a); GA_NEED(plan->rank);
    GA_TRY((sort_pairs<u64, u32>(sc, stamps0, stamps1, ids0, plan->node_a, nn, key_bits(host[2]))));
    if (nn) scatter_rank_kernel<<<scan_grid(nn), 256, 0, st>>>(plan->node_a, nn, plan->rank);
    ga_note_launches(1);
    GA_TRY(cudaGetLastError());
    plan->n_nodes = (long long)nn;
    plan->n_edges = (long long)host[1];
    *n_nodes = plan->n_nodes;
    *n_edges = plan->n_edges;
    *plan_out = plan;
    sc.release();
    return GA_OK;
}

extern "C" int ga_csr_plan_unpaired_dna(const uint64_t* node_stamp_dev, const uint64_t* edge_stamp_dev,
                                        uint64_t n_solid, const void* solid_keys_dev, int key_words, int k,
                                        int sym_bits, const void* solid_dev, uint64_t solid_capacity,
                                        ga_stream stream, ga_csr_plan** plan_out, int64_t* n_nodes,
                                        int64_t* n_edges) {
    if (!plan_out || !n_nodes || !n_edges || !solid_dev || solid_capacity == 0 || sym_bits > 2 ||
        (key_words != 1 && key_words != 2) || (n_solid && (!node_stamp_dev || !edge_stamp_dev || !solid_keys_dev))) {
        ga_set_error("ga_csr_plan_unpaired_dna: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    ga_csr_plan* plan = new ga_csr_plan();
    plan->key_words = key_words;
    plan->sym_bits = sym_bits;
    plan->solid_keys = solid_keys_dev;
    plan->stream = (cudaStream_t)stream;
    plan->edge_stamp4 = (const u64*)edge_stamp_dev;
    plan->solid = solid_dev;
    plan->solid_cap = solid_capacity;
    plan->w = k - 1;
    Scratch sc{plan->stream, &plan->owned};
    cudaStream_t st = plan->stream;
    u64* counters = sc.get<u64>(3);
    u64* stamps0 = sc.get<u64>(n_solid);
    u32* ids0 = sc.get<u32>(n_solid);
    GA_NEED(counters); GA_NEED(stamps0); GA_NEED(ids0);
    GA_TRY(cudaMemsetAsync(counters, 0, 3 * sizeof(u64), st));
    if (n_solid) {
        compact_nodes_kernel<<<scan_grid(n_solid), 256, 0, st>>>((const u64*)node_stamp_dev, n_solid, stamps0, ids0, counters);
        count_stamps_kernel<<<scan_grid(4 * n_solid), 256, 0, st>>>((const u64*)edge_stamp_dev, 4 * n_solid, counters + 1);
        ga_note_launches(2);
    }
    GA_TRY(cudaGetLastError());
    u64 host[3];
    GA_TRY(cudaMemcpyAsync(host, counters, sizeof(host), cudaMemcpyDeviceToHost, st));
    GA_TRY(cudaStreamSynchronize(st));
    const u64 nn = host[0];
    u64* stamps1 = sc.get<u64>(nn);
    plan->node_a = sc.get<u32>(nn, true);
    plan->rank = sc.get<u32>(n_solid, true);
    GA_NEED(stamps1); GA_NEED(plan->node_a); GA_NEED(plan->rank);
    GA_TRY((sort_pairs<u64, u32>(sc, stamps0, stamps1, ids0, plan->node_a, nn, key_bits(host[2]))));
    if (nn) scatter_rank_kernel<<<scan_grid(nn), 256, 0, st>>>(plan->node_a, nn, plan->rank);
    ga_note_launches(1);
    GA_TRY(cudaGetLastError());
    plan->n_nodes = (long long)nn;
    plan->n_edges = (long long)host[1];
    *n_nodes = plan->n_nodes;
    *n_edges = plan->n_edges;
    *plan_out = plan;
    sc.release();
    return GA_OK;
}

extern "C" int ga_csr_plan_paired(const void* solid_dev, uint64_t solid_capacity, const void* solid_keys_dev,
                                  uint64_t n_solid, int key_words, int k, int sym_bits,
                                  const void* query_table_dev, uint64_t query_capacity,
                                  const void* qedge_table_dev, uint64_t qedge_capacity, const uint64_t* dh_dev,
                                  ga_stream stream, ga_csr_plan** plan_out, int64_t* n_nodes, int64_t* n_edges,
                                  int64_t* num_edges_attr) {
    (void)n_solid;
    if (!plan_out || !n_nodes || !n_edges || !num_edges_attr || !solid_dev || !solid_keys_dev ||
        !query_table_dev || !qedge_table_dev || !dh_dev || query_capacity == 0 || qedge_capacity == 0 ||
        (key_words != 1 && key_words != 2)) {
        ga_set_error("ga_csr_plan_paired: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    ga_csr_plan* plan = new ga_csr_plan();
    plan->paired = 1;
    plan->key_words = key_words;
    plan->sym_bits = sym_bits;
    plan->solid_keys = solid_keys_dev;
    plan->stream = (cudaStream_t)stream;
    Scratch sc{plan->stream, &plan->owned};
    int rc;
    if (key_words == 1)
        rc = plan_paired_impl<u64>(plan, sc, solid_dev, solid_capacity, k - 1, sym_bits,
                                   (const StampSlot*)query_table_dev, query_capacity,
                                   (const StampSlot*)qedge_table_dev, qedge_capacity, (const u64*)dh_dev,
                                   n_nodes, n_edges, num_edges_attr);
    else
        rc = plan_paired_impl<u128>(plan, sc, solid_dev, solid_capacity, k - 1, sym_bits,
                                    (const StampSlot*)query_table_dev, query_capacity,
                                    (const StampSlot*)qedge_table_dev, qedge_capacity, (const u64*)dh_dev,
                                    n_nodes, n_edges, num_edges_attr);
    if (rc == GA_OK) *plan_out = plan;
    return rc;
}

extern "C" int ga_csr_emit(ga_csr_plan* plan, int32_t* rowptr_dev, int32_t* col_dev, int32_t* indeg_dev,
                           uint8_t* branching_dev, uint8_t* last_sym_dev, void* node_keys_a_dev,
                           void* node_keys_b_dev, ga_stream stream) {
    if (!plan || !rowptr_dev || !col_dev || !indeg_dev || !branching_dev || !last_sym_dev) {
        ga_set_error("ga_csr_emit: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<void*> none;
    Scratch sc{st, &none};
    const u64 nn = (u64)plan->n_nodes, ne = (u64)plan->n_edges;
    int* outdeg = sc.get<int>(nn + 1);
    u64* estamp = sc.get<u64>(ne);
    if (!outdeg || !estamp) {
        sc.release();
        return ga_cuda_fail(sc.err, "cudaMallocAsync");
    }
#define GA_TRY2(expr)                               \
    do {                                            \
        cudaError_t _e = (expr);                    \
        if (_e != cudaSuccess) {                    \
            sc.release();                           \
            return ga_cuda_fail(_e, #expr);         \
        }                                           \
    } while (0)
    if (plan->edge_stamp4) {
        GA_TRY2(cudaMemsetAsync(outdeg, 0, (nn + 1) * sizeof(int), st));
        if (nn) GA_TRY2(cudaMemsetAsync(indeg_dev, 0, nn * sizeof(int), st));
        if (nn) dna4_outdeg_kernel<<<scan_grid(nn), 256, 0, st>>>(plan->edge_stamp4, plan->node_a, nn, outdeg);
        GA_TRY2(exclusive_sum(sc, outdeg, rowptr_dev, nn + 1));
        if (nn) {
            u32 smask = (1u << plan->sym_bits) - 1u;
            if (plan->key_words == 1) {
                dna4_rows_kernel<u64><<<scan_grid(nn), 256, 0, st>>>(plan->edge_stamp4, plan->node_a, nn,
                    (const u64*)plan->solid_keys, (const Slot<u64>*)plan->solid, plan->solid_cap, plan->rank, plan->w,
                    plan->sym_bits, rowptr_dev, col_dev, indeg_dev);
                dna4_finish_kernel<u64><<<scan_grid(nn), 256, 0, st>>>(nn, rowptr_dev, indeg_dev, (const u64*)plan->solid_keys,
                    plan->node_a, smask, branching_dev, last_sym_dev, (u64*)node_keys_a_dev);
            } else {
                dna4_rows_kernel<u128><<<scan_grid(nn), 256, 0, st>>>(plan->edge_stamp4, plan->node_a, nn,
                    (const u128*)plan->solid_keys, (const Slot<u128>*)plan->solid, plan->solid_cap, plan->rank, plan->w,
                    plan->sym_bits, rowptr_dev, col_dev, indeg_dev);
                dna4_finish_kernel<u128><<<scan_grid(nn), 256, 0, st>>>(nn, rowptr_dev, indeg_dev, (const u128*)plan->solid_keys,
                    plan->node_a, smask, branching_dev, last_sym_dev, (u128*)node_keys_a_dev);
            }
            ga_note_launches(3);
        }
        GA_TRY2(cudaGetLastError());
        sc.release();
        return GA_OK;
    }
    GA_TRY2(cudaMemsetAsync(outdeg, 0, (nn + 1) * sizeof(int), st));
    if (plan->extra_in && nn)
        GA_TRY2(cudaMemcpyAsync(indeg_dev, plan->extra_in, nn * sizeof(int), cudaMemcpyDeviceToDevice, st));
    else if (nn)
        GA_TRY2(cudaMemsetAsync(indeg_dev, 0, nn * sizeof(int), st));
    if (ne) degree_kernel<<<scan_grid(plan->ecap), 256, 0, st>>>(plan->etable, plan->ecap, plan->rank, outdeg, indeg_dev);
    ga_note_launches(1);
    GA_TRY2(exclusive_sum(sc, outdeg, rowptr_dev, nn + 1));
    GA_TRY2(cudaMemsetAsync(outdeg, 0, (nn + 1) * sizeof(int), st));
    if (ne) place_edges_kernel<<<scan_grid(plan->ecap), 256, 0, st>>>(plan->etable, plan->ecap, plan->rank, rowptr_dev, outdeg, col_dev, estamp);
    ga_note_launches(1);
    if (nn) {
        u32 smask = (1u << plan->sym_bits) - 1u;
        if (plan->key_words == 1)
            finish_nodes_kernel<u64><<<scan_grid(nn), 256, 0, st>>>(plan->n_nodes, rowptr_dev, col_dev, estamp, indeg_dev,
                                                                   (const u64*)plan->solid_keys, plan->node_a, plan->node_b,
                                                                   smask, branching_dev, last_sym_dev,
                                                                   (u64*)node_keys_a_dev, (u64*)node_keys_b_dev);
        else
            finish_nodes_kernel<u128><<<scan_grid(nn), 256, 0, st>>>(plan->n_nodes, rowptr_dev, col_dev, estamp, indeg_dev,
                                                                    (const u128*)plan->solid_keys, plan->node_a, plan->node_b,
                                                                    smask, branching_dev, last_sym_dev,
                                                                    (u128*)node_keys_a_dev, (u128*)node_keys_b_dev);
    ga_note_launches(1);
    }
    GA_TRY2(cudaGetLastError());
    sc.release();
    return GA_OK;
}
