// ga_prefilter.cu -- exact counting restricted to the windows that can pass the filter.
//
// The reference counts every (k-1)-mer (debruijn_graph.py:144-152, 349-367) and then only ever asks
// "count > threshold" (:127-128, 275-278).  On real reads most distinct (k-1)-mers are sequencing
// errors seen once; giving each a slot makes the count table several times the 126 MB L2 and every
// probe a DRAM miss (profiles/r01/ncu_summary_v1.md).  Here pass A bumps one small saturating
// counter per occurrence in a sketch sized to stay L2 resident; a cell can only over-estimate, so
// "cell > threshold" is a superset of the solid windows.  Pass B counts exactly those candidates in
// a table a fraction of the size.  The filter outcome is bit-identical to counting everything.
#include "ga_common.cuh"

namespace {

template <class K, int SB>
__global__ void __launch_bounds__(256) prefilter_update_kernel(ReadsView rv, int w, PrefilterView pf) {
    const K mask = ga_key_mask<K>(w, rv.sym_bits);
    const u32 cmask = (1u << pf.cell_bits) - 1u;
    GA_FOR_EACH_READ_WARP(rv, r, valid) {
        u32 len = valid ? ga_read_len(rv, rv.paired ? (r & ~1ull) : r) : 0u;
        if (len < (u32)w) len = 0;
        ga_for_each_window<K, SB>(valid ? ga_read_ptr(rv, r) : rv.words, len, w, rv.sym_bits, mask, [&](u32, K key) {
            u64 wi;
            u32 sh;
            u32 v = ga_prefilter_value(pf, ga_key_hash(key), wi, sh);
            if (v < pf.limit) {   // cells stop counting once they mark a candidate
                // Carry-free bump: the field is incremented by a CAS on its word only while it is below the
                // limit (<= the field's maximum), so racing adders can neither wrap the cell nor carry into
                // its neighbour -- a cell never under-estimates, whatever the contention.
                u32 cur = __ldcg(pf.words + wi);
                while (((cur >> sh) & cmask) < pf.limit) {
                    const u32 seen = atomicCAS(pf.words + wi, cur, cur + (1u << sh));
                    if (seen == cur) break;
                    cur = seen;
                }
            }
        });
    }
}

__global__ void prefilter_hot_kernel(PrefilterView pf, u64 n_words, u64* n_hot) {
    const u32 per = 1u << pf.lg_per, cmask = (1u << pf.cell_bits) - 1u;
    u64 n = 0;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n_words; i += (u64)gridDim.x * blockDim.x) {
        u32 word = pf.words[i];
        for (u32 j = 0; j < per; ++j) n += ((word >> (j * pf.cell_bits)) & cmask) >= pf.limit;
    }
    for (int off = 16; off > 0; off >>= 1) n += __shfl_down_sync(0xFFFFFFFFu, n, off);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(n_hot, n);
}

template <class K, int SB>
__global__ void __launch_bounds__(256)
count_candidates_kernel(ReadsView rv, int w, PrefilterView pf, Slot<K>* __restrict__ table, u64 capacity,
                        u32* status) {
    const K mask = ga_key_mask<K>(w, rv.sym_bits);
    bool full = false;
    GA_FOR_EACH_READ_WARP(rv, r, valid) {
        u32 len = valid ? ga_read_len(rv, rv.paired ? (r & ~1ull) : r) : 0u;
        if (len < (u32)w) len = 0;
        ga_for_each_window<K, SB>(valid ? ga_read_ptr(rv, r) : rv.words, len, w, rv.sym_bits, mask, [&](u32, K key) {
            u64 wi;
            u32 sh;
            if (ga_prefilter_value(pf, ga_key_hash(key), wi, sh) < pf.limit) return;
            u64 s = ga_table_upsert(table, capacity, key);
            if (s == GA_NONE64) full = true;
            else atomicAdd(&table[s].val, 1u);
        });
    }
    if (full) atomicOr(status, GA_ST_TABLE_FULL);
}

int check_prefilter(const char* fn, const ga_prefilter* pf, long long threshold) {
    if (!pf || !pf->words || pf->n_cells == 0 || (pf->cell_bits != 4 && pf->cell_bits != 8) ||
        threshold + 1 > (1ll << pf->cell_bits) - 1) {
        ga_set_error("%s: bad pre-filter (cells=%llu bits=%d threshold=%lld)", fn,
                     pf ? (unsigned long long)pf->n_cells : 0ull, pf ? pf->cell_bits : 0, threshold);
        return GA_ERR_BAD_ARG;
    }
    return GA_OK;
}

}  // namespace

extern "C" int ga_prefilter_update(const ga_reads* reads, int k, const ga_prefilter* pf, int64_t threshold,
                                   ga_stream stream) {
    int rc = check_prefilter("ga_prefilter_update", pf, threshold);
    if (rc) return rc;
    if (!reads) return GA_ERR_BAD_ARG;
    int kw = ga_key_words(k, reads->sym_bits);
    if (!kw) {
        ga_set_error("ga_prefilter_update: unsupported key width");
        return GA_ERR_BAD_ARG;
    }
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    PrefilterView pv = ga_prefilter_view(pf, threshold);
    unsigned grid = ga_grid(rv.n_reads, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (kw == 1 && rv.storage_bits == 2) prefilter_update_kernel<u64, 2><<<grid, 256, 0, st>>>(rv, k - 1, pv);
    else if (kw == 1) prefilter_update_kernel<u64, 8><<<grid, 256, 0, st>>>(rv, k - 1, pv);
    else if (rv.storage_bits == 2) prefilter_update_kernel<u128, 2><<<grid, 256, 0, st>>>(rv, k - 1, pv);
    else prefilter_update_kernel<u128, 8><<<grid, 256, 0, st>>>(rv, k - 1, pv);
    GA_LAUNCH_CHECK("prefilter_update");
    return GA_OK;
}

extern "C" int ga_prefilter_hot(const ga_prefilter* pf, int64_t threshold, uint64_t* n_hot_dev, ga_stream stream) {
    int rc = check_prefilter("ga_prefilter_hot", pf, threshold);
    if (rc) return rc;
    PrefilterView pv = ga_prefilter_view(pf, threshold);
    u64 n_words = (pf->n_cells + (1u << pv.lg_per) - 1) >> pv.lg_per;
    unsigned grid = ga_grid(n_words, 256);
    if (grid > 148u * 16u) grid = 148u * 16u;
    prefilter_hot_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pv, n_words, (u64*)n_hot_dev);
    GA_LAUNCH_CHECK("prefilter_hot");
    return GA_OK;
}

extern "C" int ga_count_candidates(const ga_reads* reads, int k, const ga_prefilter* pf, int64_t threshold,
                                   void* table_dev, uint64_t capacity, uint32_t* status_dev, ga_stream stream) {
    int rc = check_prefilter("ga_count_candidates", pf, threshold);
    if (rc) return rc;
    if (!reads || !table_dev || capacity == 0) return GA_ERR_BAD_ARG;
    int kw = ga_key_words(k, reads->sym_bits);
    if (!kw) {
        ga_set_error("ga_count_candidates: unsupported key width");
        return GA_ERR_BAD_ARG;
    }
    if (reads->n_reads == 0) return GA_OK;
    ReadsView rv = ga_view(reads);
    PrefilterView pv = ga_prefilter_view(pf, threshold);
    unsigned grid = ga_grid(rv.n_reads, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define GA_CAND(K, SB) \
    count_candidates_kernel<K, SB><<<grid, 256, 0, st>>>(rv, k - 1, pv, (Slot<K>*)table_dev, capacity, status_dev)
    if (kw == 1 && rv.storage_bits == 2) GA_CAND(u64, 2);
    else if (kw == 1) GA_CAND(u64, 8);
    else if (rv.storage_bits == 2) GA_CAND(u128, 2);
    else GA_CAND(u128, 8);
#undef GA_CAND
    GA_LAUNCH_CHECK("count_candidates");
    return GA_OK;
}
