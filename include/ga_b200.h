/*
 * ga_b200.h -- C ABI of libga_b200.so: the B200 (sm_100a) implementation of the k-mer
 * counting / CountMinSketch / de Bruijn graph construction path of
 * tonycheang/genome-assembler.
 *
 * The reference is pure Python and has no FFI; the functions below are what a ctypes
 * binding for that path binds.  Each entry names the reference
 * code it replaces (paths relative to the upstream checkout).
 *
 * Conventions
 *  - plain C, no C++/torch types; every pointer named *_dev is a CUDA device pointer
 *    owned by the caller; `stream` is a cudaStream_t passed as void*.
 *  - every function returns 0 (GA_OK) or a negative GA_ERR_* code and never throws or
 *    exits; ga_last_error() gives the message (thread-local).
 *  - functions are asynchronous on `stream` unless documented as synchronising.
 *  - device-side conditions (a full hash table, a symbol outside the alphabet, a sketch
 *    cell above 65535) are OR-ed into the caller's `status_dev[0]` (GA_ST_* bits); the
 *    caller reads it at its next synchronisation point.
 *  - there is no CPU fallback: without a CUDA device every compute entry returns
 *    GA_ERR_CUDA.
 */
#ifndef GA_B200_H
#define GA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GA_OK 0
#define GA_ERR_BAD_ARG (-1)
#define GA_ERR_CAPACITY (-2)
#define GA_ERR_CUDA (-3)
#define GA_ERR_NCCL (-4)
#define GA_ERR_OVERFLOW_U16 (-5)
#define GA_ERR_ALPHABET (-6)

#define GA_STATUS_TABLE_FULL 1u
#define GA_STATUS_BAD_SYMBOL 2u
#define GA_STATUS_STAMP_FULL 4u
#define GA_STATUS_U16_OVERFLOW 8u

#define GA_MAX_SKETCH_ROWS 20
#define GA_PEER_MAX_RANKS 16    /* ranks of one node that exchange records over NVLink peer memory */

typedef void* ga_stream;

/* Packed reads resident on the device.  A read is a run of 64-bit words; symbol i sits at bits
 * [storage_bits*(i % spw), +storage_bits) of word i / spw, spw = 64 / storage_bits.
 * storage_bits = 2: A,C,G,T = 0..3 (DNA); storage_bits = 8: one code byte per symbol
 * (arbitrary alphabets, codes < 2^sym_bits).  Paired input stores mate 1 of pair p as read 2p
 * and mate 2 as read 2p+1; windows are taken over mate 1's length (debruijn_graph.py:369-374)
 * so mate 2 must be at least as long. */
typedef struct ga_reads {
    const void* words;       /* device */
    const uint64_t* offsets; /* device, first word of each read; NULL = r * stride_words */
    const uint32_t* lengths; /* device, symbols per read; NULL = uniform_len */
    uint64_t n_reads;        /* reads, or 2 * pairs when paired */
    uint64_t first_read;     /* global index of the first read / pair of this shard (stamps) */
    uint32_t uniform_len;
    uint32_t stride_words;
    int32_t storage_bits;    /* 2 or 8 */
    int32_t sym_bits;        /* bits per symbol inside a key: 2, or 1..8 for storage_bits 8 */
    int32_t paired;
    uint32_t estride;        /* stamp stride per read, >= longest read; same on all shards */
} ga_reads;

/* CountMinSketch geometry (countminsketch.py:26-32): `rows` rows, row i has width[i] cells.
 * cells_dev holds 32-bit working cells, rows back to back; ga_sketch_narrow() produces the
 * reference's unsigned-16 rows and reports cells above 65535. */
typedef struct ga_sketch {
    void* cells;             /* device, uint32_t[sum(width)] */
    uint32_t width[GA_MAX_SKETCH_ROWS];
    int32_t rows;
} ga_sketch;

/* Pre-filter sketch: one row of n_cells saturating counters of cell_bits (4 or 8) bits packed in
 * 32-bit words (zero-filled by the caller).  Internal to the exact path: a cell only ever
 * over-estimates the occurrences hashed to it, so "cell > threshold" selects a superset of the
 * windows whose exact count is > threshold; the superset is then counted exactly. */
typedef struct ga_prefilter {
    void* words;             /* device, uint32_t[ceil(n_cells * cell_bits / 32)] */
    uint64_t n_cells;
    int32_t cell_bits;       /* 4 (threshold <= 14) or 8 (threshold <= 254) */
} ga_prefilter;

/* ---- housekeeping ------------------------------------------------------------------------- */
int ga_version(void);
const char* ga_last_error(void);
int ga_device_count(void);                 /* number of CUDA devices, 0 without a GPU */
uint64_t ga_launch_count(void);            /* kernels of this library launched so far (process-wide) */
int ga_set_l2_fetch_granularity(int bytes); /* 32, 64 or 128: DRAM bytes fetched per L2 miss (hint) */
int ga_fill_bytes(void* dev, int value, uint64_t bytes, ga_stream stream);
int ga_copy_bytes(void* dst, const void* src, uint64_t bytes, ga_stream stream);   /* cudaMemcpyAsync, any direction (peer-mapped too) */

/* key geometry: 1 word (64-bit keys) when (k-1)*sym_bits <= 63, 2 words when <= 127,
 * otherwise 0 (unsupported on this build). */
int ga_key_words(int k, int sym_bits);
/* bytes per slot of a count / id table with that key width (16 or 32) */
int ga_slot_bytes(int key_words);

/* ---- read ingestion: replaces IOHandler.read_input's string list (assemble.py:40-71) ------ */
/* ASCII symbols -> packed reads.  ascii_dev: all reads back to back; in_offsets_dev[n+1] byte
 * offsets (NULL: read r at r*uniform_len).  lut_dev[256]: byte -> code, 0xFF = not in the
 * alphabet (sets GA_STATUS_BAD_SYMBOL).  Output layout as described at ga_reads. */
int ga_pack_reads(const uint8_t* ascii_dev, const uint64_t* in_offsets_dev, uint64_t n_reads,
                  uint32_t uniform_len, const uint8_t* lut_dev, int storage_bits, void* words_dev,
                  const uint64_t* out_offsets_dev, uint32_t stride_words, uint32_t* status_dev,
                  ga_stream stream);

/* Inverse of ga_pack_reads for uniform-length reads: packed words -> n_reads*uniform_len bytes.
 * inv_lut_dev[256]: symbol code -> byte. */
int ga_unpack_reads(const void* words_dev, uint64_t n_reads, uint32_t uniform_len, uint32_t stride_words,
                    int storage_bits, const uint8_t* inv_lut_dev, uint8_t* ascii_dev, ga_stream stream);

/* Synthetic reads on the device (our replacement for generate_reads.py:42-74 where it cannot
 * produce the shape: any length, per-base substitutions, seeded; arithmetic documented in
 * oracle/readgen.py splitmix_*).  genome_codes_dev: one 2-bit code per byte.  first_read /
 * n_reads count stored reads (mates when paired: read r is mate r&1 of pair r>>1, mate 2 drawn
 * mate_distance bases after mate 1). */
int ga_gen_genome(uint8_t* genome_codes_dev, uint64_t size, uint64_t seed, ga_stream stream);
int ga_gen_reads(const uint8_t* genome_codes_dev, uint64_t genome_size, uint64_t first_read,
                 uint64_t n_reads, uint32_t read_len, uint64_t seed, uint32_t sub_per_10k,
                 void* words_dev, uint32_t stride_words, int paired, uint32_t mate_distance,
                 ga_stream stream);

/* ---- exact counting: replaces _count_kmers (debruijn_graph.py:144-152, 349-367) ----------- */
int ga_table_clear(void* table_dev, uint64_t capacity, int key_words, ga_stream stream);
/* one increment per window occurrence of every read (both mates when paired) */
int ga_count_kmers(const ga_reads* reads, int k, void* table_dev, uint64_t capacity,
                   uint32_t* status_dev, ga_stream stream);
/* add `amounts_dev[i]` (NULL: 1) for explicit packed keys (multi-GPU owner side; dict upload) */
int ga_count_keys(const void* keys_dev, const uint32_t* amounts_dev, uint64_t n, int key_words,
                  void* table_dev, uint64_t capacity, uint32_t* status_dev, ga_stream stream);
/* owner_dev[i] = rank in [0, n_parts) that owns keys_dev[i] in the hash-partitioned exchange */
int ga_key_owner(const void* keys_dev, uint64_t n, int key_words, uint32_t n_parts, int32_t* owner_dev,
                 ga_stream stream);
/* out4_dev = { distinct keys, keys with count > threshold, sum of counts, max count } */
int ga_table_summary(const void* table_dev, uint64_t capacity, int key_words, int64_t threshold,
                     uint64_t* out4_dev, ga_stream stream);
/* compact (key, count) of every slot with count > min_exclusive (-1: all) into keys_out_dev /
 * counts_out_dev (either may be NULL); *n_out_dev must be zeroed by the caller */
int ga_table_export(const void* table_dev, uint64_t capacity, int key_words, int64_t min_exclusive,
                    void* keys_out_dev, uint32_t* counts_out_dev, uint64_t* n_out_dev,
                    ga_stream stream);
/* counts_out_dev[i] = count of keys_dev[i], 0 when absent (dict __getitem__ of a defaultdict) */
int ga_table_lookup(const void* table_dev, uint64_t capacity, int key_words, const void* keys_dev,
                    uint64_t n, uint32_t* counts_out_dev, ga_stream stream);
/* id table: insert keys_dev[i] -> id_base + i (keys must be distinct) */
int ga_table_insert_ids(const void* keys_dev, uint64_t n, int key_words, uint32_t id_base,
                        void* table_dev, uint64_t capacity, uint32_t* status_dev, ga_stream stream);

/* Two-pass exact counting of the windows that can pass the filter (same results as
 * ga_count_kmers + "count > threshold", without giving every singleton a table slot):
 * pass A bumps one pre-filter cell per window occurrence; ga_prefilter_hot counts the cells that
 * reached threshold+1 (sizes the candidate table); pass B counts exactly, in `table_dev`, the
 * windows whose cell reached threshold+1. */
int ga_prefilter_update(const ga_reads* reads, int k, const ga_prefilter* pf, int64_t threshold,
                        ga_stream stream);
int ga_prefilter_hot(const ga_prefilter* pf, int64_t threshold, uint64_t* n_hot_dev, ga_stream stream);
int ga_count_candidates(const ga_reads* reads, int k, const ga_prefilter* pf, int64_t threshold,
                        void* table_dev, uint64_t capacity, uint32_t* status_dev, ga_stream stream);

/* ---- bucketed count + build (unpaired DNA reads, 64-bit keys): replaces BOTH hot loops,
 *      _count_kmers (debruijn_graph.py:144-152) and _build_graph (:113-142), without one random
 *      global-memory access per occurrence (csrc/ga_superkmer.cu) ------------------ */
/* m-mer length used to pick a window's bucket for --kmer_length k */
int ga_sk_minimizer_len(int k);
/* Paired input (reads->paired) is taken as 2*pairs plain reads whose windows run over mate 1's length
 * (debruijn_graph.py:349-374): enough for counting, i.e. for ga_sk_count_build without edge stamps.
 * Cut every read into records (runs of consecutive windows that share a bucket: 16 bytes of bases + one
 * meta word) and scatter them to 2^l1_bits level-1 buckets of l1_capacity 32-byte slots each in records_dev
 * (slot = {bases hi, bases lo, meta, 0}, one full-sector store per record; bucket b starts at slot
 * b*l1_capacity; records_dev 32-byte aligned).  l1_cursors_dev[b * ga_sk_cursor_stride()] = records written
 * to bucket b (2^l1_bits * stride words, zeroed by the caller; the cursors sit 128 bytes apart so that their
 * atomics spread over the L2 slices).  hist_dev[2^(l1_bits+l2_bits)] (zeroed by the caller) accumulates, per
 * final bucket, records << 32 | windows.  GA_STATUS_TABLE_FULL: a level-1 bucket overflowed (its cursor kept
 * counting), retry with a larger capacity.  GA_SK_SCATTER=warp|lane|lane128 in the environment picks
 * the kernel by hand (same records from each; tests and A/B runs). */
int ga_sk_cursor_stride(void);
int ga_sk_scatter_reads(const ga_reads* reads, int k, int l1_bits, int l2_bits, void* records_dev,
                        uint64_t l1_capacity, uint64_t* l1_cursors_dev, uint64_t* hist_dev,
                        uint32_t* status_dev, ga_stream stream);
/* offsets_dev[n_buckets+1] = exclusive prefix sum of the record counts; cursors_dev[n_buckets] = a copy */
int ga_sk_offsets(const uint64_t* hist_dev, uint64_t n_buckets, uint64_t* offsets_dev,
                  uint64_t* cursors_dev, ga_stream stream);
/* level-1 buckets -> final buckets at offsets_dev (cursors_dev is consumed), in one of two forms:
 *  dense (out_bases_dev + out_meta_dev, out_index_dev NULL): the records themselves, densely packed --
 *    what the multi-GPU exchange sends;
 *  index (out_index_dev, the other two NULL): per record only its 32-bit position inside its level-1
 *    bucket; ga_sk_count_build then gathers the records from the level-1 buckets (a quarter of the
 *    traffic on one GPU, where a level-1 bucket stays L2 resident while its final buckets are done). */
int ga_sk_scatter_buckets(const void* records_dev, uint64_t l1_capacity, const uint64_t* l1_cursors_dev,
                          int l1_bits, int l2_bits, uint64_t* cursors_dev, void* out_bases_dev,
                          uint64_t* out_meta_dev, uint32_t* out_index_dev, ga_stream stream);
/* offsets_dev holds n_segments rows of n_buckets+1 positions: the records of bucket b are the union of
 * [offsets[s][b], offsets[s][b+1]) over the segments s (one segment on a single GPU; after the
 * multi-GPU exchange, one per source rank, each sorted by bucket).  hist_dev[b] & 0xFFFFFFFF = windows
 * of bucket b over all segments.  index_dev != NULL selects the index form (single segment): bases_dev is
 * then records_dev of ga_sk_scatter_reads (32-byte slots; meta_dev is not used and may be NULL), offsets
 * address index_dev, and the record of entry e of bucket b is slot (b >> l2_bits) * l1_capacity + index_dev[e].
 * A bucket must hold fewer than 2^25 records and l1_capacity must be below 2^25 (a table slot names a record
 * in 25 bits; the host checks).
 * One CTA per bucket and ONE walk over its records: exact counts in a shared-memory table of at most
 * table_slots 16-byte slots {key, state} (a power of two, 256..4096); a window seen twice
 * becomes a candidate with 4 stamp slots (at most max_solid candidates per pass, bounded by what is left of
 * the 110 KB pool), the occurrence seen before that is remembered in the slot's state word, so no record is
 * ever walked twice.  Every window with count > threshold is appended to solid_keys_out_dev together with 4 candidate edge
 * stamps (edge_stamp_out_dev[4*i + c] = smallest occurrence ordinal of "window i followed by symbol
 * c", all-ones if never; edge_stamp_out_dev == NULL: counting only, just the solid windows).  counters_dev[8] (zeroed by the caller): [0] scheduling cursor, [1] solid
 * windows found (may exceed out_capacity: nothing is written beyond it, the caller retries with
 * that many), [2] passes that did not fit and were listed in spill_list_dev.  A bucket whose distinct
 * or candidate windows exceed the pool is done in 2, 4, ... 32 passes over disjoint hash ranges of its
 * windows, still in shared memory; only what would need more is listed (entry = bucket | passes << 32
 * | pass << 48). */
int ga_sk_count_build(const void* bases_dev, const uint64_t* meta_dev, const uint64_t* offsets_dev,
                      uint32_t n_segments, const uint64_t* hist_dev, uint64_t n_buckets, int k, int64_t threshold, uint32_t table_slots,
                      uint32_t max_solid, uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev,
                      uint64_t out_capacity, uint64_t* counters_dev, uint64_t* spill_list_dev,
                      uint64_t spill_capacity, uint32_t* status_dev, const uint32_t* index_dev,
                      uint64_t l1_capacity, int l2_bits, ga_stream stream);
/* Sources form (multi-GPU, the exchange fused into the count: nothing is copied): the records stay in the level-1
 * slots of the rank that cut them, and the owner of a bucket range gathers them while it counts -- from its own
 * memory or over NVLink from buffers mapped with ga_peer_open.  Segment s of a bucket is source rank s:
 * offsets_dev row s = rank s's bucket-sorted positions of this launch's buckets (n_buckets + 1 words, addressing
 * index[s]); the record of entry e is slot (global bucket id >> l2_bits) * l1_capacity[s] + index[s][e] of
 * records[s], global bucket id = first_bucket + bucket number.  HOST struct; pointers are device pointers valid
 * in this process.  Everything else as ga_sk_count_build / ga_sk_count_build_spill. */
typedef struct ga_sk_sources {
    const void* records[GA_PEER_MAX_RANKS];     /* 32-byte slots of ga_sk_scatter_reads, one array per source rank */
    const uint32_t* index[GA_PEER_MAX_RANKS];   /* index form of ga_sk_scatter_buckets, one array per source rank */
    uint64_t l1_capacity[GA_PEER_MAX_RANKS];
    uint64_t first_bucket;
    uint64_t* solid_counter;  /* optional device word shared by all ranks: solid windows appended so far to
                               * solid_keys_out_dev / edge_stamp_out_dev, which are then the SAME buffers on every rank
                               * (one rank's, mapped with ga_peer_open) -- results land where the graph is built, no
                               * gather afterwards; zeroed by the owner before any rank launches.  NULL: counters_dev[1] */
    uint32_t n_sources;
} ga_sk_sources;
int ga_sk_count_build_from(const ga_sk_sources* sources, const uint64_t* offsets_dev, const uint64_t* hist_dev,
                           uint64_t n_buckets, int k, int64_t threshold, uint32_t table_slots, uint32_t max_solid,
                           uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev, uint64_t out_capacity,
                           uint64_t* counters_dev, uint64_t* spill_list_dev, uint64_t spill_capacity,
                           uint32_t* status_dev, int l2_bits, ga_stream stream);
int ga_sk_count_build_spill_from(const ga_sk_sources* sources, const uint64_t* offsets_dev, uint64_t n_buckets,
                                 const uint64_t* spill_list_dev, uint64_t n_spill, int k, int64_t threshold,
                                 uint32_t table_slots, void* scratch_dev, uint32_t n_ctas,
                                 uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev, uint64_t out_capacity,
                                 uint64_t* counters_dev, uint32_t* status_dev, int l2_bits, ga_stream stream);
/* The listed buckets again with tables in global scratch (n_ctas slices of
 * ga_sk_spill_scratch_bytes(table_slots) bytes; table_slots >= twice the windows of the largest). */
uint64_t ga_sk_spill_scratch_bytes(uint32_t table_slots);
int ga_sk_count_build_spill(const void* bases_dev, const uint64_t* meta_dev, const uint64_t* offsets_dev,
                            uint32_t n_segments, uint64_t n_buckets, const uint64_t* spill_list_dev, uint64_t n_spill, int k, int64_t threshold,
                            uint32_t table_slots, void* scratch_dev, uint32_t n_ctas,
                            uint64_t* solid_keys_out_dev, uint64_t* edge_stamp_out_dev,
                            uint64_t out_capacity, uint64_t* counters_dev, uint32_t* status_dev,
                            const uint32_t* index_dev, uint64_t l1_capacity, int l2_bits, ga_stream stream);
/* Candidate stamps -> the reference's graph: clears edge_stamp_dev[4*i + c] when the successor of
 * window i through c is not solid (solid_dev: id table over solid_keys_dev, id = index) and folds
 * 2e / 2e+1 into node_stamp_dev[n_solid] (0xFF filled by the caller).  The arrays then feed
 * ga_csr_plan_unpaired_dna. */
int ga_sk_resolve(const uint64_t* solid_keys_dev, uint64_t n_solid, int k, const void* solid_dev,
                  uint64_t solid_capacity, uint64_t* edge_stamp_dev, uint64_t* node_stamp_dev,
                  ga_stream stream);

/* ---- CountMinSketch: replaces countminsketch.py:34-44 and _make_sketch --------------------- */
/* cells[row][murmur3(window) % width[row]] += count for every key of the count table
 * (debruijn_graph.py:181-188, 398-405).  lut_dev[256]: symbol code -> byte. */
int ga_sketch_update_table(const void* table_dev, uint64_t capacity, int key_words, int k,
                           int sym_bits, const uint8_t* lut_dev, const ga_sketch* sketch,
                           ga_stream stream);
/* update(string, amount) / estimate(string) in bulk on raw bytes: string i is
 * bytes_dev[offsets_dev[i] .. offsets_dev[i+1]) */
int ga_sketch_update_bytes(const uint8_t* bytes_dev, const uint64_t* offsets_dev,
                           const uint32_t* amounts_dev, uint64_t n, const ga_sketch* sketch,
                           ga_stream stream);
int ga_sketch_estimate_bytes(const uint8_t* bytes_dev, const uint64_t* offsets_dev, uint64_t n,
                             const ga_sketch* sketch, uint32_t* est_out_dev, ga_stream stream);
/* 32-bit working cells -> the reference's array('H') rows; a cell above 65535 sets
 * GA_STATUS_U16_OVERFLOW (the reference raises OverflowError there) */
int ga_sketch_narrow(const ga_sketch* sketch, uint16_t* rows_out_dev, uint32_t* status_dev,
                     ga_stream stream);

/* ---- filter: the strict `> threshold` test of _build_graph (debruijn_graph.py:127-128,
 *      275-278) evaluated once per distinct window ---------------------------------------- */
/* Compacts the keys whose exact count (sketch == NULL) or sketch estimate is > threshold.
 * *n_out_dev must be zeroed by the caller.  counts_out_dev may be NULL. */
int ga_select_solid(const void* table_dev, uint64_t capacity, int key_words, int k, int sym_bits,
                    int64_t threshold, const ga_sketch* sketch, const uint8_t* lut_dev,
                    void* keys_out_dev, uint32_t* counts_out_dev, uint64_t* n_out_dev,
                    ga_stream stream);

/* ---- graph build: replaces _build_graph (debruijn_graph.py:113-142 unpaired, 269-317 paired) */
/* Unpaired.  solid_dev: id table (key -> dense id < n_solid).  node_stamp_dev[n_solid] and the
 * edge stamp table (16-byte slots, key = src_id << 32 | dst_id) must be filled with 0xFF. */
int ga_build_unpaired(const ga_reads* reads, int k, const void* solid_dev, uint64_t solid_capacity,
                      uint64_t* node_stamp_dev, void* edge_table_dev, uint64_t edge_capacity,
                      uint32_t* status_dev, ga_stream stream);
/* Unpaired with at most 4 symbols (DNA): no edge table.  solid_dev is an id table whose aux words
 * are 0xFFFFFFFF (ga_table_clear); edge_stamp_dev[4*n_solid] (0xFF filled) takes the first
 * occurrence of edge (node id, next symbol).  The aux word of a slot caches, per stamp, the
 * earliest read epoch that already folded it in, so repeat occurrences stop at the slot. */
int ga_build_unpaired_dna(const ga_reads* reads, int k, void* solid_dev, uint64_t solid_capacity,
                          uint64_t* node_stamp_dev, uint64_t* edge_stamp_dev, uint32_t* status_dev,
                          ga_stream stream);
/* Two-phase variant for read sets whose id table does not fit the L2: after ga_build_unpaired_dna
 * has run over a PREFIX of the reads, almost every node and edge already holds its final (earliest)
 * stamp.  ga_unstamped_scan finds what is still open: mask_out_dev[id] = bit c set iff the
 * successor of node id through symbol c is solid and edge (id, c) has no stamp yet (*n_open_dev =
 * number of nodes with a non-zero mask; zeroed by the caller).  ga_unstamped_table_build puts
 * those nodes in a small table (key -> id, mask) fronted by a blocked Bloom filter
 * (bloom_dev: uint32 words, zero-filled).  ga_build_unpaired_dna_tail then walks the remaining
 * reads: one Bloom probe (L2 resident) per window, and only the rare hits touch the tables. */
int ga_unstamped_scan(const void* solid_dev, uint64_t solid_capacity, const void* solid_keys_dev,
                      uint64_t n_solid, int key_words, const uint64_t* edge_stamp_dev, int k, int sym_bits,
                      uint8_t* mask_out_dev, uint64_t* n_open_dev, ga_stream stream);
int ga_unstamped_table_build(const void* solid_keys_dev, const uint8_t* mask_dev, uint64_t n_solid,
                             int key_words, void* open_table_dev, uint64_t open_capacity,
                             uint32_t* bloom_dev, uint64_t bloom_words, uint32_t* status_dev,
                             ga_stream stream);
int ga_build_unpaired_dna_tail(const ga_reads* reads, int k, const uint32_t* bloom_dev, uint64_t bloom_words,
                               const void* open_table_dev, uint64_t open_capacity, const void* solid_dev,
                               uint64_t solid_capacity, uint64_t* node_stamp_dev, uint64_t* edge_stamp_dev,
                               ga_stream stream);
/* Paired.  query table: key = idA << 32 | idB -> min stamp; query-edge table: key =
 * query_slot(P) << 32 | query_slot(S) -> min occurrence; dh_dev: 256*256*2 uint64 (0xFF filled),
 * the two smallest occurrences of "prefix pair == suffix pair" per (symbol A, symbol B)
 * (SURVEY App. A-9). */
int ga_build_paired(const ga_reads* reads, int k, const void* solid_dev, uint64_t solid_capacity,
                    void* query_table_dev, uint64_t query_capacity, void* qedge_table_dev,
                    uint64_t qedge_capacity, uint64_t* dh_dev, uint32_t* status_dev,
                    ga_stream stream);

/* Paired build across GPUs (each rank runs ga_build_paired over its read shard against the same replicated solid
 * table; stamps carry global read indices).  ga_stamp_table_export lists the occupied slots of a query table
 * (query_table_dev == NULL: key_out = idA << 32 | idB) or of a query-edge table (query_table_dev = the rank's own
 * query table: key_out / key2_out = the keys of the two queries the edge joins -- slot numbers do not travel);
 * *n_out_dev (zeroed by the caller) may exceed out_capacity, nothing is written beyond it.  ga_paired_merge folds
 * such lists (of any number of ranks) into one pair of 0xFF-filled tables with min(stamp): first every query, then
 * the edges re-keyed by the merged table's slots.  GA_ST_STAMP_FULL in status_dev: a table was too small. */
int ga_stamp_table_export(const void* table_dev, uint64_t capacity, const void* query_table_dev,
                          uint64_t* key_out_dev, uint64_t* key2_out_dev, uint64_t* stamp_out_dev,
                          uint64_t out_capacity, uint64_t* n_out_dev, ga_stream stream);
int ga_paired_merge(const uint64_t* query_keys_dev, const uint64_t* query_stamps_dev, uint64_t n_queries,
                    const uint64_t* edge_pkeys_dev, const uint64_t* edge_skeys_dev, const uint64_t* edge_stamps_dev,
                    uint64_t n_edges, void* query_table_dev, uint64_t query_capacity, void* qedge_table_dev,
                    uint64_t qedge_capacity, uint32_t* status_dev, ga_stream stream);

/* ---- CSR emission in the reference's insertion order --------------------- */
typedef struct ga_csr_plan ga_csr_plan;
/* Both plan calls synchronise `stream` and report the graph size; ga_csr_emit then fills
 * caller-allocated arrays:
 *   rowptr[n_nodes+1], col[n_edges] (successor ids, each row in edge-insertion order),
 *   indeg[n_nodes], branching[n_nodes], last_sym[n_nodes] (code of the node's last symbol),
 *   node_keys_a[n_nodes*key_words] (+ node_keys_b when paired) packed window keys.
 * num_edges_attr = the reference's graph.num_edges (n_edges plus orphaned self-loops). */
int ga_csr_plan_unpaired(const uint64_t* node_stamp_dev, uint64_t n_solid, const void* solid_keys_dev,
                         int key_words, int sym_bits, const void* edge_table_dev,
                         uint64_t edge_capacity, ga_stream stream, ga_csr_plan** plan_out,
                         int64_t* n_nodes, int64_t* n_edges);
int ga_csr_plan_unpaired_dna(const uint64_t* node_stamp_dev, const uint64_t* edge_stamp_dev,
                             uint64_t n_solid, const void* solid_keys_dev, int key_words, int k,
                             int sym_bits, const void* solid_dev, uint64_t solid_capacity,
                             ga_stream stream, ga_csr_plan** plan_out, int64_t* n_nodes,
                             int64_t* n_edges);
int ga_csr_plan_paired(const void* solid_dev, uint64_t solid_capacity, const void* solid_keys_dev,
                       uint64_t n_solid, int key_words, int k, int sym_bits,
                       const void* query_table_dev, uint64_t query_capacity,
                       const void* qedge_table_dev, uint64_t qedge_capacity, const uint64_t* dh_dev,
                       ga_stream stream, ga_csr_plan** plan_out, int64_t* n_nodes, int64_t* n_edges,
                       int64_t* num_edges_attr);
int ga_csr_emit(ga_csr_plan* plan, int32_t* rowptr_dev, int32_t* col_dev, int32_t* indeg_dev,
                uint8_t* branching_dev, uint8_t* last_sym_dev, void* node_keys_a_dev,
                void* node_keys_b_dev, ga_stream stream);
void ga_csr_plan_free(ga_csr_plan* plan);

/* ---- multi-GPU record exchange over NVLink peer memory ("k-mers routed to their owning GPU by a
 *      hash-partition all-to-all over NVLink"; the reference is single-process, debruijn_graph.py:113-152 is what
 *      every rank's share must add up to) -------------------------------------------------------------------- */
#define GA_PEER_HANDLE_BYTES 64
/* A receive buffer other processes of this node can map: cudaMalloc + CUDA IPC handle (64 opaque HOST bytes
 * written to handle_out; ship them to the peers by any means).  ga_peer_open maps a peer's buffer into this
 * process (peer access over NVLink is enabled on demand); ga_peer_close unmaps it; ga_peer_free releases the
 * owner's allocation once every peer has closed it.  All four synchronise the device. */
int ga_peer_alloc(uint64_t bytes, void** ptr_out, void* handle_out);
int ga_peer_open(const void* handle, void** ptr_out);
int ga_peer_close(void* ptr);
int ga_peer_free(void* ptr);
/* Sort + send in one kernel.  Input: the index form of ga_sk_scatter_buckets (records_dev = 32-byte slots of
 * ga_sk_scatter_reads, index_dev, offsets_dev over all 2^(l1_bits+l2_bits) buckets).  cut[world+1] (HOST):
 * bucket-sorted positions [cut[g], cut[g+1]) belong to rank g (cut[g] = offsets[first bucket of rank g]).
 * dst_bases[g] / dst_meta[g] (HOST arrays of device pointers, local or ga_peer_open'ed): where this rank's
 * segment starts in rank g's receive arrays (16 bytes of bases and one meta word per record, the dense
 * form ga_sk_count_build reads with one segment per source rank).  The caller orders the kernel after every
 * peer has finished reading its buffer, and the peers' reads after this kernel, with stream-ordered
 * collectives (a barrier on either side). */
int ga_sk_push_records(const void* records_dev, uint64_t l1_capacity, const uint32_t* index_dev,
                       const uint64_t* offsets_dev, int l1_bits, int l2_bits, uint32_t world,
                       const uint64_t* cut, void* const* dst_bases, void* const* dst_meta, ga_stream stream);
/* Level-2 split + send in one pass, no index: input is what ga_sk_scatter_reads + ga_sk_offsets leave behind
 * (records_dev, l1_cursors_dev, and cursors_dev = the exact-offset cursors, consumed).  A chunk of consecutive
 * slots is counting-sorted by final bucket in shared memory and its records are stored straight into the owners'
 * arrays in runs; every slot is streamed from DRAM once.  Same cut / dst_* contract and the same receive layout as
 * ga_sk_push_records (the order of the records inside a final bucket differs; nothing depends on it). */
int ga_sk_push_sorted(const void* records_dev, uint64_t l1_capacity, const uint64_t* l1_cursors_dev, int l1_bits,
                      int l2_bits, uint64_t* cursors_dev, uint32_t world, const uint64_t* cut,
                      void* const* dst_bases, void* const* dst_meta, ga_stream stream);

/* ---- raw ingest (replaces IOHandler.read_input, assemble.py:40-71) --------------------------------------- */
/* HOST pointers.  Parses the bytes of stdin with the reference's rules (first line = number of reads n;
 * max(n, 1) read lines, each stripped; "read" or "read1|read2|distance", the kind decided by the first read
 * line; missing lines are empty reads, or GA_ERR_BAD_ARG for pairs; trailing lines ignored) into one buffer
 * of symbols (mates of a pair back to back) and one length per read / mate -- no per-read objects.
 * symbols_out == NULL: sizing call, only *n_reads_out, *paired_out and *n_symbols_out (an upper bound) are
 * set.  GA_ERR_ALPHABET (parsing call): the input is not plain ASCII, or the distance field of the last pair line
 * is not a plain decimal integer -- the caller parses the input as text instead, so that Python's own rules decide
 * (int("1_0") is 10, int("x") raises, exactly as upstream).
 * Inputs of several MB whose only line break is "\n" are parsed by up to 16 host threads (GA_PARSE_THREADS), one
 * byte range each; the result is that of the serial scan. */
int ga_parse_reads(const uint8_t* text, uint64_t n_bytes, uint8_t* symbols_out, int32_t* lens_out,
                   uint64_t lens_capacity, uint64_t* n_reads_out, int* paired_out, int64_t* distance_out,
                   uint64_t* n_symbols_out);

/* ---- host-side contig traversal over the CSR (debruijn_graph.py:72-111, 222-267) ----------- */
/* Arrays are HOST pointers.  last_char: the byte each node contributes.  Allocates *text_out
 * (all contigs back to back) and *offsets_out[n_contigs+1]; release both with ga_free_host.
 * left_out (optional, [n_nodes]): edges remaining per node afterwards (each row is consumed
 * from its end, as dict.popitem does). */
int ga_traverse_contigs(const int32_t* rowptr, const int32_t* col, const int32_t* indeg,
                        const uint8_t* branching, const uint8_t* last_char, int64_t n_nodes,
                        int64_t num_edges_attr, int paired, uint8_t** text_out,
                        uint64_t** offsets_out, uint64_t* n_contigs, int32_t* left_out);
void ga_free_host(void* p);
/* Which routine answered this thread's last ga_traverse_contigs call: 1 = chains walked piecewise on
 * several host threads (graphs from 65536 nodes on whose edges all go in the reference's first sweep),
 * 0 = the serial edge-by-edge sweep.  Same output either way.  GA_TRAVERSE_THREADS=0 forces the serial one. */
int ga_traverse_last_route(void);

#ifdef __cplusplus
}
#endif
#endif /* GA_B200_H */
