/*
 * c_oracle.c -- CPU oracle (plain C) for the k-mer counting / CountMinSketch / de Bruijn
 * graph build path of tonycheang/genome-assembler.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_build/libga_oracle.so and
 * loaded (ctypes, oracle/c_oracle.py) by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  The product (genome-assembler_b200/) never links or loads it.
 *
 * It restates the reference's *sequential* algorithm on byte strings (any alphabet), so
 * that inputs of the size of BASELINE configs C2/C3 (10^8 occurrences) finish in seconds:
 *   - windows / counting         debruijn_graph.py:144-157 (unpaired), :349-374 (paired)
 *   - MurmurHash3_x86_32, sketch countminsketch.py:26-95
 *   - unpaired build             debruijn_graph.py:113-142
 *   - paired build + fuzzy key   debruijn_graph.py:269-347
 *   - contig traversal           debruijn_graph.py:72-111, :222-267
 * Parity pin: tests/test_oracle.py checks it against tests/golden/golden.json (outputs of
 * the unmodified reference) and against oracle/py_oracle.py; tests/test_reference_differential.py runs it
 * next to the live reference (oracle/_ref/, where present) on fresh random inputs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

static const u32 PRIMES_1_10_7[20] = {
    9999889, 9999901, 9999907, 9999929, 9999931, 9999937, 9999943, 9999971, 9999973, 9999991,
    10000019, 10000079, 10000103, 10000121, 10000139, 10000141, 10000169, 10000189, 10000223, 10000229};

#define NONE 0xFFFFFFFFu

/* ------------------------------------------------------------------ murmur (countminsketch.py:46-95) */
static u32 rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }

u32 orc_murmur3_32(const u8 *p, u64 n, u32 seed) {
    u32 h = seed;
    u64 body = n & ~(u64)3;
    for (u64 i = 0; i < body; i += 4) {
        u32 b = (u32)p[i] | ((u32)p[i + 1] << 8) | ((u32)p[i + 2] << 16) | ((u32)p[i + 3] << 24);
        b *= 0xcc9e2d51u; b = rotl32(b, 15); b *= 0x1b873593u;
        h ^= b; h = rotl32(h, 13); h = h * 5 + 0xe6546b64u;
    }
    u32 t = 0;
    switch (n & 3) {
    case 3: t |= (u32)p[body + 2] << 16; /* fallthrough */
    case 2: t |= (u32)p[body + 1] << 8;  /* fallthrough */
    case 1: t |= (u32)p[body];
        t *= 0xcc9e2d51u; t = rotl32(t, 15); t *= 0x1b873593u; h ^= t;
    }
    h ^= (u32)n;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

/* ------------------------------------------------------------------ result object */
typedef struct {
    int k, w, F, paired, sketch_rows, error; /* error: 1 = sketch cell overflow, 2 = alloc */
    /* distinct (k-1)-mers in first-occurrence order */
    u64 n_entries, cap_entries;
    const u8 **entry_ptr; u32 *entry_count; u8 *entry_solid;
    u64 map_cap; u32 *map; /* open addressing: entry index or NONE */
    u16 *sketch[20];
    /* graph */
    u64 n_nodes, n_edges_attr, n_csr_edges;
    int64_t *rowptr; int32_t *col; int32_t *indeg; u8 *branching; u8 *lastchar;
    u32 *node_a, *node_b; /* entry ids of each node's strings (node_b unused when unpaired) */
    /* contigs */
    u64 n_contigs, contig_bytes; u8 *contig_text; u64 *contig_off;
} orc_result;

static u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static u64 hash_bytes(const u8 *p, int w) {
    u64 h = 0x9e3779b97f4a7c15ull;
    for (int i = 0; i < w; i++) h = (h ^ p[i]) * 0x100000001b3ull;
    return mix64(h);
}

static int grow_entries(orc_result *R) {
    u64 cap = R->cap_entries ? R->cap_entries * 2 : 1024;
    R->entry_ptr = (const u8 **)realloc(R->entry_ptr, cap * sizeof(*R->entry_ptr));
    R->entry_count = (u32 *)realloc(R->entry_count, cap * sizeof(u32));
    if (!R->entry_ptr || !R->entry_count) return -1;
    R->cap_entries = cap;
    return 0;
}

static int rehash(orc_result *R) {
    u64 cap = R->map_cap ? R->map_cap * 2 : 4096;
    u32 *m = (u32 *)malloc(cap * sizeof(u32));
    if (!m) return -1;
    memset(m, 0xFF, cap * sizeof(u32));
    for (u64 e = 0; e < R->n_entries; e++) {
        u64 s = hash_bytes(R->entry_ptr[e], R->w) & (cap - 1);
        while (m[s] != NONE) s = (s + 1) & (cap - 1);
        m[s] = (u32)e;
    }
    free(R->map);
    R->map = m; R->map_cap = cap;
    return 0;
}

/* entry id of window p (length w); insert when absent and `insert` is set */
static u32 entry_of(orc_result *R, const u8 *p, int insert) {
    if (insert && (R->n_entries + 1) * 2 > R->map_cap && rehash(R)) { R->error = 2; return NONE; }
    u64 s = hash_bytes(p, R->w) & (R->map_cap - 1);
    for (;;) {
        u32 e = R->map[s];
        if (e == NONE) break;
        if (memcmp(R->entry_ptr[e], p, (size_t)R->w) == 0) return e;
        s = (s + 1) & (R->map_cap - 1);
    }
    if (!insert) return NONE;
    if (R->n_entries == R->cap_entries && grow_entries(R)) { R->error = 2; return NONE; }
    u32 e = (u32)R->n_entries++;
    R->entry_ptr[e] = p; R->entry_count[e] = 0;
    R->map[s] = e;
    return e;
}

/* ------------------------------------------------------------------ growable int vectors */
typedef struct { u32 *v; u64 n, cap; } vec32;
static int push32(vec32 *a, u32 x) {
    if (a->n == a->cap) {
        u64 c = a->cap ? a->cap * 2 : 1024;
        u32 *nv = (u32 *)realloc(a->v, c * sizeof(u32));
        if (!nv) return -1;
        a->v = nv; a->cap = c;
    }
    a->v[a->n++] = x;
    return 0;
}

/* fuzzy second-key rule (debruijn_graph.py:336-347): a suffix of `text` of >= 3 symbols
 * equals the prefix of `pattern` of that length */
static int overlap(const u8 *pattern, const u8 *text, int w) {
    for (int s = 0; s < w - 2; s++)
        if (memcmp(text + s, pattern, (size_t)(w - s)) == 0) return w - s;
    return 0;
}

/* ------------------------------------------------------------------ builds */
typedef struct {
    vec32 a, b;            /* node -> entry ids */
    vec32 indeg;
    vec32 head, tail;      /* per node: first / last out-edge record */
    vec32 e_next, e_ta, e_tb; /* edge records: next, target key (entry a, entry b) */
    u64 n_edges_attr;
} builder;

static u32 new_node(builder *B, u32 a, u32 b) {
    u32 id = (u32)B->a.n;
    push32(&B->a, a); push32(&B->b, b); push32(&B->indeg, 0);
    push32(&B->head, NONE); push32(&B->tail, NONE);
    return id;
}
static int has_edge(builder *B, u32 node, u32 ta, u32 tb) {
    for (u32 e = B->head.v[node]; e != NONE; e = B->e_next.v[e])
        if (B->e_ta.v[e] == ta && B->e_tb.v[e] == tb) return 1;
    return 0;
}
static void add_edge(builder *B, u32 node, u32 ta, u32 tb) {
    u32 e = (u32)B->e_next.n;
    push32(&B->e_next, NONE); push32(&B->e_ta, ta); push32(&B->e_tb, tb);
    if (B->head.v[node] == NONE) B->head.v[node] = e; else B->e_next.v[B->tail.v[node]] = e;
    B->tail.v[node] = e;
}

static void finish_graph(orc_result *R, builder *B, const u32 *order, u64 n_nodes,
                         const u32 *node_of_key_a /*unpaired: entry->node*/,
                         u32 (*resolve)(void *, u32, u32), void *ctx) {
    /* order[i] = builder node id of the i-th node in iteration order */
    u32 *rank = (u32 *)malloc((B->a.n + 1) * sizeof(u32));
    for (u64 i = 0; i < B->a.n; i++) rank[i] = NONE;
    for (u64 i = 0; i < n_nodes; i++) rank[order[i]] = (u32)i;
    R->n_nodes = n_nodes;
    R->rowptr = (int64_t *)calloc(n_nodes + 1, sizeof(int64_t));
    R->indeg = (int32_t *)calloc(n_nodes + 1, sizeof(int32_t));
    R->branching = (u8 *)calloc(n_nodes + 1, 1);
    R->lastchar = (u8 *)calloc(n_nodes + 1, 1);
    R->node_a = (u32 *)calloc(n_nodes + 1, sizeof(u32));
    R->node_b = (u32 *)calloc(n_nodes + 1, sizeof(u32));
    u64 m = 0;
    for (u64 i = 0; i < n_nodes; i++)
        for (u32 e = B->head.v[order[i]]; e != NONE; e = B->e_next.v[e]) m++;
    R->col = (int32_t *)calloc(m + 1, sizeof(int32_t));
    m = 0;
    for (u64 i = 0; i < n_nodes; i++) {
        u32 nd = order[i];
        R->rowptr[i] = (int64_t)m;
        for (u32 e = B->head.v[nd]; e != NONE; e = B->e_next.v[e]) {
            u32 tgt = resolve ? resolve(ctx, B->e_ta.v[e], B->e_tb.v[e]) : node_of_key_a[B->e_ta.v[e]];
            R->col[m++] = (int32_t)rank[tgt];
        }
        R->indeg[i] = (int32_t)B->indeg.v[nd];
        R->node_a[i] = B->a.v[nd]; R->node_b[i] = B->b.v[nd];
        R->lastchar[i] = R->entry_ptr[B->a.v[nd]][R->w - 1];
    }
    R->rowptr[n_nodes] = (int64_t)m;
    R->n_csr_edges = m;
    for (u64 i = 0; i < n_nodes; i++)
        R->branching[i] = (R->rowptr[i + 1] - R->rowptr[i] > 1) || R->indeg[i] > 1;
    R->n_edges_attr = B->n_edges_attr;
    free(rank);
}

static void build_unpaired(orc_result *R, const u8 *m1, const u64 *off1, u64 n_reads) {
    builder B; memset(&B, 0, sizeof(B));
    int w = R->w;
    u32 *node_of = (u32 *)malloc((R->n_entries + 1) * sizeof(u32));
    for (u64 i = 0; i < R->n_entries; i++) node_of[i] = NONE;
    for (u64 r = 0; r < n_reads; r++) {
        const u8 *p = m1 + off1[r];
        int64_t len = (int64_t)(off1[r + 1] - off1[r]);
        u32 prev = NONE;
        for (int64_t i = 0; i + w <= len; i++) {
            u32 cur = entry_of(R, p + i, 0);
            if (i > 0 && R->entry_solid[prev] && R->entry_solid[cur]) {
                u32 np = node_of[prev], ns = node_of[cur];
                if (!(np != NONE && ns != NONE && has_edge(&B, np, cur, 0))) {
                    if (np == NONE) np = node_of[prev] = new_node(&B, prev, 0);
                    ns = node_of[cur];             /* prev == cur (homopolymer) */
                    if (ns == NONE) ns = node_of[cur] = new_node(&B, cur, 0);
                    add_edge(&B, np, cur, 0);
                    B.indeg.v[ns]++;
                    B.n_edges_attr++;
                }
            }
            prev = cur;
        }
    }
    u64 n = B.a.n;
    u32 *order = (u32 *)malloc((n + 1) * sizeof(u32));
    for (u64 i = 0; i < n; i++) order[i] = (u32)i;
    finish_graph(R, &B, order, n, node_of, NULL, NULL);
    free(order); free(node_of);
    free(B.a.v); free(B.b.v); free(B.indeg.v); free(B.head.v); free(B.tail.v);
    free(B.e_next.v); free(B.e_ta.v); free(B.e_tb.v);
}

/* paired: groups keyed by entry A; members in insertion order */
typedef struct {
    orc_result *R;
    u32 *g_head, *g_tail;       /* per entry A: first/last member record */
    vec32 m_next, m_b, m_node;  /* member records */
    vec32 g_order;              /* entry A ids in group-creation order */
} groups;

static u32 group_exact(groups *G, u32 a, u32 b) {
    for (u32 m = G->g_head[a]; m != NONE; m = G->m_next.v[m])
        if (G->m_b.v[m] == b) return m;
    return NONE;
}
static u32 group_lookup(groups *G, u32 a, u32 b) { /* -> node id or NONE (debruijn_graph.py:319-334) */
    if (G->g_head[a] == NONE) return NONE;
    u32 m = group_exact(G, a, b);
    if (m != NONE) return G->m_node.v[m];
    const u8 *bs = G->R->entry_ptr[b];
    int w = G->R->w;
    for (m = G->g_head[a]; m != NONE; m = G->m_next.v[m]) {
        const u8 *ks = G->R->entry_ptr[G->m_b.v[m]];
        if (overlap(ks, bs, w) || overlap(bs, ks, w)) return G->m_node.v[m];
    }
    return NONE;
}
static void group_store(groups *G, u32 a, u32 b, u32 node) {
    u32 m = group_exact(G, a, b);
    if (m != NONE) { G->m_node.v[m] = node; return; }   /* overwrite keeps the key's position */
    m = (u32)G->m_next.n;
    push32(&G->m_next, NONE); push32(&G->m_b, b); push32(&G->m_node, node);
    if (G->g_head[a] == NONE) { G->g_head[a] = m; push32(&G->g_order, a); }
    else G->m_next.v[G->g_tail[a]] = m;
    G->g_tail[a] = m;
}
static u32 resolve_pair(void *ctx, u32 a, u32 b) {
    groups *G = (groups *)ctx;
    return G->m_node.v[group_exact(G, a, b)];
}

static void build_paired(orc_result *R, const u8 *m1, const u64 *off1, const u8 *m2, const u64 *off2,
                         u64 n_reads) {
    builder B; memset(&B, 0, sizeof(B));
    groups G; memset(&G, 0, sizeof(G));
    int w = R->w;
    G.R = R;
    G.g_head = (u32 *)malloc((R->n_entries + 1) * sizeof(u32));
    G.g_tail = (u32 *)malloc((R->n_entries + 1) * sizeof(u32));
    for (u64 i = 0; i < R->n_entries; i++) G.g_head[i] = G.g_tail[i] = NONE;
    for (u64 r = 0; r < n_reads; r++) {
        const u8 *p = m1 + off1[r], *q = m2 + off2[r];
        int64_t len = (int64_t)(off1[r + 1] - off1[r]);
        u32 pa = NONE, pb = NONE;
        for (int64_t i = 0; i + w <= len; i++) {
            u32 ca = entry_of(R, p + i, 0), cb = entry_of(R, q + i, 0);
            if (i > 0 && R->entry_solid[pa] && R->entry_solid[ca] && R->entry_solid[pb] && R->entry_solid[cb]) {
                u32 src = group_lookup(&G, pa, pb), dst = group_lookup(&G, ca, cb);
                int both = (src != NONE && dst != NONE);
                if (src == NONE) { src = new_node(&B, pa, pb); group_store(&G, pa, pb, src); }
                if (dst == NONE) { dst = new_node(&B, ca, cb); group_store(&G, ca, cb, dst); }
                u32 ta = B.a.v[dst], tb = B.b.v[dst];
                if (!(both && has_edge(&B, src, ta, tb))) {
                    if (!has_edge(&B, src, ta, tb)) add_edge(&B, src, ta, tb); /* dict key set */
                    B.indeg.v[dst]++;
                    B.n_edges_attr++;
                }
            }
            pa = ca; pb = cb;
        }
    }
    vec32 order; memset(&order, 0, sizeof(order));
    for (u64 g = 0; g < G.g_order.n; g++)
        for (u32 m = G.g_head[G.g_order.v[g]]; m != NONE; m = G.m_next.v[m]) push32(&order, G.m_node.v[m]);
    finish_graph(R, &B, order.v, order.n, NULL, resolve_pair, &G);
    free(order.v); free(G.g_head); free(G.g_tail); free(G.m_next.v); free(G.m_b.v); free(G.m_node.v);
    free(G.g_order.v);
    free(B.a.v); free(B.b.v); free(B.indeg.v); free(B.head.v); free(B.tail.v);
    free(B.e_next.v); free(B.e_ta.v); free(B.e_tb.v);
}

/* ------------------------------------------------------------------ traversal (debruijn_graph.py:72-111, 222-267) */
typedef struct { u8 *v; u64 n, cap; } vec8;
static void push8(vec8 *a, u8 x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 4096; a->v = (u8 *)realloc(a->v, a->cap); }
    a->v[a->n++] = x;
}
typedef struct { u64 *v; u64 n, cap; } vec64;
static void push64(vec64 *a, u64 x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 256; a->v = (u64 *)realloc(a->v, a->cap * 8); }
    a->v[a->n++] = x;
}

static void traverse(orc_result *R) {
    u64 n = R->n_nodes;
    int64_t *left = (int64_t *)malloc((n + 1) * sizeof(int64_t)); /* remaining out-edges per node */
    for (u64 i = 0; i < n; i++) left[i] = R->rowptr[i + 1] - R->rowptr[i];
    int64_t remaining = (int64_t)R->n_edges_attr;
    vec8 text; memset(&text, 0, sizeof(text));
    vec64 offs; memset(&offs, 0, sizeof(offs));
    push64(&offs, 0);
#define WALK(i) do { \
        u64 j = (u64)R->col[R->rowptr[i] + --left[i]]; remaining--; \
        push8(&text, R->lastchar[j]); \
        while (left[j] > 0 && !R->branching[j]) { \
            u64 nx = (u64)R->col[R->rowptr[j] + --left[j]]; remaining--; \
            push8(&text, R->lastchar[nx]); j = nx; } \
        push64(&offs, text.n); } while (0)
    int done = 0;
    for (u64 i = 0; i < n && !done; i++) {
        while (left[i] > 0 && (R->branching[i] || R->indeg[i] == 0)) WALK(i);
        if (remaining == 0) done = 1;
    }
    if (!done && n) {
        if (R->paired) {
            for (u64 i = 0; i < n && !done; i++) {
                while (left[i] > 0) WALK(i);
                if (remaining == 0) done = 1;
            }
        } else {
            u64 last = n - 1;
            while (left[last] > 0) WALK(last);
        }
    }
#undef WALK
    R->n_contigs = offs.n - 1; R->contig_bytes = text.n;
    R->contig_text = text.v; R->contig_off = offs.v;
    free(left);
}

/* ------------------------------------------------------------------ entry point */
orc_result *orc_assemble(const u8 *m1, const u64 *off1, const u8 *m2, const u64 *off2, u64 n_reads,
                         int k, int F, int sketch_rows, int want_graph) {
    orc_result *R = (orc_result *)calloc(1, sizeof(orc_result));
    if (!R) return NULL;
    R->k = k; R->w = k - 1; R->F = F; R->paired = (m2 != NULL); R->sketch_rows = sketch_rows;
    int w = R->w;
    if (rehash(R)) { R->error = 2; return R; }
    /* count (debruijn_graph.py:144-152 / 349-367): mate-2 windows share mate-1's index range */
    for (u64 r = 0; r < n_reads; r++) {
        const u8 *p = m1 + off1[r];
        int64_t len = (int64_t)(off1[r + 1] - off1[r]);
        for (int64_t i = 0; i + w <= len; i++) {
            u32 e = entry_of(R, p + i, 1);
            if (e == NONE) return R;
            R->entry_count[e]++;
            if (m2) {
                e = entry_of(R, m2 + off2[r] + i, 1);
                if (e == NONE) return R;
                R->entry_count[e]++;
            }
        }
    }
    R->entry_solid = (u8 *)calloc(R->n_entries + 1, 1);
    if (sketch_rows > 0) {
        /* countminsketch.py:34-44 with array('H') overflow semantics */
        u32 *hv = (u32 *)malloc((R->n_entries + 1) * sizeof(u32));
        for (int d = 0; d < sketch_rows; d++) R->sketch[d] = (u16 *)calloc(PRIMES_1_10_7[d], sizeof(u16));
        for (u64 e = 0; e < R->n_entries; e++) {
            u32 h = hv[e] = orc_murmur3_32(R->entry_ptr[e], (u64)w, 0);
            for (int d = 0; d < sketch_rows; d++) {
                u32 cell = (u32)R->sketch[d][h % PRIMES_1_10_7[d]] + R->entry_count[e];
                if (cell > 65535u) { R->error = 1; free(hv); return R; }
                R->sketch[d][h % PRIMES_1_10_7[d]] = (u16)cell;
            }
        }
        for (u64 e = 0; e < R->n_entries; e++) {
            u32 est = 0xFFFFFFFFu;
            for (int d = 0; d < sketch_rows; d++) {
                u32 c = R->sketch[d][hv[e] % PRIMES_1_10_7[d]];
                if (c < est) est = c;
            }
            R->entry_solid[e] = (int64_t)est > (int64_t)F;
        }
        free(hv);
    } else {
        for (u64 e = 0; e < R->n_entries; e++) R->entry_solid[e] = (int64_t)R->entry_count[e] > (int64_t)F;
    }
    if (!want_graph) return R;
    if (m2) build_paired(R, m1, off1, m2, off2, n_reads);
    else build_unpaired(R, m1, off1, n_reads);
    traverse(R);
    return R;
}

int orc_error(const orc_result *R) { return R->error; }
u64 orc_n_distinct(const orc_result *R) { return R->n_entries; }
u64 orc_n_nodes(const orc_result *R) { return R->n_nodes; }
u64 orc_num_edges(const orc_result *R) { return R->n_edges_attr; }
u64 orc_n_csr_edges(const orc_result *R) { return R->n_csr_edges; }
u64 orc_n_contigs(const orc_result *R) { return R->n_contigs; }
u64 orc_contig_bytes(const orc_result *R) { return R->contig_bytes; }

void orc_copy_counts(const orc_result *R, u8 *keys, u32 *counts) {
    for (u64 e = 0; e < R->n_entries; e++) {
        memcpy(keys + e * (u64)R->w, R->entry_ptr[e], (size_t)R->w);
        counts[e] = R->entry_count[e];
    }
}
void orc_copy_csr(const orc_result *R, int64_t *rowptr, int32_t *col, int32_t *indeg, u8 *branching,
                  u8 *lastchar) {
    memcpy(rowptr, R->rowptr, (R->n_nodes + 1) * sizeof(int64_t));
    memcpy(col, R->col, R->n_csr_edges * sizeof(int32_t));
    memcpy(indeg, R->indeg, R->n_nodes * sizeof(int32_t));
    memcpy(branching, R->branching, R->n_nodes);
    memcpy(lastchar, R->lastchar, R->n_nodes);
}
void orc_copy_node_keys(const orc_result *R, u8 *a, u8 *b) {
    for (u64 i = 0; i < R->n_nodes; i++) {
        memcpy(a + i * (u64)R->w, R->entry_ptr[R->node_a[i]], (size_t)R->w);
        if (b && R->paired) memcpy(b + i * (u64)R->w, R->entry_ptr[R->node_b[i]], (size_t)R->w);
    }
}
void orc_copy_sketch_row(const orc_result *R, int row, u16 *out) {
    memcpy(out, R->sketch[row], (size_t)PRIMES_1_10_7[row] * sizeof(u16));
}
void orc_copy_contigs(const orc_result *R, u8 *text, u64 *offsets) {
    memcpy(text, R->contig_text, R->contig_bytes);
    memcpy(offsets, R->contig_off, (R->n_contigs + 1) * sizeof(u64));
}
void orc_free(orc_result *R) {
    if (!R) return;
    free(R->entry_ptr); free(R->entry_count); free(R->entry_solid); free(R->map);
    for (int d = 0; d < 20; d++) free(R->sketch[d]);
    free(R->rowptr); free(R->col); free(R->indeg); free(R->branching); free(R->lastchar);
    free(R->node_a); free(R->node_b); free(R->contig_text); free(R->contig_off);
    free(R);
}
