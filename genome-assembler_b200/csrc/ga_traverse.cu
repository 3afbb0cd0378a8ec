// ga_traverse.cu -- contig traversal over the CSR.  It stays on the host by design (north star item 4;
// debruijn_graph.py:72-111 unpaired, :222-267 paired; Node.pop_edge = dict.popitem, debruijn_node.py:24-26).
//
// Two routines with the same result:
//  * traverse_serial: the reference's two sweeps, edge by edge, in one thread.  Always right; used for small
//    graphs and whenever the parallel routine declines.
//  * traverse_parallel: the first sweep only.  Every contig of that sweep starts with an edge of a node that
//    is branching or has in-degree 0 and then runs along nodes that are not branching (out <= 1, in <= 1) until
//    a branching node or a dead end, so the contigs do not share non-branching nodes and may be walked in
//    any order.  A walk is one dependent cache miss per node; to have many of them in flight the chains are
//    also cut at pseudo-random "splitter" nodes (one node in 512), every piece is walked on its own -- several
//    host threads, several pieces per thread, the next record prefetched while the other pieces move -- and
//    the pieces are stitched back together in the reference's order (nodes in self.nodes order, a node's edges
//    from the last one backwards).  Every node a walk pops an edge from is marked; meeting a marked node, a
//    non-branching node with two edges, an edge count that differs from the reference's own counter, or edges
//    left over after the first sweep (cycles without a branching node: the reference's second sweep) makes the
//    routine decline, and the serial one runs instead.
#include <atomic>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <sys/mman.h>

#include "ga_common.cuh"

namespace {

// Tables that are walked at random: 2 MB pages where the kernel grants them (one TLB entry per 2 MB instead of
// per 4 KB, and 512 times fewer first-touch faults).  Released with free().
void* alloc_big(size_t bytes) {
    const size_t huge = (size_t)2 << 20;
    if (bytes < 2 * huge) return malloc(bytes ? bytes : 1);
    void* p = aligned_alloc(huge, (bytes + huge - 1) / huge * huge);
#ifdef MADV_HUGEPAGE
    if (p) madvise(p, (bytes + huge - 1) / huge * huge, MADV_HUGEPAGE);
#endif
    return p;
}

struct CsrView {
    const int32_t* rowptr;
    const int32_t* col;
    const int32_t* indeg;
    const uint8_t* branching;
    const uint8_t* last_char;
    int64_t n_nodes;
    int64_t num_edges_attr;
    int paired;
};

// Walk = pop the last remaining edge of the start node, then keep going while the current node
// still has edges and was not branching; every step contributes the successor's last symbol.
void traverse_serial(const CsrView& g, std::vector<uint8_t>& text, std::vector<uint64_t>& offs,
                     std::vector<int32_t>& left) {
    const int32_t* rowptr = g.rowptr;
    const int32_t* col = g.col;
    const uint8_t* branching = g.branching;
    const uint8_t* last_char = g.last_char;
    const int64_t n_nodes = g.n_nodes;
    left.resize((size_t)n_nodes);
    for (int64_t i = 0; i < n_nodes; ++i) left[(size_t)i] = rowptr[i + 1] - rowptr[i];
    text.clear();
    offs.assign(1, 0);
    int64_t remaining = g.num_edges_attr;
    auto walk = [&](int64_t start) {
        int64_t j = col[rowptr[start] + --left[(size_t)start]];
        --remaining;
        text.push_back(last_char[j]);
        while (left[(size_t)j] > 0 && !branching[j]) {
            int64_t nx = col[rowptr[j] + --left[(size_t)j]];
            --remaining;
            text.push_back(last_char[nx]);
            j = nx;
        }
        offs.push_back(text.size());
    };
    bool done = false;
    for (int64_t i = 0; i < n_nodes && !done; ++i) {
        while (left[(size_t)i] > 0 && (branching[i] || g.indeg[i] == 0)) walk(i);
        if (remaining == 0) done = true;   // (:80-81, 231-232)
    }
    if (!done && n_nodes > 0) {
        if (g.paired) {
            for (int64_t i = 0; i < n_nodes && !done; ++i) {
                while (left[(size_t)i] > 0) walk(i);
                if (remaining == 0) done = true;
            }
        } else {
            // the reference's second loop keeps testing the last node of the first loop
            // (debruijn_graph.py:85-86): only that node's cycle is ever emitted
            int64_t last = n_nodes - 1;
            while (left[(size_t)last] > 0) walk(last);
        }
    }
}

// ---------------------------------------------------------------------------------------- parallel
struct Rec {             // what a walk needs to know about a node, in one 8-byte word
    int32_t next;        // target of the node's last edge (the one popitem takes first); -1 without edges
    uint8_t last;        // the symbol the node contributes
    uint8_t flags;
    uint8_t marked;      // an edge of this (non-branching) node has been popped
    uint8_t pad;
};
enum : uint8_t { R_BRANCH = 1, R_OUT = 2, R_SPLIT = 4 };

// The record table is kept between calls (grow-only, like the device workspace): first touch of fresh pages
// costs as much as the walk itself.  A call that finds it taken by another thread allocates its own.
struct RecScratch {
    std::atomic<int> taken{0};
    Rec* ptr = nullptr;
    size_t cap = 0;
} g_rec_scratch;

struct RecLease {
    Rec* ptr = nullptr;
    bool shared = false;
    explicit RecLease(size_t n) {
        if (g_rec_scratch.taken.exchange(1, std::memory_order_acquire) == 0) {
            shared = true;
            if (g_rec_scratch.cap < n) {
                free(g_rec_scratch.ptr);
                g_rec_scratch.cap = n + n / 8;
                g_rec_scratch.ptr = (Rec*)alloc_big(g_rec_scratch.cap * sizeof(Rec));
                if (!g_rec_scratch.ptr) g_rec_scratch.cap = 0;
            }
            ptr = g_rec_scratch.ptr;
        } else {
            ptr = (Rec*)alloc_big(n * sizeof(Rec));
        }
    }
    ~RecLease() {
        if (shared)
            g_rec_scratch.taken.store(0, std::memory_order_release);
        else
            free(ptr);
    }
    RecLease(const RecLease&) = delete;
    RecLease& operator=(const RecLease&) = delete;
};

struct Head {            // where a piece starts
    int32_t node;        // splitter: the node itself; contig start: the node that owns the edge
    int32_t edge;        // contig start: position in col; splitter: -1
};

struct Piece {           // a walked piece: bytes [off, off + len) of its thread's buffer
    uint32_t thread;
    uint32_t len;
    uint64_t off;
    int32_t cont;        // the splitter node it stopped at (the contig goes on with that node's piece); -1 = ends here
};

#ifndef GA_TRV_LANES
#define GA_TRV_LANES 8
#endif
#ifndef GA_TRV_STEPS
#define GA_TRV_STEPS 4
#endif
constexpr int kLanes = GA_TRV_LANES;            // pieces in flight per host thread
constexpr uint32_t kSplitShift = 23; // one node in 2^(32-23) = 512 is a splitter

inline bool is_splitter_id(uint32_t j) { return ((j * 0x9E3779B1u) >> kSplitShift) == 0; }

template <class F>
void run_threads(unsigned threads, F&& body) {
    if (threads <= 1) {
        body(0u);
        return;
    }
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    for (unsigned t = 1; t < threads; ++t) pool.emplace_back([&body, t] { body(t); });
    body(0u);
    for (auto& th : pool) th.join();
}

// true: text / offs hold every contig and every edge of the graph was consumed.  false: declined.
bool traverse_parallel(const CsrView& g, unsigned threads, uint8_t** text_out, std::vector<uint64_t>& offs) {
    const int64_t n = g.n_nodes;
    const int64_t n_edges = g.rowptr[n];
    if (n_edges != g.num_edges_attr || n_edges <= 0 || n >= (int64_t)0x7fffffff) return false;
    RecLease lease((size_t)n);
    Rec* rec = lease.ptr;
    if (!rec) return false;
    std::atomic<int> decline{0};

    // 1. records + heads, by node range (the heads of a range come out in node order)
    std::vector<std::vector<Head>> starts(threads), splits(threads);
    run_threads(threads, [&](unsigned t) {
        const int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
        std::vector<Head>& st = starts[t];
        std::vector<Head>& sp = splits[t];
        for (int64_t j = lo; j < hi; ++j) {
            const int32_t r0 = g.rowptr[j], r1 = g.rowptr[j + 1];
            const bool br = g.branching[j] != 0;
            Rec r;
            r.next = r1 > r0 ? g.col[r1 - 1] : -1;
            r.last = g.last_char[j];
            r.flags = (uint8_t)((br ? R_BRANCH : 0) | (r1 > r0 ? R_OUT : 0));
            r.marked = 0;
            r.pad = 0;
            if (r1 > r0) {
                if (br || g.indeg[j] == 0) {
                    if (!br && r1 - r0 > 1) decline.store(1, std::memory_order_relaxed);
                    for (int32_t e = r1 - 1; e >= r0; --e) st.push_back(Head{(int32_t)j, e});
                } else if (r1 - r0 > 1) {
                    decline.store(1, std::memory_order_relaxed);     // "not branching" with two edges: not this routine's case
                } else if (is_splitter_id((uint32_t)j)) {
                    r.flags |= R_SPLIT;
                    sp.push_back(Head{(int32_t)j, -1});
                }
            }
            rec[j] = r;
        }
    });
    if (decline.load()) return false;
    std::vector<Head> heads;
    size_t n_start = 0, n_split = 0;
    for (unsigned t = 0; t < threads; ++t) n_start += starts[t].size(), n_split += splits[t].size();
    heads.reserve(n_start + n_split);
    for (unsigned t = 0; t < threads; ++t) heads.insert(heads.end(), starts[t].begin(), starts[t].end());
    for (unsigned t = 0; t < threads; ++t) heads.insert(heads.end(), splits[t].begin(), splits[t].end());
    starts.clear();
    splits.clear();
    const size_t n_heads = heads.size();
    if (n_start == 0 || n_heads >= 0xffffffffull) return false;

    // 2. walk the pieces
    std::vector<Piece> pieces(n_heads);
    std::vector<std::vector<uint8_t>> bufs(threads);
    std::atomic<size_t> next_head{0};
    // A mark is a plain byte store (a locked exchange per step costs more than the walk itself); that no node was
    // popped twice is checked afterwards: as many marked nodes as marks made.
    std::vector<uint64_t> marks_made(threads, 0), marks_found(threads, 0);
    run_threads(threads, [&](unsigned t) {
        uint64_t made = 0;
        auto mark = [&](int32_t j) {
            __atomic_store_n(&rec[j].marked, (uint8_t)1, __ATOMIC_RELAXED);
            ++made;
        };
        struct Lane {
            size_t piece;
            int32_t at;                  // node whose record is looked at next (already asked for)
            bool busy;
            std::vector<uint8_t> text;
        };
        Lane lanes[kLanes];
        std::vector<uint8_t>& buf = bufs[t];
        buf.reserve((size_t)(n_edges / threads) + 4096);
        int busy = 0;
        bool drained = false;
        auto begin = [&](Lane& ln) -> bool {
            if (drained) return false;
            const size_t h = next_head.fetch_add(1, std::memory_order_relaxed);
            if (h >= n_heads) {
                drained = true;
                return false;
            }
            const Head hd = heads[h];
            ln.piece = h;
            ln.text.clear();
            if (hd.edge >= 0) {
                if (!(rec[hd.node].flags & R_BRANCH)) mark(hd.node);
                ln.at = g.col[hd.edge];
            } else {
                mark(hd.node);
                ln.at = rec[hd.node].next;
            }
            __builtin_prefetch(&rec[ln.at]);
            ln.busy = true;
            return true;
        };
        auto finish = [&](Lane& ln, int32_t cont) {
            Piece& p = pieces[ln.piece];
            p.thread = t;
            p.off = buf.size();
            p.len = (uint32_t)ln.text.size();
            p.cont = cont;
            buf.insert(buf.end(), ln.text.begin(), ln.text.end());
            ln.busy = false;
        };
        for (int l = 0; l < kLanes; ++l) {
            lanes[l].busy = false;
            if (begin(lanes[l])) ++busy;
        }
        while (busy > 0) {
            if (decline.load(std::memory_order_relaxed)) break;
            for (int l = 0; l < kLanes; ++l) {
                Lane& ln = lanes[l];
                if (!ln.busy) continue;
                // a few steps per visit: consecutive nodes of a chain often share a cache line
                for (int step = 0; step < GA_TRV_STEPS; ++step) {
                    const int32_t j = ln.at;
                    // field by field: `marked` of this record may be written by another thread right now (a piece
                    // that starts at this splitter), the other fields do not change during the walk
                    const int32_t next = rec[j].next;
                    const uint8_t flags = rec[j].flags;
                    ln.text.push_back(rec[j].last);
                    if ((flags & R_BRANCH) || !(flags & R_OUT)) {
                        finish(ln, -1);
                    } else if (flags & R_SPLIT) {
                        finish(ln, j);
                    } else {
                        mark(j);
                        ln.at = next;
                        __builtin_prefetch(&rec[next]);
                        if ((int64_t)ln.text.size() > n) decline.store(1, std::memory_order_relaxed);   // going round in circles
                        continue;
                    }
                    if (!begin(ln)) --busy;
                    break;
                }
            }
        }
        marks_made[t] = made;
    });
    if (!decline.load()) {
        run_threads(threads, [&](unsigned t) {
            const int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
            uint64_t found = 0;
            for (int64_t j = lo; j < hi; ++j) found += rec[j].marked;
            marks_found[t] = found;
        });
        uint64_t made = 0, found = 0;
        for (unsigned t = 0; t < threads; ++t) made += marks_made[t], found += marks_found[t];
        if (made != found) decline.store(1);          // some node was popped twice
    }
    if (decline.load()) return false;

    // 3. stitch: contig starts in order, each followed through the pieces of the splitters it meets
    const Head* split_heads = heads.data() + n_start;      // sorted by node
    auto piece_of_splitter = [&](int32_t node) -> size_t {
        const Head* it = std::lower_bound(split_heads, split_heads + n_split, node,
                                          [](const Head& h, int32_t v) { return h.node < v; });
        return n_start + (size_t)(it - split_heads);
    };
    struct Copy {
        uint64_t dst;
        size_t piece;
    };
    std::vector<Copy> copies;
    copies.reserve(n_heads);
    offs.assign(1, 0);
    offs.reserve(n_start + 1);
    uint64_t total = 0;
    bool ok = true;
    for (size_t s = 0; s < n_start && ok; ++s) {
        size_t p = s;
        for (;;) {
            copies.push_back(Copy{total, p});
            total += pieces[p].len;
            if (copies.size() > n_heads) {       // a piece used twice: cannot be, but never loop
                ok = false;
                break;
            }
            if (pieces[p].cont < 0) break;
            p = piece_of_splitter(pieces[p].cont);
        }
        offs.push_back(total);
    }
    if (!ok || total != (uint64_t)n_edges) return false;    // edges left: the second sweep is the serial routine's
    uint8_t* text = (uint8_t*)alloc_big(total ? total : 1);
    if (!text) return false;
    run_threads(threads, [&](unsigned t) {
        const size_t lo = copies.size() * t / threads, hi = copies.size() * (t + 1) / threads;
        for (size_t c = lo; c < hi; ++c) {
            const Piece& p = pieces[copies[c].piece];
            memcpy(text + copies[c].dst, bufs[p.thread].data() + p.off, p.len);
        }
    });
    *text_out = text;
    return true;
}

thread_local int g_last_route = 0;

}  // namespace

// 1 when this thread's last ga_traverse_contigs call was answered by the piecewise routine, 0 by the serial sweep
extern "C" int ga_traverse_last_route(void) { return g_last_route; }

extern "C" int ga_traverse_contigs(const int32_t* rowptr, const int32_t* col, const int32_t* indeg,
                                   const uint8_t* branching, const uint8_t* last_char, int64_t n_nodes,
                                   int64_t num_edges_attr, int paired, uint8_t** text_out,
                                   uint64_t** offsets_out, uint64_t* n_contigs, int32_t* left_out) {
    if (!text_out || !offsets_out || !n_contigs || n_nodes < 0 ||
        (n_nodes > 0 && (!rowptr || !indeg || !branching || !last_char))) {
        ga_set_error("ga_traverse_contigs: bad arguments");
        return GA_ERR_BAD_ARG;
    }
    const CsrView g{rowptr, col, indeg, branching, last_char, n_nodes, num_edges_attr, paired};
    unsigned threads = std::thread::hardware_concurrency();
    threads = threads < 1 ? 1 : (threads > 16 ? 16 : threads);
    int64_t min_nodes = 1 << 16;       // below this the serial sweep is done before the threads have started
    if (const char* e = getenv("GA_TRAVERSE_THREADS")) threads = (unsigned)(atoi(e) > 0 ? atoi(e) : 0);
    if (const char* e = getenv("GA_TRAVERSE_MIN_NODES")) min_nodes = atoll(e);
    std::vector<uint64_t> offs(1, 0);
    uint8_t* t = nullptr;
    if (threads >= 1 && n_nodes >= min_nodes && n_nodes > 0 && col && traverse_parallel(g, threads, &t, offs)) {
        if (left_out) memset(left_out, 0, (size_t)n_nodes * sizeof(int32_t));   // every edge was popped
        g_last_route = 1;
    } else {
        g_last_route = 0;
        std::vector<uint8_t> text;
        std::vector<int32_t> left;
        traverse_serial(g, text, offs, left);
        t = (uint8_t*)malloc(text.size() ? text.size() : 1);
        if (t && !text.empty()) memcpy(t, text.data(), text.size());
        if (t && left_out && n_nodes > 0) memcpy(left_out, left.data(), (size_t)n_nodes * sizeof(int32_t));
    }
    uint64_t* o = (uint64_t*)malloc(offs.size() * sizeof(uint64_t));
    if (!t || !o) {
        free(t);
        free(o);
        ga_set_error("ga_traverse_contigs: out of host memory");
        return GA_ERR_BAD_ARG;
    }
    memcpy(o, offs.data(), offs.size() * sizeof(uint64_t));
    *text_out = t;
    *offsets_out = o;
    *n_contigs = offs.size() - 1;
    return GA_OK;
}
