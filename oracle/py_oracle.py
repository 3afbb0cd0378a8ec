"""CPU oracle (pure Python) for the k-mer counting / CountMinSketch / de Bruijn build path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``genome-assembler_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs use it, and there only as the checker / the timed CPU arm.

This is a *restatement* of the reference's algorithm, written against plain dicts and
lists the way the reference's CPython path works, so that its speed is representative
of the reference's own pure-Python implementation (it is the ``cpu_baseline`` "port").
Every function cites the reference lines it follows (paths are into the upstream
checkout, ``/root/reference`` in the build container).

Parity pin: ``tests/golden/golden.json`` holds digests produced by the *unmodified*
reference (``tests/golden/make_golden.py``); ``tests/test_oracle.py`` checks this
module against every one of them, plus the README known-answer test and the
MurmurHash3 vectors of SURVEY App. B.1; where ``oracle/_ref/`` (the unmodified reference) is present,
``tests/test_reference_differential.py`` also runs it next to the reference on fresh random inputs.
"""
from __future__ import annotations

import hashlib
from array import array
from collections import defaultdict

# countminsketch.py:15-18 -- the only prime table the sketch ever indexes.
PRIMES_1_10_7 = (9999889, 9999901, 9999907, 9999929, 9999931,
                 9999937, 9999943, 9999971, 9999973, 9999991,
                 10000019, 10000079, 10000103, 10000121, 10000139,
                 10000141, 10000169, 10000189, 10000223, 10000229)

MIN_FUZZY_OVERLAP = 3  # debruijn_graph.py:202,339: len(text) - 2 start positions


# --------------------------------------------------------------------------- read breaking
def windows(k: int, read: str):
    """(k-1)-length windows of ``read`` (debruijn_graph.py:154-157)."""
    w = k - 1
    return [read[i:i + w] for i in range(len(read) - w + 1)]


def paired_windows(k: int, pair):
    """Zipped windows of both mates; range from mate 1 (debruijn_graph.py:369-374)."""
    w = k - 1
    a, b = pair[0], pair[1]
    return [(a[i:i + w], b[i:i + w]) for i in range(len(a) - w + 1)]


# --------------------------------------------------------------------------- counting
def count_unpaired(k: int, reads):
    """Occurrences per distinct (k-1)-mer (debruijn_graph.py:144-152)."""
    tally = defaultdict(int)
    for read in reads:
        for piece in windows(k, read):
            tally[piece] += 1
    return tally


def count_paired(k: int, pairs):
    """Both mates feed one table (debruijn_graph.py:349-367)."""
    tally = defaultdict(int)
    for pair in pairs:
        for left, right in paired_windows(k, pair):
            tally[left] += 1
            tally[right] += 1
    return tally


# --------------------------------------------------------------------------- sketch
def murmur3_32(text: str, seed: int = 0) -> int:
    """MurmurHash3_x86_32 over ``ord(c) & 0xff`` bytes (countminsketch.py:46-95).

    (The reference leaves the 4th byte of a block unmasked, ``ord(data[i+3]) << 24``;
    after the final ``& 0xffffffff`` only code points > 255 could differ, and those are
    rejected by the GPU path -- SURVEY App. A-18.)
    """
    m = 0xFFFFFFFF
    data = bytes(ord(c) & 0xFF for c in text)
    n = len(data)
    h = seed & m
    body = n & ~3
    for i in range(0, body, 4):
        blk = int.from_bytes(data[i:i + 4], "little")
        blk = (blk * 0xCC9E2D51) & m
        blk = ((blk << 15) | (blk >> 17)) & m
        blk = (blk * 0x1B873593) & m
        h ^= blk
        h = ((h << 13) | (h >> 19)) & m
        h = (h * 5 + 0xE6546B64) & m
    rest = n & 3
    if rest:
        blk = int.from_bytes(data[body:], "little")
        blk = (blk * 0xCC9E2D51) & m
        blk = ((blk << 15) | (blk >> 17)) & m
        blk = (blk * 0x1B873593) & m
        h ^= blk
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & m
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & m
    h ^= h >> 16
    return h


class Sketch:
    """d rows of unsigned-16 cells with prime widths (countminsketch.py:26-44, 97-99)."""

    def __init__(self, num_rows: int):
        assert num_rows < len(PRIMES_1_10_7)
        self.num_rows = num_rows
        self.rows = [array("H", bytes(2 * PRIMES_1_10_7[i])) for i in range(num_rows)]

    def update(self, text: str, amount: int) -> None:
        h = murmur3_32(text)
        for row in self.rows:
            row[h % len(row)] += amount  # array('H') raises OverflowError past 65535

    def estimate(self, text: str) -> int:
        h = murmur3_32(text)
        return min(row[h % len(row)] for row in self.rows)

    __getitem__ = estimate


def fill_sketch(tally, num_rows: int) -> Sketch:
    """Pour the exact table into a sketch (debruijn_graph.py:181-188, 398-405)."""
    sk = Sketch(num_rows)
    for piece, n in tally.items():
        sk.update(piece, n)
    return sk


# --------------------------------------------------------------------------- graph records
class Graph:
    """Order-preserving result of a build, convertible to the CSR contract.

    keys[i]       node key in ``self.nodes`` iteration order (str, or (A, B) when paired)
    succ[i]       successor node indices in edge-insertion order
    indeg[i]      distinct accepted in-edges
    num_edges     the reference's ``graph.num_edges`` after the build
    """

    def __init__(self, keys, succ, indeg, num_edges, paired):
        self.keys, self.succ, self.indeg = keys, succ, indeg
        self.num_edges, self.paired = num_edges, paired
        self.branching = [len(s) > 1 or d > 1 for s, d in zip(succ, indeg)]

    def last_char(self, i):
        key = self.keys[i]
        return (key[0] if self.paired else key)[-1]

    def digest(self) -> str:
        """sha256 in the format of SURVEY App. B.3 (first 16 hex digits)."""
        h = hashlib.sha256()
        for i, key in enumerate(self.keys):
            edges = [self.keys[j] for j in self.succ[i]]
            h.update(repr((key, edges, self.indeg[i], self.branching[i])).encode())
        return h.hexdigest()[:16]


def build_unpaired(tally, reads, k: int, threshold: int) -> Graph:
    """Sequential edge insertion with the strict ``>`` filter (debruijn_graph.py:113-142)."""
    index = {}          # key -> node id, in first-insertion order
    keys, succ, indeg = [], [], []
    edge_seen = set()
    total = 0
    for read in reads:
        pieces = windows(k, read)
        for j in range(len(pieces) - 1):
            left, right = pieces[j], pieces[j + 1]
            if (left, right) in edge_seen:
                continue
            if tally[left] > threshold and tally[right] > threshold:
                for piece in (left, right):      # prefix is created before suffix
                    if piece not in index:
                        index[piece] = len(keys)
                        keys.append(piece)
                        succ.append([])
                        indeg.append(0)
                edge_seen.add((left, right))
                succ[index[left]].append(index[right])
                indeg[index[right]] += 1
                total += 1
    return Graph(keys, succ, indeg, total, paired=False)


def fuzzy_overlap(pattern: str, text: str) -> int:
    """Length of a suffix of ``text`` (>= 3 chars) that prefixes ``pattern``, else 0
    (debruijn_graph.py:336-347)."""
    for start in range(len(text) - (MIN_FUZZY_OVERLAP - 1)):
        span = min(len(text) - start, len(pattern))
        if text[start:start + span] == pattern[:span]:
            return span
    return 0


def build_paired(tally, pairs, k: int, threshold: int) -> Graph:
    """Sequential paired build: 4-way filter, two look-ups *then* two inserts, fuzzy
    second key (debruijn_graph.py:269-334).  Node objects are lists
    ``[A, B, out_keys(dict), indeg]`` so the orphan-overwrite corner (SURVEY A-9) falls
    out of the same mechanics as in the reference."""
    groups = {}         # A -> {B -> node}
    total = 0

    def lookup(a, b):
        inner = groups.get(a)
        if inner is None:
            return None
        node = inner.get(b)
        if node is not None:
            return node
        for other, cand in inner.items():
            if fuzzy_overlap(other, b) or fuzzy_overlap(b, other):
                return cand
        return None

    for pair in pairs:
        steps = paired_windows(k, pair)
        for j in range(len(steps) - 1):
            (pa, pb), (sa, sb) = steps[j], steps[j + 1]
            if not (tally[pa] > threshold and tally[sa] > threshold and
                    tally[pb] > threshold and tally[sb] > threshold):
                continue
            src, dst = lookup(pa, pb), lookup(sa, sb)
            both_known = src is not None and dst is not None
            if src is None:
                src = [pa, pb, {}, 0]
                groups.setdefault(pa, {})[pb] = src
            if dst is None:
                dst = [sa, sb, {}, 0]
                groups.setdefault(sa, {})[sb] = dst
            target = (dst[0], dst[1])
            if both_known and target in src[2]:
                continue
            src[2][target] = True
            dst[3] += 1
            total += 1

    keys, nodes = [], []
    for a, inner in groups.items():      # outer-then-inner iteration (:225-226)
        for b, node in inner.items():
            keys.append((a, b))
            nodes.append(node)
    where = {key: i for i, key in enumerate(keys)}
    succ = [[where[t] for t in node[2]] for node in nodes]
    indeg = [node[3] for node in nodes]
    return Graph(keys, succ, indeg, total, paired=True)


# --------------------------------------------------------------------------- traversal
def contigs(graph: Graph):
    """Contig enumeration over the ordered adjacency (debruijn_graph.py:72-111 unpaired,
    :222-267 paired).  Pops edges LIFO (``dict.popitem``, debruijn_node.py:24-26).  The
    unpaired second sweep only ever revisits the last node of the first sweep
    (SURVEY A-11); the paired one revisits every node."""
    left = [list(s) for s in graph.succ]
    remaining = graph.num_edges
    out = []

    def walk(i):
        nonlocal remaining
        text = []
        j = left[i].pop()
        remaining -= 1
        text.append(graph.last_char(j))
        while left[j] and not graph.branching[j]:
            nxt = left[j].pop()
            remaining -= 1
            text.append(graph.last_char(nxt))
            j = nxt
        return "".join(text)

    n = len(graph.keys)
    for i in range(n):
        while left[i] and (graph.branching[i] or graph.indeg[i] == 0):
            out.append(walk(i))
        if remaining == 0:
            return out
    if graph.paired:
        for i in range(n):
            while left[i]:
                out.append(walk(i))
            if remaining == 0:
                return out
    elif n:
        last = n - 1                       # stale loop variable of the first sweep
        for _ in range(n):
            while left[last]:
                out.append(walk(last))
            if remaining == 0:
                return out
    return out


def contig_digest(lines) -> str:
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()[:16]


# --------------------------------------------------------------------------- whole path
def assemble(reads, k: int, threshold: int, paired: bool, sketch_rows: int = 0):
    """count -> [sketch] -> build, returning (tally, sketch|None, Graph)."""
    tally = (count_paired if paired else count_unpaired)(k, reads)
    sk = fill_sketch(tally, sketch_rows) if sketch_rows else None
    source = sk if sk is not None else tally
    graph = (build_paired if paired else build_unpaired)(source, reads, k, threshold)
    return tally, sk, graph
