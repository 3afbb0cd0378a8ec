"""De Bruijn graph classes with the reference's Python interface, built on the B200.

Drop-in for the upstream ``debruijn_graph.py``: ``DeBruijnGraph``, ``CMSDeBruijnGraph``,
``PairedDeBruijnGraph``, ``CMSPairedDeBruijnGraph`` take the same constructor arguments
(:54, :161, :204, :378), expose ``KMER_LEN`` / ``HAMMING_DIST`` /
``ALLOWED_PAIRED_DIST_ERROR``, ``nodes``, ``num_edges`` and ``enumerate_contigs()``, call the
overridable hooks ``_count_kmers`` -> [``_make_sketch``] -> ``_build_graph`` through ``self``
in that order (so ``debug_graph.DebugGraph`` composes), and raise the same ``ValueError``.

What changed underneath: ``_count_kmers`` fills a lock-free hash table on the GPU and returns
a dict-like view of it; ``_make_sketch`` pours it into a device CountMinSketch;
``_build_graph`` runs the stamp-based build kernels and receives the graph as a CSR whose
row / column order equals the reference's dict insertion order; ``enumerate_contigs`` walks
that CSR on the host.  ``nodes`` materialises real ``Node`` objects only when looked at.
There is no CPU fallback for the build: without the CUDA library these classes raise.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from collections import defaultdict
from itertools import tee

from countminsketch import CountMinSketch
from debruijn_node import Node, PairedNode


class AbstractDeBruijnGraph(ABC):
    @abstractmethod
    def enumerate_contigs(self):
        ...

    @abstractmethod
    def _get_longest_contig(self, start_node):
        ...

    @abstractmethod
    def _build_graph(self, kmer_counts, reads: list):
        ...

    @staticmethod
    @abstractmethod
    def _count_kmers(k: int, reads: list):
        ...

    @staticmethod
    @abstractmethod
    def _break_read_into_k_minus_one_mers(k: int, read):
        ...

    @staticmethod
    def valid_allowed_error(KMER_LEN, ALLOWED_PAIRED_DIST_ERROR):
        return KMER_LEN > ALLOWED_PAIRED_DIST_ERROR

    @staticmethod
    def _pairwise(iterable):
        first, second = tee(iterable)
        next(second, None)
        return zip(first, second)


class _GpuGraphBase(AbstractDeBruijnGraph):
    """Shared constructor flow, lazy ``nodes`` and the CSR traversal."""

    KMER_LEN = 29
    HAMMING_DIST = 3
    ALLOWED_PAIRED_DIST_ERROR = 2
    _PAIRED = False
    _SKETCHED = False

    def __init__(self, reads: list, k=None, hamming_dist=None, paired_error=None):
        if k is not None:
            self.KMER_LEN = k
        if hamming_dist is not None:
            self.HAMMING_DIST = hamming_dist
        if paired_error is not None:
            self.ALLOWED_PAIRED_DIST_ERROR = paired_error
        if not self.valid_allowed_error(self.KMER_LEN, self.ALLOWED_PAIRED_DIST_ERROR):
            raise ValueError("Allowed error must be less than the kmer length.")
        if self.KMER_LEN < 2:
            raise ValueError("kmer length must be at least 2 (nodes are (k-1)-mers)")
        self.num_edges = 0
        self._csr = None
        self._nodes = None
        self._nodes_dirty = False
        self._left = None
        kmer_counts = self._count_kmers(self.KMER_LEN, reads)
        if self._SKETCHED:
            exact = kmer_counts
            kmer_counts = self._make_sketch(exact)
            del exact
        self._build_graph(kmer_counts, reads)

    # -- nodes: real dict / defaultdict(dict) of Node objects, built on first access ------------
    @property
    def nodes(self):
        if self._nodes is None:
            self._nodes = self._materialise_nodes()
            self._nodes_dirty = True      # the caller may mutate them: traverse objects from now on
        return self._nodes

    @nodes.setter
    def nodes(self, value):
        self._nodes = value
        self._nodes_dirty = True

    def _materialise_nodes(self):
        csr = self._csr
        if csr is None or csr.n_nodes == 0:
            return defaultdict(dict) if self._PAIRED else dict()
        keys = csr.node_strings()
        rowptr, col = csr.rowptr, csr.col
        objs = [PairedNode(k[0], k[1]) for k in keys] if self._PAIRED else [Node(k) for k in keys]
        left = getattr(self, "_left", None)
        for i, node in enumerate(objs):
            stop = rowptr[i + 1] if left is None else rowptr[i] + left[i]
            for j in col[rowptr[i]:stop]:
                node.edges[keys[j]] = True
            node.num_edges_in = int(csr.indeg[i])
            node.was_branching = bool(csr.branching[i])
        if not self._PAIRED:
            return dict(zip(keys, objs))
        table = defaultdict(dict)
        for (a, b), node in zip(keys, objs):
            table[a][b] = node
        return table

    # -- hooks ----------------------------------------------------------------------------------
    @staticmethod
    def _device_reads(reads, paired):
        from ga_device import DeviceReads
        return DeviceReads(reads, paired)

    def _build_graph(self, kmer_counts, reads: list):
        """Graph construction on the GPU (replaces debruijn_graph.py:113-142 / :269-317)."""
        import ga_device as gd
        sketch, counts, keep = None, None, None
        if isinstance(kmer_counts, gd.KmerCounts):
            counts = kmer_counts
        elif isinstance(kmer_counts, CountMinSketch):
            kmer_counts._flush()
            sketch = kmer_counts._struct()
            counts = kmer_counts.source_counts
        else:
            threshold = self.HAMMING_DIST
            keep = lambda s, _m=kmer_counts: _m[s] > threshold   # noqa: E731  foreign mapping
        if counts is None or counts.reads.source is not reads or counts.k != self.KMER_LEN:
            counts = gd.KmerCounts(self.KMER_LEN, self._device_reads(reads, self._PAIRED))
        self._csr = gd.build_graph(counts, counts.reads, self.HAMMING_DIST, sketch=sketch, keep_fn=keep)
        self.num_edges = self._csr.num_edges_attr
        self._nodes = None
        self._nodes_dirty = False
        self._left = None

    # -- traversal --------------------------------------------------------------------------------
    def enumerate_contigs(self) -> list:
        if self._nodes_dirty:
            return self._enumerate_on_objects()
        if self._csr is None:
            return []
        contigs, remaining, left = self._csr.contigs()
        self.num_edges = remaining
        # the graph is consumed like the reference's: later looks at ``nodes`` (or a second
        # call) see only the edges that were not popped, and walk Node objects
        self._left = left
        self._nodes = None
        self._nodes_dirty = True
        return contigs

    def _all_nodes(self):
        if self._PAIRED:
            for inner in self.nodes.values():
                yield from inner.values()
        else:
            yield from self.nodes.values()

    def _enumerate_on_objects(self):
        """Same sweeps on materialised Node objects (debruijn_graph.py:72-92, 222-244)."""
        contigs = []
        last = None
        for node in self._all_nodes():
            last = node
            while node.outdegree > 0 and (node.was_branching or node.indegree == 0):
                contigs.append(self._get_longest_contig(node))
            if self.num_edges == 0:
                return contigs
        if self._PAIRED:
            for node in self._all_nodes():
                while node.outdegree > 0:
                    contigs.append(self._get_longest_contig(node))
                if self.num_edges == 0:
                    return contigs
        elif last is not None:
            while last.outdegree > 0:       # upstream quirk: only the last node's cycle (:85-86)
                contigs.append(self._get_longest_contig(last))
        return contigs

    def _successor(self, edge):
        if self._PAIRED:
            return edge[0][-1], self.nodes[edge[0]][edge[1]]
        return edge[-1], self.nodes[edge]

    def _get_longest_contig(self, cur_node) -> str:
        pieces = []
        while True:
            edge, _ = cur_node.pop_edge()
            self.num_edges -= 1
            char, cur_node = self._successor(edge)
            pieces.append(char)
            if cur_node.outdegree == 0 or cur_node.was_branching:
                return "".join(pieces)


class DeBruijnGraph(_GpuGraphBase):
    """Unpaired graph: nodes are (k-1)-mers, edges adjacent (k-1)-mers of a read."""

    KMER_LEN = 29
    HAMMING_DIST = 3
    ALLOWED_PAIRED_DIST_ERROR = 2

    @staticmethod
    def _count_kmers(k: int, reads: list):
        """Exact (k-1)-mer counts (debruijn_graph.py:144-152) as a device-backed mapping."""
        from ga_device import DeviceReads, KmerCounts
        return KmerCounts(k, DeviceReads(reads, False))

    @staticmethod
    def _break_read_into_k_minus_one_mers(k: int, read: str, paired=False) -> list:
        """'ACTGAC', k=4 -> ['ACT', 'CTG', 'TGA', 'GAC'] (host helper; the kernels roll keys)."""
        width = k - 1
        return [read[start:start + width] for start in range(len(read) - width + 1)]


class CMSDeBruijnGraph(DeBruijnGraph):
    _SKETCHED = True

    @staticmethod
    def _make_sketch(kmer_counts_dict) -> CountMinSketch:
        return _pour(kmer_counts_dict, 10)


class PairedDeBruijnGraph(_GpuGraphBase):
    """Paired graph: a node is (A, B) with A exact and B matched up to a >= 3 symbol overlap."""

    KMER_LEN = 23
    HAMMING_DIST = 3
    ALLOWED_PAIRED_DIST_ERROR = 2
    _PAIRED = True

    @staticmethod
    def _count_kmers(k: int, reads: list):
        """Both mates counted into one table (debruijn_graph.py:349-367)."""
        from ga_device import DeviceReads, KmerCounts
        return KmerCounts(k, DeviceReads(reads, True))

    @staticmethod
    def _break_read_into_k_minus_one_mers(k: int, read) -> list:
        """('ACTGAC', 'TCGATC'), k=4 -> [('ACT','TCG'), ('CTG','CGA'), ('TGA','GAT'), ('GAC','ATC')]"""
        width = k - 1
        first, second = read[0], read[1]
        return [(first[s:s + width], second[s:s + width]) for s in range(len(first) - width + 1)]

    def _find_matching_node(self, paired_strings) -> tuple:
        """(found, node) as the reference's lookup (debruijn_graph.py:319-334) on ``nodes``."""
        inner = self.nodes.get(paired_strings[0])
        if inner:
            node = inner.get(paired_strings[1])
            if node is not None:
                return (True, node)
            for other, candidate in inner.items():
                if self._find_longest_overlap_brute(other, paired_strings[1]) or \
                        self._find_longest_overlap_brute(paired_strings[1], other):
                    return (True, candidate)
        return (False, None)

    @staticmethod
    def _find_longest_overlap_brute(pattern: str, text: str) -> int:
        """Length of the suffix of ``text`` (at least 3 symbols) that is a prefix of ``pattern``."""
        slack = PairedDeBruijnGraph.ALLOWED_PAIRED_DIST_ERROR
        for shift in range(len(text) - slack):
            span = min(len(text) - shift, len(pattern))
            if text[shift:shift + span] == pattern[:span]:
                return span
        return 0


class CMSPairedDeBruijnGraph(PairedDeBruijnGraph):
    _SKETCHED = True

    @staticmethod
    def _make_sketch(kmer_counts_dict) -> CountMinSketch:
        return _pour(kmer_counts_dict, 8)


def _pour(kmer_counts_dict, num_rows: int) -> CountMinSketch:
    """dict -> sketch (debruijn_graph.py:181-188, 398-405); one kernel when the dict is ours."""
    from ga_device import KmerCounts
    sketch = CountMinSketch(num_rows)
    if isinstance(kmer_counts_dict, KmerCounts):
        sketch.pour_counts(kmer_counts_dict)
    else:
        items = list(kmer_counts_dict.items())
        sketch.update_many([kmer for kmer, _ in items], [count for _, count in items])
    return sketch
