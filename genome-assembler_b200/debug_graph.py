"""Timing / memory instrumentation mixin, same surface as the reference's debug_graph.py.

``DebugGraph`` wraps the four hooks of the graph classes (``_count_kmers``, ``_make_sketch``,
``_build_graph``, ``enumerate_contigs``) by cooperative ``super()`` and prints the same
``-t`` / ``-m`` lines (debug_graph.py:20-85).  Composed with the GPU-backed classes of
``debruijn_graph`` exactly as upstream composes it with the Python ones (:88-105).  As
upstream, the mixin's ``_make_sketch`` always uses 10 rows (it shadows the 8-row paired
variant in the MRO).
"""
import sys
import time

from countminsketch import CountMinSketch
from debruijn_graph import DeBruijnGraph, PairedDeBruijnGraph
from debruijn_graph import CMSDeBruijnGraph, CMSPairedDeBruijnGraph, _pour


class DebugGraph:
    """Mixin; not meant to be instantiated on its own."""

    def __init__(self, print_syssizeof=False, print_runtime=False, start_time=0, **kwargs):
        self.print_syssizeof = print_syssizeof
        self.print_runtime = print_runtime
        self.start_time = start_time
        super().__init__(**kwargs)

    def _elapsed(self):
        return time.time() - self.start_time

    def _banner(self, what):
        if self.print_runtime:
            print("\n>--- STARTING TO {0} AT T = {1:.2f} ---".format(what, self._elapsed()))

    def enumerate_contigs(self):
        self._banner("ENUMERATE CONTIGS")
        contigs = super().enumerate_contigs()
        if self.print_runtime:
            print(">FINISHED ENUMERATING CONTIGS AT T = {:.2f} ---".format(self._elapsed()))
        return contigs

    def _build_graph(self, kmer_counts, reads):
        self._banner("BUILD GRAPH")
        super()._build_graph(kmer_counts, reads)
        if self.print_runtime:
            print(">FINISHED BUILDING GRAPH AT T = {:.2f}".format(self._elapsed()))
        if self.print_syssizeof:
            container = sys.getsizeof(self.nodes)
            payload = 0
            for key, node in self.nodes.items():
                container += sys.getsizeof(key)
                payload += sys.getsizeof(node)
            print(">SIZE OF GRAPH CONTAINER: {:,}".format(container))
            print(">SIZE OF ALL NODES: {:,}".format(payload))

    def _count_kmers(self, k, reads):
        self._banner("COUNT KMERS")
        counts = super()._count_kmers(k, reads)
        if self.print_runtime:
            print(">FINISHED COUNTING KMERS AT T = {:.2f}".format(self._elapsed()))
        if self.print_syssizeof:
            container = sys.getsizeof(counts)
            values = 0
            if counts:
                for kmer, count in counts.items():
                    container += sys.getsizeof(kmer)
                    values += sys.getsizeof(count)
            print(">SIZE OF COUNTS CONTAINER: {:,}".format(container))
            print(">SIZE OF STRINGS IN COUNTS: {:,}".format(values))   # label as upstream prints it
        return counts

    def _make_sketch(self, kmer_counts_dict) -> CountMinSketch:
        self._banner("MAKE COUNTMIN SKETCH")
        sketch = _pour(kmer_counts_dict, 10)
        if self.print_runtime:
            # upstream reports progress every 50,000 k-mers of its Python loop; the pour is one
            # kernel here, so the same lines are printed with the time it finished
            for done in range(0, len(kmer_counts_dict), 50000):
                print(">Processed {0} kmers by time T={1:.2f}".format(done, self._elapsed()))
            print(">FINISHED MAKING COUNTMIN SKETCH AT T = {:.2f}".format(self._elapsed()))
        if self.print_syssizeof:
            print(">SIZE OF COUNTMIN SKETCH: {:,}".format(sys.getsizeof(sketch)))
        return sketch


class DebugDeBruijnGraph(DebugGraph, DeBruijnGraph):
    pass


class DebugCMSDeBruijnGraph(DebugGraph, CMSDeBruijnGraph):
    pass


class DebugPairedDeBruijnGraph(DebugGraph, PairedDeBruijnGraph):
    pass


class DebugCMSPairedDeBruijnGraph(DebugGraph, CMSPairedDeBruijnGraph):
    pass
