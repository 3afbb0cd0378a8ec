"""Node records of the de Bruijn graph -- host-side only.

Same observable interface as the reference's debruijn_node.py (Node :4-53, PairedNode
:56-70): ``data``, ``edges`` (insertion-ordered mapping successor-key -> True),
``was_branching``, ``num_edges_in``, ``outdegree`` / ``indegree``, ``pop_edge`` (LIFO, as
``dict.popitem``) and ``append_edge``.  The GPU build never creates these; they are
materialised from the CSR only when a caller looks at ``graph.nodes``.
"""
import sys


class Node:
    # plain instances (no __slots__): object.__sizeof__ enters the reference's -m totals (debruijn_node.py:35-53)
    def __init__(self, data):
        self.data = data
        self.edges = {}
        self.was_branching = False
        self.num_edges_in = 0

    @property
    def outdegree(self):
        """Shrinks while the traversal pops edges."""
        return len(self.edges)

    @property
    def indegree(self):
        """Fixed once the graph is built."""
        return self.num_edges_in

    def pop_edge(self):
        """Remove and return the most recently added edge as (key, True)."""
        return self.edges.popitem()

    def append_edge(self, edge):
        self.edges[edge] = True

    def __repr__(self):
        return "Data: {0} | Edges: {1}".format(self.data, self.edges)

    def _edge_bytes(self):
        total = 0
        if self.edges:
            first = next(iter(self.edges))
            if isinstance(first, tuple):
                total += sum(sys.getsizeof(second) for _, second in self.edges)
            elif isinstance(first, str):
                total += sum(sys.getsizeof(edge) for edge in self.edges)
        return total

    def __sizeof__(self):
        # same accounting as the reference's -m report (debruijn_node.py:35-53)
        if hasattr(self, "total_mem"):
            return self.total_mem
        total = object.__sizeof__(self) + sys.getsizeof(self.data) + sys.getsizeof(self.edges)
        total += self._edge_bytes()
        total += sys.getsizeof(self.was_branching) + sys.getsizeof(self.num_edges_in)
        self.total_mem = total + sys.getsizeof(total)
        return total


class PairedNode(Node):
    def __init__(self, data, paired_data):
        self.paired_data = paired_data
        super().__init__(data)

    def __repr__(self):
        return "Data: {0} | Pair: {1} | Edges: {2}".format(self.data, self.paired_data, self.edges)

    def __sizeof__(self):
        if hasattr(self, "total_mem"):
            return self.total_mem
        total = super().__sizeof__() + sys.getsizeof(self.paired_data)
        self.total_mem = total
        return total
