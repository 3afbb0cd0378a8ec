"""Size-independent properties of an unpaired DNA de Bruijn graph built from high-coverage reads of a known
circular genome (BASELINE config C4's input class).  TEST INFRASTRUCTURE: plain torch tensor arithmetic that runs
on the CPU (checked against the C oracle's CSR in tests/test_graph_properties.py) and on the GPU, where it is what
looks at the FULL-SIZE C4 graph (100 M reads, 57 M nodes) that no CPU oracle finishes.

What the reference's construction (debruijn_graph.py:113-142) guarantees and this module checks:
  P1  every (k-1)-window of the genome is a node (each is covered ~240 times at 300x, F = 3);
  P2  consecutive genome windows are joined by an edge;
  P3  every edge u -> v is a one-symbol shift: v = u[1:] + last(v);
  P4  node.num_edges_in equals the number of edges that point at the node, was_branching is
      "out-degree > 1 or in-degree > 1" (:139-142), a row holds at most one edge per symbol, rows are contiguous;
  P5  nodes are distinct.
"""
from __future__ import annotations

import torch


def genome_window_keys(codes: torch.Tensor, w: int) -> torch.Tensor:
    """Packed key (first symbol most significant, 2 bits per base) of the window that starts at every position
    of the circular genome; codes: uint8 in 0..3, w <= 31."""
    g = codes.numel()
    ext = torch.cat([codes, codes[:w]]).to(torch.int64)
    acc = torch.zeros(g, dtype=torch.int64, device=codes.device)
    for j in range(w):
        acc = (acc << 2) | ext[j:j + g]
    return acc


def check_unpaired_dna_graph(codes, k, rowptr, col, indeg, branching, last_sym, keys, n_nodes, n_edges):
    """All tensors on one device; keys: int64[n] packed node keys in node order; last_sym: symbol codes 0..3.
    Returns a dict of counts for the caller's report; raises AssertionError on the first broken property."""
    w = k - 1
    assert 2 <= w <= 31
    n, m = int(n_nodes), int(n_edges)
    rowptr = rowptr[:n + 1].to(torch.int64)
    col = col[:m].to(torch.int64)
    indeg = indeg[:n].to(torch.int64)
    branching = branching[:n].to(torch.bool)
    last_sym = last_sym[:n].to(torch.int64)
    keys = keys[:n].to(torch.int64)
    dev = keys.device
    mask = (1 << (2 * w)) - 1

    # P4a: rows
    outdeg = rowptr[1:] - rowptr[:-1]
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == m, "rowptr does not span the edges"
    assert bool((outdeg >= 0).all()) and bool((outdeg <= 4).all()), "a row with more than four edges"
    assert bool(((col >= 0) & (col < n)).all()), "edge to a node that does not exist"
    # P5: distinct nodes
    sorted_keys, order = torch.sort(keys)
    assert bool((sorted_keys[1:] != sorted_keys[:-1]).all()), "the same window is two nodes"
    assert bool(((keys >= 0) & (keys <= mask)).all()), "key wider than the window"
    assert bool(((keys & 3) == last_sym).all()), "last symbol does not match the key"
    # P3: every edge is a one-symbol shift
    src = torch.repeat_interleave(torch.arange(n, device=dev), outdeg)
    assert bool((keys[col] == (((keys[src] << 2) & mask) | last_sym[col])).all()), "an edge that is not a shift"
    # one edge per symbol inside a row: (src, last symbol of dst) pairs are distinct
    pair = src * 4 + last_sym[col]
    pair_sorted, _ = torch.sort(pair)
    assert bool((pair_sorted[1:] != pair_sorted[:-1]).all()), "two edges of a node with the same symbol"
    # P4b: degree bookkeeping
    true_in = torch.bincount(col, minlength=n)
    assert bool((true_in == indeg).all()), "num_edges_in differs from the edges that arrive"
    assert bool((branching == ((outdeg > 1) | (indeg > 1))).all()), "was_branching differs from the degrees"
    # P1: the genome's windows are nodes
    gkeys = genome_window_keys(codes, w)
    pos = torch.searchsorted(sorted_keys, gkeys).clamp(max=n - 1)
    found = sorted_keys[pos] == gkeys
    missing = int((~found).sum())
    assert missing == 0, "%d genome windows are not nodes" % missing
    node_of = order[pos]
    # P2: consecutive genome windows are joined
    u, v = node_of, torch.roll(node_of, -1)
    joined = torch.zeros(u.numel(), dtype=torch.bool, device=dev)
    for t in range(4):
        e = rowptr[u] + t
        ok = e < rowptr[u + 1]
        joined |= ok & (col[e.clamp(max=max(m - 1, 0))] == v)
    unjoined = int((~joined).sum())
    assert unjoined == 0, "%d consecutive genome windows without an edge" % unjoined
    genome_nodes = int(torch.unique(node_of).numel())
    return {"nodes": n, "edges": m, "genome_nodes": genome_nodes, "other_nodes": n - genome_nodes,
            "branching": int(branching.sum())}
