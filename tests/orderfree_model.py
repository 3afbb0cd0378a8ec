"""Executable model of the ORDER-FREE build the CUDA path implements ("stamps").

Test infrastructure: a slow, dictionary-based statement of exactly the data flow of the
kernels (first-occurrence stamps -> sort -> CSR; paired: distinct queries -> per-A-group
greedy -> representative map -> distinct edges), so the algorithm itself -- including
the homopolymer corners of SURVEY App. A-9 -- is pinned against the reference's golden
digests on the CPU before any kernel runs.  It shares no code with the product.
"""
from __future__ import annotations

from oracle import py_oracle as po

INF = float("inf")


def _match(x: str, y: str) -> bool:
    return bool(po.fuzzy_overlap(x, y) or po.fuzzy_overlap(y, x))


def build_unpaired(solid, reads, k):
    """solid: set of accepted (k-1)-mers.  Returns po.Graph."""
    node_stamp, edge_stamp = {}, {}
    e = 0
    for read in reads:
        win = po.windows(k, read)
        for j in range(len(win) - 1):
            p, s = win[j], win[j + 1]
            if p in solid and s in solid:
                node_stamp[p] = min(node_stamp.get(p, INF), 2 * e)
                node_stamp[s] = min(node_stamp.get(s, INF), 2 * e + 1)
                edge_stamp[(p, s)] = min(edge_stamp.get((p, s), INF), e)
            e += 1
    keys = sorted(node_stamp, key=node_stamp.get)
    rank = {x: i for i, x in enumerate(keys)}
    rows = [[] for _ in keys]
    indeg = [0] * len(keys)
    for (p, s), st in edge_stamp.items():
        rows[rank[p]].append((st, rank[s]))
        indeg[rank[s]] += 1
    succ = [[t for _, t in sorted(r)] for r in rows]
    return po.Graph(keys, succ, indeg, len(edge_stamp), paired=False)


def build_paired(solid, pairs, k):
    qstamp, qedge = {}, {}
    dh = {}                       # (A,B) with P == S -> sorted list of event stamps (two smallest kept)
    e = 0
    for pair in pairs:
        win = po.paired_windows(k, pair)
        for j in range(len(win) - 1):
            P, S = win[j], win[j + 1]
            if P[0] in solid and S[0] in solid and P[1] in solid and S[1] in solid:
                qstamp[P] = min(qstamp.get(P, INF), 2 * e)
                qstamp[S] = min(qstamp.get(S, INF), 2 * e + 1)
                qedge[(P, S)] = min(qedge.get((P, S), INF), e)
                if P == S:
                    dh[P] = sorted(dh.get(P, []) + [e])[:2]
            e += 1
    # group by A, members by stamp; greedy representative choice
    groups = {}
    for q in sorted(qstamp, key=qstamp.get):
        groups.setdefault(q[0], []).append(q)
    rep, key_stamp, group_stamp = {}, {}, {}
    for a, members in groups.items():
        keys = []
        group_stamp[a] = qstamp[members[0]]
        for q in members:
            st = qstamp[q]
            found = None
            for kq in keys:
                if (st & 1) and qstamp[kq] == st - 1:
                    continue          # inserted by the same occurrence: not visible yet
                if _match(kq[1], q[1]):
                    found = kq
                    break
            if found is None:
                keys.append(q)
                key_stamp[q] = st
                rep[q] = q
            else:
                rep[q] = found
    nodes = sorted(key_stamp, key=lambda q: (group_stamp[q[0]], key_stamp[q]))
    rank = {q: i for i, q in enumerate(nodes)}
    edge_stamp = {}
    for (P, S), st in qedge.items():
        key = (rank[rep[P]], rank[rep[S]])
        edge_stamp[key] = min(edge_stamp.get(key, INF), st)
    extra_in = [0] * len(nodes)
    extra_edges = 0
    for q, evs in dh.items():
        if rep[q] == q and qstamp[q] == 2 * evs[0]:
            # first self-loop occurrence created the node twice: the edge hangs off an orphan
            i = rank[q]
            loop = (i, i)
            others = [st for (P, S), st in qedge.items()
                      if (P, S) != (q, q) and rank[rep[P]] == i and rank[rep[S]] == i]
            second = evs[1] if len(evs) > 1 else INF
            best = min(others + [second])
            if best == INF:
                del edge_stamp[loop]
            else:
                edge_stamp[loop] = best
            extra_in[i] += 1
            extra_edges += 1
    rows = [[] for _ in nodes]
    indeg = list(extra_in)
    for (i, j), st in edge_stamp.items():
        rows[i].append((st, j))
        indeg[j] += 1
    succ = [[t for _, t in sorted(r)] for r in rows]
    return po.Graph(nodes, succ, indeg, len(edge_stamp) + extra_edges, paired=True)
