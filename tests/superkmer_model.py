"""Executable model of the BUCKETED (super-k-mer) count + build the CUDA path uses for unpaired DNA
reads with 64-bit keys (genome-assembler_b200/csrc/ga_superkmer.cu, DESIGN.md "super-k-mer buckets").

Test infrastructure, pure Python, no code shared with the product.  It states the data flow the
kernels implement and is pinned on the CPU against the reference's golden digests:

  1. every window ((k-1)-mer occurrence) gets a bucket that depends on the window's CONTENT only
     (the smallest hash among its m-mers), so all occurrences of a window meet in one bucket;
  2. runs of consecutive windows of a read that share a bucket travel as one record: the bases of
     the run, the base that follows it (if the read has one) and the occurrence ordinal of its
     first window;
  3. per bucket: exact counts; for every window with count > F ("solid") and every next symbol c
     the CANDIDATE edge stamp  cand[p][c] = min ordinal over occurrences of p followed by c;
  4. resolve (global): edge (p, c) exists iff p and s = p[1:] + c are both solid -- solidity is a
     property of the string, so either every occurrence of "p followed by c" is accepted or none is
     and the candidate IS the reference's first-insertion ordinal (debruijn_graph.py:121-137);
     node_stamp[p] = min 2e over its out-edges, node_stamp[s] = min 2e+1 over its in-edges
     (SURVEY App. C.1).
"""
from __future__ import annotations

from oracle import py_oracle as po

INF = float("inf")
MAX_RUN = 32          # windows per record (one 32-lane group of the warp that cuts the read)


def _hash32(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    return x ^ (x >> 16)


def minimizer_len(w: int) -> int:
    """m-mer length for windows of w symbols: at most 16 m-mers per window, m <= 16."""
    return max(min(w, 11), min(w - 15, 16))


def window_bucket(window: str, m: int, bits: int) -> int:
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    best = 0xFFFFFFFF
    for i in range(len(window) - m + 1):
        x = 0
        for j, ch in enumerate(window[i:i + m]):       # first symbol in the low bits, as the packed reads
            x |= code[ch] << (2 * j)
        best = min(best, _hash32(x))
    return best >> (32 - bits) if bits else 0


def cut_records(reads, k: int, bits: int):
    """[(bucket, bases, n_windows, has_next, first ordinal)] -- step 2 above."""
    w = k - 1
    m = minimizer_len(w)
    out = []
    base_e = 0
    stride = max((len(r) for r in reads), default=1) or 1
    for index, read in enumerate(reads):
        n_win = len(read) - w + 1
        base_e = index * stride
        if n_win <= 0:
            continue
        buckets = [window_bucket(read[p:p + w], m, bits) for p in range(n_win)]
        p = 0
        while p < n_win:
            q = p + 1
            while q < n_win and buckets[q] == buckets[p] and q % MAX_RUN != 0:
                q += 1
            has_next = q - 1 + w < len(read)
            out.append((buckets[p], read[p:q - 1 + w + (1 if has_next else 0)], q - p, has_next, base_e + p))
            p = q
    return out, stride


def bucket_pass(records, k: int, F: int):
    """Steps 3: per bucket exact counts -> solid windows and candidate edge stamps."""
    w = k - 1
    by_bucket = {}
    for rec in records:
        by_bucket.setdefault(rec[0], []).append(rec)
    solid_keys, cand = [], {}
    for bucket in sorted(by_bucket):
        tally = {}
        for _, bases, n_win, _, _ in by_bucket[bucket]:
            for j in range(n_win):
                x = bases[j:j + w]
                tally[x] = tally.get(x, 0) + 1
        solid = {x for x, c in tally.items() if c > F}
        solid_keys.extend(sorted(solid))
        for _, bases, n_win, has_next, e0 in by_bucket[bucket]:
            for j in range(n_win):
                if j == n_win - 1 and not has_next:
                    continue
                x = bases[j:j + w]
                if x in solid:
                    key = (x, bases[j + w])
                    cand[key] = min(cand.get(key, INF), e0 + j)
    return solid_keys, cand


def resolve(solid_keys, cand):
    """Step 4 -> (node_stamp, edge_stamp) keyed like tests/orderfree_model.py."""
    solid = set(solid_keys)
    node_stamp, edge_stamp = {}, {}
    for (p, c), e in cand.items():
        s = p[1:] + c
        if s not in solid:
            continue
        edge_stamp[(p, s)] = e
        node_stamp[p] = min(node_stamp.get(p, INF), 2 * e)
        node_stamp[s] = min(node_stamp.get(s, INF), 2 * e + 1)
    return node_stamp, edge_stamp


def build_unpaired(reads, k: int, F: int, bits: int = 4):
    records, stride = cut_records(reads, k, bits)
    solid_keys, cand = bucket_pass(records, k, F)
    node_stamp, edge_stamp = resolve(solid_keys, cand)
    keys = sorted(node_stamp, key=node_stamp.get)
    rank = {x: i for i, x in enumerate(keys)}
    rows = [[] for _ in keys]
    indeg = [0] * len(keys)
    for (p, s), st in edge_stamp.items():
        rows[rank[p]].append((st, rank[s]))
        indeg[rank[s]] += 1
    succ = [[t for _, t in sorted(r)] for r in rows]
    return po.Graph(keys, succ, indeg, len(edge_stamp), paired=False), records
