#!/usr/bin/env python3
"""Runs the UNMODIFIED reference (the directory given as argv[1]: oracle/_ref, a verbatim copy of /root/reference made
by oracle/make_ref.py) on jobs read as JSON from stdin and prints the answers as JSON.  A separate process because
the reference's modules carry the same names as the product's.  Test infrastructure: used by
tests/test_reference_differential.py only, never by the product."""
import contextlib
import hashlib
import io
import json
import sys

REF = sys.argv[1]
sys.path.insert(0, REF)
sys.setrecursionlimit(10000)

import assemble as ref_cli                 # noqa: E402
import countminsketch as ref_cms           # noqa: E402
import debruijn_graph as ref_dbg           # noqa: E402
import debruijn_node as ref_node           # noqa: E402


def sha16(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()[:16]


def graph_digest(graph, paired):
    """Same digest as tests/golden/make_golden.py: nodes in dict order, each with its edges in dict order."""
    h = hashlib.sha256()
    n = 0
    if paired:
        it = (((a, b), node) for a, inner in graph.nodes.items() for b, node in inner.items())
    else:
        it = iter(graph.nodes.items())
    for key, node in it:
        h.update(repr((key, list(node.edges), node.num_edges_in, node.was_branching)).encode())
        n += 1
    return n, h.hexdigest()[:16]


def job_args(job):
    argv = job["argv"]
    old = sys.argv
    sys.argv = ["assemble.py"] + argv
    err = io.StringIO()
    try:
        with contextlib.redirect_stderr(err), contextlib.redirect_stdout(io.StringIO()):
            ns = ref_cli.IOHandler.read_args()
        return {"ok": vars(ns)}
    except SystemExit as exc:
        return {"exit": exc.code}
    finally:
        sys.argv = old


def job_graph(job):
    reads = [tuple(r) for r in job["reads"]] if job["paired"] else list(job["reads"])
    cls = ref_dbg.PairedDeBruijnGraph if job["paired"] else ref_dbg.DeBruijnGraph
    try:
        counts = cls._count_kmers(job["k"], reads)
        items = sorted(counts.items())
        graph = cls(reads, k=job["k"], hamming_dist=job["F"], paired_error=job.get("e"))
    except Exception as exc:        # noqa: BLE001 -- the exception type is the answer
        return {"error": type(exc).__name__, "message": str(exc)}
    n_nodes, digest = graph_digest(graph, job["paired"])
    num_edges = graph.num_edges
    contigs = graph.enumerate_contigs()
    after = graph_digest(graph, job["paired"])[1]          # what is left of the nodes once the contigs are out
    return {"digest_after": after, "edges_left": graph.num_edges,
            "counts_sha": sha16("".join("%s:%d\n" % kc for kc in items).encode()), "n_distinct": len(items),
            "n_nodes": n_nodes, "num_edges": num_edges, "graph_digest": digest, "contigs": contigs,
            "constants": [graph.KMER_LEN, graph.HAMMING_DIST, graph.ALLOWED_PAIRED_DIST_ERROR]}


def job_helpers(job):
    out = {}
    out["hash"] = [ref_cms.CountMinSketch._hash(s, seed) for s, seed in job["hash"]]
    out["overlap"] = [ref_dbg.PairedDeBruijnGraph._find_longest_overlap_brute(a, b) for a, b in job["overlap"]]
    out["break"] = [ref_dbg.DeBruijnGraph._break_read_into_k_minus_one_mers(k, r) for k, r in job["break"]]
    out["break_paired"] = [[list(t) for t in ref_dbg.PairedDeBruijnGraph._break_read_into_k_minus_one_mers(k, tuple(r))]
                           for k, r in job["break_paired"]]
    out["pairwise"] = [[list(t) for t in ref_dbg.AbstractDeBruijnGraph._pairwise(x)] for x in job["pairwise"]]
    out["valid"] = [bool(ref_dbg.AbstractDeBruijnGraph.valid_allowed_error(k, e)) for k, e in job["valid"]]
    return out


def job_nodes(job):
    """A script of operations on one Node / PairedNode; the observable state after every step."""
    node = ref_node.PairedNode(*job["data"]) if job["paired"] else ref_node.Node(job["data"])
    trace = []
    for op, arg in job["ops"]:
        if op == "append":
            node.append_edge(tuple(arg) if job["paired"] else arg)
            got = None
        elif op == "pop":
            try:
                got = node.pop_edge()
                got = [list(got[0]) if isinstance(got[0], tuple) else got[0], got[1]]
            except Exception as exc:        # noqa: BLE001
                got = "raised " + type(exc).__name__
        elif op == "in":
            node.num_edges_in += arg
            got = None
        edges = [list(e) if isinstance(e, tuple) else e for e in node.edges]
        trace.append([got, node.outdegree, node.indegree, edges, node.num_edges_in, node.was_branching])
    return trace


JOBS = {"args": job_args, "graph": job_graph, "helpers": job_helpers, "nodes": job_nodes}


def main():
    jobs = json.load(sys.stdin)
    json.dump([JOBS[j["job"]](j) for j in jobs], sys.stdout)


if __name__ == "__main__":
    main()
