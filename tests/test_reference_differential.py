"""Live differential tests against the UNMODIFIED reference (oracle/_ref, a verbatim copy of /root/reference made by
oracle/make_ref.py; run in a separate process by tests/ref_worker.py because its modules carry the product's names).

Fresh random inputs every time the seeds below change -- the committed golden vectors (tests/golden/golden.json) pin
the named cases and 206 fuzz graphs; these tests widen that to inputs nobody has looked at: both oracles on new
graphs (ragged reads, 2-5 letter alphabets, pairs with jitter, k from 3 up) and the product's host contig traversal
over them, the CLI's argument parser, the static
helpers, and the node records.  CPU only.  Skipped where oracle/_ref is absent (the GPU box)."""
import json
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "debruijn_graph.py")),
                                reason="oracle/_ref is only present where /root/reference is")


def ask_reference(jobs):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_worker.py"), REF_DIR],
                         input=json.dumps(jobs), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout)


def _random_case(i: int):
    """Small read sets that force repeats, branching, fuzzy groups, ragged reads and perfect cycles; other shapes
    than tests/golden/recipes.fuzz_recipe (longer reads, larger k, more coverage spread, circular and linear)."""
    rng = random.Random(910000 + i)
    alphabet = rng.choice(["AC", "ACG", "ACGT", "ACGT", "ACGTN", "ab_;"])
    glen = rng.randint(8, 160)
    genome = "".join(rng.choice(alphabet) for _ in range(glen))
    if rng.random() < 0.2:                                   # tandem repeat: cycles and branching
        unit = genome[:rng.randint(2, 7)]
        genome = (unit * (glen // len(unit) + 1))[:glen]
    paired = bool(i % 3 == 0)
    L = rng.randint(5, 24)
    k = rng.randint(3, min(L, 16))
    F = rng.choice([0, 0, 1, 2, 3, 5])
    N = rng.randint(1, 160)
    circ = genome * (2 + (L + 12) // glen + 1)
    reads = []
    for _ in range(N):
        s = rng.randrange(glen)
        r1 = circ[s:s + L]
        if rng.random() < 0.25:
            p = rng.randrange(L)
            r1 = r1[:p] + rng.choice(alphabet) + r1[p + 1:]
        if paired:
            s2 = (s + rng.randint(2, 10) + rng.randint(-2, 2)) % glen
            r2 = circ[s2:s2 + L]
            if rng.random() < 0.1:
                p = rng.randrange(L)
                r2 = r2[:p] + rng.choice(alphabet) + r2[p + 1:]
            reads.append([r1, r2])
        else:
            if rng.random() < 0.12:
                r1 = r1[:rng.randint(0, L)]
            reads.append(r1)
    return {"job": "graph", "reads": reads, "paired": paired, "k": k, "F": F}


def _random_case_long(i: int):
    """Longer reads and windows of 16 ... 71 symbols (64-bit, 128-bit and -- beyond what the GPU path takes -- wider
    keys), pairs a few dozen symbols apart."""
    rng = random.Random(555000 + i)
    alphabet = rng.choice(["ACGT", "ACGT", "AC", "ACGTN"])
    glen = rng.randint(60, 400)
    genome = "".join(rng.choice(alphabet) for _ in range(glen))
    paired = i % 2 == 0
    L = rng.randint(30, 90)
    k = rng.randint(17, min(L, 72))
    F = rng.choice([0, 1, 2, 3])
    circ = genome * (3 + (L + 40) // glen)
    reads = []
    for _ in range(rng.randint(20, 200)):
        s = rng.randrange(glen)
        r1 = circ[s:s + L]
        if rng.random() < 0.2:
            p = rng.randrange(L)
            r1 = r1[:p] + rng.choice(alphabet) + r1[p + 1:]
        if paired:
            s2 = (s + rng.randint(5, 30) + rng.randint(-2, 2)) % glen
            reads.append([r1, circ[s2:s2 + L]])
        else:
            if rng.random() < 0.1:
                r1 = r1[:rng.randint(0, L)]
            reads.append(r1)
    return {"job": "graph", "reads": reads, "paired": paired, "k": k, "F": F}


def test_both_oracles_on_fresh_graphs():
    from oracle import c_oracle as co
    from oracle import py_oracle as po
    from helpers import counts_sha
    from test_host_side import _traverse
    jobs = [_random_case(i) for i in range(800)] + [_random_case_long(i) for i in range(200)]
    answers = ask_reference(jobs)
    shapes = set()
    for job, want in zip(jobs, answers):
        assert "error" not in want, want
        reads = [tuple(r) for r in job["reads"]] if job["paired"] else job["reads"]
        tally, _, graph = po.assemble(reads, job["k"], job["F"], job["paired"])
        assert counts_sha(tally.items()) == want["counts_sha"] and len(tally) == want["n_distinct"]
        assert (len(graph.keys), graph.num_edges, graph.digest()) == \
            (want["n_nodes"], want["num_edges"], want["graph_digest"]), job
        assert po.contigs(graph) == want["contigs"], job
        # the product's host traversal (ga_traverse_contigs: serial sweep and forced piecewise walk) over the same graph
        assert _traverse(graph) == want["contigs"], job
        res = co.assemble(reads, job["k"], job["F"], job["paired"])
        assert counts_sha(res.counts_dict().items()) == want["counts_sha"]
        assert (res.n_nodes, res.num_edges, res.digest()) == (want["n_nodes"], want["num_edges"], want["graph_digest"]), job
        assert res.contigs() == want["contigs"], job
        res.close()
        shapes.add((job["paired"], want["n_nodes"] > 0, len(want["contigs"]) > 1))
    assert len(shapes) >= 6            # empty and non-empty graphs, one and several contigs, both kinds


def test_cli_argument_parser():
    """assemble.py's flags, abbreviations and refusals (assemble.py:13-37 of the reference).  `-p/--paired` is the one
    deliberate addition (the reference exits with status 2 on it; BASELINE.json's north star names it)."""
    import assemble
    rng = random.Random(4)
    pool = ["-t", "-m", "-c", "-s", "--time", "--memory", "--count_min_sketch", "--stdout", "--std", "--mem", "--count",
            "-k", "--kmer_length", "--kmer", "--k", "-f", "--filter_threshold", "--filter", "--fil", "--f", "-e",
            "--error", "--err", "--e", "-ts", "-tmcs", "-k31", "-f3", "--kmer_length=28", "--filter=2", "-x", "--bogus",
            "extra", "--t", "--m", "--s", "--c"]
    values = ["31", "3", "0", "-1", "x", "2.5", "", "28"]
    vectors = [[], ["-k", "28", "-f", "3", "-s"], ["--kmer_length", "28", "--filter", "3", "--stdout"], ["-h"]]
    for _ in range(400):
        argv = []
        for _piece in range(rng.randint(0, 5)):
            flag = rng.choice(pool)
            argv.append(flag)
            if flag.lstrip("-") and flag in ("-k", "--kmer_length", "--kmer", "--k", "-f", "--filter_threshold", "--filter",
                                             "--fil", "--f", "-e", "--error", "--err", "--e") and rng.random() < 0.85:
                argv.append(rng.choice(values))
        vectors.append(argv)
    answers = ask_reference([{"job": "args", "argv": v} for v in vectors])
    import contextlib
    import io
    for argv, want in zip(vectors, answers):
        try:
            with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
                ns = vars(assemble.IOHandler.read_args(argv))
            assert ns.pop("paired") is False
            got = {"ok": ns}
        except SystemExit as exc:
            got = {"exit": exc.code}
        assert got == want, argv
    with contextlib.redirect_stderr(io.StringIO()):
        assert vars(assemble.IOHandler.read_args(["-p", "-k", "5"]))["paired"] is True
        assert vars(assemble.IOHandler.read_args(["--paired"]))["paired"] is True


def test_static_helpers():
    """_hash (MurmurHash3_x86_32 incl. seeds and bytes above 0x7f), the overlap rule, read breaking (also reads shorter
    than a window and ragged mates), _pairwise, valid_allowed_error -- countminsketch.py:46-95,
    debruijn_graph.py:36-44, 154-157, 336-347, 369-374 of the reference."""
    import countminsketch
    import debruijn_graph as dg
    rng = random.Random(8)

    def word(alphabet, lo, hi):
        return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))

    job = {"job": "helpers",
           "hash": [[word("ACGT", 0, 70), 0] for _ in range(300)] +
                   [[word("ACGTN_;ab\x7f\x80\xe9\xff", 0, 40), rng.choice([0, 1, 42, 0xFFFFFFFF, 0x9747B28C])] for _ in range(300)],
           "overlap": [[word("AC", 0, 9), word("AC", 0, 9)] for _ in range(600)] +
                      [[word("ACGT", 3, 30), word("ACGT", 3, 30)] for _ in range(100)],
           "break": [[rng.randint(2, 12), word("ACGT", 0, 20)] for _ in range(300)],
           "break_paired": [[rng.randint(2, 12), [word("ACGT", 0, 20), word("ACGT", 0, 20)]] for _ in range(300)],
           "pairwise": [[rng.randrange(100) for _ in range(rng.randint(0, 9))] for _ in range(50)],
           "valid": [[rng.randint(-2, 40), rng.randint(-2, 40)] for _ in range(100)]}
    same_len = [[word("ACGT", n, n), word("ACGT", n, n)] for n in range(3, 40) for _ in range(4)]
    job["overlap"] += same_len
    (want,) = ask_reference([job])
    assert [countminsketch.CountMinSketch._hash(s, seed) for s, seed in job["hash"]] == want["hash"]
    assert [dg.PairedDeBruijnGraph._find_longest_overlap_brute(a, b) for a, b in job["overlap"]] == want["overlap"]
    assert [dg.DeBruijnGraph._break_read_into_k_minus_one_mers(k, r) for k, r in job["break"]] == want["break"]
    assert [[list(t) for t in dg.PairedDeBruijnGraph._break_read_into_k_minus_one_mers(k, tuple(r))]
            for k, r in job["break_paired"]] == want["break_paired"]
    assert [[list(t) for t in dg.AbstractDeBruijnGraph._pairwise(x)] for x in job["pairwise"]] == want["pairwise"]
    assert [bool(dg.AbstractDeBruijnGraph.valid_allowed_error(k, e)) for k, e in job["valid"]] == want["valid"]


def test_node_records_behave_alike():
    """Node / PairedNode under random scripts of append_edge / pop_edge / in-degree changes: the same observable
    state after every step (debruijn_node.py:4-33, 56-63), popping from an empty node included."""
    import debruijn_node as ours
    rng = random.Random(15)
    jobs = []
    for i in range(120):
        paired = bool(i & 1)
        ops = []
        for _ in range(rng.randint(1, 14)):
            r = rng.random()
            if r < 0.5:
                e = "".join(rng.choice("AC") for _ in range(3))
                ops.append(["append", [e, e[::-1]] if paired else e])
            elif r < 0.85:
                ops.append(["pop", None])
            else:
                ops.append(["in", rng.randint(0, 3)])
        jobs.append({"job": "nodes", "paired": paired, "data": ["ACG", "TTA"] if paired else "ACG", "ops": ops})
    answers = ask_reference(jobs)
    for job, want in zip(jobs, answers):
        node = ours.PairedNode(*job["data"]) if job["paired"] else ours.Node(job["data"])
        trace = []
        for op, arg in job["ops"]:
            got = None
            if op == "append":
                node.append_edge(tuple(arg) if job["paired"] else arg)
            elif op == "pop":
                try:
                    got = node.pop_edge()
                    got = [list(got[0]) if isinstance(got[0], tuple) else got[0], got[1]]
                except Exception as exc:        # noqa: BLE001
                    got = "raised " + type(exc).__name__
            else:
                node.num_edges_in += arg
            edges = [list(e) if isinstance(e, tuple) else e for e in node.edges]
            trace.append([got, node.outdegree, node.indegree, edges, node.num_edges_in, node.was_branching])
        assert trace == want, job


def _graph_digest(graph, paired):
    """tests/ref_worker.graph_digest, on the product's graph object."""
    import hashlib
    h = hashlib.sha256()
    if paired:
        it = (((a, b), node) for a, inner in graph.nodes.items() for b, node in inner.items())
    else:
        it = iter(graph.nodes.items())
    n = 0
    for key, node in it:
        h.update(repr((key, list(node.edges), node.num_edges_in, node.was_branching)).encode())
        n += 1
    return n, h.hexdigest()[:16]


def _product_graph(job, graph):
    """The product's graph object (debruijn_graph.DeBruijnGraph / PairedDeBruijnGraph) around a CSR that was NOT built
    on the GPU: the arrays come from the oracle's graph, packed the way ga_csr_emit hands them over.
    Everything above the CSR -- lazy `nodes`, Node objects, traversal on the CSR and on objects -- is the product's."""
    import numpy as np
    import debruijn_graph as dg
    import ga_device as gd
    from test_host_side import _graph_arrays
    paired = job["paired"]
    symbols = sorted({ord(ch) for r in job["reads"] for ch in ("".join(r) if paired else r)})
    alphabet = gd.Alphabet(np.array(symbols))
    w = job["k"] - 1
    kw = 1 if w * alphabet.sym_bits <= 64 else 2
    csr = gd.BuiltGraph(paired, w, alphabet, kw)
    rowptr, col, indeg, branching, last = _graph_arrays(graph)
    csr.n_nodes, csr.n_edges, csr.num_edges_attr = len(graph.keys), len(col), graph.num_edges
    csr.rowptr, csr.col, csr.indeg, csr.branching, csr.last_char = rowptr, col, indeg, branching, last

    def pack(strings):
        out = np.zeros((len(strings), kw), dtype=np.uint64)
        for i, text in enumerate(strings):
            lo, hi = alphabet.pack_key(text, kw)
            out[i, 0] = lo
            if kw > 1:
                out[i, 1] = hi
        return out

    if paired:
        csr.keys_a, csr.keys_b = pack([a for a, _ in graph.keys]), pack([b for _, b in graph.keys])
    else:
        csr.keys_a = pack(list(graph.keys))
    cls = dg.PairedDeBruijnGraph if paired else dg.DeBruijnGraph
    g = cls.__new__(cls)
    g.KMER_LEN, g.HAMMING_DIST = job["k"], job["F"]
    g.num_edges = csr.num_edges_attr
    g._csr, g._nodes, g._nodes_dirty, g._left = csr, None, False, None
    return g


def test_node_facade_and_both_traversals_on_fresh_graphs():
    """What a user of the reference sees above the CSR: `graph.nodes` (dict / dict of dicts of Node objects, in the
    reference's order, with edges, in-degrees and was_branching), `enumerate_contigs()` on the CSR and on the
    objects (after somebody looked at `nodes`), `num_edges` and what is left of `nodes` afterwards."""
    from oracle import py_oracle as po
    jobs = [_random_case(i) for i in range(5000, 5400)]
    answers = ask_reference(jobs)
    for job, want in zip(jobs, answers):
        reads = [tuple(r) for r in job["reads"]] if job["paired"] else job["reads"]
        _, _, graph = po.assemble(reads, job["k"], job["F"], job["paired"])
        # 1. contigs from the CSR first, then the nodes: only what was not popped is left
        g = _product_graph(job, graph)
        assert g.num_edges == want["num_edges"]
        assert g.enumerate_contigs() == want["contigs"], job
        assert g.num_edges == want["edges_left"]
        assert _graph_digest(g, job["paired"]) == (want["n_nodes"], want["digest_after"]), job
        assert g.enumerate_contigs() == []          # as upstream: nothing is left to walk
        # 2. nodes first (materialised Node objects), then the traversal on those objects
        g = _product_graph(job, graph)
        assert _graph_digest(g, job["paired"]) == (want["n_nodes"], want["graph_digest"]), job
        assert g.enumerate_contigs() == want["contigs"], job
        assert g.num_edges == want["edges_left"]
        assert _graph_digest(g, job["paired"]) == (want["n_nodes"], want["digest_after"]), job


def _install_cpu_shim(monkeypatch):
    """TEST-ONLY stand-ins for the two GPU hooks, so that everything AROUND them (the CLI, the -t / -m mixin, hook
    order, output writers) can be run next to the reference on a machine without a GPU: `_count_kmers` answers from
    the Python oracle, `_build_graph` installs a CSR taken from the oracle's graph (`_product_graph`).  The product
    itself has no such path."""
    import debruijn_graph as dg
    from oracle import py_oracle as po

    def count_for(paired):
        def count(k, reads):
            reads = list(reads)
            return po.count_paired(k, reads) if paired else po.count_unpaired(k, reads)
        return staticmethod(count)

    def build(self, kmer_counts, reads):
        reads = list(reads)
        _, _, graph = po.assemble(reads, self.KMER_LEN, self.HAMMING_DIST, self._PAIRED)
        job = {"paired": self._PAIRED, "reads": reads, "k": self.KMER_LEN, "F": self.HAMMING_DIST}
        made = _product_graph(job, graph)
        self._csr, self.num_edges = made._csr, made.num_edges
        self._nodes, self._nodes_dirty, self._left = None, False, None

    monkeypatch.setattr(dg.DeBruijnGraph, "_count_kmers", count_for(False))
    monkeypatch.setattr(dg.PairedDeBruijnGraph, "_count_kmers", count_for(True))
    monkeypatch.setattr(dg._GpuGraphBase, "_build_graph", build)


def _mask(lines, sizes_too=()):
    import re
    out = []
    for line in lines:
        if line.startswith(">Time"):
            continue
        line = re.sub(r"T ?= ?[0-9]+\.[0-9]+", "T=#", line)
        if any(line.startswith(label) for label in sizes_too):
            line = re.sub(r"[0-9,]+$", "#", line)
        out.append(line)
    return out


@pytest.mark.parametrize("flags", [["-s"], ["-s", "-t"], ["-s", "-m"], ["-s", "-t", "-m"], [], ["-t"]])
def test_cli_around_the_hooks(flags, monkeypatch, capsys, tmp_path):
    """The product's assemble.main() next to the reference's own CLI (a subprocess of oracle/_ref/assemble.py) on the
    same stdin: every printed line -- banners of -t, size reports of -m, the contig report on stdout or in
    ./output/*.FASTQ -- with wall-clock values masked.  Of the -m sizes only those of the k-mer counts are masked (the test-only
    shim returns the oracle's dict; the product's device facade is weighed in tests/test_host_side.py); read container,
    read strings, graph and node sizes must agree."""
    import io
    import assemble
    _install_cpu_shim(monkeypatch)
    hidden = (">SIZE OF COUNTS CONTAINER", ">SIZE OF STRINGS IN COUNTS")      # the shim's dict, not the product's facade
    rng = random.Random(77)
    for i in (3, 4, 6, 7, 9, 12):                  # pairs and plain reads
        job = _random_case(7000 + i)
        if job["k"] < 4:
            job["k"] = 4
        lines = ["|".join(r) + "|7" if job["paired"] else r for r in job["reads"]]
        text = "%d\n%s\n" % (len(lines), "\n".join(lines))
        argv = flags + ["-k", str(job["k"]), "-f", str(job["F"])] + (["-e", "1"] if rng.random() < 0.5 else [])
        ref_dir = tmp_path / ("ref%d" % i)
        our_dir = tmp_path / ("our%d" % i)
        ref_dir.mkdir()
        our_dir.mkdir()
        ref = subprocess.run([sys.executable, os.path.join(REF_DIR, "assemble.py")] + argv, input=text,
                             capture_output=True, text=True, cwd=ref_dir, timeout=300)
        assert ref.returncode == 0, ref.stderr[-1500:]
        monkeypatch.chdir(our_dir)
        monkeypatch.setattr(sys, "stdin", io.TextIOWrapper(io.BytesIO(text.encode("ascii")), encoding="ascii"))
        capsys.readouterr()
        assemble.main(argv)
        ours = capsys.readouterr().out
        assert _mask(ours.split("\n"), hidden) == _mask(ref.stdout.split("\n"), hidden), (argv, job)
        if "-s" not in flags:
            (theirs,) = list((ref_dir / "output").iterdir())
            (mine,) = list((our_dir / "output").iterdir())
            assert mine.name.split("_k")[1] == theirs.name.split("_k")[1]            # ..._k{K}_f{F}_e{E}.FASTQ
            assert _mask(mine.read_text().split("\n")) == _mask(theirs.read_text().split("\n"))
