"""CPU-only checks of the product's host side: the C ABI loads and exports every symbol the
header declares, key packing round-trips, input parsing follows the reference's rules, and the
host contig traversal (ga_traverse_contigs) reproduces the oracle on oracle-built graphs.
No kernel is launched here."""
import ctypes as C
import io
import os
import re

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, reads_for
from oracle import py_oracle as po
import recipes


def _lib():
    import ga_native as gn
    if not os.path.exists(gn.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    return gn.lib()


def test_abi_exports_every_declared_symbol():
    import ga_native as gn
    header = open(os.path.join(ROOT, "include", "ga_b200.h")).read()
    declared = set(re.findall(r"^(?:int|void|uint64_t|const char\*)\s+(ga_\w+)\(", header, flags=re.M))
    assert len(declared) >= 25
    lib = _lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(gn.SIGNATURES), declared ^ set(gn.SIGNATURES)
    assert lib.ga_version() >= 100


def test_key_geometry():
    lib = _lib()
    assert [lib.ga_key_words(k, 2) for k in (2, 32, 33, 64, 65)] == [1, 1, 2, 2, 0]
    assert [lib.ga_key_words(k, 5) for k in (5, 13, 14, 26, 27)] == [1, 1, 2, 2, 0]
    assert (lib.ga_slot_bytes(1), lib.ga_slot_bytes(2)) == (16, 32)


def test_no_gpu_fails_loudly():
    import ga_native as gn
    lib = _lib()
    if lib.ga_device_count() > 0:
        pytest.skip("a GPU is visible")
    import debruijn_graph as dg
    with pytest.raises(gn.GaError, match="no CPU fallback"):
        dg.DeBruijnGraph(["ACGTACGT"], k=4, hamming_dist=0)


def test_alphabet_roundtrip():
    from ga_device import Alphabet
    rng = np.random.default_rng(1)
    for symbols, w in ((b"ACGT", 31), (b"ACGT", 63), (b"_abcdefgh;", 9), (bytes(range(48, 48 + 33)), 20)):
        alpha = Alphabet(np.array(sorted(symbols)))
        kw = 1 if w * alpha.sym_bits <= 63 else 2
        words = ["".join(chr(symbols[c]) for c in rng.integers(0, len(symbols), w)) for _ in range(50)]
        keys = np.zeros((len(words), kw), dtype=np.uint64)
        for i, word in enumerate(words):
            lo, hi = alpha.pack_key(word, kw)
            keys[i, 0] = lo
            if kw > 1:
                keys[i, 1] = hi
        assert alpha.decode_strings(keys, w) == words
    assert Alphabet(np.array(sorted(b"ACGT"))).pack_key("ACGN", 1) is None


def test_read_input_rules():
    import assemble
    parse = lambda text: assemble.IOHandler.read_input(io.StringIO(text))   # noqa: E731
    assert parse("3\nACGT\n  CCC \r\nGG\nignored\n") == (["ACGT", "CCC", "GG"], False, 0, 9)
    assert parse("2\nAC|GT|5\nCC|GG|7\n") == ([("AC", "GT"), ("CC", "GG")], True, 7, 8)
    assert parse("0\nACGT\n")[0] == ["ACGT"]                  # n = 0 still consumes one read
    assert parse("3\nACGT\n")[0] == ["ACGT", "", ""]          # missing lines become empty reads
    with pytest.raises(ValueError):
        parse("2\nAC|GT|5\nCCGG\n")
    args = assemble.IOHandler.read_args(["--kmer_length", "28", "--filter", "3", "--stdout", "--paired"])
    assert (args.kmer_length, args.filter_threshold, args.stdout, args.paired) == (28, 3, True, True)


def _traverse_arrays(rowptr, col, indeg, br, last, num_edges, paired, mode=None):
    """ga_traverse_contigs on host arrays -> (contigs, left).  mode: None = the library's own choice,
    "serial" = the edge-by-edge sweep, "parallel" = the piecewise routine forced (it may still decline
    and hand over to the serial one -- the result must be the same either way)."""
    lib = _lib()
    n = len(rowptr) - 1
    env = {"serial": {"GA_TRAVERSE_THREADS": "0"},
           "parallel": {"GA_TRAVERSE_THREADS": "3", "GA_TRAVERSE_MIN_NODES": "0"}}.get(mode, {})
    saved = {key: os.environ.get(key) for key in ("GA_TRAVERSE_THREADS", "GA_TRAVERSE_MIN_NODES")}
    for key in saved:
        os.environ.pop(key, None)
    os.environ.update(env)
    try:
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        col = np.ascontiguousarray(col if len(col) else [0], dtype=np.int32)
        indeg = np.ascontiguousarray(indeg if len(indeg) else [0], dtype=np.int32)
        br = np.ascontiguousarray(br if len(br) else [0], dtype=np.uint8)
        last = np.ascontiguousarray(last if len(last) else [0], dtype=np.uint8)
        text, offs, cnt = C.c_void_p(), C.c_void_p(), C.c_uint64()
        left = np.full(max(n, 1), -7, dtype=np.int32)
        rc = lib.ga_traverse_contigs(rowptr.ctypes.data, col.ctypes.data, indeg.ctypes.data, br.ctypes.data,
                                     last.ctypes.data, n, int(num_edges), int(paired), C.byref(text),
                                     C.byref(offs), C.byref(cnt), left.ctypes.data)
        assert rc == 0
        route = lib.ga_traverse_last_route()
    finally:
        for key, value in saved.items():
            os.environ.pop(key, None)
            if value is not None:
                os.environ[key] = value
    off = np.ctypeslib.as_array(C.cast(offs, C.POINTER(C.c_uint64)), shape=(cnt.value + 1,)).copy()
    raw = C.string_at(text, int(off[-1]))
    lib.ga_free_host(text)
    lib.ga_free_host(offs)
    _traverse_arrays.route = route
    return [raw[int(off[i]):int(off[i + 1])].decode("latin-1") for i in range(cnt.value)], left[:n]


def _graph_arrays(graph):
    n = len(graph.keys)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum([len(s) for s in graph.succ])
    col = np.array([j for s in graph.succ for j in s], dtype=np.int32)
    last = np.array([ord(graph.last_char(i)) for i in range(n)], dtype=np.uint8)
    return rowptr, col, np.array(graph.indeg, dtype=np.int32), np.array(graph.branching, dtype=np.uint8), last


def _traverse(graph):
    """Contigs of an oracle-built graph through ga_traverse_contigs: the serial sweep and the forced
    piecewise routine must agree on the contigs and on what is left of every node."""
    arrays = _graph_arrays(graph)
    serial, left_s = _traverse_arrays(*arrays, graph.num_edges, graph.paired, "serial")
    forced, left_p = _traverse_arrays(*arrays, graph.num_edges, graph.paired, "parallel")
    assert forced == serial and np.array_equal(left_s, left_p)
    return serial


def test_host_traversal_matches_oracle():
    for key in list(GOLDEN["fuzz"])[:120]:
        recipe, k, F = recipes.fuzz_recipe(int(key))
        _, _, graph = po.assemble(reads_for(recipe), k, F, recipe["paired"])
        assert _traverse(graph) == po.contigs(graph), key
    for name in ("two-circles", "homopoly-AC-paired", "kat-f1", "nd-paired-jitter2"):
        gold = GOLDEN["cases"][name]
        _, _, graph = po.assemble(reads_for(gold["recipe"]), gold["k"], gold["F"], gold["recipe"]["paired"])
        got = _traverse(graph)
        assert po.contig_digest(got) == gold["contig_digest"], name


def _chain_graph(rng, n, n_branch, cycles=0):
    """A random graph of the shape the traversal sees: `n_branch` branching nodes joined by chains of
    1-in-1-out nodes (ids shuffled), plus `cycles` rings without any branching node."""
    ids = rng.permutation(n).tolist()
    hubs = [ids.pop() for _ in range(n_branch)]
    rings = []
    for _ in range(cycles):
        size = int(rng.integers(1, 40))
        rings.append([ids.pop() for _ in range(size)])
    succ = [[] for _ in range(n)]
    while ids:
        length = min(len(ids), int(rng.integers(0, 3000)))
        chain = [ids.pop() for _ in range(length)]
        src = hubs[int(rng.integers(len(hubs)))]
        path = [src] + chain
        if rng.random() < 0.8:
            path.append(hubs[int(rng.integers(len(hubs)))])      # else: dead end
        for a, b in zip(path, path[1:]):
            succ[a].append(b)
    for ring in rings:
        for a, b in zip(ring, ring[1:] + ring[:1]):
            succ[a].append(b)
    indeg = np.zeros(n, dtype=np.int32)
    for s in succ:
        for j in s:
            indeg[j] += 1
    br = np.array([len(succ[i]) > 1 or indeg[i] > 1 for i in range(n)], dtype=np.uint8)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum([len(s) for s in succ])
    col = np.array([j for s in succ for j in s], dtype=np.int32)
    last = rng.integers(65, 91, n).astype(np.uint8)
    return rowptr, col, indeg, br, last


def test_piecewise_traversal_equals_the_serial_sweep():
    """Long chains cut at splitter nodes, walked piece by piece on several threads and stitched
    (csrc/ga_traverse.cu) give the serial sweep's contigs; shapes the piecewise routine does not
    take -- rings without a branching node, an edge counter that disagrees with the CSR, in-degree
    bookkeeping that hides a second way into a chain -- fall back to the serial sweep."""
    rng = np.random.default_rng(11)
    for n, hubs, cycles, paired in ((60000, 40, 0, False), (60000, 40, 0, True), (30000, 3, 0, False),
                                    (20000, 25, 3, False), (20000, 25, 3, True), (5000, 1, 0, False)):
        arrays = _chain_graph(rng, n, hubs, cycles)
        m = int(arrays[0][-1])
        serial, left_s = _traverse_arrays(*arrays, m, paired, "serial")
        forced, left_p = _traverse_arrays(*arrays, m, paired, "parallel")
        assert _traverse_arrays.route == (1 if cycles == 0 else 0)      # taken, not declined (rings: declined)
        auto, left_a = _traverse_arrays(*arrays, m, paired)
        assert forced == serial and auto == serial, (n, hubs, cycles, paired)
        assert np.array_equal(left_s, left_p) and np.array_equal(left_s, left_a)
        if cycles == 0:
            assert sum(map(len, serial)) == m and not left_s.any()
        # the reference's own edge counter may differ from the CSR (paired corner cases): early exit
        for attr in (m - 5, m + 5, 3):
            a, la = _traverse_arrays(*arrays, attr, paired, "serial")
            b, lb = _traverse_arrays(*arrays, attr, paired, "parallel")
            assert a == b and np.array_equal(la, lb)
    # bookkeeping that hides a second entry into a chain: node x has in-degree 2 but says 1
    rowptr, col, indeg, br, last = _chain_graph(rng, 4000, 6)
    chain_nodes = [i for i in range(4000) if not br[i] and rowptr[i + 1] - rowptr[i] == 1 and indeg[i] == 1]
    a, b = chain_nodes[10], chain_nodes[500]
    col = col.copy()
    col[rowptr[a]] = b                      # a now also leads to b; b keeps "in-degree 1, not branching"
    m = int(rowptr[-1])
    for paired in (False, True):
        s, ls = _traverse_arrays(rowptr, col, indeg, br, last, m, paired, "serial")
        p, lp = _traverse_arrays(rowptr, col, indeg, br, last, m, paired, "parallel")
        assert s == p and np.array_equal(ls, lp) and _traverse_arrays.route == 0
    # a "not branching" node with two edges
    rowptr, col, indeg, br, last = _chain_graph(rng, 3000, 5)
    hub = int(np.flatnonzero(br)[0])
    br = br.copy()
    br[hub] = 0
    s, ls = _traverse_arrays(rowptr, col, indeg, br, last, int(rowptr[-1]), False, "serial")
    p, lp = _traverse_arrays(rowptr, col, indeg, br, last, int(rowptr[-1]), False, "parallel")
    assert s == p and np.array_equal(ls, lp)


def test_prime_tables_and_hash_helper():
    from countminsketch import CountMinSketch
    for name, want in GOLDEN["prime_tables"].items():
        assert getattr(CountMinSketch, name) == want, name
    for text, want in GOLDEN["murmur3"]:
        assert CountMinSketch._hash(text) == want
    with pytest.raises(AssertionError):
        CountMinSketch(20)


def test_break_helpers():
    import debruijn_graph as dg
    for read, k, want in GOLDEN["break"]["unpaired"]:
        assert dg.DeBruijnGraph._break_read_into_k_minus_one_mers(k, read) == want
    for pair, k, want in GOLDEN["break"]["paired"]:
        got = dg.PairedDeBruijnGraph._break_read_into_k_minus_one_mers(k, tuple(pair))
        assert [list(t) for t in got] == want
    for a, b, want in GOLDEN["overlap"]:
        assert dg.PairedDeBruijnGraph._find_longest_overlap_brute(a, b) == want


def test_bucketed_entries_reject_bad_arguments():
    """Argument validation of the ga_sk_* entries happens before any CUDA call: checkable without a GPU."""
    import ga_native as gn
    L = _lib()
    assert L.ga_sk_minimizer_len(31) == 15 and L.ga_sk_minimizer_len(32) == 16 and L.ga_sk_minimizer_len(5) == 4
    assert L.ga_sk_minimizer_len(21) == 11 and L.ga_sk_minimizer_len(1) == 0
    reads = gn.GaReads()
    reads.n_reads, reads.uniform_len, reads.stride_words, reads.estride = 4, 100, 4, 100
    reads.storage_bits = reads.sym_bits = 2
    dummy = C.create_string_buffer(128)
    ptr = C.c_void_p((C.addressof(dummy) + 31) & ~31)          # the record slots are 32-byte aligned
    assert L.ga_sk_cursor_stride() == 16
    # bucket bits out of range, k beyond 64-bit keys
    assert L.ga_sk_scatter_reads(C.byref(reads), 31, 11, 10, ptr, 16, ptr, ptr, ptr, None) == gn.GA_ERR_BAD_ARG
    assert b"bucket bits" in L.ga_last_error()
    assert L.ga_sk_scatter_reads(C.byref(reads), 40, 2, 2, ptr, 16, ptr, ptr, ptr, None) == gn.GA_ERR_BAD_ARG
    # ordinals beyond 47 bits
    reads.first_read = 1 << 46
    assert L.ga_sk_scatter_reads(C.byref(reads), 31, 2, 2, ptr, 16, ptr, ptr, ptr, None) == gn.GA_ERR_BAD_ARG
    assert b"47 bits" in L.ga_last_error()
    # level-2 pass: exactly one of the dense / index outputs
    assert L.ga_sk_scatter_buckets(ptr, 16, ptr, 2, 2, ptr, ptr, ptr, ptr, None) == gn.GA_ERR_BAD_ARG
    assert L.ga_sk_scatter_buckets(ptr, 16, ptr, 2, 2, ptr, None, None, None, None) == gn.GA_ERR_BAD_ARG
    # bucket pass: table size must be a power of two within the shared-memory pool, threshold within the counter
    args = (ptr, ptr, ptr, 1, ptr, 4, 31, 3)
    tail = (ptr, ptr, 16, ptr, ptr, 16, ptr, None, 0, 0, None)
    assert L.ga_sk_count_build(*args, 3000, 16, *tail) == gn.GA_ERR_BAD_ARG
    assert L.ga_sk_count_build(*args, 1 << 20, 16, *tail) == gn.GA_ERR_BAD_ARG
    assert L.ga_sk_count_build(ptr, ptr, ptr, 1, ptr, 4, 31, 70000, 4096, 16, *tail) == gn.GA_ERR_BAD_ARG
    assert L.ga_sk_count_build(ptr, ptr, ptr, 99, ptr, 4, 31, 3, 4096, 16, *tail) == gn.GA_ERR_BAD_ARG
    assert L.ga_sk_resolve(None, 5, 31, ptr, 16, ptr, ptr, None) == gn.GA_ERR_BAD_ARG
    assert L.ga_sk_spill_scratch_bytes(1024) == 1024 * 84


def test_peer_exchange_entries_reject_bad_arguments():
    """ga_peer_* / ga_sk_push_* / ga_sk_count_build_from validate their arguments before touching CUDA."""
    import ga_native as gn
    L = _lib()
    dummy = C.create_string_buffer(256)
    ptr = C.c_void_p((C.addressof(dummy) + 31) & ~31)
    out = C.c_void_p()
    handle = (C.c_uint8 * 64)()
    assert L.ga_peer_alloc(0, C.byref(out), handle) == gn.GA_ERR_BAD_ARG
    assert L.ga_peer_alloc(1024, None, handle) == gn.GA_ERR_BAD_ARG
    assert L.ga_peer_open(None, C.byref(out)) == gn.GA_ERR_BAD_ARG
    assert L.ga_peer_close(None) == gn.GA_OK and L.ga_peer_free(None) == gn.GA_OK      # nothing to do
    cut = (C.c_uint64 * 3)(0, 10, 20)
    down = (C.c_uint64 * 3)(0, 10, 5)
    targets = (C.c_void_p * 2)(ptr.value, ptr.value)
    holes = (C.c_void_p * 2)(ptr.value, None)
    # more ranks than a node holds, descending cut, a non-empty range without a target
    assert L.ga_sk_push_records(ptr, 16, ptr, ptr, 2, 2, 17, cut, targets, targets, None) == gn.GA_ERR_BAD_ARG
    assert b"1..16 ranks" in L.ga_last_error()
    assert L.ga_sk_push_records(ptr, 16, ptr, ptr, 2, 2, 2, down, targets, targets, None) == gn.GA_ERR_BAD_ARG
    assert L.ga_sk_push_records(ptr, 16, ptr, ptr, 2, 2, 2, cut, holes, targets, None) == gn.GA_ERR_BAD_ARG
    assert b"cut must ascend" in L.ga_last_error()
    assert L.ga_sk_push_sorted(ptr, 16, ptr, 2, 11, ptr, 2, cut, targets, targets, None) == gn.GA_ERR_BAD_ARG
    assert L.ga_sk_push_sorted(ptr, 16, ptr, 2, 2, ptr, 2, down, targets, targets, None) == gn.GA_ERR_BAD_ARG
    # nothing to send: returns before any launch
    empty = (C.c_uint64 * 3)(7, 7, 7)
    assert L.ga_sk_push_records(ptr, 16, ptr, ptr, 2, 2, 2, empty, holes, holes, None) == gn.GA_OK
    # sources form: one source per segment, capacities below 2^25, aligned slots
    src = gn.GaSkSources()
    tail = (ptr, ptr, 4, 31, 3, 4096, 16, ptr, ptr, 16, ptr, ptr, 16, ptr, 2, None)
    assert L.ga_sk_count_build_from(None, *tail) == gn.GA_ERR_BAD_ARG
    src.n_sources = 0
    assert L.ga_sk_count_build_from(C.byref(src), *tail) == gn.GA_ERR_BAD_ARG
    src.n_sources = 2
    src.records[0], src.index[0], src.l1_capacity[0] = ptr.value, ptr.value, 1 << 25
    src.records[1], src.index[1], src.l1_capacity[1] = ptr.value, ptr.value, 16
    assert L.ga_sk_count_build_from(C.byref(src), *tail) == gn.GA_ERR_BAD_ARG
    assert b"sources" in L.ga_last_error()
    assert L.ga_sk_count_build_spill_from(None, ptr, 4, ptr, 1, 31, 3, 4096, ptr, 1, ptr, ptr, 16, ptr, ptr, 2,
                                          None) == gn.GA_ERR_BAD_ARG


def test_output_writers_keep_the_reference_format(tmp_path, monkeypatch, capsys):
    """IOHandler.write_FASTQ / write_stdout (assemble.py:74-99 of the reference): file name, line layout, no trailing
    newline in the file, two blanks after the colon on stdout, exclusive creation.  Where the unmodified reference is at
    hand (oracle/_ref, this container) its own writer must produce the same bytes."""
    import importlib.util
    import time
    import assemble
    monkeypatch.chdir(tmp_path)
    contigs, constants, start = ["ACGTACGT", "TTGA"], (28, 3, 2), 1_700_000_000.0
    assemble.IOHandler.write_FASTQ(contigs, constants, start)
    stamp = time.strftime("%b_%d_%H:%M:%S_%Y", time.localtime(start))
    path = tmp_path / "output" / ("%s_k28_f3_e2.FASTQ" % stamp)
    text = path.read_text()
    lines = text.split("\n")
    assert lines[0] == ">Time started: " + time.strftime("%c", time.localtime(start))
    assert lines[1:6] == [">Number of contigs: 2", ">CONTIG1", "ACGTACGT", ">CONTIG2", "TTGA"]
    assert lines[6].startswith(">Time finished: ") and len(lines) == 7 and not text.endswith("\n")
    with pytest.raises(FileExistsError):                       # mode 'x', as upstream
        assemble.IOHandler.write_FASTQ(contigs, constants, start)
    assemble.IOHandler.write_stdout(contigs, constants, start)
    out = capsys.readouterr().out.split("\n")
    assert out[0] == ">Time started:  " + time.strftime("%c", time.localtime(start))
    assert out[1:6] == [">Number of contigs:  2", ">CONTIG1", "ACGTACGT", ">CONTIG2", "TTGA"] and out[7] == ""
    ref_path = os.path.join(ROOT, "oracle", "_ref", "assemble.py")
    if os.path.exists(ref_path):
        spec = importlib.util.spec_from_file_location("reference_assemble", ref_path)
        ref = importlib.util.module_from_spec(spec)
        monkeypatch.syspath_prepend(os.path.dirname(ref_path))
        spec.loader.exec_module(ref)
        other = tmp_path / "ref"
        other.mkdir()
        monkeypatch.chdir(other)
        ref.IOHandler.write_FASTQ(contigs, constants, start)
        theirs = (other / "output" / path.name).read_text().split("\n")
        assert theirs[:6] == lines[:6] and theirs[6].startswith(">Time finished: ") and len(theirs) == 7
        ref.IOHandler.write_stdout(contigs, constants, start)
        assert capsys.readouterr().out.split("\n")[:6] == out[:6]


def test_node_records_weigh_what_the_reference_says():
    """The -m report sums sys.getsizeof over the Node objects (debruijn_node.py:35-53 of the reference, incl. the
    cached second answer that is one int larger than the first).  Where the unmodified reference is at hand
    (oracle/_ref, this container) its classes must report the same numbers in the same interpreter."""
    import importlib.util
    import sys
    import debruijn_node as ours
    node = ours.Node("ACGT")
    node.append_edge("CGTA")
    node.append_edge("CGTC")
    first = sys.getsizeof(node)
    assert sys.getsizeof(node) == first + sys.getsizeof(first)       # the cached answer: one int more
    assert node.pop_edge() == ("CGTC", True) and node.outdegree == 1 and node.indegree == 0
    ref_path = os.path.join(ROOT, "oracle", "_ref", "debruijn_node.py")
    if not os.path.exists(ref_path):
        return
    spec = importlib.util.spec_from_file_location("reference_debruijn_node", ref_path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    def weigh(mod, paired, edges):
        n = mod.PairedNode("ACGT", "TTGA") if paired else mod.Node("ACGT")
        for e in edges:
            n.append_edge(e)
        n.num_edges_in = 2
        return sys.getsizeof(n), sys.getsizeof(n)

    for paired in (False, True):
        for edges in ([], ["CGTA"], ["CGTA", "CGTC", "CGTG"], [("CGTA", "TGAC")], [("CGTA", "TGAC"), ("CGTT", "TGAA")]):
            assert weigh(ours, paired, edges) == weigh(ref, paired, edges), (paired, edges)


def test_graph_checksum_host_and_tensor_paths_agree():
    """BuiltGraph.checksum(): the same value from the host arrays and from tensors holding the same numbers (the
    form a `to_host=False` build keeps on the device), keys with the top bit set included; sensitive to order."""
    import torch
    import ga_device as gd
    rng = np.random.default_rng(0)
    n = 1000
    rowptr = np.zeros(n + 1, np.int32)
    rowptr[1:] = np.cumsum(rng.integers(0, 3, n))
    m = int(rowptr[-1])
    g = gd.BuiltGraph(False, 30, None, 1)
    g.n_nodes, g.n_edges = n, m
    g.rowptr, g.col, g.indeg = rowptr, rng.integers(0, n, m).astype(np.int32), rng.integers(0, 3, n).astype(np.int32)
    g.branching = rng.integers(0, 2, n).astype(np.uint8)
    g.keys_a = rng.integers(0, 2 ** 62, (n, 1)).astype(np.uint64)
    g.keys_a[3, 0] = 2 ** 64 - 5
    want = g.checksum()
    t = gd.BuiltGraph(False, 30, None, 1)
    t.n_nodes, t.n_edges = n, m
    t.device = dict(rowptr=torch.from_numpy(rowptr), col=torch.from_numpy(g.col), indeg=torch.from_numpy(g.indeg),
                    branching=torch.from_numpy(g.branching), keys_a=torch.from_numpy(g.keys_a.view(np.int64)), keys_b=None)
    assert t.checksum() == want and len(want) == 16
    i = int(np.flatnonzero(g.col[:-1] != g.col[1:])[0])
    g.col[[i, i + 1]] = g.col[[i + 1, i]]
    assert g.checksum() != want
    assert gd.BuiltGraph(True, 27, None, 1).checksum() == gd.BuiltGraph(True, 27, None, 1).checksum()     # empty graph


def test_container_sizes_follow_the_interpreter():
    """py_sizes: what sys.getsizeof reports for a list grown by appends and a defaultdict(int) grown by string-key
    insertions -- the containers the reference's -m report weighs -- against the real objects, entry by entry; and
    the product's stand-ins (RawReads, KmerCounts) answering with those numbers."""
    import sys
    from collections import defaultdict
    import py_sizes
    import ga_ingest
    import ga_device as gd
    grown, counts = [], defaultdict(int)
    assert py_sizes.appended_list_sizeof(0) == grown.__sizeof__()
    assert py_sizes.grown_str_dict_sizeof(0) == counts.__sizeof__()
    for n in range(1, 70001):
        grown.append(None)
        counts["ACGT%d" % n] += 1
        assert py_sizes.appended_list_sizeof(n) == grown.__sizeof__(), n
        assert py_sizes.grown_str_dict_sizeof(n) == counts.__sizeof__(), n
    big = []
    for i in range(3 * 10 ** 6):                          # one large point each (C3 has 1.4 M reads, 15 M distinct windows)
        big.append(None)
    assert py_sizes.appended_list_sizeof(len(big)) == big.__sizeof__()
    big = defaultdict(int)
    for i in range(1500000):
        big[str(i)] += 1
    assert py_sizes.grown_str_dict_sizeof(1500000) == big.__sizeof__()
    reads, _, _, _ = ga_ingest.parse(b"5\nAC\nGT\nAA\nCC\nGG\n")
    five = []
    for r in ("AC", "GT", "AA", "CC", "GG"):
        five.append(r)
    assert sys.getsizeof(reads) == sys.getsizeof(five)

    class Counted(gd.KmerCounts):
        def __init__(self, n):
            self._n = n

        def __len__(self):
            return self._n

    assert sys.getsizeof(Counted(70000)) == sys.getsizeof(counts)
    assert sys.getsizeof(Counted(0)) == sys.getsizeof(defaultdict(int))


def test_bucket_limit_switches_to_the_table_kernels(monkeypatch):
    """A bucket beyond what the shared-memory pass can name (GaBucketLimit: one window repeated tens of millions of
    times) must not end the build: build_graph frees the buckets' workspace and goes on with the global-table
    kernels.  Control flow only (the kernels themselves are stubbed; no GPU here)."""
    import types
    import torch
    import ga_device as gd
    import ga_native as gn

    class Reached(Exception):
        pass

    calls = []

    def refuse(*a, **kw):
        calls.append("buckets")
        raise gn.GaBucketLimit("bucketed count: a bucket holds 2^25 records or more (one repeated window?)")

    def table_route(counts, threshold, sketch):
        calls.append("tables")
        raise Reached()

    monkeypatch.setattr(gd, "_dev", lambda: torch.device("cpu"))
    monkeypatch.setattr(gd, "superkmer_stamps", refuse)
    monkeypatch.setattr(gd, "superkmer_solid", refuse)
    monkeypatch.setattr(gd, "_solid_keys", table_route)
    monkeypatch.setattr(gd, "release_workspace", lambda: calls.append("release"))
    assert issubclass(gn.GaBucketLimit, gn.GaError)           # callers that catch GaError still see it
    alphabet = gd.Alphabet(np.zeros(0))
    for paired, n_occ in ((False, 1 << 23), (True, 1 << 29)):
        reads = types.SimpleNamespace(paired=paired, alphabet=alphabet, first_read=0, n_reads=1000, estride=150,
                                      status=None)
        counts = types.SimpleNamespace(k=31, w=30, key_words=1, _table=None, _cand={}, n_occ=n_occ, slot_bytes=16)
        calls.clear()
        with pytest.raises(Reached):
            gd.build_graph(counts, reads, 3)
        assert calls == ["buckets", "release", "tables"]
    # the ordinary case: the bucketed route answers, the table route is not asked
    calls.clear()
    monkeypatch.setattr(gd, "superkmer_stamps", lambda *a, **kw: (calls.append("buckets"), (torch.zeros((1, 1), dtype=torch.int64), 0, None))[1])
    monkeypatch.setattr(gd, "superkmer_solid", lambda *a, **kw: (calls.append("solid"), (torch.zeros((1, 1), dtype=torch.int64), 0))[1])
    for paired, n_occ, want in ((False, 1 << 23, ["buckets"]), (True, 1 << 29, ["solid"]), (False, 1 << 10, ["tables"])):
        reads = types.SimpleNamespace(paired=paired, alphabet=alphabet, first_read=0, n_reads=1000, estride=150, status=None)
        counts = types.SimpleNamespace(k=31, w=30, key_words=1, _table=None, _cand={}, n_occ=n_occ, slot_bytes=16)
        calls.clear()
        if want == ["tables"]:
            with pytest.raises(Reached):
                gd.build_graph(counts, reads, 3)
        else:
            graph = gd.build_graph(counts, reads, 3)
            assert graph.n_nodes == 0 and graph.paired == paired
        assert calls == want
    # any other failure of the bucketed route is not swallowed
    monkeypatch.setattr(gd, "superkmer_stamps", lambda *a, **kw: (_ for _ in ()).throw(gn.GaError("something else")))
    reads = types.SimpleNamespace(paired=False, alphabet=alphabet, first_read=0, n_reads=1000, estride=150, status=None)
    counts = types.SimpleNamespace(k=31, w=30, key_words=1, _table=None, _cand={}, n_occ=1 << 23, slot_bytes=16)
    with pytest.raises(gn.GaError, match="something else"):
        gd.build_graph(counts, reads, 3)
