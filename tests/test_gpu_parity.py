"""Parity of the CUDA path (through the C ABI / the reference-shaped classes) with the oracle and
with the golden vectors the unmodified reference produced.  Bit-exact: counts, sketch cells,
node order, edge order, in-degrees, was_branching, contigs."""
import hashlib
import io
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, counts_sha, reads_for, sha16
from oracle import py_oracle as po
from oracle import c_oracle as co
import recipes

pytestmark = pytest.mark.gpu

CLASS_BY_NAME = {}


def _classes():
    if not CLASS_BY_NAME:
        import debruijn_graph as dg
        for name in ("DeBruijnGraph", "CMSDeBruijnGraph", "PairedDeBruijnGraph", "CMSPairedDeBruijnGraph"):
            CLASS_BY_NAME[name] = getattr(dg, name)
    return CLASS_BY_NAME


def digest_of(graph) -> str:
    """Graph digest from the CSR a product graph object holds."""
    csr = graph._csr
    keys = csr.node_strings()
    h = hashlib.sha256()
    rowptr, col = csr.rowptr, csr.col
    for i, key in enumerate(keys):
        edges = [keys[j] for j in col[rowptr[i]:rowptr[i + 1]]]
        h.update(repr((key, edges, int(csr.indeg[i]), bool(csr.branching[i]))).encode())
    return h.hexdigest()[:16]


def check_against_gold(name, check_counts=True):
    gold = GOLDEN["cases"][name]
    reads = reads_for(gold["recipe"])
    cls = _classes()[gold["cls"]]
    if check_counts and "counts_sha" in gold:
        counts = cls._count_kmers(gold["k"], reads)
        assert len(counts) == gold["n_distinct"]
        assert counts.summary(gold["F"])[1] == gold["n_solid"]
        assert counts.summary(gold["F"])[2] == gold["n_occ"]
        assert counts_sha(counts.items()) == gold["counts_sha"]
        if "sketch_rows" in gold:
            sk = cls._make_sketch(counts)
            assert sk.num_rows == gold["sketch_rows"]
            assert [sha16(row.tobytes()) for row in sk.hash_values] == gold["sketch_row_sha"]
    graph = cls(reads, k=gold["k"], hamming_dist=gold["F"])
    assert graph._csr.n_nodes == gold["n_nodes"]
    assert graph.num_edges == gold["num_edges"]
    assert digest_of(graph) == gold["graph_digest"]
    contigs = graph.enumerate_contigs()
    assert len(contigs) == gold["n_contigs"]
    assert po.contig_digest(contigs) == gold["contig_digest"]
    assert graph.num_edges == gold["edges_left"]
    if "contigs" in gold:
        assert contigs == gold["contigs"]


SMALL = ["kat-f1", "kat-f0", "homopoly-A-paired", "homopoly-AC-paired", "homopoly-unpaired", "two-circles",
         "ragged", "toy-unpaired", "toy-paired"]
MEDIUM = ["nd-unpaired", "nd-paired", "nd-paired-jitter2", "nd-unpaired-s1", "nd-unpaired-k32",
          "nd-unpaired-k33", "nd-unpaired-k41", "nd-unpaired-k64", "nd-paired-k35", "nd-paired-k64",
          # the bench's input class; the reference itself produced these digests (make_golden.py, kind "splitmix")
          "mix-c4-sample", "mix-c4-sample-k32", "mix-c3-pairs", "mix-c5-pairs-k41"]
SKETCHED = ["kat-f1-cms", "toy-unpaired-cms", "nd-unpaired-s1-cms", "nd-paired-cms8", "mix-c4-sample-cms"]


@pytest.mark.parametrize("name", SMALL)
def test_small_golden(name):
    check_against_gold(name)


@pytest.mark.parametrize("name", MEDIUM)
def test_medium_golden(name):
    check_against_gold(name)


@pytest.mark.parametrize("name", SKETCHED)
def test_sketch_golden(name):
    check_against_gold(name)


@pytest.mark.parametrize("name", ["mix-c4-sample", "mix-c4-sample-k32"])
@pytest.mark.parametrize("shape", ["default", "multi-pass"])
def test_bucketed_kernels_on_the_bench_input_class_against_the_reference(name, shape, monkeypatch):
    """BASELINE config C4's input class (150-bp reads, 1 % per-base substitutions, 300x coverage, k = 31 / 32, F = 3)
    through the kernels bench.py times (sk_scatter_reads -> level-2 index -> sk_bucket -> sk_resolve -> CSR) against
    the digests the UNMODIFIED reference produced on the same reads -- no oracle in between."""
    import ga_device as gd
    monkeypatch.setattr(gd, "SUPERKMER_MIN_OCC", 0)
    if shape == "multi-pass":
        monkeypatch.setattr(gd, "SUPERKMER_TABLE_SLOTS", 1024)
    calls = []
    real = gd.sk_bucket_pass
    monkeypatch.setattr(gd, "sk_bucket_pass", lambda *a, **kw: calls.append(1) or real(*a, **kw))
    check_against_gold(name, check_counts=False)
    assert calls, "the bucketed kernels did not run"


def test_fuzz_golden():
    """206 small random graphs (tiny alphabets, jittered pairs, homopolymers, cycles)."""
    bad = []
    for key, gold in GOLDEN["fuzz"].items():
        recipe, k, F = recipes.fuzz_recipe(int(key))
        cls = _classes()["PairedDeBruijnGraph" if recipe["paired"] else "DeBruijnGraph"]
        graph = cls(reads_for(recipe), k=k, hamming_dist=F)
        got = (graph._csr.n_nodes, graph.num_edges, digest_of(graph))
        contigs = graph.enumerate_contigs()
        got += (len(contigs), po.contig_digest(contigs))
        want = (gold["n_nodes"], gold["num_edges"], gold["graph_digest"], gold["n_contigs"], gold["contig_digest"])
        if got != want:
            bad.append((key, got, want))
    assert not bad, bad[:5]


def test_fuzz_counts_vs_oracle():
    for i in range(0, 40):
        recipe, k, F = recipes.fuzz_recipe(i)
        reads = reads_for(recipe)
        cls = _classes()["PairedDeBruijnGraph" if recipe["paired"] else "DeBruijnGraph"]
        want = (po.count_paired if recipe["paired"] else po.count_unpaired)(k, reads)
        got = dict(cls._count_kmers(k, reads).items())
        assert got == dict(want), i


@pytest.mark.parametrize("name", ["c2-s-aureus", "c3-ecoli-standin"])
def test_full_size_configs(name):
    """BASELINE configs C2 and C3 at full size against the reference's own digests."""
    check_against_gold(name, check_counts=False)


def test_full_size_count_properties():
    """Size-independent invariants on C2: occurrences conserved, table deterministic."""
    gold = GOLDEN["cases"]["c2-s-aureus"]
    reads = reads_for(gold["recipe"])
    cls = _classes()["DeBruijnGraph"]
    a = cls._count_kmers(gold["k"], reads)
    distinct, solid, occ, top = a.summary(gold["F"])
    assert occ == len(reads) * (100 - gold["k"] + 2)
    b = cls._count_kmers(gold["k"], reads)
    assert b.summary(gold["F"]) == (distinct, solid, occ, top)
    ka, ca = a.export(gold["F"])
    kb, cb = b.export(gold["F"])
    oa, ob = np.argsort(ka[:, 0], kind="stable"), np.argsort(kb[:, 0], kind="stable")
    assert np.array_equal(ka[oa], kb[ob]) and np.array_equal(ca[oa], cb[ob])


# ------------------------------------------------------------------------------ edge cases
def test_empty_and_short_inputs():
    dg = _classes()
    for reads in ([], [""], ["AC"], ["ACG", "", "A"]):
        g = dg["DeBruijnGraph"](reads, k=5, hamming_dist=0)
        assert g.num_edges == 0 and g.enumerate_contigs() == [] and len(g.nodes) == 0
    g = dg["PairedDeBruijnGraph"]([], k=5, hamming_dist=0)
    assert g.enumerate_contigs() == [] and len(g.nodes) == 0
    g = dg["DeBruijnGraph"](["ACGTA"], k=6, hamming_dist=0)      # exactly one window, no edge
    assert dict(g._count_kmers(6, ["ACGTA"]).items()) == {"ACGTA": 1}
    assert g.num_edges == 0


def test_error_behaviour():
    dg = _classes()
    with pytest.raises(ValueError, match="Allowed error must be less than the kmer length."):
        dg["DeBruijnGraph"](["ACGT"], k=2)
    with pytest.raises(ValueError):
        dg["DeBruijnGraph"](["ACĀT" * 4], k=3, hamming_dist=0)
    with pytest.raises(ValueError):
        dg["PairedDeBruijnGraph"]([("ACGTACGT", "ACG")], k=4, hamming_dist=0)
    with pytest.raises(ValueError):
        dg["DeBruijnGraph"](["ACGT" * 40], k=66, hamming_dist=0)   # 130 key bits
    from countminsketch import CountMinSketch
    with pytest.raises(AssertionError):
        CountMinSketch(20)


def test_arbitrary_alphabet_against_oracle():
    text = "It was many and many a year ago, In a kingdom by the sea, That a maiden there lived whom you may know"
    reads = [text[i:i + 14] for i in range(0, len(text) - 14)] * 3
    for k in (4, 7, 12):
        _, _, want = po.assemble(reads, k, 1, False)
        g = _classes()["DeBruijnGraph"](reads, k=k, hamming_dist=1)
        assert digest_of(g) == want.digest()
        assert g.enumerate_contigs() == po.contigs(want)


def test_against_c_oracle_random_dna():
    rng = np.random.default_rng(5)
    genome = "".join("ACGT"[c] for c in rng.integers(0, 4, 3000))
    for paired in (False, True):
        reads = []
        for _ in range(4000):
            s = int(rng.integers(0, 3000))
            r = (genome * 2)[s:s + 60]
            if rng.random() < 0.5:
                p = int(rng.integers(0, 60))
                r = r[:p] + "ACGT"[int(rng.integers(0, 4))] + r[p + 1:]
            if paired:
                s2 = (s + 80 + int(rng.integers(-2, 3))) % 3000
                reads.append((r, (genome * 2)[s2:s2 + 60]))
            else:
                reads.append(r)
        for k in (12, 21, 33, 40):
            want = co.assemble(reads, k, 2, paired)
            cls = _classes()["PairedDeBruijnGraph" if paired else "DeBruijnGraph"]
            g = cls(reads, k=k, hamming_dist=2)
            assert (g._csr.n_nodes, g.num_edges) == (want.n_nodes, want.num_edges), (paired, k)
            assert digest_of(g) == want.digest(), (paired, k)
            assert g.enumerate_contigs() == want.contigs(), (paired, k)
            want.close()


@pytest.mark.parametrize("name", ["nd-unpaired", "nd-unpaired-s1", "nd-unpaired-k41", "two-circles", "toy-unpaired"])
def test_two_phase_build_matches_golden(name, monkeypatch):
    """Prefix + Bloom-filtered tail (the route C4-sized inputs take) gives the same graph."""
    import ga_device as gd
    monkeypatch.setattr(gd, "TWO_PHASE_MIN_TABLE_BYTES", 0)
    monkeypatch.setattr(gd, "TWO_PHASE_MIN_READS", 1)
    monkeypatch.setattr(gd, "TWO_PHASE_MIN_STEP", 1)
    check_against_gold(name, check_counts=False)


BUCKET_SHAPES = {
    "one-bucket": dict(SUPERKMER_MIN_OCC=0),
    "many-buckets": dict(SUPERKMER_MIN_OCC=0, SUPERKMER_TARGET=48),
    "two-levels": dict(SUPERKMER_MIN_OCC=0, SUPERKMER_TARGET=2),
    "spill": dict(SUPERKMER_MIN_OCC=0, SUPERKMER_TARGET=4096, SUPERKMER_TABLE_SLOTS=256, SUPERKMER_MAX_SOLID=16),
}


@pytest.mark.parametrize("shape", sorted(BUCKET_SHAPES))
@pytest.mark.parametrize("name", ["nd-unpaired", "nd-unpaired-s1", "nd-unpaired-k32", "two-circles",
                                  "homopoly-unpaired"])
def test_bucketed_build_matches_golden(name, shape, monkeypatch):
    """The super-k-mer bucket route (what C2/C4-sized unpaired DNA inputs take) gives the same graph,
    whatever the bucket geometry, including buckets that spill out of shared memory."""
    import ga_device as gd
    for key, value in BUCKET_SHAPES[shape].items():
        monkeypatch.setattr(gd, key, value)
    calls = []
    real = gd.superkmer_stamps
    monkeypatch.setattr(gd, "superkmer_stamps", lambda *a, **kw: calls.append(1) or real(*a, **kw))
    check_against_gold(name, check_counts=False)
    assert calls, "the bucketed path did not run"


@pytest.mark.parametrize("shape", ["one-bucket", "many-buckets", "spill"])
def test_bucketed_build_fuzz_and_ragged(shape, monkeypatch):
    import ga_device as gd
    for key, value in BUCKET_SHAPES[shape].items():
        monkeypatch.setattr(gd, key, value)
    test_fuzz_golden()
    test_against_c_oracle_random_dna()
    rng = np.random.default_rng(17)
    genome = "".join("ACGT"[c] for c in rng.integers(0, 4, 5000))
    reads = []
    for _ in range(3000):                         # ragged lengths: shorter than a window up to 300 symbols
        s, n = int(rng.integers(0, 5000)), int(rng.integers(0, 301))
        reads.append((genome * 2)[s:s + n])
    for k in (4, 17, 31, 32):
        want = co.assemble(reads, k, 2, False)
        g = _classes()["DeBruijnGraph"](reads, k=k, hamming_dist=2)
        assert digest_of(g) == want.digest(), k
        assert g.enumerate_contigs() == want.contigs(), k
        want.close()


@pytest.mark.parametrize("name", ["nd-paired", "nd-paired-jitter2", "toy-paired", "kat-f1", "homopoly-A-paired"])
def test_bucketed_counting_for_pairs_matches_golden(name, monkeypatch):
    """Read pairs take the buckets for counting (the solid set) and the paired build afterwards."""
    import ga_device as gd
    monkeypatch.setattr(gd, "SUPERKMER_MIN_OCC", 0)
    monkeypatch.setattr(gd, "SUPERKMER_MIN_OCC_PAIRS", 0)
    monkeypatch.setattr(gd, "SUPERKMER_TARGET", 3000)
    calls = []
    real = gd.superkmer_solid
    monkeypatch.setattr(gd, "superkmer_solid", lambda *a, **kw: calls.append(1) or real(*a, **kw))
    check_against_gold(name, check_counts=False)
    gold = GOLDEN["cases"][name]
    dna = all(set(a) | set(b) <= set("ACGT") for a, b in reads_for(gold["recipe"]))
    assert bool(calls) == (dna and gold["k"] <= 32)


def test_bucketed_dense_form_matches_golden(monkeypatch):
    """The dense level-2 form (what the multi-GPU exchange uses) on one GPU."""
    import ga_device as gd
    monkeypatch.setattr(gd, "SUPERKMER_MIN_OCC", 0)
    monkeypatch.setattr(gd, "SUPERKMER_INDEX_FORM", False)
    monkeypatch.setattr(gd, "SUPERKMER_TARGET", 2000)
    for name in ("nd-unpaired", "nd-unpaired-k32", "two-circles"):
        check_against_gold(name, check_counts=False)


def test_bucket_segments_match_whole(monkeypatch):
    """Buckets gathered from several sources (one segment per source rank after the multi-GPU exchange):
    two read shards scattered separately and presented as two segments give the same solid set and the
    same candidate stamps as one pass over all reads."""
    import torch
    import ga_device as gd
    gold = GOLDEN["cases"]["nd-unpaired"]
    reads = reads_for(gold["recipe"])
    k, F = gold["k"], gold["F"]
    monkeypatch.setattr(gd, "SUPERKMER_TARGET", 3000)
    whole = gd.DeviceReads(reads, False)
    n_occ = whole.windows_total(k)
    l1_bits, l2_bits = gd.sk_geometry(n_occ)
    n_buckets = 1 << (l1_bits + l2_bits)
    cut = len(reads) // 3
    parts = []
    for lo, hi in ((0, cut), (cut, len(reads))):
        shard = gd.DeviceReads(reads[lo:hi], False, first_read=lo, estride=whole.estride)
        bases, meta, offsets, hist, total = gd.sk_scatter_local(shard, k, l1_bits, l2_bits, dense=True)
        parts.append((bases[:2 * total].clone(), meta[:total].clone(), offsets.clone(), hist.clone(), total))
    bases = torch.cat([p[0] for p in parts])
    meta = torch.cat([p[1] for p in parts])
    seg = torch.stack([parts[0][2], parts[1][2] + parts[0][4]]).contiguous()
    hist = (parts[0][3] + parts[1][3]).contiguous()
    keys2, n2, st2 = gd.sk_bucket_pass(bases, meta, seg, 2, hist, n_buckets, k, F, n_occ, whole.status)
    keys2, st2 = keys2[:n2, 0].clone(), st2[:4 * n2].clone().view(-1, 4)
    keys1, n1, st1 = gd.superkmer_stamps(whole, k, F)
    keys1, st1 = keys1[:n1, 0], st1[:4 * n1].view(-1, 4)
    assert n1 == n2 == gold["n_solid"]
    o1, o2 = torch.argsort(keys1), torch.argsort(keys2)
    assert torch.equal(keys1[o1], keys2[o2])
    assert torch.equal(st1[o1], st2[o2])


def _two_shard_exchange_case(monkeypatch, target=3000):
    """Two read shards cut separately with ONE geometry (what two ranks hold before the exchange), plus the whole."""
    import torch
    import ga_device as gd
    gold = GOLDEN["cases"]["nd-unpaired"]
    reads = reads_for(gold["recipe"])
    k, F = gold["k"], gold["F"]
    monkeypatch.setattr(gd, "SUPERKMER_TARGET", target)
    whole = gd.DeviceReads(reads, False)
    n_occ = whole.windows_total(k)
    l1_bits, l2_bits = gd.sk_geometry(n_occ)
    cut = len(reads) // 3
    shards = [gd.DeviceReads(reads[lo:hi], False, first_read=lo, estride=whole.estride)
              for lo, hi in ((0, cut), (cut, len(reads)))]
    cap1 = max(gd.sk_l1_capacity(sh, k, l1_bits) for sh in shards)
    keys1, n1, st1 = gd.superkmer_stamps(whole, k, F)
    o1 = torch.argsort(keys1[:n1, 0])
    want = (keys1[:n1, 0][o1].clone(), st1[:4 * n1].view(-1, 4)[o1].clone())
    assert n1 == gold["n_solid"]
    return gd, whole, shards, k, F, n_occ, l1_bits, l2_bits, cap1, want


def test_bucket_sources_form_matches_whole(monkeypatch):
    """The exchange fused into the count (ga_sk_count_build_from, multi-GPU "peer" route) on one GPU: two read shards
    stay in their own level-1 slots with their own index, the bucket kernel gathers every bucket from both, and the
    results are appended through a shared counter -- same solid set and candidate stamps as one pass over all reads.
    Also with the loads two spans ahead of the walk (GA_SK_DEEP)."""
    import torch
    import ga_native as gn
    gd, whole, shards, k, F, n_occ, l1_bits, l2_bits, cap1, want = _two_shard_exchange_case(monkeypatch)
    dev = whole.words.device
    n_buckets = 1 << (l1_bits + l2_bits)
    held = []
    for sh in shards:
        keep = {}

        def alloc(name, n, dtype, keep=keep):
            keep[name] = torch.empty(int(n), dtype=dtype, device=dev)
            return keep[name]
        rec, _, offsets, hist, total, index, got_cap = gd.sk_scatter_local(sh, k, l1_bits, l2_bits, dense=False, cap1=cap1,
                                                                       alloc=alloc)
        assert got_cap == cap1
        held.append((rec, index, offsets.clone(), hist.clone(), total))
    seg = torch.stack([h[2] for h in held]).contiguous()
    hist = (held[0][3] + held[1][3]).contiguous()
    for deep in ("", "1"):
        monkeypatch.setenv("GA_SK_DEEP", deep) if deep else monkeypatch.delenv("GA_SK_DEEP", raising=False)
        rows = want[0].shape[0] + 100
        shared = torch.zeros(8 + rows * 5, dtype=torch.int64, device=dev)       # counter | keys | 4 stamps per key
        src = gn.GaSkSources()
        for g, h in enumerate(held):
            src.records[g], src.index[g], src.l1_capacity[g] = h[0].data_ptr(), h[1].data_ptr(), cap1
        src.first_bucket, src.n_sources, src.solid_counter = 0, 2, shared.data_ptr()
        keys_out, stamps_out = shared[8:8 + rows], shared[8 + rows:]
        gd.sk_bucket_pass(None, None, seg, 2, hist, n_buckets, k, F, n_occ, whole.status, l2_bits=l2_bits, sources=src,
                          out=(keys_out, stamps_out, rows))
        n2 = int(shared[0].item())
        assert n2 == want[0].shape[0]
        o2 = torch.argsort(keys_out[:n2])
        assert torch.equal(keys_out[:n2][o2], want[0])
        assert torch.equal(stamps_out[:4 * n2].view(-1, 4)[o2], want[1])


@pytest.mark.parametrize("kernel", ["gather", "sorted"])
def test_push_kernels_deliver_every_record(kernel, monkeypatch):
    """ga_sk_push_records (index + gather) and ga_sk_push_sorted (level-2 split + send in one pass) with two
    "ranks" on one GPU: both shards push the records of both owners' bucket ranges into the owners' receive arrays
    (plain device buffers here, NVLink-mapped ones on a node), and each owner's bucket pass over its two segments
    gives the solid set and stamps of its bucket range."""
    import ctypes as C
    import torch
    import ga_native as gn
    gd, whole, shards, k, F, n_occ, l1_bits, l2_bits, cap1, want = _two_shard_exchange_case(monkeypatch)
    L = gn.lib()
    dev = whole.words.device
    world = 2
    n_buckets = 1 << (l1_bits + l2_bits)
    bounds = [g * n_buckets // world for g in range(world + 1)]
    local = []
    for sh in shards:
        if kernel == "gather":
            rec, _, offsets, hist, total, index, _ = gd.sk_scatter_local(sh, k, l1_bits, l2_bits, dense=False, cap1=cap1,
                                                                     alloc=lambda name, n, dt: torch.empty(int(n), dtype=dt, device=dev))
            local.append((rec, index, offsets, hist, None, None))
        else:
            rec, _, offsets, hist, total, _, _, cursors1, cursors2 = gd.sk_scatter_local(
                sh, k, l1_bits, l2_bits, dense=False, level2=False, cap1=cap1,
                alloc=lambda name, n, dt: torch.empty(int(n), dtype=dt, device=dev))
            local.append((rec, None, offsets, hist, cursors1, cursors2))
    cuts = [[int(x) for x in l[2][torch.tensor(bounds, device=dev)].tolist()] for l in local]
    matrix = [[cuts[s][g + 1] - cuts[s][g] for g in range(world)] for s in range(world)]
    import ga_multi
    recv = [(torch.zeros(2 * sum(matrix[s][g] for s in range(world)) + 2, dtype=torch.int64, device=dev),
             torch.zeros(sum(matrix[s][g] for s in range(world)) + 1, dtype=torch.int64, device=dev)) for g in range(world)]
    plans = [ga_multi.push_plan(matrix, s) for s in range(world)]
    for s, (rec, index, offsets, hist, cursors1, cursors2) in enumerate(local):
        dst_start = plans[s][0]
        cut = (C.c_uint64 * (world + 1))(*cuts[s])
        dst_b = (C.c_void_p * world)(*[recv[g][0].data_ptr() + 16 * dst_start[g] for g in range(world)])
        dst_m = (C.c_void_p * world)(*[recv[g][1].data_ptr() + 8 * dst_start[g] for g in range(world)])
        if kernel == "gather":
            gn.check(L.ga_sk_push_records(gn.ptr(rec), cap1, gn.ptr(index), gn.ptr(offsets), l1_bits, l2_bits, world, cut,
                                          dst_b, dst_m, gd._stream()))
        else:
            gn.check(L.ga_sk_push_sorted(gn.ptr(rec), cap1, gn.ptr(cursors1), l1_bits, l2_bits, gn.ptr(cursors2), world,
                                         cut, dst_b, dst_m, gd._stream()))
    got_keys, got_stamps = [], []
    for g in range(world):
        mine = bounds[g + 1] - bounds[g]
        per_source = torch.stack([l[3][bounds[g]:bounds[g + 1]] for l in local])
        seg = torch.zeros((world, mine + 1), dtype=torch.int64, device=dev)
        seg[:, 1:] = torch.cumsum(per_source >> 32, dim=1)
        seg += torch.tensor(plans[g][1], dtype=torch.int64, device=dev).view(world, 1)
        summed = per_source.sum(dim=0).contiguous()
        keys, n, stamps = gd.sk_bucket_pass(recv[g][0], recv[g][1], seg.contiguous(), world, summed, mine, k, F,
                                            max(int((summed & 0xFFFFFFFF).sum().item()), 1), whole.status)
        got_keys.append(keys[:n, 0].clone())
        got_stamps.append(stamps[:4 * n].view(-1, 4).clone())
    keys, stamps = torch.cat(got_keys), torch.cat(got_stamps)
    o = torch.argsort(keys)
    assert torch.equal(keys[o], want[0])
    assert torch.equal(stamps[o], want[1])


@pytest.mark.parametrize("k", [27, 31, 32])
def test_scatter_kernels_cut_identical_records(k, monkeypatch):
    """The lane-per-read scatter kernel (both CTA shapes) and the warp-per-read one cut the same reads into the
    same records in the same buckets (ga_sk_scatter_reads, GA_SK_SCATTER picks the kernel): uniform reads and
    ragged ones -- shorter than a window, exactly 32 / 33 / 64 / 65 windows, several chunks."""
    import random
    import torch
    import ga_device as gd
    monkeypatch.setattr(gd, "SUPERKMER_TARGET", 300)
    rng = random.Random(k)
    w = k - 1
    genome = "".join(rng.choice("ACGT") for _ in range(30000))
    lengths = [w - 1, w, w + 1, w + 30, w + 31, w + 32, w + 63, w + 64, w + 95, w + 96, 5, 1, 400, 1000]
    ragged = []
    for i in range(4000):
        n = lengths[i % len(lengths)] if i % 3 else rng.randrange(1, 330)
        at = rng.randrange(len(genome) - n)
        ragged.append(genome[at:at + n] if i % 50 else "A" * n)          # some reads of one repeated symbol
    uniform = [genome[at:at + 150] for at in (rng.randrange(len(genome) - 150) for _ in range(5000))]
    for reads in (uniform, ragged):
        dr = gd.DeviceReads(reads, False)
        l1_bits, l2_bits = gd.sk_geometry(dr.windows_total(k))
        assert l1_bits > 0 and l2_bits > 0
        got = {}
        for variant in ("warp", "lane", "lane128"):
            monkeypatch.setenv("GA_SK_SCATTER", variant)
            bases, meta, offsets, hist, total = gd.sk_scatter_local(dr, k, l1_bits, l2_bits, dense=True)
            meta = meta[:total].clone()
            bases = bases[:2 * total].clone().view(-1, 2)
            counts = offsets[1:] - offsets[:-1]
            bucket = torch.repeat_interleave(torch.arange(counts.numel(), device=meta.device), counts)
            order = torch.argsort(meta)
            got[variant] = (meta[order], bases[order], bucket[order], hist.clone(), total)
        assert got["warp"][4] > len(reads) // 2
        for variant in ("lane", "lane128"):
            assert got[variant][4] == got["warp"][4], variant
            for a, b in zip(got[variant][:4], got["warp"][:4]):
                assert torch.equal(a, b), variant


def test_host_buffer_entries_match_device_path(monkeypatch):
    """host_step (ASCII in pinned memory, streamed in by chunks) and host_step_packed (2-bit words in
    pinned memory) give the CSR of the device-resident path, chunk boundaries included."""
    import torch
    import ga_native as gn
    import ga_device as gd
    monkeypatch.setattr(gd, "SUPERKMER_MIN_OCC", 0)
    monkeypatch.setattr(gd, "_STREAM_CHUNK_BYTES", 1 << 16)          # many chunks
    L = gn.lib()
    dev = torch.device("cuda", 0)
    G, n, rl, k, F = 40000, 9000, 100, 31, 2
    stride = (rl + 31) // 32
    genome = torch.empty(G, dtype=torch.uint8, device=dev)
    gn.check(L.ga_gen_genome(gn.ptr(genome), G, 3, None))
    words = torch.empty(n * stride, dtype=torch.int64, device=dev)
    gn.check(L.ga_gen_reads(gn.ptr(genome), G, 0, n, rl, 3, 100, gn.ptr(words), stride, 0, 0, None))
    reads = gd.DeviceReads.from_packed(words, n, rl, False, estride=rl)
    want = gd.build_graph(gd.KmerCounts(k, reads), reads, F, to_host=True)
    ascii_dev = torch.empty(n * rl, dtype=torch.uint8, device=dev)
    gn.check(L.ga_unpack_reads(gn.ptr(words), n, rl, stride, 2, gn.ptr(reads.alphabet.inv_dev), gn.ptr(ascii_dev), None))
    pinned = torch.empty(n * rl, dtype=torch.uint8, pin_memory=True)
    pinned.copy_(ascii_dev)
    packed = torch.empty(n * stride, dtype=torch.int64, pin_memory=True)
    packed.copy_(words)
    torch.cuda.synchronize()
    for got in (gd.host_step(pinned, n, rl, False, k, F), gd.host_step_packed(packed, n, rl, k, F)):
        assert got.n_nodes == want.n_nodes > 0 and got.n_edges == want.n_edges
        for field in ("rowptr", "col", "indeg", "branching", "last_char", "keys_a"):
            assert np.array_equal(getattr(got, field), getattr(want, field)), field


def test_deterministic_across_runs():
    gold = GOLDEN["cases"]["nd-paired-jitter2"]
    reads = reads_for(gold["recipe"])
    cls = _classes()["PairedDeBruijnGraph"]
    a = cls(reads, k=gold["k"], hamming_dist=gold["F"])._csr
    b = cls(reads, k=gold["k"], hamming_dist=gold["F"])._csr
    for field in ("rowptr", "col", "indeg", "branching", "last_char", "keys_a", "keys_b"):
        assert np.array_equal(getattr(a, field), getattr(b, field)), field


# ------------------------------------------------------------------------------ API surface
def test_counts_mapping_surface():
    reads = ["ACGTACGTAC", "CGTACGTTTT"]
    counts = _classes()["DeBruijnGraph"]._count_kmers(5, reads)
    want = po.count_unpaired(5, reads)
    assert len(counts) == len(want) and bool(counts)
    assert counts["ACGT"] == want["ACGT"] and counts["GGGG"] == 0 and counts["AC"] == 0
    assert "ACGT" in counts and "NNNN" not in counts
    assert dict(counts.items()) == dict(want)


def test_sketch_scalar_interface():
    from countminsketch import CountMinSketch
    sk = CountMinSketch(3)
    ref = po.Sketch(3)
    rng = np.random.default_rng(3)
    words = ["".join("ACGT"[c] for c in rng.integers(0, 4, int(rng.integers(1, 40)))) for _ in range(300)]
    for i, word in enumerate(words):
        sk.update(word, i % 7 + 1)
        ref.update(word, i % 7 + 1)
    assert [sk.estimate(wd) for wd in words[:50]] == [ref.estimate(wd) for wd in words[:50]]
    assert sk["ACGT"] == ref["ACGT"]
    assert [row.tobytes() for row in sk.hash_values] == [row.tobytes() for row in ref.rows]
    assert CountMinSketch._hash("ACGT") == po.murmur3_32("ACGT")
    assert sys.getsizeof(CountMinSketch(10)) == GOLDEN["sketch_sizeof_10"]
    sk.update("ACGT", 65535)
    with pytest.raises(OverflowError):
        sk.update("ACGT", 1)
        sk.estimate("ACGT")


def test_nodes_materialise_like_reference():
    gold = GOLDEN["cases"]["toy-paired"]
    reads = reads_for(gold["recipe"])
    g = _classes()["PairedDeBruijnGraph"](reads, k=gold["k"], hamming_dist=gold["F"])
    _, _, want = po.assemble(reads, gold["k"], gold["F"], True)
    flat = [((a, b), node) for a, inner in g.nodes.items() for b, node in inner.items()]
    assert [key for key, _ in flat] == want.keys
    assert [list(node.edges) for _, node in flat] == [[want.keys[j] for j in row] for row in want.succ]
    assert [node.indegree for _, node in flat] == want.indeg
    # traversal over the materialised objects gives the same contigs as the CSR walk
    assert g.enumerate_contigs() == po.contigs(want)


def test_debug_mixin_composes(capsys):
    import debug_graph
    gold = GOLDEN["cases"]["kat-f1"]
    reads = reads_for(gold["recipe"])
    g = debug_graph.DebugCMSPairedDeBruijnGraph(print_runtime=True, print_syssizeof=True, start_time=0.0,
                                                reads=reads, k=5, hamming_dist=1, paired_error=None)
    assert g.enumerate_contigs() == gold["contigs"]
    out = capsys.readouterr().out
    for needle in ("STARTING TO COUNT KMERS", "FINISHED COUNTING KMERS", "SIZE OF COUNTS CONTAINER",
                   "STARTING TO MAKE COUNTMIN SKETCH", "SIZE OF COUNTMIN SKETCH: 199,998,988",
                   "STARTING TO BUILD GRAPH", "SIZE OF ALL NODES", "STARTING TO ENUMERATE CONTIGS"):
        assert needle in out, needle


def _run_cli(args, stdin_text):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "genome-assembler_b200"))
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "genome-assembler_b200", "assemble.py")] + args,
                          input=stdin_text, capture_output=True, text=True, env=env, timeout=600)
    assert proc.returncode == 0, proc.stderr
    return [ln for ln in proc.stdout.split("\n") if not ln.startswith(">Time")]


def test_cli_readme_known_answer():
    text = ("5\nnnabe_Lee;|_Lee;By_th|5\nBy_the_nam|e_name_of_|5\nabe_Lee;By|ee;By_the_|5\n"
            "e_of_Annab|Annabe_Lee|5\ne_name_of_|e_of_Annab|5\n")
    want = [">Number of contigs:  2", ">CONTIG1", "he_", ">CONTIG2", "me_of_Annabe_Lee;By", ""]
    assert _run_cli(["--kmer_length", "5", "--filter", "1", "--stdout"], text) == want
    assert _run_cli(["-k", "5", "-f", "1", "-s", "-c", "--paired"], text) == want
    one = _run_cli(["--kmer_length", "5", "--filter", "0", "--stdout"], text)
    assert one[2] == "he_name_of_Annabe_Lee;By"


def test_synthetic_generator_matches_numpy_model():
    import ctypes as C
    import torch
    import ga_native as gn
    from oracle import readgen
    L = gn.lib()
    dev = torch.device("cuda", 0)
    G, n, rl, seed = 5003, 700, 150, 11
    genome = torch.empty(G, dtype=torch.uint8, device=dev)
    gn.check(L.ga_gen_genome(gn.ptr(genome), G, seed, None))
    want_g = readgen.splitmix_genome_codes(G, seed)
    assert np.array_equal(genome.cpu().numpy(), want_g)
    stride = (rl + 31) // 32
    words = torch.empty(n * stride, dtype=torch.int64, device=dev)
    for paired, dist in ((0, 0), (1, 125)):
        gn.check(L.ga_gen_reads(gn.ptr(genome), G, 40, n, rl, seed, 100, gn.ptr(words), stride, paired, dist,
                                None))
        torch.cuda.synchronize()
        w = words.cpu().numpy().view(np.uint64).reshape(n, stride)
        codes = np.stack([(w[:, i // 32] >> np.uint64(2 * (i % 32))) & np.uint64(3) for i in range(rl)], axis=1)
        want = readgen.splitmix_reads_codes(want_g, rl, 40, n, seed, 100, bool(paired), dist)
        assert np.array_equal(codes.astype(np.uint8), want)
