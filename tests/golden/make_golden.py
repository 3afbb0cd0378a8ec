#!/usr/bin/env python3
"""Generate tests/golden/golden.json by running the UNMODIFIED upstream reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # small + medium cases (~3 min)
    python tests/golden/make_golden.py --big      # adds C2 (S. aureus) and C3 (E. coli-sized)

It imports the reference's own modules, feeds them deterministic inputs (its generator
with ``seed`` patched, SURVEY App. B.3, or explicit read lists) and records digests of
everything the hot path produces: the exact count table, sketch cells, node order,
edge order, in-degrees, ``was_branching`` and the contig list.  Also writes the 2-bit
packed genome fixtures the GPU-side tests regenerate their reads from.  Nothing from
the reference's *source* is copied; only outputs are stored.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.setrecursionlimit(10000)

sys.path.insert(0, HERE)
from recipes import fuzz_recipe, TOY  # noqa: E402
import countminsketch as ref_cms          # noqa: E402
import debruijn_graph as ref_dbg          # noqa: E402
import generate_reads as ref_gen          # noqa: E402



def sha16(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()[:16]


def genome_text(name: str) -> str:
    if name == "toy":
        return TOY
    if name == "n_delto":
        return open(os.path.join(REF, "reference_genomes/n_deltocephalinicola.txt")).readline().strip()
    if name == "s_aureus":
        return open(os.path.join(REF, "reference_genomes/s_aureus_USA300_FPR3757.txt")).readline().strip()
    if name == "ecoli_standin":
        return "".join(random.Random(20261018).choices("ACGT", k=4641652))
    raise KeyError(name)


def pack_2bit(text: str) -> bytes:
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    out = bytearray((len(text) + 3) // 4)
    for i, ch in enumerate(text):
        out[i >> 2] |= code[ch] << (2 * (i & 3))
    return bytes(out)


def make_reads(recipe):
    if recipe["kind"] == "explicit":
        reads = recipe["reads"]
        return [tuple(r) for r in reads] if recipe["paired"] else list(reads)
    if recipe["kind"] == "splitmix":
        # the bench's input class (150-bp reads, per-base substitutions): made by the oracle-side numpy model of the
        # device generator -- the reference only ever sees the resulting strings
        for extra in (os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)):    # repo root (oracle/), tests/
            if extra not in sys.path:
                sys.path.append(extra)
        from helpers import splitmix_reads
        return splitmix_reads(recipe)
    genome = genome_text(recipe["genome"])
    ref_gen.seed = lambda *a: random.seed(recipe["seed"])
    lines = ref_gen.generate_reads(genome, read_len=recipe["L"], num_reads=recipe["N"],
                                   paired=recipe["paired"], d=recipe.get("d", 125),
                                   delta=recipe.get("delta", 0))
    if recipe["paired"]:
        return [tuple(line.split("|")[:2]) for line in lines]
    return lines


def graph_digest(graph, paired):
    h = hashlib.sha256()
    n = 0
    if paired:
        it = ((((a, b)), node) for a, inner in graph.nodes.items() for b, node in inner.items())
    else:
        it = iter(graph.nodes.items())
    for key, node in it:
        h.update(repr((key, list(node.edges), node.num_edges_in, node.was_branching)).encode())
        n += 1
    return n, h.hexdigest()[:16]


def run_case(name, recipe, k, F, cls_name, want_counts=True, keep_contigs=False):
    t0 = time.time()
    reads = make_reads(recipe)
    paired = recipe["paired"]
    cls = getattr(ref_dbg, cls_name)
    rec = {"recipe": recipe, "k": k, "F": F, "cls": cls_name,
           "reads_sha": sha16(repr(reads).encode())}
    if want_counts:
        counts = cls._count_kmers(k, reads)
        items = sorted(counts.items())
        rec["n_distinct"] = len(items)
        rec["n_occ"] = sum(c for _, c in items)
        rec["n_solid"] = sum(1 for _, c in items if c > F)
        rec["counts_sha"] = sha16("".join("%s:%d\n" % kc for kc in items).encode())
        if cls_name.startswith("CMS"):
            rows = 10 if cls_name == "CMSDeBruijnGraph" else 8
            sk = cls._make_sketch(counts)
            assert sk.num_rows == rows
            rec["sketch_rows"] = rows
            rec["sketch_row_sha"] = [sha16(row.tobytes()) for row in sk.hash_values]
            rec["sketch_nonzero"] = [sum(1 for v in row if v) for row in sk.hash_values]
            del sk
        del counts, items
    graph = cls(reads, k=k, hamming_dist=F)
    rec["num_edges"] = graph.num_edges
    rec["n_nodes"], rec["graph_digest"] = graph_digest(graph, paired)
    contigs = graph.enumerate_contigs()
    rec["n_contigs"] = len(contigs)
    rec["contig_digest"] = sha16("\n".join(contigs).encode())
    rec["edges_left"] = graph.num_edges
    if keep_contigs:
        rec["contigs"] = contigs
    rec["ref_seconds"] = round(time.time() - t0, 2)
    print("%-28s nodes=%d edges=%d contigs=%d  %.1fs" %
          (name, rec["n_nodes"], rec["num_edges"], rec["n_contigs"], rec["ref_seconds"]), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    path = os.path.join(HERE, "golden.json")
    gold = json.load(open(path)) if os.path.exists(path) else {}
    gold.setdefault("cases", {})
    gold.setdefault("fuzz", {})

    # --- hash known answers straight from the reference's _hash
    probes = ["", "A", "AC", "ACG", "ACGT", "ACGTACGTACGTACGTACGTACGTACG", "A" * 30, "It_wa",
              "TTAAAAAACATAAATTTAATTATATTTA", "\xff\xfe\xfd", "e_of_", "GATTACA" * 9]
    rng = random.Random(99)
    probes += ["".join(rng.choice("ACGT") for _ in range(rng.randint(1, 64))) for _ in range(40)]
    gold["murmur3"] = [[p, ref_cms.CountMinSketch._hash(p)] for p in probes]
    gold["overlap"] = []
    for _ in range(300):
        n = rng.randint(3, 9)
        a = "".join(rng.choice("AC") for _ in range(n))
        b = "".join(rng.choice("AC") for _ in range(n))
        gold["overlap"].append([a, b, ref_dbg.PairedDeBruijnGraph._find_longest_overlap_brute(a, b)])
    gold["break"] = {
        "unpaired": [["ACTGAC", 4, ref_dbg.DeBruijnGraph._break_read_into_k_minus_one_mers(4, "ACTGAC")],
                     ["ACT", 5, ref_dbg.DeBruijnGraph._break_read_into_k_minus_one_mers(5, "ACT")],
                     ["ACTG", 5, ref_dbg.DeBruijnGraph._break_read_into_k_minus_one_mers(5, "ACTG")]],
        "paired": [[["ACTGAC", "TCGATC"], 4,
                    [list(t) for t in ref_dbg.PairedDeBruijnGraph._break_read_into_k_minus_one_mers(4, ("ACTGAC", "TCGATC"))]]],
    }

    gold["prime_tables"] = {name: list(getattr(ref_cms.CountMinSketch, name))
                            for name in ("primes_6_10_7", "primes_1_10_7", "primes_5_10_6")}
    gold["sketch_sizeof_10"] = 199998988  # sys.getsizeof(CountMinSketch(10)), SURVEY App. B.1

    kat = [["nnabe_Lee;", "_Lee;By_th"], ["By_the_nam", "e_name_of_"], ["abe_Lee;By", "ee;By_the_"],
           ["e_of_Annab", "Annabe_Lee"], ["e_name_of_", "e_of_Annab"]]
    circles = ["ACGTTGCAAC" * 3] * 5 + ["GGATCCTAGG" * 3] * 5

    def refgen(genome, L, N, paired, seed, delta=0):
        return {"kind": "refgen", "genome": genome, "L": L, "N": N, "paired": paired,
                "seed": seed, "d": 125, "delta": delta}

    def explicit(reads, paired):
        return {"kind": "explicit", "paired": paired, "reads": reads}

    small = [
        ("kat-f1", explicit(kat, True), 5, 1, "PairedDeBruijnGraph", True),
        ("kat-f0", explicit(kat, True), 5, 0, "PairedDeBruijnGraph", True),
        ("kat-f1-cms", explicit(kat, True), 5, 1, "CMSPairedDeBruijnGraph", True),
        ("homopoly-A-paired", explicit([["AAAAAAAAAA", "CGTACCGTTA"]] * 5, True), 6, 3, "PairedDeBruijnGraph", True),
        ("homopoly-AC-paired", explicit([["A" * 10, "C" * 10]] * 5, True), 6, 3, "PairedDeBruijnGraph", True),
        ("homopoly-unpaired", explicit(["AAAAAAAAAA"] * 5, False), 4, 3, "DeBruijnGraph", True),
        ("two-circles", explicit(circles, False), 6, 3, "DeBruijnGraph", True),
        ("ragged", explicit(["ACGTACGTAC", "", "ACG", "ACGTA", "CGTACGTACGTT", "ACGTACGTAC", "ACGTACGTAC",
                             "ACGTACGTAC", "CGTACGTACGTT", "CGTACGTACGTT", "CGTACGTACGTT"], False), 5, 2,
         "DeBruijnGraph", True),
        ("toy-unpaired", refgen("toy", 8, 120, False, 1), 6, 3, "DeBruijnGraph", True),
        ("toy-paired", refgen("toy", 8, 60, True, 1), 6, 3, "PairedDeBruijnGraph", True),
        ("toy-unpaired-cms", refgen("toy", 8, 120, False, 1), 6, 3, "CMSDeBruijnGraph", True),
    ]
    medium = [
        ("nd-unpaired", refgen("n_delto", 100, 34000, False, 1234), 31, 3, "DeBruijnGraph"),
        ("nd-paired", refgen("n_delto", 100, 19000, True, 1234), 28, 3, "PairedDeBruijnGraph"),
        ("nd-paired-jitter2", refgen("n_delto", 100, 19000, True, 1234, 2), 28, 3, "PairedDeBruijnGraph"),
        ("nd-unpaired-s1", refgen("n_delto", 100, 34000, False, 1), 31, 3, "DeBruijnGraph"),
        ("nd-unpaired-s1-cms", refgen("n_delto", 100, 34000, False, 1), 31, 3, "CMSDeBruijnGraph"),
        ("nd-paired-cms8", refgen("n_delto", 100, 19000, True, 1234), 28, 3, "CMSPairedDeBruijnGraph"),
        ("nd-unpaired-k33", refgen("n_delto", 100, 20000, False, 5), 33, 2, "DeBruijnGraph"),
        ("nd-unpaired-k41", refgen("n_delto", 100, 20000, False, 5), 41, 2, "DeBruijnGraph"),
        ("nd-unpaired-k65", refgen("n_delto", 100, 20000, False, 5), 65, 1, "DeBruijnGraph"),
        ("nd-unpaired-k64", refgen("n_delto", 100, 20000, False, 5), 64, 1, "DeBruijnGraph"),
        ("nd-unpaired-k32", refgen("n_delto", 100, 20000, False, 5), 32, 2, "DeBruijnGraph"),
        ("nd-paired-k64", refgen("n_delto", 100, 12000, True, 6, 1), 64, 1, "PairedDeBruijnGraph"),
        ("nd-paired-k35", refgen("n_delto", 100, 12000, True, 6, 1), 35, 2, "PairedDeBruijnGraph"),
    ]
    def splitmix(G, N, L, paired, seed, dist=0):
        return {"kind": "splitmix", "G": G, "N": N, "L": L, "paired": paired, "seed": seed, "sub_per_10k": 100,
                "dist": dist}

    # BASELINE config C4's input class at 300x coverage (and C3 / C5's as bench.py generates them): reads only the
    # device generator and its numpy model can make (generate_reads.py:54 refuses L > 100)
    medium += [
        ("mix-c4-sample", splitmix(10000, 20000, 150, False, 4), 31, 3, "DeBruijnGraph"),
        ("mix-c4-sample-cms", splitmix(10000, 20000, 150, False, 4), 31, 3, "CMSDeBruijnGraph"),
        ("mix-c4-sample-k32", splitmix(8000, 12000, 150, False, 5), 32, 3, "DeBruijnGraph"),
        ("mix-c3-pairs", splitmix(40000, 12000, 100, True, 6, 125), 29, 3, "PairedDeBruijnGraph"),
        ("mix-c5-pairs-k41", splitmix(40000, 12000, 100, True, 6, 125), 41, 3, "PairedDeBruijnGraph"),
    ]
    big = [
        ("c2-s-aureus", refgen("s_aureus", 100, 861831, False, 1234), 31, 3, "DeBruijnGraph"),
        ("c3-ecoli-standin", refgen("ecoli_standin", 100, 700000, True, 1234), 29, 3, "PairedDeBruijnGraph"),
    ]

    def wanted(name):
        return not args.only or args.only in name

    for name, recipe, k, F, cls, keep in small:
        if wanted(name):
            gold["cases"][name] = run_case(name, recipe, k, F, cls, keep_contigs=keep)
    for name, recipe, k, F, cls in medium:
        if wanted(name):
            gold["cases"][name] = run_case(name, recipe, k, F, cls)
    if args.big:
        for name, recipe, k, F, cls in big:
            if wanted(name):
                gold["cases"][name] = run_case(name, recipe, k, F, cls, want_counts=False)

    if wanted("fuzz"):
        for i in range(240):
            recipe, k, F = fuzz_recipe(i)
            cls = "PairedDeBruijnGraph" if recipe["paired"] else "DeBruijnGraph"
            if i % 7 == 3:
                continue  # keep ids stable but skip some to bound the file size
            reads = make_reads(recipe)
            graph = getattr(ref_dbg, cls)(reads, k=k, hamming_dist=F)
            ne = graph.num_edges
            n, gd = graph_digest(graph, recipe["paired"])
            contigs = graph.enumerate_contigs()
            gold["fuzz"][str(i)] = {"k": k, "F": F, "paired": recipe["paired"], "n_nodes": n,
                                    "num_edges": ne, "graph_digest": gd, "n_contigs": len(contigs),
                                    "contig_digest": sha16("\n".join(contigs).encode())}
        print("fuzz cases:", len(gold["fuzz"]))

    # --- packed genome fixtures (data, not source): 2 bits per base, A,C,G,T = 0..3
    for gname, fname in (("n_delto", "n_delto.2bit"), ("s_aureus", "s_aureus.2bit")):
        out = os.path.join(HERE, fname)
        if not os.path.exists(out):
            text = genome_text(gname)
            with open(out, "wb") as fh:
                fh.write(len(text).to_bytes(8, "little"))
                fh.write(pack_2bit(text))
            print("wrote", out, len(text))

    with open(path, "w") as fh:
        json.dump(gold, fh, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
