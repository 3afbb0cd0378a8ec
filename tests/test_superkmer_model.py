"""The bucketed (super-k-mer) restatement of count + build, against the goldens and the
order-free model -- pins the algorithm of ga_superkmer.cu on the CPU."""
import random

import pytest

from helpers import GOLDEN, reads_for
from oracle import py_oracle as po
import orderfree_model as om
import recipes
import superkmer_model as sk


def _is_dna(reads):
    return all(set(r) <= set("ACGT") for r in reads)


def test_fuzz_unpaired_dna():
    seen = 0
    for key, gold in GOLDEN["fuzz"].items():
        recipe, k, F = recipes.fuzz_recipe(int(key))
        if recipe["paired"]:
            continue
        reads = reads_for(recipe)
        if not _is_dna(reads):
            continue
        for bits in (0, 3):
            g, _ = sk.build_unpaired(reads, k, F, bits)
            assert (len(g.keys), g.num_edges, g.digest()) == \
                (gold["n_nodes"], gold["num_edges"], gold["graph_digest"]), (key, bits)
            lines = po.contigs(g)
            assert (len(lines), po.contig_digest(lines)) == (gold["n_contigs"], gold["contig_digest"]), key
        seen += 1
    assert seen > 40


@pytest.mark.parametrize("name", ["homopoly-unpaired", "two-circles"])
def test_cases(name):
    gold = GOLDEN["cases"][name]
    reads = reads_for(gold["recipe"])
    assert _is_dna(reads)
    g, _ = sk.build_unpaired(reads, gold["k"], gold["F"], 2)
    assert (len(g.keys), g.num_edges, g.digest()) == (gold["n_nodes"], gold["num_edges"], gold["graph_digest"])
    assert po.contig_digest(po.contigs(g)) == gold["contig_digest"]


def test_random_genome_with_errors_matches_orderfree_model():
    rng = random.Random(99)
    genome = "".join(rng.choice("ACGT") for _ in range(1500))
    reads = []
    for _ in range(900):
        s = rng.randrange(len(genome))
        r = (genome * 2)[s:s + rng.choice([70, 70, 70, 45, 20])]
        if rng.random() < 0.6:
            p = rng.randrange(len(r))
            r = r[:p] + rng.choice("ACGT") + r[p + 1:]
        reads.append(r)
    for k, F in ((31, 3), (21, 2), (12, 1)):
        tally = po.count_unpaired(k, reads)
        solid = {x for x, c in tally.items() if c > F}
        want = om.build_unpaired(solid, reads, k)
        got, records = sk.build_unpaired(reads, k, F, 5)
        assert got.digest() == want.digest(), k
        assert po.contigs(got) == po.contigs(want)
        # every window travels exactly once
        assert sum(r[2] for r in records) == sum(tally.values())
        assert all(r[2] <= sk.MAX_RUN for r in records)
