"""The order-free (stamp based) restatement the kernels implement, against the goldens."""
import pytest

from helpers import GOLDEN, reads_for
from oracle import py_oracle as po
import orderfree_model as om
import recipes


def run(reads, k, F, paired):
    tally = (po.count_paired if paired else po.count_unpaired)(k, reads)
    solid = {x for x, c in tally.items() if c > F}
    return (om.build_paired if paired else om.build_unpaired)(solid, reads, k)


def test_fuzz():
    for key, gold in GOLDEN["fuzz"].items():
        recipe, k, F = recipes.fuzz_recipe(int(key))
        g = run(reads_for(recipe), k, F, recipe["paired"])
        assert (len(g.keys), g.num_edges, g.digest()) == \
            (gold["n_nodes"], gold["num_edges"], gold["graph_digest"]), key
        lines = po.contigs(g)
        assert (len(lines), po.contig_digest(lines)) == (gold["n_contigs"], gold["contig_digest"]), key


@pytest.mark.parametrize("name", ["kat-f1", "kat-f0", "homopoly-A-paired", "homopoly-AC-paired",
                                  "homopoly-unpaired", "two-circles", "ragged", "toy-paired",
                                  "toy-unpaired", "nd-paired-jitter2"])
def test_cases(name):
    gold = GOLDEN["cases"][name]
    g = run(reads_for(gold["recipe"]), gold["k"], gold["F"], gold["recipe"]["paired"])
    assert (len(g.keys), g.num_edges, g.digest()) == (gold["n_nodes"], gold["num_edges"], gold["graph_digest"])
    lines = po.contigs(g)
    assert po.contig_digest(lines) == gold["contig_digest"]
