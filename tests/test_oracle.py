"""The Python oracle against everything the unmodified reference produced
(tests/golden/golden.json, written by tests/golden/make_golden.py)."""
import pytest

from helpers import GOLDEN, counts_sha, reads_for, sha16
from oracle import py_oracle as po
import recipes

SMALL = [n for n, c in GOLDEN["cases"].items() if c.get("ref_seconds", 0) < 3 and "cms" not in n]
MEDIUM = ["nd-unpaired", "nd-paired-jitter2", "nd-unpaired-k65",
          # the bench's input class (150-bp reads, 1 % per-base substitutions; pairs as bench.py makes them), digests
          # from the unmodified reference run on these very reads
          "mix-c4-sample", "mix-c3-pairs", "mix-c5-pairs-k41"]


def test_murmur_known_answers():
    for text, want in GOLDEN["murmur3"]:
        assert po.murmur3_32(text) == want, text


def test_overlap_rule():
    for a, b, want in GOLDEN["overlap"]:
        assert po.fuzzy_overlap(a, b) == want, (a, b)


def test_break_docstring_examples():
    for read, k, want in GOLDEN["break"]["unpaired"]:
        assert po.windows(k, read) == want
    for pair, k, want in GOLDEN["break"]["paired"]:
        assert [list(t) for t in po.paired_windows(k, tuple(pair))] == want


def check_case(name, with_counts=True):
    gold = GOLDEN["cases"][name]
    reads = reads_for(gold["recipe"])
    assert sha16(repr(reads).encode()) == gold["reads_sha"], "input generator drifted"
    paired = gold["recipe"]["paired"]
    rows = gold.get("sketch_rows", 0)
    tally, sk, graph = po.assemble(reads, gold["k"], gold["F"], paired, sketch_rows=rows)
    if with_counts and "counts_sha" in gold:
        assert len(tally) == gold["n_distinct"]
        assert sum(tally.values()) == gold["n_occ"]
        assert sum(1 for c in tally.values() if c > gold["F"]) == gold["n_solid"]
        assert counts_sha(tally.items()) == gold["counts_sha"]
    if rows:
        assert [sha16(r.tobytes()) for r in sk.rows] == gold["sketch_row_sha"]
    assert len(graph.keys) == gold["n_nodes"]
    assert graph.num_edges == gold["num_edges"]
    assert graph.digest() == gold["graph_digest"]
    lines = po.contigs(graph)
    assert len(lines) == gold["n_contigs"]
    assert po.contig_digest(lines) == gold["contig_digest"]
    if "contigs" in gold:
        assert lines == gold["contigs"]


@pytest.mark.parametrize("name", SMALL)
def test_small_cases(name):
    check_case(name)


@pytest.mark.parametrize("name", MEDIUM)
def test_medium_cases(name):
    check_case(name)


def test_readme_known_answer():
    assert GOLDEN["cases"]["kat-f1"]["contigs"] == ["he_", "me_of_Annabe_Lee;By"]
    assert GOLDEN["cases"]["kat-f0"]["contigs"] == ["he_name_of_Annabe_Lee;By"]


def test_sketch_small():
    check_case("kat-f1-cms")


def test_fuzz_cases():
    for key, gold in GOLDEN["fuzz"].items():
        recipe, k, F = recipes.fuzz_recipe(int(key))
        assert (k, F, recipe["paired"]) == (gold["k"], gold["F"], gold["paired"])
        reads = reads_for(recipe)
        _, _, graph = po.assemble(reads, k, F, recipe["paired"])
        assert (len(graph.keys), graph.num_edges, graph.digest()) == \
            (gold["n_nodes"], gold["num_edges"], gold["graph_digest"]), key
        lines = po.contigs(graph)
        assert (len(lines), po.contig_digest(lines)) == (gold["n_contigs"], gold["contig_digest"]), key


# ---------------------------------------------------------------- C oracle (oracle/c_oracle.c)
from oracle import c_oracle as co  # noqa: E402

C_CASES = [n for n in GOLDEN["cases"] if not n.startswith("c")]


def check_case_c(name):
    gold = GOLDEN["cases"][name]
    reads = reads_for(gold["recipe"])
    res = co.assemble(reads, gold["k"], gold["F"], gold["recipe"]["paired"],
                      sketch_rows=gold.get("sketch_rows", 0))
    if "counts_sha" in gold:
        assert res.n_distinct == gold["n_distinct"]
        assert counts_sha(res.counts_dict().items()) == gold["counts_sha"]
    for row in range(gold.get("sketch_rows", 0)):
        assert sha16(res.sketch_row(row).tobytes()) == gold["sketch_row_sha"][row]
    assert (res.n_nodes, res.num_edges) == (gold["n_nodes"], gold["num_edges"])
    assert res.digest() == gold["graph_digest"]
    lines = res.contigs()
    assert (len(lines), po.contig_digest(lines)) == (gold["n_contigs"], gold["contig_digest"])
    res.close()


@pytest.mark.parametrize("name", C_CASES)
def test_c_oracle_cases(name):
    check_case_c(name)


def test_c_oracle_murmur():
    for text, want in GOLDEN["murmur3"]:
        assert co.murmur3_32(text) == want


def test_c_oracle_fuzz():
    for key, gold in GOLDEN["fuzz"].items():
        recipe, k, F = recipes.fuzz_recipe(int(key))
        res = co.assemble(reads_for(recipe), k, F, recipe["paired"])
        assert (res.n_nodes, res.num_edges, res.digest()) == \
            (gold["n_nodes"], gold["num_edges"], gold["graph_digest"]), key
        lines = res.contigs()
        assert (len(lines), po.contig_digest(lines)) == (gold["n_contigs"], gold["contig_digest"]), key
        res.close()


def test_c_oracle_sketch_overflow():
    with pytest.raises(OverflowError):
        co.assemble(["A" * 70000], 4, 3, False, sketch_rows=2)
