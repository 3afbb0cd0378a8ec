"""The size-independent graph properties of tests/graph_properties.py hold on the C oracle's graph (so they are
properties of the reference's construction, not of our kernels) and notice a graph that breaks them.  The GPU
suite applies the same checker to the full-size C4 graph (tests/test_gpu_c4_parity.py)."""
import numpy as np
import pytest
import torch

from graph_properties import check_unpaired_dna_graph, genome_window_keys
from oracle import c_oracle as co
from oracle import readgen


def _oracle_graph(G, n, rl, seed, k, F):
    genome = readgen.splitmix_genome_codes(G, seed)
    codes = readgen.splitmix_reads_codes(genome, rl, 0, n, seed, 100, False, 0)
    res = co.assemble_codes(codes, k, F)
    rowptr, col, indeg, br, last = res.csr()
    a, _ = res.node_keys()
    lut = np.full(256, 255, dtype=np.uint8)
    lut[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)
    sym = lut[a].astype(np.int64)                       # (n, w) codes, first symbol first
    keys = np.zeros(len(sym), dtype=np.int64)
    for j in range(k - 1):
        keys = (keys << 2) | sym[:, j]
    out = dict(codes=torch.from_numpy(genome), k=k, rowptr=torch.from_numpy(rowptr), col=torch.from_numpy(col),
               indeg=torch.from_numpy(indeg), branching=torch.from_numpy(br), last_sym=torch.from_numpy(lut[last]),
               keys=torch.from_numpy(keys), n_nodes=res.n_nodes, n_edges=res.n_csr_edges)
    res.close()
    return out


@pytest.fixture(scope="module")
def sample():
    return _oracle_graph(20000, 40000, 150, 4, 31, 3)      # C4's shape at 300x coverage


def test_window_keys():
    codes = torch.tensor([0, 1, 2, 3, 3, 0], dtype=torch.uint8)
    assert genome_window_keys(codes, 3).tolist() == [0b000110, 0b011011, 0b101111, 0b111100, 0b110000, 0b000001]


def test_properties_hold_on_the_oracle_graph(sample):
    report = check_unpaired_dna_graph(**sample)
    assert report["genome_nodes"] == 20000 and report["nodes"] >= 20000
    other = _oracle_graph(5000, 10000, 150, 9, 32, 3)      # 62-bit keys
    assert check_unpaired_dna_graph(**other)["genome_nodes"] == 5000


@pytest.mark.parametrize("what", ["edge", "indeg", "branching", "node", "duplicate"])
def test_a_broken_graph_is_noticed(sample, what):
    bad = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in sample.items()}
    n = bad["n_nodes"]
    if what == "edge":
        bad["col"][5] = (bad["col"][5] + 7) % n
    elif what == "indeg":
        bad["indeg"][n // 2] += 1
    elif what == "branching":
        bad["branching"][n // 3] ^= 1
    elif what == "node":                                  # a genome window dropped from the node list: its key changes
        bad["keys"][n // 4] ^= 1 << 20
    else:
        bad["keys"][7] = bad["keys"][8]
    with pytest.raises(AssertionError):
        check_unpaired_dna_graph(**bad)
