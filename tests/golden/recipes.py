"""Input recipes shared by make_golden.py (runs the reference) and the tests (run the
oracle / the CUDA path).  Imports nothing from the reference or from the product."""
from __future__ import annotations

import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
TOY = "It_was_many_and_many_a_year_ago_"


def unpack_2bit(path: str) -> str:
    import numpy as np
    raw = open(path, "rb").read()
    n = int.from_bytes(raw[:8], "little")
    b = np.frombuffer(raw, dtype=np.uint8, offset=8)
    codes = np.stack([(b >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)[:n]
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes().decode("ascii")


_GENOMES = {}


def genome_text(name: str) -> str:
    """Genome by recipe name, from the committed fixtures (never from /root/reference)."""
    if name not in _GENOMES:
        if name == "toy":
            _GENOMES[name] = TOY
        elif name in ("n_delto", "s_aureus"):
            _GENOMES[name] = unpack_2bit(os.path.join(HERE, name + ".2bit"))
        elif name == "ecoli_standin":
            _GENOMES[name] = "".join(random.Random(20261018).choices("ACGT", k=4641652))
        else:
            raise KeyError(name)
    return _GENOMES[name]


def fuzz_recipe(i):
    """Small random cases: alphabets of 2-5 symbols force repeats, branching, fuzzy
    groups, homopolymers and perfect cycles."""
    rng = random.Random(7000 + i)
    sigma = rng.choice([2, 3, 4, 4, 5])
    alphabet = "ACGT_"[:sigma] if sigma <= 4 else "ACGTN"
    glen = rng.randint(12, 90)
    genome = "".join(rng.choice(alphabet) for _ in range(glen))
    paired = bool(i & 1)
    L = rng.randint(6, 14)
    N = rng.randint(10, 120)
    k = rng.randint(3, min(L, 9))
    F = rng.choice([0, 0, 1, 2, 3])
    reads = []
    for _ in range(N):
        s = rng.randrange(glen)
        circ = (genome * 3)
        r1 = circ[s:s + L]
        if rng.random() < 0.3:
            p = rng.randrange(L)
            r1 = r1[:p] + rng.choice(alphabet) + r1[p + 1:]
        if paired:
            s2 = (s + rng.randint(3, 8) + rng.randint(-2, 2)) % glen
            r2 = circ[s2:s2 + L]
            reads.append([r1, r2])
        else:
            if rng.random() < 0.1:                     # ragged / short reads
                r1 = r1[:rng.randint(0, L)]
            reads.append(r1)
    return {"kind": "explicit", "paired": paired, "reads": reads}, k, F


