"""Deterministic read generators used by the oracle, the tests and the bench.

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/py_oracle.py header).

* ``reference_style_reads`` restates the upstream generator (generate_reads.py:42-74)
  call-for-call on a seeded ``random.Random`` so that, for a given seed, it emits the
  very same lines as the upstream function with its ``seed()`` patched
  (SURVEY App. B.3).  Pinned by the ``reads_sha`` entries of tests/golden/golden.json.
* ``splitmix_*`` is our own counter-based generator for shapes the upstream one cannot
  make (L > 100, per-base substitutions, 10^8 reads); the CUDA generator in
  ``csrc/ga_readgen.cu`` implements the same arithmetic and is checked against this.
"""
from __future__ import annotations

import random

import numpy as np

_BASES = ("A", "T", "C", "G")          # generate_reads.py:44 (order matters)


def reference_style_reads(genome: str, read_len: int, num_reads: int, paired: bool,
                          d: int = 125, delta: int = 0, seed: int = 0):
    """Reads (str) or read pairs ((str, str)) drawn like generate_reads.py:42-74."""
    rng = random.Random(seed)
    size = len(genome)
    out = []

    def circular(start):
        end = start + read_len
        piece = genome[start:min(end, size)]
        if end > size:
            piece += genome[0:end % size]
        return piece

    for _ in range(num_reads):
        start = rng.randint(0, size)                       # inclusive upper end (:47)
        read = circular(start)
        if rng.randint(0, 100 // read_len - 1) == 0:       # (:54) raises for L > 100
            base = _BASES[rng.randint(0, 3)]               # right-hand side is drawn first
            where = rng.randint(0, len(read) - 1)
            read = read[:where] + base + read[where + 1:]
        if len(read) != read_len:
            raise AssertionError("short read")             # (:59)
        if not paired:
            out.append(read)
            continue
        mate_start = (start + d + rng.randint(-delta, delta)) % size
        mate = circular(mate_start)
        if len(mate) != read_len:
            raise AssertionError("short mate")
        out.append((read, mate))
    return out


def random_genome(size: int, seed: int) -> str:
    """Uniform ACGT stand-in genome: random.Random(seed).choices."""
    return "".join(random.Random(seed).choices("ACGT", k=size))


# ------------------------------------------------------------------ counter-based generator
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """SplitMix64 finaliser on uint64 numpy arrays (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def splitmix_genome_codes(size: int, seed: int) -> np.ndarray:
    """Genome as 2-bit codes (0..3 = A,C,G,T): code[i] = splitmix64(seed*2^40 + i) & 3."""
    idx = np.arange(size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return (splitmix64((np.uint64(seed) << np.uint64(40)) + idx) & np.uint64(3)).astype(np.uint8)


def splitmix_reads_codes(genome_codes: np.ndarray, read_len: int, first_read: int,
                         num_reads: int, seed: int, sub_per_10k: int = 100,
                         paired: bool = False, mate_distance: int = 0) -> np.ndarray:
    """``num_reads`` x ``read_len`` uint8 codes for reads ``first_read..``: circular start
    uniform in [0, G), every base independently replaced with probability
    ``sub_per_10k / 10000`` by a uniform draw from ACGT (may equal the original).

    start(r)   = splitmix64(seed*2^40 + 2^39 + r) % G      [paired: r>>1 is drawn, and mate
                 r&1 == 1 starts mate_distance later; rows are stored reads = mates]
    h(r, i)    = splitmix64((seed+1)*2^40 ^ (r*read_len + i))   [xor; r*L+i < 2^40]
    replace    iff h % 10000 < sub_per_10k, with code (h >> 32) & 3
    """
    size = np.uint64(len(genome_codes))
    r = np.arange(first_read, first_read + num_reads, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = (np.uint64(seed) << np.uint64(40)) + (np.uint64(1) << np.uint64(39))
        draw = (r >> np.uint64(1)) if paired else r
        start = splitmix64(base + draw) % size
        if paired:
            start = (start + (r & np.uint64(1)) * np.uint64(mate_distance)) % size
        pos = (start[:, None] + np.arange(read_len, dtype=np.uint64)[None, :]) % size
        codes = genome_codes[pos.astype(np.int64)]
        ctr = (r[:, None] * np.uint64(read_len)) + np.arange(read_len, dtype=np.uint64)[None, :]
        h = splitmix64(((np.uint64(seed) + np.uint64(1)) << np.uint64(40)) ^ ctr)
        hit = (h % np.uint64(10000)) < np.uint64(sub_per_10k)
        sub = ((h >> np.uint64(32)) & np.uint64(3)).astype(np.uint8)
    return np.where(hit, sub, codes).astype(np.uint8)


def codes_to_strings(codes: np.ndarray):
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    asc = lut[codes]
    return [row.tobytes().decode("ascii") for row in asc]
