"""CountMinSketch with the reference's interface (countminsketch.py:6-111), kept on the GPU.

Same observable behaviour: ``num_rows`` rows of unsigned-16 cells whose widths are the
primes of ``primes_1_10_7``; one MurmurHash3_x86_32 (seed 0) of the string's bytes, row i
indexed by ``h % width[i]``; ``update`` adds to every row, ``estimate`` / ``[]`` take the
minimum; a cell pushed past 65,535 raises ``OverflowError`` like ``array('H')``;
``num_rows >= 20`` fails the same assertion.  Cells live in device memory (32-bit while
accumulating) and are updated by libga_b200's atomics; ``hash_values`` downloads them as
``array('H')`` rows on demand.  Scalar ``update`` calls are buffered and flushed in bulk.
"""
from __future__ import annotations

import ctypes as C
import sys
from array import array

import numpy as np


def _is_prime(n: int) -> bool:
    if n < 2:
        return False
    if n % 2 == 0:
        return n == 2
    f = 3
    while f * f <= n:
        if n % f == 0:
            return False
        f += 2
    return True


def _primes_around(centre: int, each_side: int = 10):
    """`each_side` largest primes below `centre` and smallest primes above it, ascending."""
    below, above = [], []
    n = centre - 1
    while len(below) < each_side:
        if _is_prime(n):
            below.append(n)
        n -= 1
    n = centre + 1
    while len(above) < each_side:
        if _is_prime(n):
            above.append(n)
        n += 1
    return below[::-1] + above


class CountMinSketch:
    # Row widths: 20 primes centred on 6e7, 1e7 and 5e6 (countminsketch.py:9-24); only
    # primes_1_10_7 is ever indexed (:31-32).
    primes_6_10_7 = _primes_around(6 * 10 ** 7)
    primes_1_10_7 = _primes_around(10 ** 7)
    primes_5_10_6 = _primes_around(5 * 10 ** 6)

    _FLUSH_AT = 1 << 16

    def __init__(self, num_rows, widths=None):
        table = CountMinSketch.primes_1_10_7 if widths is None else list(widths)
        assert num_rows < len(table), \
            "Requested number of rows in CountMinSketch exceeds number of primes in list"
        self.num_rows = num_rows
        self.widths = [int(x) for x in table[:num_rows]]
        self._cells = None          # torch int32 tensor on the device, rows back to back
        self._pending_text, self._pending_amount = [], []
        self._rows_cache = None
        self.source_counts = None   # KmerCounts this sketch was poured from (keeps the table alive)

    # -- device plumbing ---------------------------------------------------------------------
    def _device_cells(self):
        if self._cells is None:
            import torch
            import ga_native as gn
            gn.require_gpu()
            self._cells = torch.zeros(max(1, sum(self.widths)), dtype=torch.int32,
                                      device=torch.device("cuda", torch.cuda.current_device()))
        return self._cells

    def _struct(self):
        from ga_device import sketch_struct
        return sketch_struct(self._device_cells(), self.widths)

    @staticmethod
    def _flatten(strings):
        import torch
        raw = [bytes(ord(c) & 0xFF for c in s) for s in strings]
        off = np.zeros(len(raw) + 1, dtype=np.int64)
        np.cumsum([len(r) for r in raw], out=off[1:])
        blob = np.frombuffer(b"".join(raw) or b"\0", dtype=np.uint8).copy()
        dev = torch.device("cuda", torch.cuda.current_device())
        return torch.from_numpy(blob).to(dev), torch.from_numpy(off).to(dev)

    def _flush(self):
        if not self._pending_text:
            return
        import torch
        import ga_native as gn
        from ga_device import _stream
        blob, off = self._flatten(self._pending_text)
        # two's-complement wrap: a negative amount subtracts, and a cell pushed below zero shows
        # up as > 65535 in the overflow check, as array('H') would refuse it
        amounts = torch.from_numpy((np.array(self._pending_amount, dtype=np.int64) & 0xFFFFFFFF)
                                   .astype(np.uint32).view(np.int32)).to(blob.device)
        sk = self._struct()
        gn.check(gn.lib().ga_sketch_update_bytes(gn.ptr(blob), gn.ptr(off), gn.ptr(amounts),
                                                 len(self._pending_text), C.byref(sk), _stream()))
        self._pending_text, self._pending_amount = [], []
        self._rows_cache = None
        self._check_overflow()

    def _narrow(self):
        """Device cells as one uint16 numpy array; raises where array('H') would have."""
        import torch
        import ga_native as gn
        from ga_device import _stream
        cells = self._device_cells()
        out = torch.empty(cells.numel(), dtype=torch.int16, device=cells.device)
        status = torch.zeros(4, dtype=torch.int32, device=cells.device)
        sk = self._struct()
        gn.check(gn.lib().ga_sketch_narrow(C.byref(sk), gn.ptr(out), gn.ptr(status), _stream()))
        if int(status[0].item()) & gn.ST_U16_OVERFLOW:
            raise OverflowError("unsigned short is greater than maximum")
        return out

    def _check_overflow(self):
        self._narrow()

    # -- bulk interface used by the graph classes ---------------------------------------------
    def pour_counts(self, counts):
        """update(kmer, count) for every entry of a ga_device.KmerCounts, on the device
        (the loop of _make_sketch, debruijn_graph.py:186-187)."""
        import ga_native as gn
        from ga_device import _stream, _timed
        self._flush()
        sk = self._struct()
        table = counts.table                       # counts every window if that has not happened yet
        with _timed("sketch_update", counts.n_occ):
            gn.check(gn.lib().ga_sketch_update_table(gn.ptr(table), counts.capacity, counts.key_words,
                                                     counts.k, counts.alphabet.sym_bits,
                                                     gn.ptr(counts.alphabet.inv_dev), C.byref(sk), _stream()))
        self._rows_cache = None
        self.source_counts = counts
        self._check_overflow()

    def update_many(self, strings, amounts):
        self._pending_text.extend(strings)
        self._pending_amount.extend(int(a) for a in amounts)
        self._flush()

    def estimate_many(self, strings):
        import torch
        import ga_native as gn
        from ga_device import _stream
        self._flush()
        strings = list(strings)
        if not strings:
            return np.zeros(0, dtype=np.int64)
        blob, off = self._flatten(strings)
        out = torch.empty(len(strings), dtype=torch.int32, device=blob.device)
        sk = self._struct()
        gn.check(gn.lib().ga_sketch_estimate_bytes(gn.ptr(blob), gn.ptr(off), len(strings), C.byref(sk),
                                                   gn.ptr(out), _stream()))
        return out.cpu().numpy().astype(np.int64)

    # -- the reference's interface ------------------------------------------------------------
    def update(self, string, amount):
        self._pending_text.append(string)
        self._pending_amount.append(int(amount))
        if len(self._pending_text) >= self._FLUSH_AT:
            self._flush()

    def estimate(self, string):
        return int(self.estimate_many([string])[0])

    def __getitem__(self, key):
        return self.estimate(key)

    @property
    def hash_values(self):
        """List of ``array('H')`` rows (downloaded; cached until the next update)."""
        self._flush()
        if self._rows_cache is None:
            flat = self._narrow().cpu().numpy().view(np.uint16)
            rows, start = [], 0
            for width in self.widths:
                row = array("H")
                row.frombytes(flat[start:start + width].tobytes())
                rows.append(row)
                start += width
            self._rows_cache = rows
        return self._rows_cache

    @staticmethod
    def _hash(data, seed=0):
        """MurmurHash3_x86_32 of the low bytes of ``data`` (host helper for API parity;
        the device kernels carry their own implementation)."""
        m = 0xFFFFFFFF
        raw = bytes(ord(c) & 0xFF for c in data)
        h = seed & m
        for i in range(0, len(raw) & ~3, 4):
            h ^= CountMinSketch._scramble(int.from_bytes(raw[i:i + 4], "little"))
            h = ((h << 13) | (h >> 19)) & m
            h = (h * 5 + 0xE6546B64) & m
        if len(raw) & 3:
            h ^= CountMinSketch._scramble(int.from_bytes(raw[len(raw) & ~3:], "little"))
        h ^= len(raw)
        h ^= h >> 16
        h = (h * 0x85EBCA6B) & m
        h ^= h >> 13
        h = (h * 0xC2B2AE35) & m
        return h ^ (h >> 16)

    @staticmethod
    def _scramble(block):
        block = (block * 0xCC9E2D51) & 0xFFFFFFFF
        block = ((block << 15) | (block >> 17)) & 0xFFFFFFFF
        return (block * 0x1B873593) & 0xFFFFFFFF

    def __sizeof__(self):
        # the figure the reference's -m report prints (countminsketch.py:101-111)
        if hasattr(self, "total_mem"):
            return self.total_mem
        total = sys.getsizeof(self.num_rows) + sys.getsizeof([None for _ in range(self.num_rows)])
        total += sum(2 * width for width in self.widths)
        self.total_mem = total + sys.getsizeof(total)
        return total
