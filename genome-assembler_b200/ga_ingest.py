"""Raw ingest: the bytes of stdin -> one symbol buffer + one length per read, parsed by libga_b200's
``ga_parse_reads`` (the reference's ``IOHandler.read_input`` rules, assemble.py:40-71) -- no Python string
per read.  The buffer is pinned host memory when a GPU is present, so ``DeviceReads`` moves it with one
async copy and packs it on the device.

``RawReads`` quacks like the ``list[str]`` / ``list[(str, str)]`` the reference builds: ``len``, indexing,
iteration (strings are decoded on demand, e.g. for ``-m`` or for user code that looks at the reads).
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Sequence

import numpy as np

import ga_native as gn


class RawReads(Sequence):
    """Reads (or read pairs) held as one byte buffer: ``symbols`` (all reads back to back, mates of a pair
    next to each other) and ``lens`` (one int32 per read / mate)."""

    def __init__(self, symbols: np.ndarray, lens: np.ndarray, paired: bool, keep=None):
        self.symbols, self.lens, self.paired = symbols, lens, bool(paired)
        self._keep = keep                      # the pinned tensor the arrays alias
        self._offsets = None
        self._strings = None

    @property
    def offsets(self) -> np.ndarray:
        if self._offsets is None:
            self._offsets = np.zeros(self.lens.size + 1, dtype=np.int64)
            np.cumsum(self.lens, out=self._offsets[1:])
        return self._offsets

    def __len__(self):
        return self.lens.size // 2 if self.paired else self.lens.size

    def _one(self, i: int) -> str:
        off = self.offsets
        return self.symbols[off[i]:off[i + 1]].tobytes().decode("latin-1")

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return (self._one(2 * i), self._one(2 * i + 1)) if self.paired else self._one(i)

    def __iter__(self):
        if self._strings is None:              # one bulk decode, kept (callers that iterate tend to do it again)
            text = self.symbols[:int(self.offsets[-1])].tobytes().decode("latin-1")
            off = self.offsets.tolist()
            flat = [text[off[j]:off[j + 1]] for j in range(self.lens.size)]
            self._strings = list(zip(flat[0::2], flat[1::2])) if self.paired else flat
        return iter(self._strings)

    def __eq__(self, other):
        return list(self) == list(other)

    def __sizeof__(self):
        """The -m report prints sys.getsizeof(reads) (assemble.py:107-114 of the reference): answer with the size of
        the list of as many appended entries."""
        from py_sizes import appended_list_sizeof, for_instance
        return for_instance(appended_list_sizeof(len(self)))


def parse(raw: bytes):
    """(reads: RawReads, paired, distance, number of bases), or None when the bytes are not plain ASCII or the
    header is not a plain integer (the caller then parses them as text, with Python's own rules and errors).
    Raises ValueError for a malformed read-pair line, as the reference's tuple unpacking does."""
    L = gn.lib()
    n_reads, paired, distance, n_sym = C.c_uint64(), C.c_int(), C.c_int64(), C.c_uint64()
    view = np.frombuffer(raw, dtype=np.uint8)
    text = view.ctypes.data if view.size else None
    if text is None:
        return None
    rc = L.ga_parse_reads(text, view.size, None, None, 0, C.byref(n_reads), C.byref(paired), C.byref(distance),
                          C.byref(n_sym))
    if rc != gn.GA_OK or n_reads.value > (1 << 31):
        return None
    mates = 2 if paired.value else 1
    keep = None
    try:
        import torch
        if L.ga_device_count() > 0:
            keep = torch.empty(max(view.size, 1), dtype=torch.uint8, pin_memory=True)
            symbols = keep.numpy()
        else:
            symbols = np.empty(max(view.size, 1), dtype=np.uint8)
    except ImportError:
        symbols = np.empty(max(view.size, 1), dtype=np.uint8)
    lens = np.zeros(n_reads.value * mates, dtype=np.int32)
    rc = L.ga_parse_reads(text, view.size, symbols.ctypes.data, lens.ctypes.data, lens.size, C.byref(n_reads),
                          C.byref(paired), C.byref(distance), C.byref(n_sym))
    if rc == gn.GA_ERR_BAD_ARG:
        raise ValueError(gn.last_error())
    if rc != gn.GA_OK:
        return None
    reads = RawReads(symbols[:n_sym.value], lens, bool(paired.value), keep)
    return reads, bool(paired.value), int(distance.value), int(n_sym.value)
