"""Host orchestration of the GPU path: read packing, the count table, the sketch-backed or exact
filter, the stamp-based graph build and CSR emission.  Everything numeric happens in
libga_b200.so (see include/ga_b200.h); this module owns device buffers (torch tensors), sizes
hash tables, retries when one fills up, and converts between Python strings and packed keys.

Reference behaviour reproduced here (paths into the upstream checkout):
  read breaking / counting     debruijn_graph.py:144-157, 349-374
  filter (strict >)            debruijn_graph.py:127-128, 275-278
  node / edge insertion order  debruijn_graph.py:113-142, 269-347  (via stamps)
"""
from __future__ import annotations

import ctypes as C
import math

import os as _os_early

import numpy as np
import torch

import ga_native as gn

_DNA = b"ACGT"
# two-phase build thresholds (tests lower them to exercise the path on small inputs)
TWO_PHASE_MIN_TABLE_BYTES = 64 << 20    # id table + stamps beyond this no longer live in the L2
TWO_PHASE_MIN_READS = 1 << 22
TWO_PHASE_MIN_STEP = 1 << 20
# bucketed (super-k-mer) count + build, csrc/ga_superkmer.cu: unpaired DNA, 64-bit keys
SUPERKMER_MIN_OCC = 1 << 22             # smaller inputs stay on the table path (tests set 0 to force buckets)
SUPERKMER_MIN_OCC_PAIRS = 1 << 28       # read pairs: bucketed counting only pays once the tables outgrow the L2
                                        # (C3, 1.0e8 occurrences: 4.5 ms through the buckets, 3.9 ms through the tables)
SUPERKMER_TARGET = int(_os_early.environ.get("GA_SK_TARGET", "8192"))    # windows per bucket aimed for (C2 sweep: 8192 beats 16384 by 38 % on the bucket kernel; C4 sits at the 2^20-bucket cap either way)
SUPERKMER_TABLE_SLOTS = int(_os_early.environ.get("GA_SK_SLOTS", "8192"))   # shared-memory table slots (16 bytes each) per bucket pass (tests shrink it to force spills)
SUPERKMER_INDEX_FORM = _os_early.environ.get("GA_SK_DENSE", "0") != "1"   # single GPU: sort 32-bit indices, not records
SUPERKMER_MAX_SOLID = 16000             # candidate windows (seen twice) per bucket pass (further bounded by the shared-memory pool)
TIMERS = None   # bench.py sets this to {"count": [], "build": []}: CUDA-event pairs around the two hot kernels


import os as _os
import sys as _sys
import time as _time
_TRACE = int(_os.environ.get("GA_TRACE", "0") or 0)
_last = [0.0]


def _mark(name):
    """Stage boundary.  With TIMERS set (bench.py) a CUDA event is recorded, so the GPU-timeline length of
    every stage -- kernels AND the idle gaps in between -- can be reported.  GA_TRACE=1 additionally
    prints host wall time between stages (with a device sync) -- debugging aid."""
    if TIMERS is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        TIMERS.setdefault("_marks", []).append((name, ev))
    if _TRACE:
        if _TRACE == 1:
            torch.cuda.synchronize()
        now = _time.perf_counter()
        print("  [trace r%s] %-28s %8.3f ms" % (_os.environ.get("RANK", "0"), name, (now - _last[0]) * 1e3),
              file=_sys.stderr, flush=True)
        _last[0] = _time.perf_counter()


def _timed(name, occurrences=0):
    """Context manager recording a CUDA-event pair (and the occurrences the launch covers) around a
    kernel launch when TIMERS is set."""
    class _Span:
        def __enter__(self):
            self.on = TIMERS is not None
            if self.on:
                self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                self.a.record()

        def __exit__(self, *exc):
            if self.on:
                self.b.record()
                TIMERS.setdefault(name, []).append((self.a, self.b, occurrences))
    return _Span()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


_TOTAL_MEM = {}


def _free_bytes() -> int:
    """Device memory this process can still take: total minus what torch has reserved, plus what sits
    unused in torch's cache.  cudaMemGetInfo costs milliseconds per call (3.7 ms measured on B200 with
    a large heap) -- more than the kernels of the small workloads it was sizing tables for."""
    dev = torch.cuda.current_device()
    if dev not in _TOTAL_MEM:
        _TOTAL_MEM[dev] = torch.cuda.get_device_properties(dev).total_memory
    # (total - reserved) + (reserved - allocated) = total - allocated.  The allocator's counters come as a nested
    # dict straight from C++; torch.cuda.memory_allocated() flattens all of them in Python first (~0.1 ms per call,
    # twice per graph: 7 % of a C1 step)
    try:
        allocated = int(torch.cuda.memory_stats_as_nested_dict(dev)["allocated_bytes"]["all"]["current"])
    except (KeyError, TypeError):
        allocated = torch.cuda.memory_allocated(dev)
    return max(_TOTAL_MEM[dev] - allocated, 0)


# Persistent device workspace for the large intermediates of the bucketed path (records, bucket-sorted
# records, solid keys + stamps, id table).  They are tens of GB at BASELINE config C4; allocating them
# anew every call makes the caching allocator release and re-acquire segments next to the 180 GB limit,
# which costs more than some of the kernels.  A buffer is reused by the next call on the same device and
# only ever grows.  All users run on the current stream, so reuse is stream-ordered.
_WORKSPACE = {}


def workspace(name: str, shape, dtype) -> torch.Tensor:
    dev = _dev()
    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = 1
    for x in shape:
        n *= x
    need = max(n, 1) * torch.empty(0, dtype=dtype).element_size()
    key = (dev.index, name)
    buf = _WORKSPACE.get(key)
    if buf is None or buf.numel() < need:
        _WORKSPACE.pop(key, None)
        buf = None                                       # drop the old buffer before growing
        buf = torch.empty(need + need // 16 + 512, dtype=torch.uint8, device=dev)
        _WORKSPACE[key] = buf
    return buf[:n * torch.empty(0, dtype=dtype).element_size()].view(dtype).view(shape)


def release_workspace():
    """Give the persistent workspace back to the allocator."""
    _WORKSPACE.clear()


# ----------------------------------------------------------------------------------- alphabet
class Alphabet:
    """Symbol coding of a read set: DNA packs 2 bits per base, anything else 8 bits per stored
    symbol with the smallest code width that separates the symbols present."""

    def __init__(self, present: np.ndarray):
        present = np.asarray(present, dtype=np.int64)
        self.lut = np.full(256, 0xFF, dtype=np.uint8)          # byte -> code
        if present.size == 0 or set(present.tolist()) <= set(_DNA):
            self.storage_bits, self.sym_bits = 2, 2
            symbols = np.frombuffer(_DNA, dtype=np.uint8)
        else:
            self.storage_bits = 8
            self.sym_bits = max(1, int(math.ceil(math.log2(present.size))))
            symbols = present.astype(np.uint8)
        self.lut[symbols] = np.arange(symbols.size, dtype=np.uint8)
        self.inv = np.zeros(256, dtype=np.uint8)               # code -> byte
        self.inv[:symbols.size] = symbols
        self._lut_dev = self._inv_dev = None

    @property
    def lut_dev(self):
        if self._lut_dev is None:
            self._lut_dev = torch.from_numpy(self.lut).to(_dev())
        return self._lut_dev

    @property
    def inv_dev(self):
        if self._inv_dev is None:
            self._inv_dev = torch.from_numpy(self.inv).to(_dev())
        return self._inv_dev

    # -- packed keys <-> strings (first symbol most significant, as in the kernels)
    def pack_key(self, text: str, key_words: int):
        """Packed key of a window as (lo, hi) or None if a symbol is outside the alphabet."""
        value = 0
        for ch in text:
            o = ord(ch)
            if o > 255 or self.lut[o] == 0xFF:
                return None
            value = (value << self.sym_bits) | int(self.lut[o])
        if value >> (64 * key_words):
            return None
        return value & 0xFFFFFFFFFFFFFFFF, value >> 64

    def decode_keys(self, keys: np.ndarray, w: int) -> np.ndarray:
        """keys: (n, key_words) uint64 -> (n, w) uint8 bytes."""
        n = keys.shape[0]
        out = np.empty((n, w), dtype=np.uint8)
        b = self.sym_bits
        mask = np.uint64((1 << b) - 1)
        lo = keys[:, 0]
        hi = keys[:, 1] if keys.shape[1] > 1 else None
        for i in range(w):
            p = (w - 1 - i) * b
            if p >= 64:
                v = hi >> np.uint64(p - 64)
            elif p == 0:
                v = lo
            else:
                v = lo >> np.uint64(p)
                if hi is not None and p + b > 64:
                    v = v | (hi << np.uint64(64 - p))
            out[:, i] = self.inv[(v & mask).astype(np.int64)]
        return out

    def decode_strings(self, keys: np.ndarray, w: int):
        if keys.shape[0] == 0:
            return []
        raw = self.decode_keys(keys, w)
        return [s.decode("latin-1") for s in raw.view("S%d" % w).ravel().tolist()] if w else [""] * len(raw)


# ----------------------------------------------------------------------------------- reads
class DeviceReads:
    """Reads (or read pairs, mates interleaved) packed on the device."""

    def __init__(self, reads, paired: bool, first_read: int = 0, estride: int | None = None):
        gn.require_gpu()
        self.source = reads
        self.paired = bool(paired)
        pinned = None
        if type(reads).__name__ == "RawReads" and reads.paired == self.paired:
            # raw ingest (ga_ingest.parse): the symbols already sit in one (pinned) buffer, lengths in an array
            n = int(reads.lens.size)
            lens = reads.lens                  # int32, as parsed: no copy, no widening pass over a million entries
            buf = reads.symbols
            pinned = reads._keep
        else:
            if self.paired:
                flat = [s for pair in reads for s in (pair[0], pair[1])]
            else:
                flat = reads if isinstance(reads, list) else list(reads)
            n = len(flat)
            lens = np.fromiter(map(len, flat), dtype=np.int64, count=n)
            try:
                raw = "".join(flat).encode("latin-1")
            except UnicodeEncodeError:
                raise ValueError("reads contain characters above U+00FF; the GPU path hashes symbols "
                                 "as single bytes and refuses to alias them") from None
            buf = np.frombuffer(raw, dtype=np.uint8)
        self.n_reads = n
        # raw ingest: try the DNA coding first (the packer flags any other symbol) instead of a histogram of
        # every byte on the host; other inputs get the smallest code that separates the symbols present
        guess_dna = pinned is not None
        self.alphabet = Alphabet(np.zeros(0) if guess_dna or not buf.size else
                                 np.flatnonzero(np.bincount(buf, minlength=256)))
        self.max_len = int(lens.max()) if n else 0
        uniform = n > 0 and int(lens.min()) == self.max_len
        if self.paired and n and not uniform and np.any(lens[1::2] < lens[0::2]):
            raise ValueError("paired reads: mate 2 shorter than mate 1 is not supported "
                             "(the reference would slice truncated k-mers)")
        self.lens = lens
        self._windows = {}
        self.estride = int(estride) if estride is not None else max(self.max_len, 1)
        self.first_read = int(first_read)
        spw = 64 // self.alphabet.storage_bits
        dev = _dev()
        self.status = torch.zeros(4, dtype=torch.int32, device=dev)
        self.uniform = uniform
        if pinned is not None and buf.size:
            ascii_dev = pinned[:buf.size].to(dev, non_blocking=True)
        else:
            ascii_dev = _to_device(buf) if buf.size else torch.zeros(1, dtype=torch.uint8, device=dev)
        while True:
            spw = 64 // self.alphabet.storage_bits
            if uniform or n == 0:
                self.stride_words = max(1, -(-self.max_len // spw))
                self.words = torch.empty(max(1, n * self.stride_words), dtype=torch.int64, device=dev)
                self.offsets = self.lengths = None
                in_off = out_off = None
            else:
                words_per = -(-lens // spw)
                off_words = np.zeros(n + 1, dtype=np.int64)
                np.cumsum(words_per, out=off_words[1:])
                off_bytes = np.zeros(n + 1, dtype=np.int64)
                np.cumsum(lens, out=off_bytes[1:])
                self.stride_words = 0
                self.words = torch.empty(max(1, int(off_words[-1])), dtype=torch.int64, device=dev)
                self.offsets = torch.from_numpy(off_words).to(dev)
                self.lengths = torch.from_numpy(lens.astype(np.int32)).to(dev)
                in_off = torch.from_numpy(off_bytes).to(dev)
                out_off = self.offsets
            if n:
                gn.check(gn.lib().ga_pack_reads(
                    gn.ptr(ascii_dev), gn.ptr(in_off), n, self.max_len if uniform else 0,
                    gn.ptr(self.alphabet.lut_dev), self.alphabet.storage_bits, gn.ptr(self.words),
                    gn.ptr(out_off), self.stride_words, gn.ptr(self.status), _stream()))
            if not (guess_dna and n and _check_status(self.status) & gn.ST_BAD_SYMBOL):
                break
            guess_dna = False                  # not DNA after all: code the symbols that are there, pack again
            self.status.zero_()
            self.alphabet = Alphabet(np.flatnonzero(np.bincount(buf, minlength=256)))
        self._struct = None

    @classmethod
    def from_packed(cls, words: torch.Tensor, n_reads: int, read_len: int, paired: bool,
                    first_read: int = 0, estride: int | None = None, alphabet: "Alphabet | None" = None):
        """Wrap uniform-length reads that already sit packed on the device (2-bit DNA unless an
        8-bit alphabet is given): the bulk-ingest and synthetic-generator route, which never
        materialises Python strings."""
        gn.require_gpu()
        self = cls.__new__(cls)
        self.source = words
        self.paired = bool(paired)
        self.n_reads = int(n_reads)
        self.alphabet = alphabet if alphabet is not None else Alphabet(np.zeros(0))
        self.lens = None
        self._windows = {}
        self.max_len = int(read_len)
        self.uniform = True
        self.estride = int(estride) if estride is not None else max(self.max_len, 1)
        self.first_read = int(first_read)
        spw = 64 // self.alphabet.storage_bits
        self.stride_words = max(1, -(-self.max_len // spw))
        self.words = words
        self.offsets = self.lengths = None
        self.status = torch.zeros(4, dtype=torch.int32, device=words.device)
        self._struct = None
        return self

    @classmethod
    def from_ascii(cls, ascii_host: torch.Tensor, n_reads: int, read_len: int, paired: bool,
                   alphabet: "Alphabet | None" = None, **kw):
        """Uniform reads as one (pinned) host byte tensor of n_reads*read_len symbols: async H2D copy,
        then 2-bit packing on the device (replaces assemble.py:45-71's per-line string list)."""
        gn.require_gpu()
        alphabet = alphabet if alphabet is not None else Alphabet(np.zeros(0))
        dev = _dev()
        ascii_dev = ascii_host.to(dev, non_blocking=True)
        spw = 64 // alphabet.storage_bits
        stride = max(1, -(-int(read_len) // spw))
        words = torch.empty(max(1, n_reads * stride), dtype=torch.int64, device=dev)
        self = cls.from_packed(words, n_reads, read_len, paired, alphabet=alphabet, **kw)
        if n_reads:
            gn.check(gn.lib().ga_pack_reads(gn.ptr(ascii_dev), None, n_reads, read_len, gn.ptr(alphabet.lut_dev),
                                            alphabet.storage_bits, gn.ptr(words), None, stride,
                                            gn.ptr(self.status), _stream()))
        return self

    def windows_total(self, k: int) -> int:
        """Number of (k-1)-mer occurrences (both mates when paired)."""
        w = k - 1
        if self.n_reads == 0:
            return 0
        if self.lens is None or self.uniform:
            return self.n_reads * max(self.max_len - w + 1, 0)
        if k not in self._windows:             # ragged reads: one pass over the lengths per k, not per caller
            eff = np.repeat(self.lens[0::2], 2) if self.paired else self.lens
            self._windows[k] = int(np.maximum(eff.astype(np.int64) - w + 1, 0).sum())
        return self._windows[k]

    def struct(self) -> gn.GaReads:
        if self._struct is None:
            s = gn.GaReads()
            s.words = gn.ptr(self.words)
            s.offsets = gn.ptr(self.offsets)
            s.lengths = gn.ptr(self.lengths)
            s.n_reads = self.n_reads
            s.first_read = self.first_read
            s.uniform_len = self.max_len if self.uniform else 0
            s.stride_words = self.stride_words
            s.storage_bits = self.alphabet.storage_bits
            s.sym_bits = self.alphabet.sym_bits
            s.paired = int(self.paired)
            s.estride = self.estride
            self._struct = s
        return self._struct

    def struct_range(self, r0: int, r1: int) -> gn.GaReads:
        """Descriptor of reads [r0, r1) of this (unpaired) set; stamps keep their global index."""
        base = self.struct()
        s = gn.GaReads()
        for name, _ in gn.GaReads._fields_:
            setattr(s, name, getattr(base, name))
        s.n_reads = r1 - r0
        s.first_read = self.first_read + r0
        if self.offsets is None:
            s.words = self.words.data_ptr() + r0 * self.stride_words * 8
        else:
            s.offsets = self.offsets.data_ptr() + r0 * 8
            s.lengths = self.lengths.data_ptr() + r0 * 4
        return s

    def key_words(self, k: int) -> int:
        kw = gn.lib().ga_key_words(k, self.alphabet.sym_bits)
        if kw == 0:
            raise ValueError("kmer length %d with a %d-bit alphabet needs more than 127 key bits; "
                             "supported: (k-1)*bits <= 127" % (k, self.alphabet.sym_bits))
        return kw


def _to_device(arr: np.ndarray) -> torch.Tensor:
    """Host array -> device through a pinned staging buffer and an async copy."""
    pinned = torch.empty(arr.shape, dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[...] = arr
    return pinned.to(_dev(), non_blocking=True)


def _check_status(status: torch.Tensor) -> int:
    return int(status[0].item())


# ----------------------------------------------------------------------------------- counting
class KmerCounts:
    """Exact (k-1)-mer counts on the device, quacking like the dict the reference's
    ``_count_kmers`` returns (``[]``, ``items()``, truthiness, ``len``).

    Two device representations, both exact:
      * the full table (every distinct window), filled on first use by anything that needs all
        counts: ``items()``, ``len()``, ``[]``, a CountMinSketch pour;
      * the candidate table of ``candidates(threshold)``: only windows whose pre-filter cell says
        they may exceed the threshold -- all the graph build needs, at a fraction of the traffic.
    """

    def __init__(self, k: int, reads: DeviceReads):
        self.k, self.w = k, k - 1
        self.reads = reads
        self.alphabet = reads.alphabet
        self.key_words = reads.key_words(k)
        self.slot_bytes = gn.lib().ga_slot_bytes(self.key_words)
        self.n_occ = reads.windows_total(k)
        self._table = None
        self.capacity = 0
        self._summary = {}
        self._cand = {}

    @property
    def table(self):
        if self._table is None:
            self._count()
        return self._table

    def _count(self):
        L = gn.lib()
        dev = _dev()
        free = _free_bytes()
        limit = max(1024, int(free * 0.6) // self.slot_bytes)
        cap = min(max(1024, int(self.n_occ * 1.25) + 64), limit)
        while True:
            self._table = torch.empty(cap * self.slot_bytes, dtype=torch.uint8, device=dev)
            self.capacity = cap
            self.reads.status.zero_()
            gn.check(L.ga_table_clear(gn.ptr(self._table), cap, self.key_words, _stream()))
            with _timed("count_full", self.n_occ):
                gn.check(L.ga_count_kmers(C.byref(self.reads.struct()), self.k, gn.ptr(self._table), cap,
                                          gn.ptr(self.reads.status), _stream()))
            st = _check_status(self.reads.status)
            if st & gn.ST_BAD_SYMBOL:
                raise ValueError("read symbol outside the detected alphabet")
            if not st & gn.ST_TABLE_FULL:
                return
            if cap >= limit:
                raise MemoryError("k-mer count table does not fit in device memory")
            self._table = None
            cap = min(cap * 2, limit)

    def candidates(self, threshold: int):
        """(table, capacity) holding the exact count of every window that may exceed `threshold`
        (pre-filter sketch -> candidate table, ga_prefilter.cu), or None when the threshold is
        outside what the sketch cells can represent."""
        if threshold in self._cand:
            return self._cand[threshold]
        if threshold < 0 or threshold + 1 > 255:
            return None
        L = gn.lib()
        dev = _dev()
        bits = 4 if threshold + 1 <= 15 else 8
        _mark("enter candidates")
        free = _free_bytes()
        _mark("mem_get_info")
        n_cells = max(1 << 16, min(self.n_occ, int(free * 0.25) * 8 // bits))
        words = torch.zeros((n_cells * bits + 31) // 32, dtype=torch.int32, device=dev)
        pf = gn.GaPrefilter()
        pf.words, pf.n_cells, pf.cell_bits = gn.ptr(words), n_cells, bits
        status = self.reads.status
        with _timed("prefilter", self.n_occ):
            gn.check(L.ga_prefilter_update(C.byref(self.reads.struct()), self.k, C.byref(pf), threshold, _stream()))
        _mark("prefilter zero+update")
        n_hot = torch.zeros(1, dtype=torch.int64, device=dev)
        gn.check(L.ga_prefilter_hot(C.byref(pf), threshold, gn.ptr(n_hot), _stream()))
        hot = int(n_hot.item())
        _mark("prefilter_hot")
        limit = max(1024, int(free * 0.5) // self.slot_bytes)
        cap = min(limit, max(1024, int(hot * 2.2) + 1024))
        while True:
            table = torch.empty(cap * self.slot_bytes, dtype=torch.uint8, device=dev)
            status.zero_()
            gn.check(L.ga_table_clear(gn.ptr(table), cap, self.key_words, _stream()))
            with _timed("count", self.n_occ):
                gn.check(L.ga_count_candidates(C.byref(self.reads.struct()), self.k, C.byref(pf), threshold,
                                               gn.ptr(table), cap, gn.ptr(status), _stream()))
            out = torch.zeros(4, dtype=torch.int64, device=dev)
            gn.check(L.ga_table_summary(gn.ptr(table), cap, self.key_words, int(threshold), gn.ptr(out), _stream()))
            st = _check_status(status)
            _mark("clear+count_candidates+summary")
            if st & gn.ST_BAD_SYMBOL:
                raise ValueError("read symbol outside the detected alphabet")
            if not st & gn.ST_TABLE_FULL:
                break
            if cap >= limit:
                raise MemoryError("candidate table does not fit in device memory")
            cap = min(cap * 2, limit)
        summary = tuple(int(v) for v in out.cpu().tolist())
        self._cand[threshold] = (table, cap, summary)
        return self._cand[threshold]

    def summary(self, threshold: int):
        """(distinct, above threshold, occurrences, max count)."""
        if threshold not in self._summary:
            out = torch.zeros(4, dtype=torch.int64, device=_dev())
            gn.check(gn.lib().ga_table_summary(gn.ptr(self.table), self.capacity, self.key_words,
                                               int(threshold), gn.ptr(out), _stream()))
            self._summary[threshold] = tuple(int(v) for v in out.cpu().tolist())
        return self._summary[threshold]

    def export(self, min_exclusive: int = -1):
        """(keys (n, key_words) uint64, counts (n,) uint32) on the host, table order."""
        n_max = self.summary(min_exclusive)[1] if min_exclusive >= 0 else self.summary(0)[0]
        dev = _dev()
        keys = torch.empty((max(n_max, 1), self.key_words), dtype=torch.int64, device=dev)
        counts = torch.empty(max(n_max, 1), dtype=torch.int32, device=dev)
        n_out = torch.zeros(1, dtype=torch.int64, device=dev)
        gn.check(gn.lib().ga_table_export(gn.ptr(self.table), self.capacity, self.key_words,
                                          int(min_exclusive), gn.ptr(keys), gn.ptr(counts), gn.ptr(n_out),
                                          _stream()))
        n = int(n_out.item())
        return (keys[:n].cpu().numpy().view(np.uint64), counts[:n].cpu().numpy().view(np.uint32))

    # -- dict-like surface -----------------------------------------------------------------
    def __len__(self):
        return self.summary(0)[0]

    def __bool__(self):
        return self.n_occ > 0

    def __sizeof__(self):
        """The -m report prints sys.getsizeof(counts) (debug_graph.py:51-63 of the reference): answer with the size
        of its defaultdict(int) holding as many string keys."""
        from py_sizes import for_instance, grown_str_dict_sizeof
        return for_instance(grown_str_dict_sizeof(len(self)))

    def lookup_many(self, kmers):
        keys = np.zeros((len(kmers), self.key_words), dtype=np.uint64)
        known = np.zeros(len(kmers), dtype=bool)
        for i, text in enumerate(kmers):
            packed = self.alphabet.pack_key(text, self.key_words) if len(text) == self.w else None
            if packed is not None:
                keys[i, 0] = packed[0]
                if self.key_words > 1:
                    keys[i, 1] = packed[1]
                known[i] = True
        out = torch.zeros(max(len(kmers), 1), dtype=torch.int32, device=_dev())
        if len(kmers):
            kd = torch.from_numpy(keys.view(np.int64)).to(_dev())
            gn.check(gn.lib().ga_table_lookup(gn.ptr(self.table), self.capacity, self.key_words, gn.ptr(kd),
                                              len(kmers), gn.ptr(out), _stream()))
        res = out[:len(kmers)].cpu().numpy().astype(np.int64)
        res[~known] = 0
        return res

    def __getitem__(self, kmer):
        return int(self.lookup_many([kmer])[0])

    def get(self, kmer, default=None):
        v = self[kmer]
        return v if v else default

    def __contains__(self, kmer):
        return self[kmer] > 0

    def items(self):
        keys, counts = self.export()
        return zip(self.alphabet.decode_strings(keys, self.w), (int(c) for c in counts))

    def keys(self):
        return (k for k, _ in self.items())

    def values(self):
        return (v for _, v in self.items())

    def __iter__(self):
        return self.keys()


def sketch_struct(cells: torch.Tensor, widths) -> gn.GaSketch:
    s = gn.GaSketch()
    s.cells = gn.ptr(cells)
    s.rows = len(widths)
    for i, wdt in enumerate(widths):
        s.width[i] = int(wdt)
    return s


# ----------------------------------------------------------------------------------- buckets
def superkmer_supported(reads: "DeviceReads", k: int, threshold: int, counting_only: bool = False) -> bool:
    """The bucketed path covers 2-bit reads whose windows fit 62 bits: count + build for unpaired reads,
    counting alone (the solid set) for read pairs too."""
    if (reads.paired and not counting_only) or reads.alphabet.storage_bits != 2 or not 2 <= k <= 32:
        return False
    if not 0 <= threshold <= 60000:
        return False
    return (reads.first_read + reads.n_reads) * max(reads.estride, 1) < (1 << 47)


def _record_groups(reads: "DeviceReads", w: int) -> int:
    """Upper bound on the extra records caused by cutting runs at 32-window groups."""
    if reads.lens is None or reads.uniform:
        return reads.n_reads * max(-(-max(reads.max_len - w + 1, 0) // 32), 0)
    win = np.maximum(reads.lens.astype(np.int64) - w + 1, 0)
    return int((-(-win // 32)).sum())


def sk_records_estimate(reads: "DeviceReads", k: int) -> int:
    """Records this read set will be cut into: two per window span of a minimizer, plus the cuts at 32-window groups."""
    w = k - 1
    per_window = w - gn.lib().ga_sk_minimizer_len(k) + 1
    return int(reads.windows_total(k) * 2.0 / (per_window + 1)) + _record_groups(reads, w)


def sk_l1_capacity(reads: "DeviceReads", k: int, l1_bits: int, records: int = 0) -> int:
    """Slots per level-1 bucket: the expected records (of this read set, or `records`) with 8 % to spare."""
    return int((records or sk_records_estimate(reads, k)) / (1 << l1_bits) * 1.08) + 4096


def sk_geometry(n_occ: int):
    """(level-1 bits, level-2 bits) of the bucket id for `n_occ` window occurrences IN TOTAL (all
    ranks): about SUPERKMER_TARGET windows per bucket, at most 2^20 buckets."""
    bits = 0
    while bits < 20 and (n_occ >> bits) > SUPERKMER_TARGET:
        bits += 1
    l2_bits = min(10, bits)
    return bits - l2_bits, l2_bits


class ScatterOverflow(Exception):
    """A level-1 bucket outgrew a capacity the caller fixed (sk_scatter_local(..., cap1=...)); `.needed` slots would do."""

    def __init__(self, needed: int):
        super().__init__("level-1 bucket overflow: %d slots needed" % needed)
        self.needed = needed


def sk_scatter_local(reads: "DeviceReads", k: int, l1_bits: int, l2_bits: int, feed=None, dense: bool = True,
                     level2: bool = True, cap1: int = 0, alloc=None):
    """This rank's reads -> records sorted by bucket.  dense: (bases int64[2*total], meta int64[total],
    offsets int64[n_buckets+1], hist int64[n_buckets] = records << 32 | windows, total) -- the records
    themselves in bucket order, what the multi-GPU exchange sends.  Not dense: (level-1 slots int64[4 * n_l1 *
    l1_capacity], None, offsets, hist, total, index int32[total], l1_capacity) -- the records stay in the 32-byte
    slots of their level-1 buckets and only a 32-bit index per record is sorted.  level2=False stops after the
    offsets: (slots, None, offsets, hist, total, None, l1_capacity, level-1 cursors, exact-offset cursors).
    ga_sk_scatter_reads + ga_sk_offsets + ga_sk_scatter_buckets."""
    L = gn.lib()
    dev = _dev()
    w = k - 1
    n_occ = reads.windows_total(k)
    per_window = w - L.ga_sk_minimizer_len(k) + 1
    n_l1, n_buckets = 1 << l1_bits, 1 << (l1_bits + l2_bits)
    status = reads.status
    fixed = cap1 > 0                          # the caller owns the geometry (multi-GPU: the same on every rank)
    alloc = alloc or workspace
    if not fixed:
        cap1 = sk_l1_capacity(reads, k, l1_bits)
    cstride = L.ga_sk_cursor_stride()
    while True:
        rec = alloc("sk_l1_records", n_l1 * cap1 * 4, torch.int64)            # 32-byte slots: bases hi, lo, meta, 0
        cursors1 = torch.zeros(n_l1 * cstride, dtype=torch.int64, device=dev)
        hist = torch.zeros(n_buckets, dtype=torch.int64, device=dev)
        status.zero_()
        with _timed("sk_scatter1", n_occ):
            for r0, r1 in (feed if feed is not None else ((0, reads.n_reads),)):
                gn.check(L.ga_sk_scatter_reads(C.byref(reads.struct_range(r0, r1)), k, l1_bits, l2_bits,
                                               gn.ptr(rec), cap1, gn.ptr(cursors1), gn.ptr(hist), gn.ptr(status),
                                               _stream()))
            feed = None                     # a retry finds every read resident
        offsets = torch.empty(n_buckets + 1, dtype=torch.int64, device=dev)
        cursors2 = torch.empty(n_buckets, dtype=torch.int64, device=dev)
        with _timed("sk_offsets"):
            gn.check(L.ga_sk_offsets(gn.ptr(hist), n_buckets, gn.ptr(offsets), gn.ptr(cursors2), _stream()))
        total = int(offsets[n_buckets].item())
        st = _check_status(status)
        if st & gn.ST_BAD_SYMBOL:
            raise ValueError("read symbol outside the alphabet")
        if not st & gn.ST_TABLE_FULL:
            break
        needed = int(int(cursors1.max().item()) * 1.05) + 4096    # cursors kept counting past the capacity
        if fixed:
            raise ScatterOverflow(needed)
        # level-1 buckets are slabs of one size: a read set whose records crowd into one of them (a window repeated
        # tens of millions of times) would need that size n_l1 times over
        if needed >= (1 << 25) or n_l1 * needed * 32 > 0.8 * _free_bytes() + rec.numel() * 8:
            raise gn.GaBucketLimit("bucketed count: one level-1 bucket needs %d slots" % needed)
        cap1 = needed
        del rec                              # the view keeps the old slab alive: drop it before workspace() grows
    _mark("sk scatter reads")
    if not level2:        # the caller splits the level-1 buckets itself (ga_multi: split + send in one kernel)
        return rec, None, offsets, hist, total, None, cap1, cursors1, cursors2
    if not dense:
        index = alloc("sk_index", max(total, 1), torch.int32)
        with _timed("sk_scatter2", n_occ):
            gn.check(L.ga_sk_scatter_buckets(gn.ptr(rec), cap1, gn.ptr(cursors1), l1_bits, l2_bits, gn.ptr(cursors2),
                                             None, None, gn.ptr(index), _stream()))
        _mark("sk scatter buckets")
        return rec, None, offsets, hist, total, index, cap1
    bases = workspace("sk_bases", max(total, 1) * 2, torch.int64)
    meta = workspace("sk_meta", max(total, 1), torch.int64)
    with _timed("sk_scatter2", n_occ):
        gn.check(L.ga_sk_scatter_buckets(gn.ptr(rec), cap1, gn.ptr(cursors1), l1_bits, l2_bits, gn.ptr(cursors2),
                                         gn.ptr(bases), gn.ptr(meta), None, _stream()))
    _mark("sk scatter buckets")
    return bases, meta, offsets, hist, total


def sk_bucket_pass(bases, meta, offsets, n_segments: int, hist, n_buckets: int, k: int, threshold: int,
                   n_occ: int, status, index=None, l1_capacity: int = 0, l2_bits: int = 0, want_stamps: bool = True,
                   sources=None, out=None):
    """Bucket-sorted records -> (solid keys (cap, 1) int64, n_solid, candidate edge stamps int64[4*cap]).
    offsets: n_segments rows of n_buckets+1 record positions (one row on a single GPU, one per source
    rank after the multi-GPU exchange); hist[b] & 0xFFFFFFFF = windows of bucket b over all segments.
    ga_sk_count_build (+ ga_sk_count_build_spill for what does not fit shared memory).
    out = (keys buffer, stamps buffer, capacity): write the result there instead (buffers shared by all ranks with
    sources.solid_counter set -- the caller then reads the total after a barrier; returns (keys, None, stamps))."""
    L = gn.lib()
    dev = hist.device
    # the state word of a table slot names a record (its number inside the bucket, or its slot inside the level-1
    # bucket) in 25 bits; that also keeps the 31-bit hand-out counter and the packed histogram inside their bits
    if n_occ >= (1 << 25) and bool((((hist >> 32) >= (1 << 25)) | (hist < 0)).any().item()):
        raise gn.GaBucketLimit("bucketed count: a bucket holds 2^25 records or more (one repeated window?)")
    if l1_capacity >= (1 << 25):
        raise gn.GaBucketLimit("bucketed count: level-1 buckets of 2^25 slots or more")
    _mark("sk bucket: checks")
    # solid windows are at most n_occ / (threshold + 1); start from a guess and grow on demand (a second run of the
    # kernel).  Big inputs: one solid window per 48 occurrences (sequencing depth >= 60x; C4: one per 212).  Inputs
    # whose whole bound is small get room for one per 8 occurrences, so that a 30x read set such as C2 (one solid
    # window per 21 occurrences) is not counted twice.
    out_cap = max(1 << 20, min(n_occ // (int(threshold) + 1), max(n_occ // 48, min(n_occ // 8, 16 << 20))) + 1024)
    spill_cap = 1 << 16 if out is None else max(1 << 16, 2 * n_buckets)   # shared output: the pass cannot be repeated
    while True:
        counters = torch.zeros(16, dtype=torch.int64, device=dev)      # [8..14]: phase cycles of a GA_SB_PROFILE build
        spill_list = torch.empty(spill_cap, dtype=torch.int64, device=dev)
        if out is not None:
            solid_keys, edge_stamp, out_cap = out
        else:
            solid_keys = workspace("sk_solid_keys", (out_cap, 1), torch.int64)
            edge_stamp = workspace("sk_edge_stamp", 4 * out_cap, torch.int64) if want_stamps else None
        status.zero_()
        _mark("sk bucket: buffers")
        with _timed("sk_bucket", n_occ):
            if sources is not None:         # records gathered from every source rank's own slots (local or NVLink)
                gn.check(L.ga_sk_count_build_from(C.byref(sources), gn.ptr(offsets), gn.ptr(hist), n_buckets, k,
                                                  int(threshold), SUPERKMER_TABLE_SLOTS, SUPERKMER_MAX_SOLID,
                                                  gn.ptr(solid_keys), gn.ptr(edge_stamp), out_cap, gn.ptr(counters),
                                                  gn.ptr(spill_list), spill_cap, gn.ptr(status), l2_bits, _stream()))
            else:
                gn.check(L.ga_sk_count_build(gn.ptr(bases), gn.ptr(meta), gn.ptr(offsets), n_segments, gn.ptr(hist),
                                             n_buckets, k, int(threshold), SUPERKMER_TABLE_SLOTS, SUPERKMER_MAX_SOLID,
                                             gn.ptr(solid_keys), gn.ptr(edge_stamp), out_cap, gn.ptr(counters),
                                             gn.ptr(spill_list), spill_cap, gn.ptr(status), gn.ptr(index), l1_capacity,
                                             l2_bits, _stream()))
        _mark("sk bucket: kernel")
        host_counters = counters.cpu().tolist()
        _, n_solid, n_spill, n_pass, n_distinct, n_cand = (int(v) for v in host_counters[:6])
        if _TRACE and host_counters[14]:
            whole = float(host_counters[14])
            print("  [trace] bucket kernel warp-cycles: clear %.1f%% walk %.1f%% wait-after-walk %.1f%% notes %.1f%% "
                  "output %.1f%% (body %.1f%% of the kernel); %d of %d windows reached the table after the merge of "
                  "identical records" % (tuple(100.0 * host_counters[8 + i] / whole for i in range(6)) +
                                         (host_counters[15], n_occ)), file=_sys.stderr, flush=True)
        if _TRACE:
            print("  [trace] bucket passes %d (failed %d) over %d buckets, %d solid, %d spilled, %d distinct, "
                  "%d candidates" % (n_pass & 0xFFFFFFFF, n_pass >> 32, n_buckets, n_solid, n_spill, n_distinct,
                                     n_cand), file=_sys.stderr, flush=True)
        if n_spill > spill_cap:
            if out is not None:
                raise gn.GaError("bucketed count: more passes to spill than the list holds (shared output)")
            spill_cap = n_spill
            continue
        if n_spill:
            windows = int((hist[spill_list[:n_spill] & 0xFFFFFFFF] & 0xFFFFFFFF).max().item())
            slots = 256
            while slots < 2 * windows:
                slots <<= 1
            per_cta = int(L.ga_sk_spill_scratch_bytes(slots))
            free = _free_bytes()
            n_ctas = max(1, min(n_spill, 148, int(free * 0.4) // per_cta))
            scratch = torch.empty(n_ctas * per_cta, dtype=torch.uint8, device=dev)
            with _timed("sk_bucket_spill"):
                if sources is not None:
                    gn.check(L.ga_sk_count_build_spill_from(C.byref(sources), gn.ptr(offsets), n_buckets,
                                                            gn.ptr(spill_list), n_spill, k, int(threshold), slots,
                                                            gn.ptr(scratch), n_ctas, gn.ptr(solid_keys),
                                                            gn.ptr(edge_stamp), out_cap, gn.ptr(counters),
                                                            gn.ptr(status), l2_bits, _stream()))
                else:
                    gn.check(L.ga_sk_count_build_spill(gn.ptr(bases), gn.ptr(meta), gn.ptr(offsets), n_segments,
                                                       n_buckets, gn.ptr(spill_list), n_spill, k, int(threshold), slots,
                                                       gn.ptr(scratch), n_ctas, gn.ptr(solid_keys), gn.ptr(edge_stamp),
                                                       out_cap, gn.ptr(counters), gn.ptr(status), gn.ptr(index),
                                                       l1_capacity, l2_bits, _stream()))
            n_solid = int(counters[1].item())
            del scratch
        if _check_status(status) & gn.ST_TABLE_FULL:
            raise gn.GaError("bucketed count: a spilled bucket overflowed its scratch table")
        if out is not None:
            _mark("sk bucket pass")
            return solid_keys, None, edge_stamp
        if n_solid <= out_cap:
            break
        out_cap = n_solid
        del solid_keys, edge_stamp           # views of the old buffers: drop them before workspace() grows
    _mark("sk bucket pass")
    return solid_keys, n_solid, edge_stamp


def superkmer_solid(reads: "DeviceReads", k: int, threshold: int):
    """(solid keys (n, 1) int64, n) through the buckets, counting only -- paired reads too (both mates
    count as plain reads, debruijn_graph.py:349-367)."""
    n_occ = reads.windows_total(k)
    l1_bits, l2_bits = sk_geometry(n_occ)
    out = sk_scatter_local(reads, k, l1_bits, l2_bits, None, dense=not SUPERKMER_INDEX_FORM)
    bases, meta, offsets, hist = out[:4]
    index, cap1 = (out[5], out[6]) if SUPERKMER_INDEX_FORM else (None, 0)
    keys, n_solid, _ = sk_bucket_pass(bases, meta, offsets, 1, hist, 1 << (l1_bits + l2_bits), k, threshold, n_occ,
                                      reads.status, index, cap1, l2_bits, want_stamps=False)
    return keys, n_solid


def superkmer_stamps(reads: "DeviceReads", k: int, threshold: int, feed=None):
    """reads -> (solid keys (n, 1) int64, n, candidate edge stamps int64[4n]) through the bucketed
    pipeline of csrc/ga_superkmer.cu: scatter records to level-1 buckets, split into final buckets,
    count + stamp per bucket in shared memory.  Exact: same solid set and stamps as counting every
    window in one table (debruijn_graph.py:144-152, 113-142).

    `feed` (optional) yields read ranges (r0, r1) as they become resident on the device -- the
    host-buffer entry streams chunks in while earlier chunks are already being scattered."""
    n_occ = reads.windows_total(k)
    l1_bits, l2_bits = sk_geometry(n_occ)
    out = sk_scatter_local(reads, k, l1_bits, l2_bits, feed, dense=not SUPERKMER_INDEX_FORM)
    bases, meta, offsets, hist = out[:4]
    index, cap1 = (out[5], out[6]) if SUPERKMER_INDEX_FORM else (None, 0)
    return sk_bucket_pass(bases, meta, offsets, 1, hist, 1 << (l1_bits + l2_bits), k, threshold, n_occ, reads.status,
                          index, cap1, l2_bits)


def resolve_and_emit(graph, solid_keys, n_solid: int, edge_stamp, k: int, alphabet, status, to_host: bool):
    """Solid keys + candidate edge stamps (the output of the bucket pass) -> CSR: id table over the
    solid keys, ga_sk_resolve, then the shared DNA CSR emission."""
    L = gn.lib()
    dev = solid_keys.device
    kw = 1
    if n_solid >= (1 << 31) - 64:        # node ids, CSR row pointers and columns are 32-bit
        raise gn.GaError("graph too large for the 32-bit CSR: %d solid windows" % n_solid)
    solid_cap = int(1.7 * n_solid) + 64
    solid = torch.empty(solid_cap * L.ga_slot_bytes(kw), dtype=torch.uint8, device=dev)
    status.zero_()
    with _timed("id_table"):
        gn.check(L.ga_table_clear(gn.ptr(solid), solid_cap, kw, _stream()))
        gn.check(L.ga_table_insert_ids(gn.ptr(solid_keys), n_solid, kw, 0, gn.ptr(solid), solid_cap, gn.ptr(status),
                                       _stream()))
    node_stamp = torch.full((n_solid,), -1, dtype=torch.int64, device=dev)
    with _timed("sk_resolve"):
        gn.check(L.ga_sk_resolve(gn.ptr(solid_keys), n_solid, k, gn.ptr(solid), solid_cap, gn.ptr(edge_stamp),
                                 gn.ptr(node_stamp), _stream()))
    out = emit_dna4(graph, node_stamp, edge_stamp, n_solid, solid_keys, solid, solid_cap, kw, k, alphabet, to_host)
    if _check_status(status) & gn.ST_TABLE_FULL:
        raise gn.GaError("id table overflow")
    return out


# ----------------------------------------------------------------------------------- build
class BuiltGraph:
    """The CSR contract of SURVEY App. C.3 on the host."""

    def __init__(self, paired, w, alphabet, key_words):
        self.paired, self.w, self.alphabet, self.key_words = paired, w, alphabet, key_words
        self.n_nodes = self.n_edges = self.num_edges_attr = 0
        z = np.zeros(0, dtype=np.int32)
        self.rowptr, self.col, self.indeg = np.zeros(1, dtype=np.int32), z, z
        self.branching = self.last_char = np.zeros(0, dtype=np.uint8)
        self.keys_a = np.zeros((0, key_words), dtype=np.uint64)
        self.keys_b = None

    def node_strings(self):
        a = self.alphabet.decode_strings(self.keys_a, self.w)
        if not self.paired:
            return a
        return list(zip(a, self.alphabet.decode_strings(self.keys_b, self.w)))

    def checksum(self) -> str:
        """64-bit fingerprint of the graph (row pointers, columns, in-degrees, was_branching, node keys in node
        order; the last symbol is part of the key): position-weighted wrap-around sums, computed where the arrays are
        (device tensors of a `to_host=False` build, or the host arrays).  Two builds of the same reads -- one GPU or
        eight, any exchange route -- must print the same value; bench.py reports it next to the node / edge counts."""
        nn, ne = self.n_nodes, self.n_edges
        dev = getattr(self, "device", None)
        if dev is not None:
            arrs = [dev["rowptr"][:nn + 1], dev["col"][:ne], dev["indeg"][:nn], dev["branching"][:nn], dev["keys_a"][:nn]]
            if dev.get("keys_b") is not None:
                arrs.append(dev["keys_b"][:nn])
        else:
            host = [self.rowptr[:nn + 1], self.col[:ne], self.indeg[:nn], self.branching[:nn], self.keys_a[:nn]]
            if self.keys_b is not None:
                host.append(self.keys_b[:nn])
            arrs = [torch.from_numpy(np.ascontiguousarray(a).view(np.int64) if a.dtype == np.uint64
                                     else np.ascontiguousarray(a)) for a in host]
        acc, m64 = 0, (1 << 64) - 1
        for x in arrs:
            x = x.reshape(-1).to(torch.int64)
            weight = torch.arange(1, x.numel() + 1, dtype=torch.int64, device=x.device) * -0x61C8864680B583EB | 1
            part = int((x * weight).sum().item()) if x.numel() else 0          # int64 arithmetic wraps: sums mod 2^64
            acc = (acc * 0x100000001B3 + (part & m64) + x.numel()) & m64
        return "%016x" % acc

    def contigs(self):
        """Host traversal (libga_b200's ga_traverse_contigs) -> (contigs, edges left, left per node)."""
        L = gn.lib()
        text, offs, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        arrs = [np.ascontiguousarray(a) for a in (self.rowptr, self.col, self.indeg, self.branching, self.last_char)]
        left = np.zeros(max(self.n_nodes, 1), dtype=np.int32)
        gn.check(L.ga_traverse_contigs(*(a.ctypes.data for a in arrs), self.n_nodes, self.num_edges_attr,
                                       int(self.paired), C.byref(text), C.byref(offs), C.byref(n),
                                       left.ctypes.data))
        try:
            count = n.value
            off = np.ctypeslib.as_array(C.cast(offs, C.POINTER(C.c_uint64)), shape=(count + 1,)).copy()
            total = int(off[-1])
            raw = C.string_at(text, total) if total else b""
        finally:
            L.ga_free_host(text)
            L.ga_free_host(offs)
        out = [raw[int(off[i]):int(off[i + 1])].decode("latin-1") for i in range(count)]
        return out, self.num_edges_attr - total, left[:self.n_nodes]


def _solid_keys(counts: KmerCounts, threshold: int, sketch=None):
    """Device array of the keys that pass the filter, and their number."""
    L = gn.lib()
    dev = _dev()
    cand = None
    if sketch is None and counts._table is None:
        cand = counts.candidates(threshold)
    if cand is not None:
        table, capacity, summary = cand
        n_max = summary[1]
    else:
        table, capacity = counts.table, counts.capacity
        n_max = counts.summary(threshold)[1] if sketch is None else counts.summary(0)[0]
    keys = torch.empty((max(n_max, 1), counts.key_words), dtype=torch.int64, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    sk = C.byref(sketch) if sketch is not None else None
    with _timed("select_solid", counts.n_occ):
        gn.check(L.ga_select_solid(gn.ptr(table), capacity, counts.key_words, counts.k,
                                   counts.alphabet.sym_bits, int(threshold), sk, gn.ptr(counts.alphabet.inv_dev),
                                   gn.ptr(keys), None, gn.ptr(n_out), _stream()))
    if sketch is None:
        return keys, n_max           # exact filter: the summary already counted them (no sync)
    return keys, int(n_out.item())


def _solid_keys_from_flags(counts: KmerCounts, keep_fn):
    """Compat path for a foreign ``kmer_counts`` mapping: ask it once per distinct window."""
    keys, _ = counts.export()
    strings = counts.alphabet.decode_strings(keys, counts.w)
    mask = np.fromiter((bool(keep_fn(s)) for s in strings), dtype=bool, count=len(strings))
    sel = np.ascontiguousarray(keys[mask])
    if sel.shape[0] == 0:
        return torch.zeros((1, counts.key_words), dtype=torch.int64, device=_dev()), 0
    return torch.from_numpy(sel.view(np.int64)).to(_dev()), sel.shape[0]


def _to_host(tensor: torch.Tensor, n: int) -> torch.Tensor:
    """First n rows of a device tensor -> pinned host tensor (async; the caller synchronises)."""
    out = torch.empty((n,) + tuple(tensor.shape[1:]), dtype=tensor.dtype, pin_memory=True)
    if n:
        out.copy_(tensor[:n], non_blocking=True)
    return out


def build_graph(counts: KmerCounts, reads: DeviceReads, threshold: int, sketch=None, keep_fn=None,
                to_host: bool = True, feed=None):
    """count table (+ optional sketch) + reads -> BuiltGraph (or device tensors)."""
    L = gn.lib()
    dev = _dev()
    k, w, kw = counts.k, counts.w, counts.key_words
    alphabet = reads.alphabet
    graph = BuiltGraph(reads.paired, w, alphabet, kw)
    bucketed = (keep_fn is None and sketch is None and counts._table is None and not counts._cand and
                counts.n_occ >= SUPERKMER_MIN_OCC and superkmer_supported(reads, k, threshold))
    edge_stamp = None
    counted = (not bucketed and keep_fn is None and sketch is None and counts._table is None and not counts._cand and
               reads.paired and counts.n_occ >= max(SUPERKMER_MIN_OCC, SUPERKMER_MIN_OCC_PAIRS) and
               superkmer_supported(reads, k, threshold, counting_only=True))
    solid_keys = None
    if bucketed or counted:
        try:
            if bucketed:
                solid_keys, n_solid, edge_stamp = superkmer_stamps(reads, k, threshold, feed)
            else:
                solid_keys, n_solid = superkmer_solid(reads, k, threshold)   # read pairs: the solid set from the buckets
        except gn.GaBucketLimit:
            # one window repeated tens of millions of times (adapter or poly-A reads): its bucket is beyond what a
            # shared-memory pass can name.  The global-table kernels have no such limit; every read is resident by
            # now (the scatter drained `feed`), and the buckets' workspace makes room for the tables
            release_workspace()
            bucketed = counted = False
    if solid_keys is None:
        if keep_fn is not None:
            solid_keys, n_solid = _solid_keys_from_flags(counts, keep_fn)
        else:
            solid_keys, n_solid = _solid_keys(counts, threshold, sketch)
    _mark("select solid")
    if n_solid == 0 or reads.n_reads == 0:
        return graph
    solid_cap = int(1.7 * n_solid) + 64
    solid = torch.empty(solid_cap * counts.slot_bytes, dtype=torch.uint8, device=dev)
    status = reads.status
    status.zero_()
    with _timed("id_table"):
        gn.check(L.ga_table_clear(gn.ptr(solid), solid_cap, kw, _stream()))
        gn.check(L.ga_table_insert_ids(gn.ptr(solid_keys), n_solid, kw, 0, gn.ptr(solid), solid_cap,
                                       gn.ptr(status), _stream()))
    _mark("id table")
    n_nodes, n_edges, attr = C.c_int64(), C.c_int64(), C.c_int64()
    plan = C.c_void_p()
    dna4 = (not reads.paired) and alphabet.sym_bits <= 2
    if dna4:
        # <= 4 symbols: per-node edge stamps, epoch-tagged slots (ga_build_unpaired_dna)
        node_stamp = torch.full((n_solid,), -1, dtype=torch.int64, device=dev)
        if bucketed:
            with _timed("sk_resolve"):
                gn.check(L.ga_sk_resolve(gn.ptr(solid_keys), n_solid, k, gn.ptr(solid), solid_cap, gn.ptr(edge_stamp),
                                         gn.ptr(node_stamp), _stream()))
        else:
            edge_stamp = torch.full((4 * n_solid,), -1, dtype=torch.int64, device=dev)
            build_dna4(reads, k, solid, solid_cap, solid_keys, n_solid, kw, node_stamp, edge_stamp, status)
        _mark("fill stamps + build_dna")
        with _timed("csr_plan"):
            gn.check(L.ga_csr_plan_unpaired_dna(gn.ptr(node_stamp), gn.ptr(edge_stamp), n_solid, gn.ptr(solid_keys),
                                                kw, k, alphabet.sym_bits, gn.ptr(solid), solid_cap, _stream(),
                                                C.byref(plan), C.byref(n_nodes), C.byref(n_edges)))
        attr.value = n_edges.value
        _mark("csr plan dna")
        if _check_status(status) & gn.ST_TABLE_FULL:
            raise gn.GaError("id table overflow")
    free = _free_bytes()
    cap_limit = min(0xFFFFFFF0, max(4096, int(free * 0.4) // 32))
    cap = min(cap_limit, max(1024, 3 * n_solid))
    while not dna4:
        status.zero_()
        if not reads.paired:
            node_stamp = torch.full((n_solid,), -1, dtype=torch.int64, device=dev)
            edges = torch.full((cap, 2), -1, dtype=torch.int64, device=dev)
            with _timed("build"):
                gn.check(L.ga_build_unpaired(C.byref(reads.struct()), k, gn.ptr(solid), solid_cap,
                                             gn.ptr(node_stamp), gn.ptr(edges), cap, gn.ptr(status), _stream()))
        else:
            queries = torch.full((cap, 2), -1, dtype=torch.int64, device=dev)
            qedges = torch.full((cap, 2), -1, dtype=torch.int64, device=dev)
            dh = torch.full((256 * 256 * 2,), -1, dtype=torch.int64, device=dev)
            with _timed("build"):
                gn.check(L.ga_build_paired(C.byref(reads.struct()), k, gn.ptr(solid), solid_cap, gn.ptr(queries),
                                           cap, gn.ptr(qedges), cap, gn.ptr(dh), gn.ptr(status), _stream()))
        st = _check_status(status)
        if not st & gn.ST_STAMP_FULL:
            break
        if cap >= cap_limit:
            raise MemoryError("edge tables do not fit in device memory")
        cap = min(cap * 2, cap_limit)
    if dna4:
        pass
    elif not reads.paired:
        _mark("build")
        with _timed("csr_plan"):
            gn.check(L.ga_csr_plan_unpaired(gn.ptr(node_stamp), n_solid, gn.ptr(solid_keys), kw, alphabet.sym_bits,
                                            gn.ptr(edges), cap, _stream(), C.byref(plan), C.byref(n_nodes),
                                            C.byref(n_edges)))
        attr.value = n_edges.value
        _mark("csr plan")
    else:
        _mark("build")
        with _timed("csr_plan"):
            gn.check(L.ga_csr_plan_paired(gn.ptr(solid), solid_cap, gn.ptr(solid_keys), n_solid, kw, k,
                                          alphabet.sym_bits, gn.ptr(queries), cap, gn.ptr(qedges), cap, gn.ptr(dh),
                                          _stream(), C.byref(plan), C.byref(n_nodes), C.byref(n_edges),
                                          C.byref(attr)))
        _mark("csr plan")
    try:
        nn, ne = n_nodes.value, n_edges.value
        _t0 = _time.perf_counter()
        rowptr = torch.empty(nn + 1, dtype=torch.int32, device=dev)
        col = torch.empty(max(ne, 1), dtype=torch.int32, device=dev)
        indeg = torch.empty(max(nn, 1), dtype=torch.int32, device=dev)
        branching = torch.empty(max(nn, 1), dtype=torch.uint8, device=dev)
        last_sym = torch.empty(max(nn, 1), dtype=torch.uint8, device=dev)
        keys_a = torch.empty((max(nn, 1), kw), dtype=torch.int64, device=dev)
        keys_b = torch.empty((max(nn, 1), kw), dtype=torch.int64, device=dev) if reads.paired else None
        _t1 = _time.perf_counter()
        with _timed("csr_emit"):
            gn.check(L.ga_csr_emit(plan, gn.ptr(rowptr), gn.ptr(col), gn.ptr(indeg), gn.ptr(branching),
                                   gn.ptr(last_sym), gn.ptr(keys_a), gn.ptr(keys_b), _stream()))
        _t2 = _time.perf_counter()
        torch.cuda.current_stream().synchronize()
        _t3 = _time.perf_counter()
        if TIMERS is not None:
            TIMERS.setdefault("_host", []).append(("emit: alloc %.1f call %.1f sync %.1f ms" %
                                                   ((_t1 - _t0) * 1e3, (_t2 - _t1) * 1e3, (_t3 - _t2) * 1e3)))
        _mark("csr emit")
    finally:
        L.ga_csr_plan_free(plan)
    graph.n_nodes, graph.n_edges, graph.num_edges_attr = nn, ne, attr.value
    if not to_host:
        graph.device = dict(rowptr=rowptr, col=col, indeg=indeg, branching=branching, last_sym=last_sym,
                            keys_a=keys_a, keys_b=keys_b)
        return graph
    # D2H through pinned buffers; the code -> byte mapping of the last symbol happens on the device
    last_char = alphabet.inv_dev[last_sym[:nn].long()] if nn else last_sym[:0]
    host = [_to_host(rowptr, nn + 1), _to_host(col, ne), _to_host(indeg, nn), _to_host(branching, nn),
            _to_host(last_char, nn), _to_host(keys_a, nn), _to_host(keys_b, nn) if reads.paired else None]
    torch.cuda.current_stream().synchronize()
    graph._pinned = host                       # the numpy views below alias these buffers
    graph.rowptr, graph.col, graph.indeg, graph.branching, graph.last_char = (t.numpy() for t in host[:5])
    graph.keys_a = host[5].numpy().view(np.uint64)
    graph.keys_b = host[6].numpy().view(np.uint64) if reads.paired else None
    return graph


def build_dna4(reads, k, solid, solid_cap, solid_keys, n_solid, kw, node_stamp, edge_stamp, status):
    """Stamps of an unpaired read set over <= 4 symbols (ga_build_unpaired_dna).  When the id table is
    far bigger than the L2, only a prefix of the reads goes through it: what is still unstamped
    afterwards is small, and the rest of the reads is checked against a Bloom filter of it
    (ga_build_unpaired_dna_tail).  Same stamps either way."""
    L = gn.lib()
    dev = node_stamp.device
    n = reads.n_reads
    per_read = max(reads.max_len - (k - 1) + 1, 0) if reads.lens is None or reads.uniform else 0

    def run(r0, r1):
        with _timed("build", (r1 - r0) * per_read):
            gn.check(L.ga_build_unpaired_dna(C.byref(reads.struct_range(r0, r1)), k, gn.ptr(solid), solid_cap,
                                             gn.ptr(node_stamp), gn.ptr(edge_stamp), gn.ptr(status), _stream()))

    big = n_solid * 27 > TWO_PHASE_MIN_TABLE_BYTES and n >= TWO_PHASE_MIN_READS
    if not big:
        run(0, n)
        return
    done, step = 0, max(TWO_PHASE_MIN_STEP, n // 20)
    while True:
        hi = min(n, done + step)
        run(done, hi)
        done = hi
        if done >= n:
            return
        mask = torch.empty(n_solid, dtype=torch.uint8, device=dev)
        n_open = torch.zeros(1, dtype=torch.int64, device=dev)
        gn.check(L.ga_unstamped_scan(gn.ptr(solid), solid_cap, gn.ptr(solid_keys), n_solid, kw, gn.ptr(edge_stamp), k,
                                     reads.alphabet.sym_bits, gn.ptr(mask), gn.ptr(n_open), _stream()))
        opened = int(n_open.item())
        if opened <= max(TWO_PHASE_MIN_STEP, n_solid // 4) and opened <= (24 << 20):
            break
        step *= 2
    open_cap = 2 * opened + 64
    open_table = torch.empty(open_cap * L.ga_slot_bytes(kw), dtype=torch.uint8, device=dev)
    bloom_words = max(1024, opened // 2)
    bloom = torch.zeros(bloom_words, dtype=torch.int32, device=dev)
    gn.check(L.ga_table_clear(gn.ptr(open_table), open_cap, kw, _stream()))
    gn.check(L.ga_unstamped_table_build(gn.ptr(solid_keys), gn.ptr(mask), n_solid, kw, gn.ptr(open_table), open_cap,
                                        gn.ptr(bloom), bloom_words, gn.ptr(status), _stream()))
    with _timed("build_tail", (n - done) * per_read):
        gn.check(L.ga_build_unpaired_dna_tail(C.byref(reads.struct_range(done, n)), k, gn.ptr(bloom), bloom_words,
                                              gn.ptr(open_table), open_cap, gn.ptr(solid), solid_cap,
                                              gn.ptr(node_stamp), gn.ptr(edge_stamp), _stream()))


def emit_dna4(graph, node_stamp, edge_stamp, n_solid, solid_keys, solid, solid_cap, kw, k, alphabet, to_host):
    """CSR from per-node edge stamps (unpaired, <= 4 symbols); used by the multi-GPU path on rank 0."""
    L = gn.lib()
    dev = node_stamp.device
    n_nodes, n_edges, plan = C.c_int64(), C.c_int64(), C.c_void_p()
    gn.check(L.ga_csr_plan_unpaired_dna(gn.ptr(node_stamp), gn.ptr(edge_stamp), n_solid, gn.ptr(solid_keys), kw, k,
                                        alphabet.sym_bits, gn.ptr(solid), solid_cap, _stream(), C.byref(plan),
                                        C.byref(n_nodes), C.byref(n_edges)))
    try:
        nn, ne = n_nodes.value, n_edges.value
        rowptr = torch.empty(nn + 1, dtype=torch.int32, device=dev)
        col = torch.empty(max(ne, 1), dtype=torch.int32, device=dev)
        indeg = torch.empty(max(nn, 1), dtype=torch.int32, device=dev)
        branching = torch.empty(max(nn, 1), dtype=torch.uint8, device=dev)
        last_sym = torch.empty(max(nn, 1), dtype=torch.uint8, device=dev)
        keys_a = torch.empty((max(nn, 1), kw), dtype=torch.int64, device=dev)
        gn.check(L.ga_csr_emit(plan, gn.ptr(rowptr), gn.ptr(col), gn.ptr(indeg), gn.ptr(branching),
                               gn.ptr(last_sym), gn.ptr(keys_a), None, _stream()))
        torch.cuda.current_stream().synchronize()
    finally:
        L.ga_csr_plan_free(plan)
    graph.n_nodes, graph.n_edges, graph.num_edges_attr = nn, ne, ne
    if not to_host:
        graph.device = dict(rowptr=rowptr, col=col, indeg=indeg, branching=branching, last_sym=last_sym,
                            keys_a=keys_a, keys_b=None)
        return graph
    last_char = alphabet.inv_dev[last_sym[:nn].long()] if nn else last_sym[:0]
    host = [_to_host(rowptr, nn + 1), _to_host(col, ne), _to_host(indeg, nn), _to_host(branching, nn),
            _to_host(last_char, nn), _to_host(keys_a, nn)]
    torch.cuda.current_stream().synchronize()
    graph._pinned = host
    graph.rowptr, graph.col, graph.indeg, graph.branching, graph.last_char = (t.numpy() for t in host[:5])
    graph.keys_a = host[5].numpy().view(np.uint64)
    return graph


def paired_tables(reads, k, solid, solid_cap, n_solid, status):
    """This rank's query / query-edge / homopolymer tables of a paired read set (ga_build_paired), grown until
    they fit: (queries, qedges, capacity, dh)."""
    L = gn.lib()
    dev = solid.device
    free = _free_bytes()
    cap_limit = min(0xFFFFFFF0, max(4096, int(free * 0.4) // 32))
    cap = min(cap_limit, max(1024, 3 * n_solid))
    while True:
        status.zero_()
        queries = torch.full((cap, 2), -1, dtype=torch.int64, device=dev)
        qedges = torch.full((cap, 2), -1, dtype=torch.int64, device=dev)
        dh = torch.full((256 * 256 * 2,), -1, dtype=torch.int64, device=dev)
        with _timed("build"):
            gn.check(L.ga_build_paired(C.byref(reads.struct()), k, gn.ptr(solid), solid_cap, gn.ptr(queries),
                                       cap, gn.ptr(qedges), cap, gn.ptr(dh), gn.ptr(status), _stream()))
        if not _check_status(status) & gn.ST_STAMP_FULL:
            return queries, qedges, cap, dh
        if cap >= cap_limit:
            raise MemoryError("edge tables do not fit in device memory")
        del queries, qedges
        cap = min(cap * 2, cap_limit)


def emit_paired(graph, solid, solid_cap, solid_keys, n_solid, kw, k, alphabet, queries, qedges, cap, dh, to_host):
    """CSR of a paired graph from (merged) query tables; used by the multi-GPU path on rank 0."""
    L = gn.lib()
    dev = solid.device
    n_nodes, n_edges, attr, plan = C.c_int64(), C.c_int64(), C.c_int64(), C.c_void_p()
    gn.check(L.ga_csr_plan_paired(gn.ptr(solid), solid_cap, gn.ptr(solid_keys), n_solid, kw, k, alphabet.sym_bits,
                                  gn.ptr(queries), cap, gn.ptr(qedges), cap, gn.ptr(dh), _stream(), C.byref(plan),
                                  C.byref(n_nodes), C.byref(n_edges), C.byref(attr)))
    try:
        nn, ne = n_nodes.value, n_edges.value
        rowptr = torch.empty(nn + 1, dtype=torch.int32, device=dev)
        col = torch.empty(max(ne, 1), dtype=torch.int32, device=dev)
        indeg = torch.empty(max(nn, 1), dtype=torch.int32, device=dev)
        branching = torch.empty(max(nn, 1), dtype=torch.uint8, device=dev)
        last_sym = torch.empty(max(nn, 1), dtype=torch.uint8, device=dev)
        keys_a = torch.empty((max(nn, 1), kw), dtype=torch.int64, device=dev)
        keys_b = torch.empty((max(nn, 1), kw), dtype=torch.int64, device=dev)
        gn.check(L.ga_csr_emit(plan, gn.ptr(rowptr), gn.ptr(col), gn.ptr(indeg), gn.ptr(branching),
                               gn.ptr(last_sym), gn.ptr(keys_a), gn.ptr(keys_b), _stream()))
        torch.cuda.current_stream().synchronize()
    finally:
        L.ga_csr_plan_free(plan)
    graph.n_nodes, graph.n_edges, graph.num_edges_attr = nn, ne, attr.value
    if not to_host:
        graph.device = dict(rowptr=rowptr, col=col, indeg=indeg, branching=branching, last_sym=last_sym,
                            keys_a=keys_a, keys_b=keys_b)
        return graph
    last_char = alphabet.inv_dev[last_sym[:nn].long()] if nn else last_sym[:0]
    host = [_to_host(rowptr, nn + 1), _to_host(col, ne), _to_host(indeg, nn), _to_host(branching, nn),
            _to_host(last_char, nn), _to_host(keys_a, nn), _to_host(keys_b, nn)]
    torch.cuda.current_stream().synchronize()
    graph._pinned = host
    graph.rowptr, graph.col, graph.indeg, graph.branching, graph.last_char = (t.numpy() for t in host[:5])
    graph.keys_a = host[5].numpy().view(np.uint64)
    graph.keys_b = host[6].numpy().view(np.uint64)
    return graph


# ----------------------------------------------------------------------------------- whole path
def _poured_sketch(counts: KmerCounts, rows: int, widths=None):
    """The -c route (debruijn_graph.py:181-188, debug_graph.py:66-85): every distinct window's exact count
    poured into a `rows`-row CountMinSketch (uint16 overflow checked as the reference's array('H') would)."""
    from countminsketch import CountMinSketch
    sketch = CountMinSketch(rows, widths=widths)      # widths: one of the reference's prime tables (countminsketch.py:9-24)
    sketch.pour_counts(counts)
    return sketch


def device_step(reads: DeviceReads, k: int, threshold: int, timers=None, sketch_rows: int = 0, sketch_widths=None):
    """One pass of the hot path over device-resident packed reads; the CSR stays on the device.
    sketch_rows > 0: the CountMinSketch route (exact table -> sketch -> filter on the estimate -> build)."""
    global TIMERS
    TIMERS = timers
    try:
        _mark("step begin")
        counts = KmerCounts(k, reads)
        sketch = _poured_sketch(counts, sketch_rows, sketch_widths) if sketch_rows else None
        out = build_graph(counts, reads, threshold, to_host=False,
                          sketch=sketch._struct() if sketch is not None else None)
        _mark("step end")
        return out
    finally:
        TIMERS = None


_STREAM_CHUNK_BYTES = 256 << 20
_copy_stream = {}


def _stream_in(ascii_pinned: torch.Tensor, reads: DeviceReads, read_len: int):
    """Generator: copies the reads host -> device chunk by chunk on a side stream (two staging buffers),
    packs each chunk on the current stream, and yields its read range -- so the consumer's kernels on
    chunk i overlap the copy of chunk i+1."""
    L = gn.lib()
    dev = _dev()
    n = reads.n_reads
    per_chunk = max(1, _STREAM_CHUNK_BYTES // max(read_len, 1))
    side = _copy_stream.setdefault(dev.index, torch.cuda.Stream(device=dev))
    main = torch.cuda.current_stream()
    stage = [torch.empty(min(n, per_chunk) * read_len, dtype=torch.uint8, device=dev) for _ in range(2)]
    free = [None, None]
    lut = reads.alphabet.lut_dev
    side.wait_stream(main)
    for i, r0 in enumerate(range(0, n, per_chunk)):
        r1 = min(n, r0 + per_chunk)
        buf = stage[i & 1][:(r1 - r0) * read_len]
        if free[i & 1] is not None:
            side.wait_event(free[i & 1])
        with torch.cuda.stream(side):
            buf.copy_(ascii_pinned[r0 * read_len:r1 * read_len], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(side)
        main.wait_event(ready)
        gn.check(L.ga_pack_reads(gn.ptr(buf), None, r1 - r0, read_len, gn.ptr(lut), reads.alphabet.storage_bits,
                                 reads.words.data_ptr() + r0 * reads.stride_words * 8, None, reads.stride_words,
                                 gn.ptr(reads.status), _stream()))
        free[i & 1] = torch.cuda.Event()
        free[i & 1].record(main)
        yield r0, r1
    for t in stage:
        t.record_stream(main)


def _stream_in_packed(words_pinned: torch.Tensor, reads: DeviceReads):
    """Like _stream_in for reads that were already packed 2 bits per base on the host (the layout of
    ga_reads): chunks are copied straight into place on a side stream; each chunk's read range is yielded
    once the current stream has been ordered after its copy."""
    dev = _dev()
    n, stride = reads.n_reads, reads.stride_words
    per_chunk = max(1, _STREAM_CHUNK_BYTES // (stride * 8))
    side = _copy_stream.setdefault(dev.index, torch.cuda.Stream(device=dev))
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    for r0 in range(0, n, per_chunk):
        r1 = min(n, r0 + per_chunk)
        with torch.cuda.stream(side):
            reads.words[r0 * stride:r1 * stride].copy_(words_pinned[r0 * stride:r1 * stride], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(side)
        main.wait_event(ready)
        yield r0, r1


def host_step_packed(words_pinned: torch.Tensor, n_reads: int, read_len: int, k: int, threshold: int):
    """host_step for unpaired DNA reads the caller already holds packed (2 bits per base, 64-bit words,
    ga_reads layout) in pinned host memory -- the ingest format BASELINE.json's north star describes: a
    quarter of the bytes cross PCIe."""
    alphabet = Alphabet(np.zeros(0))
    stride = max(1, -(-int(read_len) // 32))
    words = torch.empty(max(1, n_reads * stride), dtype=torch.int64, device=_dev())
    reads = DeviceReads.from_packed(words, n_reads, read_len, False, estride=read_len, alphabet=alphabet)
    counts = KmerCounts(k, reads)
    if counts.n_occ >= SUPERKMER_MIN_OCC and superkmer_supported(reads, k, threshold):
        return build_graph(counts, reads, threshold, to_host=True, feed=_stream_in_packed(words_pinned, reads))
    for _ in _stream_in_packed(words_pinned, reads):
        pass
    return build_graph(counts, reads, threshold, to_host=True)


def host_step(ascii_pinned: torch.Tensor, n_reads: int, read_len: int, paired: bool, k: int, threshold: int,
              sketch_rows: int = 0):
    """The same from host memory: ASCII reads (pinned) -> H2D -> pack -> count -> filter -> build ->
    CSR arrays on the host.  This is the call bench.py times as `e2e`.  Unpaired DNA streams in by
    chunks so that the copy overlaps the scatter kernel; other inputs copy first."""
    alphabet = Alphabet(np.zeros(0))
    spw = 64 // alphabet.storage_bits
    stride = max(1, -(-int(read_len) // spw))
    probe = DeviceReads.from_packed(torch.empty(0, dtype=torch.int64, device=_dev()), n_reads, read_len, paired,
                                    estride=read_len, alphabet=alphabet)
    n_occ = probe.windows_total(k)
    if not paired and not sketch_rows and n_occ >= SUPERKMER_MIN_OCC and superkmer_supported(probe, k, threshold):
        words = torch.empty(max(1, n_reads * stride), dtype=torch.int64, device=_dev())
        reads = DeviceReads.from_packed(words, n_reads, read_len, False, estride=read_len, alphabet=alphabet)
        counts = KmerCounts(k, reads)
        return build_graph(counts, reads, threshold, to_host=True, feed=_stream_in(ascii_pinned, reads, read_len))
    reads = DeviceReads.from_ascii(ascii_pinned, n_reads, read_len, paired, estride=read_len)
    counts = KmerCounts(k, reads)
    if _check_status(reads.status) & gn.ST_BAD_SYMBOL:
        raise ValueError("read symbol outside the alphabet")
    sketch = _poured_sketch(counts, sketch_rows) if sketch_rows else None
    return build_graph(counts, reads, threshold, to_host=True, sketch=sketch._struct() if sketch is not None else None)
