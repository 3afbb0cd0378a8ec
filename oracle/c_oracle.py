"""ctypes front end of oracle/c_oracle.c (TEST INFRASTRUCTURE ONLY -- see that file)."""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libga_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
        L.orc_assemble.restype = vp
        L.orc_assemble.argtypes = [vp, vp, vp, vp, u64, i32, i32, i32, i32]
        for name in ("orc_n_distinct", "orc_n_nodes", "orc_num_edges", "orc_n_csr_edges",
                     "orc_n_contigs", "orc_contig_bytes"):
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [vp]
        L.orc_error.restype = i32
        L.orc_error.argtypes = [vp]
        L.orc_copy_counts.argtypes = [vp, vp, vp]
        L.orc_copy_csr.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_copy_node_keys.argtypes = [vp, vp, vp]
        L.orc_copy_sketch_row.argtypes = [vp, i32, vp]
        L.orc_copy_contigs.argtypes = [vp, vp, vp]
        L.orc_free.argtypes = [vp]
        L.orc_murmur3_32.restype = C.c_uint32
        L.orc_murmur3_32.argtypes = [C.c_char_p, u64, C.c_uint32]
        _lib = L
    return _lib


def murmur3_32(text: str) -> int:
    raw = bytes(ord(c) & 0xFF for c in text)
    return lib().orc_murmur3_32(raw, len(raw), 0)


def _flatten(strings):
    lens = np.fromiter((len(s) for s in strings), dtype=np.uint64, count=len(strings))
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    buf = np.frombuffer("".join(strings).encode("latin-1"), dtype=np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(buf), off


PRIMES_1_10_7 = (9999889, 9999901, 9999907, 9999929, 9999931, 9999937, 9999943, 9999971, 9999973,
                 9999991, 10000019, 10000079, 10000103, 10000121, 10000139, 10000141, 10000169,
                 10000189, 10000223, 10000229)


class Result:
    """Owns one orc_result; arrays are copied out lazily as numpy."""

    def __init__(self, handle, k, paired, rows, keep):
        self._h, self.k, self.w, self.paired, self.rows = handle, k, k - 1, paired, rows
        self._keep = keep          # input buffers the C side points into
        L = lib()
        err = L.orc_error(handle)
        if err == 1:
            self.close()
            raise OverflowError("unsigned short is greater than maximum")
        if err:
            self.close()
            raise MemoryError("oracle allocation failed")
        self.n_distinct = L.orc_n_distinct(handle)
        self.n_nodes = L.orc_n_nodes(handle)
        self.num_edges = L.orc_num_edges(handle)
        self.n_csr_edges = L.orc_n_csr_edges(handle)
        self.n_contigs = L.orc_n_contigs(handle)

    def close(self):
        if self._h:
            lib().orc_free(self._h)
            self._h = None

    __del__ = close

    def counts(self):
        keys = np.zeros((self.n_distinct, self.w), dtype=np.uint8)
        cnt = np.zeros(self.n_distinct, dtype=np.uint32)
        lib().orc_copy_counts(self._h, keys.ctypes.data, cnt.ctypes.data)
        return keys, cnt

    def counts_dict(self):
        keys, cnt = self.counts()
        return {row.tobytes().decode("latin-1"): int(c) for row, c in zip(keys, cnt)}

    def csr(self):
        n, m = self.n_nodes, self.n_csr_edges
        rowptr = np.zeros(n + 1, dtype=np.int64)
        col = np.zeros(max(m, 1), dtype=np.int32)
        indeg = np.zeros(max(n, 1), dtype=np.int32)
        br = np.zeros(max(n, 1), dtype=np.uint8)
        last = np.zeros(max(n, 1), dtype=np.uint8)
        lib().orc_copy_csr(self._h, rowptr.ctypes.data, col.ctypes.data, indeg.ctypes.data,
                           br.ctypes.data, last.ctypes.data)
        return rowptr, col[:m], indeg[:n], br[:n], last[:n]

    def node_keys(self):
        a = np.zeros((max(self.n_nodes, 1), self.w), dtype=np.uint8)
        b = np.zeros((max(self.n_nodes, 1), self.w), dtype=np.uint8) if self.paired else None
        lib().orc_copy_node_keys(self._h, a.ctypes.data, b.ctypes.data if b is not None else None)
        return a[:self.n_nodes], (b[:self.n_nodes] if b is not None else None)

    def sketch_row(self, row):
        out = np.zeros(PRIMES_1_10_7[row], dtype=np.uint16)
        lib().orc_copy_sketch_row(self._h, row, out.ctypes.data)
        return out

    def contigs(self):
        L = lib()
        text = np.zeros(max(L.orc_contig_bytes(self._h), 1), dtype=np.uint8)
        off = np.zeros(self.n_contigs + 1, dtype=np.uint64)
        L.orc_copy_contigs(self._h, text.ctypes.data, off.ctypes.data)
        raw = text.tobytes()
        return [raw[int(off[i]):int(off[i + 1])].decode("latin-1") for i in range(self.n_contigs)]

    def digest(self) -> str:
        """Graph digest in the format of SURVEY App. B.3 / py_oracle.Graph.digest."""
        rowptr, col, indeg, br, _ = self.csr()
        a, b = self.node_keys()
        sa = [row.tobytes().decode("latin-1") for row in a]
        keys = list(zip(sa, (row.tobytes().decode("latin-1") for row in b))) if self.paired else sa
        h = hashlib.sha256()
        for i, key in enumerate(keys):
            edges = [keys[j] for j in col[rowptr[i]:rowptr[i + 1]]]
            h.update(repr((key, edges, int(indeg[i]), bool(br[i]))).encode())
        return h.hexdigest()[:16]


def assemble(reads, k: int, threshold: int, paired: bool, sketch_rows: int = 0,
             want_graph: bool = True) -> Result:
    """count -> [sketch] -> build -> contigs on the C oracle."""
    if paired:
        m1, o1 = _flatten([p[0] for p in reads])
        if any(len(p[1]) < len(p[0]) for p in reads):
            raise ValueError("oracle: mate 2 shorter than mate 1 is not supported")
        m2, o2 = _flatten([p[1] for p in reads])
        h = lib().orc_assemble(m1.ctypes.data, o1.ctypes.data, m2.ctypes.data, o2.ctypes.data,
                               len(reads), k, threshold, sketch_rows, int(want_graph))
        keep = (m1, o1, m2, o2)
    else:
        m1, o1 = _flatten(list(reads))
        h = lib().orc_assemble(m1.ctypes.data, o1.ctypes.data, None, None, len(reads), k, threshold,
                               sketch_rows, int(want_graph))
        keep = (m1, o1)
    return Result(h, k, paired, sketch_rows, keep)


def assemble_codes(codes: np.ndarray, k: int, threshold: int, alphabet: bytes = b"ACGT") -> Result:
    """Unpaired, uniform-length reads given as an (n, L) uint8 code matrix."""
    asc = np.ascontiguousarray(np.frombuffer(alphabet, dtype=np.uint8)[codes])
    n, L = asc.shape
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    h = lib().orc_assemble(asc.ctypes.data, off.ctypes.data, None, None, n, k, threshold, 0, 1)
    return Result(h, k, False, 0, (asc, off))
