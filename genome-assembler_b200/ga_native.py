"""ctypes binding of libga_b200.so (include/ga_b200.h) -- the only door from the Python
host code to the sm_100a kernels.

There is deliberately no CPU implementation behind this module: if the shared library is
missing, or no CUDA device is visible, every compute entry raises.  PyTorch is used for
device buffers and streams only (tensors are handed over as raw pointers).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GA_LIB") or os.path.join(HERE, "libga_b200.so")   # GA_LIB: A/B builds of the library

GA_OK = 0
GA_ERR_BAD_ARG, GA_ERR_CAPACITY, GA_ERR_CUDA, GA_ERR_NCCL, GA_ERR_OVERFLOW_U16, GA_ERR_ALPHABET = \
    -1, -2, -3, -4, -5, -6
ST_TABLE_FULL, ST_BAD_SYMBOL, ST_STAMP_FULL, ST_U16_OVERFLOW = 1, 2, 4, 8
MAX_SKETCH_ROWS = 20


class GaError(RuntimeError):
    """A libga_b200 call failed (message from ga_last_error())."""


class GaBucketLimit(GaError):
    """The read set is outside what the bucketed kernels can name (a bucket of 2^25 records or more: one window
    repeated tens of millions of times, e.g. adapter or poly-A reads).  ga_device.build_graph then takes the
    global-table kernels, which have no such limit."""


class GaReads(C.Structure):
    _fields_ = [("words", C.c_void_p), ("offsets", C.c_void_p), ("lengths", C.c_void_p),
                ("n_reads", C.c_uint64), ("first_read", C.c_uint64),
                ("uniform_len", C.c_uint32), ("stride_words", C.c_uint32),
                ("storage_bits", C.c_int32), ("sym_bits", C.c_int32), ("paired", C.c_int32),
                ("estride", C.c_uint32)]


class GaPrefilter(C.Structure):
    _fields_ = [("words", C.c_void_p), ("n_cells", C.c_uint64), ("cell_bits", C.c_int32)]


class GaSkSources(C.Structure):
    _fields_ = [("records", C.c_void_p * 16), ("index", C.c_void_p * 16), ("l1_capacity", C.c_uint64 * 16),
                ("first_bucket", C.c_uint64), ("solid_counter", C.c_void_p), ("n_sources", C.c_uint32)]


class GaSketch(C.Structure):
    _fields_ = [("cells", C.c_void_p), ("width", C.c_uint32 * MAX_SKETCH_ROWS), ("rows", C.c_int32)]


_vp, _u64, _u32, _i64, _i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_int
_PR, _PS, _PF = C.POINTER(GaReads), C.POINTER(GaSketch), C.POINTER(GaPrefilter)

# name -> (restype, argtypes); mirrors include/ga_b200.h one to one
SIGNATURES = {
    "ga_version": (_i32, []),
    "ga_last_error": (C.c_char_p, []),
    "ga_device_count": (_i32, []),
    "ga_launch_count": (_u64, []),
    "ga_set_l2_fetch_granularity": (_i32, [_i32]),
    "ga_fill_bytes": (_i32, [_vp, _i32, _u64, _vp]),
    "ga_copy_bytes": (_i32, [_vp, _vp, _u64, _vp]),
    "ga_key_words": (_i32, [_i32, _i32]),
    "ga_slot_bytes": (_i32, [_i32]),
    "ga_parse_reads": (_i32, [_vp, _u64, _vp, _vp, _u64, C.POINTER(C.c_uint64), C.POINTER(C.c_int),
                              C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]),
    "ga_pack_reads": (_i32, [_vp, _vp, _u64, _u32, _vp, _i32, _vp, _vp, _u32, _vp, _vp]),
    "ga_unpack_reads": (_i32, [_vp, _u64, _u32, _u32, _i32, _vp, _vp, _vp]),
    "ga_gen_genome": (_i32, [_vp, _u64, _u64, _vp]),
    "ga_gen_reads": (_i32, [_vp, _u64, _u64, _u64, _u32, _u64, _u32, _vp, _u32, _i32, _u32, _vp]),
    "ga_table_clear": (_i32, [_vp, _u64, _i32, _vp]),
    "ga_count_kmers": (_i32, [_PR, _i32, _vp, _u64, _vp, _vp]),
    "ga_count_keys": (_i32, [_vp, _vp, _u64, _i32, _vp, _u64, _vp, _vp]),
    "ga_key_owner": (_i32, [_vp, _u64, _i32, _u32, _vp, _vp]),
    "ga_table_summary": (_i32, [_vp, _u64, _i32, _i64, _vp, _vp]),
    "ga_table_export": (_i32, [_vp, _u64, _i32, _i64, _vp, _vp, _vp, _vp]),
    "ga_table_lookup": (_i32, [_vp, _u64, _i32, _vp, _u64, _vp, _vp]),
    "ga_table_insert_ids": (_i32, [_vp, _u64, _i32, _u32, _vp, _u64, _vp, _vp]),
    "ga_prefilter_update": (_i32, [_PR, _i32, _PF, _i64, _vp]),
    "ga_prefilter_hot": (_i32, [_PF, _i64, _vp, _vp]),
    "ga_count_candidates": (_i32, [_PR, _i32, _PF, _i64, _vp, _u64, _vp, _vp]),
    "ga_sk_minimizer_len": (_i32, [_i32]),
    "ga_sk_cursor_stride": (_i32, []),
    "ga_sk_scatter_reads": (_i32, [_PR, _i32, _i32, _i32, _vp, _u64, _vp, _vp, _vp, _vp]),
    "ga_sk_offsets": (_i32, [_vp, _u64, _vp, _vp, _vp]),
    "ga_sk_scatter_buckets": (_i32, [_vp, _u64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ga_sk_count_build": (_i32, [_vp, _vp, _vp, _u32, _vp, _u64, _i32, _i64, _u32, _u32, _vp, _vp, _u64, _vp, _vp, _u64, _vp,
                                 _vp, _u64, _i32, _vp]),
    "ga_sk_count_build_from": (_i32, [C.POINTER(GaSkSources), _vp, _vp, _u64, _i32, _i64, _u32, _u32, _vp, _vp, _u64, _vp, _vp,
                                      _u64, _vp, _i32, _vp]),
    "ga_sk_count_build_spill_from": (_i32, [C.POINTER(GaSkSources), _vp, _u64, _vp, _u64, _i32, _i64, _u32, _vp, _u32, _vp, _vp,
                                            _u64, _vp, _vp, _i32, _vp]),
    "ga_sk_spill_scratch_bytes": (_u64, [_u32]),
    "ga_sk_count_build_spill": (_i32, [_vp, _vp, _vp, _u32, _u64, _vp, _u64, _i32, _i64, _u32, _vp, _u32, _vp, _vp, _u64, _vp, _vp,
                                       _vp, _u64, _i32, _vp]),
    "ga_sk_resolve": (_i32, [_vp, _u64, _i32, _vp, _u64, _vp, _vp, _vp]),
    "ga_peer_alloc": (_i32, [_u64, C.POINTER(_vp), _vp]),
    "ga_peer_open": (_i32, [_vp, C.POINTER(_vp)]),
    "ga_peer_close": (_i32, [_vp]),
    "ga_peer_free": (_i32, [_vp]),
    "ga_sk_push_records": (_i32, [_vp, _u64, _vp, _vp, _i32, _i32, _u32, _vp, _vp, _vp, _vp]),
    "ga_sk_push_sorted": (_i32, [_vp, _u64, _vp, _i32, _i32, _vp, _u32, _vp, _vp, _vp, _vp]),
    "ga_sketch_update_table": (_i32, [_vp, _u64, _i32, _i32, _i32, _vp, _PS, _vp]),
    "ga_sketch_update_bytes": (_i32, [_vp, _vp, _vp, _u64, _PS, _vp]),
    "ga_sketch_estimate_bytes": (_i32, [_vp, _vp, _u64, _PS, _vp, _vp]),
    "ga_sketch_narrow": (_i32, [_PS, _vp, _vp, _vp]),
    "ga_select_solid": (_i32, [_vp, _u64, _i32, _i32, _i32, _i64, _PS, _vp, _vp, _vp, _vp, _vp]),
    "ga_build_unpaired": (_i32, [_PR, _i32, _vp, _u64, _vp, _vp, _u64, _vp, _vp]),
    "ga_build_unpaired_dna": (_i32, [_PR, _i32, _vp, _u64, _vp, _vp, _vp, _vp]),
    "ga_unstamped_scan": (_i32, [_vp, _u64, _vp, _u64, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "ga_unstamped_table_build": (_i32, [_vp, _vp, _u64, _i32, _vp, _u64, _vp, _u64, _vp, _vp]),
    "ga_build_unpaired_dna_tail": (_i32, [_PR, _i32, _vp, _u64, _vp, _u64, _vp, _u64, _vp, _vp, _vp]),
    "ga_build_paired": (_i32, [_PR, _i32, _vp, _u64, _vp, _u64, _vp, _u64, _vp, _vp, _vp]),
    "ga_stamp_table_export": (_i32, [_vp, _u64, _vp, _vp, _vp, _vp, _u64, _vp, _vp]),
    "ga_paired_merge": (_i32, [_vp, _vp, _u64, _vp, _vp, _vp, _u64, _vp, _u64, _vp, _u64, _vp, _vp]),
    "ga_csr_plan_unpaired": (_i32, [_vp, _u64, _vp, _i32, _i32, _vp, _u64, _vp,
                                    C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64)]),
    "ga_csr_plan_unpaired_dna": (_i32, [_vp, _vp, _u64, _vp, _i32, _i32, _i32, _vp, _u64, _vp,
                                        C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64)]),
    "ga_csr_plan_paired": (_i32, [_vp, _u64, _vp, _u64, _i32, _i32, _i32, _vp, _u64, _vp, _u64, _vp, _vp,
                                  C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "ga_csr_emit": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ga_csr_plan_free": (None, [_vp]),
    "ga_traverse_contigs": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32,
                                   C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_u64), _vp]),
    "ga_free_host": (None, [_vp]),
    "ga_traverse_last_route": (_i32, []),
}

_lib = None


def lib():
    """The loaded shared library (raises ImportError if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libga_b200.so is missing: build it with `python __graft_entry__.py` or "
                "`make -C genome-assembler_b200` (there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                if os.environ.get("GA_LIB"):       # an older A/B build of the library: entries added since are absent
                    continue
                raise
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().ga_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != GA_OK:
        raise GaError("libga_b200 error %d: %s" % (rc, last_error()))


_tuned = False


def require_gpu() -> None:
    global _tuned
    if lib().ga_device_count() < 1:
        raise GaError("no CUDA device visible: the k-mer counting / graph build path runs on "
                      "the GPU only (no CPU fallback)")
    if not _tuned:
        _tuned = True
        gran = int(os.environ.get("GA_L2_FETCH", "0"))   # measured on B200: no effect (profiles/r01)
        if gran in (32, 64, 128):
            lib().ga_set_l2_fetch_granularity(gran)


def ptr(tensor) -> int:
    """Raw device (or host) pointer of a torch tensor, None -> NULL."""
    return None if tensor is None else tensor.data_ptr()
