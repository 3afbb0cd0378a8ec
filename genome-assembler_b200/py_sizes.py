"""What ``sys.getsizeof`` says about the reference's containers, for the ``-m`` report (assemble.py:107-114,
debug_graph.py:51-63 of the reference print ``sys.getsizeof`` of the read list and of the k-mer count dict).

Here the reads are one byte buffer (``ga_ingest.RawReads``) and the counts a device table (``ga_device.KmerCounts``);
their ``__sizeof__`` answers with the size the reference's object would have in THIS interpreter: a list grown by
``n`` appends, a ``defaultdict(int)`` grown by ``n`` insertions of string keys.  Both follow from CPython's growth
rules (listobject.c ``list_resize``, dictobject.c ``GROWTH_RATE`` / ``USABLE_FRACTION``); the constants are taken
from live objects at import, and tests/test_host_side.py compares the answers with real containers.
"""
import struct
from collections import defaultdict

import sys

_POINTER = struct.calcsize("P")
_LIST_BASE = [].__sizeof__()


class _Plain:
    """An instance of an ordinary Python class, as RawReads and KmerCounts are."""


# sys.getsizeof(o) = o.__sizeof__() + what the interpreter keeps in front of the object (GC header; for instances of
# Python classes also the managed dict / weakref words).  A stand-in that wants sys.getsizeof to print the container's
# number must take the difference of the two headers off.
_HEADER_SHIFT = (sys.getsizeof(_Plain()) - object.__sizeof__(_Plain())) - (sys.getsizeof([]) - [].__sizeof__())


def for_instance(container_sizeof: int) -> int:
    """The value an ordinary class's __sizeof__ must return so that sys.getsizeof(instance) equals
    sys.getsizeof(container) for a list / dict whose __sizeof__() is `container_sizeof`."""
    return container_sizeof - _HEADER_SHIFT


def appended_list_sizeof(n: int) -> int:
    """``l.__sizeof__()`` after ``n`` calls of ``l.append`` on an empty list."""
    allocated, size = 0, 1
    while size <= n:
        allocated = (size + (size >> 3) + 6) & ~3           # list_resize: over-allocation on append
        size = allocated + 1
    return _LIST_BASE + _POINTER * allocated


def _usable(log2: int) -> int:
    return (2 << log2) // 3                                 # USABLE_FRACTION


def _one_key():
    d = defaultdict(int)
    d["k"] += 1
    return d


_DICT_EMPTY = defaultdict(int).__sizeof__()
_DICT_ENTRY = 2 * _POINTER                                  # unicode-keys table: (key, value) per entry
_DICT_BASE = _one_key().__sizeof__() - 8 - _usable(3) * _DICT_ENTRY      # object + keys header, without the table


def grown_str_dict_sizeof(n: int) -> int:
    """``d.__sizeof__()`` of a ``defaultdict(int)`` after ``n`` distinct string keys were inserted one by one."""
    if n <= 0:
        return _DICT_EMPTY
    log2, used, free = 3, 0, _usable(3)
    while n - used > free:                                  # the table is full: resize to hold used * 3
        used += free
        log2 = 3
        while (1 << log2) < used * 3:
            log2 += 1
        free = _usable(log2) - used
    index = 1 if log2 < 8 else 2 if log2 < 16 else 4 if log2 < 32 else 8
    return _DICT_BASE + (1 << log2) * index + _usable(log2) * _DICT_ENTRY
