"""Multi-GPU build: one process per GPU, reads sharded by index, k-mers owned by hash.

Per pass (every exchange is a torch.distributed collective over NCCL / NVLink):
  1. each rank bumps a pre-filter sketch with its own reads; the clamped sketches are
     all-reduced (sum), so every rank knows which cells can hold a solid k-mer *globally*;
  2. each rank counts its own occurrences of those candidates exactly (partial counts);
  3. hash-partition all-to-all: every candidate (key, partial count) goes to the rank that owns
     the key; owners add the partial counts up and apply the strict `> threshold` filter;
  4. the solid keys are all-gathered; every rank builds the same id table (id = position in the
     gathered list), so stamp arrays are indexed identically everywhere;
  5. each rank folds the stamps of its own reads (global read indices) into node / edge stamp
     arrays; an all-reduce(min) makes them global;
  6. rank 0 emits the CSR.  Stamps depend only on global read indices, so the graph is
     bit-identical for 1, 2, 4 or 8 GPUs.
The helpers that only move tensors (`exchange_by_owner`, `all_gather_var`) are device-agnostic
and are covered on the CPU with the gloo backend (tests/test_multi_host.py).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist


def all_gather_var(local: torch.Tensor, group=None):
    """Concatenation of every rank's `local` (rows may differ per rank), plus per-rank row counts."""
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes) if sizes else 0
    padded = torch.zeros((max(width, 1),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0), sizes


def exchange_by_owner(owner: torch.Tensor, payloads, group=None):
    """All-to-all of rows: row i of every tensor in `payloads` goes to rank owner[i].
    Returns the received tensors (rows from rank 0 first, then rank 1, ...)."""
    world = dist.get_world_size(group)
    order = torch.argsort(owner.to(torch.int64), stable=True)
    send_counts = torch.bincount(owner.to(torch.int64), minlength=world)[:world]
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    send_list, recv_list = [int(x) for x in send_counts.tolist()], [int(x) for x in recv_counts.tolist()]
    received = []
    for t in payloads:
        src = t[order].contiguous()
        dst = torch.empty((sum(recv_list),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_to_all_single(dst, src, output_split_sizes=recv_list, input_split_sizes=send_list, group=group)
        received.append(dst)
    return received


_SIGN = -(1 << 63)


def all_reduce_min_u64(stamps: torch.Tensor, group=None):
    """In-place minimum of unsigned 64-bit stamps held in an int64 tensor (0xFF..FF = none):
    flipping the top bit makes signed order equal unsigned order."""
    stamps.bitwise_xor_(_SIGN)
    dist.all_reduce(stamps, op=dist.ReduceOp.MIN, group=group)
    stamps.bitwise_xor_(_SIGN)
    return stamps


def sharded_step(reads, k: int, threshold: int, timers=None, to_host: bool = False):
    """One pass of the hot path over this rank's read shard; returns the graph on rank 0 (a
    BuiltGraph whose CSR stays on the device unless to_host) and None elsewhere."""
    import ga_native as gn
    import ga_device as gd
    if reads.paired or reads.alphabet.sym_bits > 2:
        raise NotImplementedError("multi-GPU build currently covers unpaired reads over <= 4 symbols")
    if threshold < 0 or (threshold + 1) * dist.get_world_size() > 255:
        raise NotImplementedError("multi-GPU pre-filter needs 0 <= threshold and (threshold+1)*ranks <= 255")
    gd.TIMERS = timers
    try:
        return _sharded_step(reads, k, threshold, to_host, gn, gd)
    finally:
        gd.TIMERS = None


def _sharded_step(reads, k, threshold, to_host, gn, gd):
    L = gn.lib()
    dev = reads.words.device
    world, rank = dist.get_world_size(), dist.get_rank()
    stream = gd._stream
    kw = reads.key_words(k)
    slot_bytes = L.ga_slot_bytes(kw)
    status = reads.status
    status.zero_()

    # 1. pre-filter: local sketch (8-bit cells), clamp, all-reduce(sum)
    occ = torch.tensor([reads.windows_total(k)], dtype=torch.int64, device=dev)
    dist.all_reduce(occ)
    n_cells = max(1 << 16, int(occ.item()))
    n_cells = (n_cells + 3) // 4 * 4
    cells = torch.zeros(n_cells, dtype=torch.uint8, device=dev)
    pf = gn.GaPrefilter()
    pf.words, pf.n_cells, pf.cell_bits = gn.ptr(cells), n_cells, 8
    n_occ_local = reads.windows_total(k)
    with gd._timed("prefilter", n_occ_local):
        gn.check(L.ga_prefilter_update(C.byref(reads.struct()), k, C.byref(pf), threshold, stream()))
    cells.clamp_(max=threshold + 1)
    dist.all_reduce(cells)

    # 2. exact partial counts of the candidates among this rank's reads
    n_hot = torch.zeros(1, dtype=torch.int64, device=dev)
    gn.check(L.ga_prefilter_hot(C.byref(pf), threshold, gn.ptr(n_hot), stream()))
    cap = max(1024, int(int(n_hot.item()) * 2.2) + 1024)
    while True:
        table = torch.empty(cap * slot_bytes, dtype=torch.uint8, device=dev)
        status.zero_()
        gn.check(L.ga_table_clear(gn.ptr(table), cap, kw, stream()))
        with gd._timed("count", n_occ_local):
            gn.check(L.ga_count_candidates(C.byref(reads.struct()), k, C.byref(pf), threshold, gn.ptr(table), cap,
                                           gn.ptr(status), stream()))
        out4 = torch.zeros(4, dtype=torch.int64, device=dev)
        gn.check(L.ga_table_summary(gn.ptr(table), cap, kw, -1, gn.ptr(out4), stream()))
        if not gd._check_status(status) & gn.ST_TABLE_FULL:
            break
        cap *= 2
    n_cand = int(out4[0].item())
    keys = torch.empty((max(n_cand, 1), kw), dtype=torch.int64, device=dev)
    counts = torch.empty(max(n_cand, 1), dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    gn.check(L.ga_table_export(gn.ptr(table), cap, kw, -1, gn.ptr(keys), gn.ptr(counts), gn.ptr(n_out), stream()))
    keys, counts = keys[:n_cand], counts[:n_cand]
    del table

    # 3. hash-partition all-to-all of (key, partial count); owners sum and filter
    owner = torch.empty(max(n_cand, 1), dtype=torch.int32, device=dev)
    gn.check(L.ga_key_owner(gn.ptr(keys), n_cand, kw, world, gn.ptr(owner), stream()))
    got_keys, got_counts = exchange_by_owner(owner[:n_cand], [keys, counts])
    n_got = got_keys.shape[0]
    mcap = max(1024, 2 * n_got + 64)
    merged = torch.empty(mcap * slot_bytes, dtype=torch.uint8, device=dev)
    gn.check(L.ga_table_clear(gn.ptr(merged), mcap, kw, stream()))
    got_keys = got_keys.contiguous()
    got_counts = got_counts.contiguous()
    gn.check(L.ga_count_keys(gn.ptr(got_keys), gn.ptr(got_counts), n_got, kw, gn.ptr(merged), mcap, gn.ptr(status),
                             stream()))
    out4 = torch.zeros(4, dtype=torch.int64, device=dev)
    gn.check(L.ga_table_summary(gn.ptr(merged), mcap, kw, int(threshold), gn.ptr(out4), stream()))
    n_mine = int(out4[1].item())
    mine = torch.empty((max(n_mine, 1), kw), dtype=torch.int64, device=dev)
    n_out.zero_()
    gn.check(L.ga_select_solid(gn.ptr(merged), mcap, kw, k, reads.alphabet.sym_bits, int(threshold), None, None,
                               gn.ptr(mine), None, gn.ptr(n_out), stream()))
    # 4. replicated id table: id = position in the rank-ordered concatenation
    solid_keys, _ = all_gather_var(mine[:n_mine])
    solid_keys = solid_keys.contiguous()
    n_solid = solid_keys.shape[0]
    graph = gd.BuiltGraph(False, k - 1, reads.alphabet, kw)
    if n_solid == 0:
        return graph if rank == 0 else None
    solid_cap = int(1.7 * n_solid) + 64
    solid = torch.empty(solid_cap * slot_bytes, dtype=torch.uint8, device=dev)
    gn.check(L.ga_table_clear(gn.ptr(solid), solid_cap, kw, stream()))
    gn.check(L.ga_table_insert_ids(gn.ptr(solid_keys), n_solid, kw, 0, gn.ptr(solid), solid_cap, gn.ptr(status),
                                   stream()))
    # 5. stamps of the local shard, then the global minimum
    stamps = torch.full((5 * n_solid,), -1, dtype=torch.int64, device=dev)
    node_stamp, edge_stamp = stamps[:n_solid], stamps[n_solid:]
    gd.build_dna4(reads, k, solid, solid_cap, solid_keys, n_solid, kw, node_stamp, edge_stamp, status)
    all_reduce_min_u64(stamps)
    if gd._check_status(status) & (gn.ST_TABLE_FULL | gn.ST_BAD_SYMBOL):
        raise gn.GaError("table overflow or bad symbol in the sharded build")
    if rank != 0:
        return None
    # 6. CSR on rank 0
    return gd.emit_dna4(graph, node_stamp, edge_stamp, n_solid, solid_keys, solid, solid_cap, kw, k, reads.alphabet,
                        to_host)
