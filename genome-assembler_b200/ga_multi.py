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
The helpers that only move tensors (`exchange_by_owner`, `all_gather_var`, `exchange_ranges`,
`gather_rows`) are device-agnostic and are covered on the CPU with the gloo backend
(tests/test_multi_host.py).

Unpaired DNA with k <= 32 takes the bucketed route instead (`_sharded_step_buckets`, csrc/ga_superkmer.cu):
  1. each rank cuts its read shard into super-k-mer records sorted by bucket (the bucket of a window
     depends on its content only, so all its occurrences -- on any rank -- share one bucket id);
  2. hash-partition all-to-all: rank g owns a contiguous range of bucket ids and receives every rank's
     records of that range.  EXCHANGE = "auto" (default) takes "push" on 2-3 ranks and "peer" from 4 on.
     "push": the owners' receive buffers are mapped into every rank over NVLink (CUDA IPC, `PeerBuffers`) and
     ONE kernel per rank (csrc/ga_peer.cu) stores the records straight into the owners' buffers -- no dense
     local copy, no NCCL call on the data path (NCCL carries the 8 MB of histograms and the barriers).
     PUSH = "gather" (default): index pass, then `ga_sk_push_records` gathers through the index; PUSH = "sorted":
     `ga_sk_push_sorted` splits the level-1 buckets and sends in one pass, no index (measured slower).
     "peer": nothing is copied -- the owner's bucket kernel gathers the records from every rank's own level-1
     slots over NVLink while it counts (`_count_in_place`, ga_sk_count_build_from).
     "nccl": the round-1 route, a dense local copy + `all_to_all_single` per record array (PHASES = 2 would
     send a second half while the first is counted -- no gain measured); also what every rank falls back to
     when peer memory cannot be mapped (`peers_available`);
  3. the owner counts and stamps each of its buckets in shared memory, reading the bucket as one
     segment per source rank;
  4. solid keys + candidate edge stamps are gathered on rank 0, which resolves them into the CSR.
Ordinals are global (read index * stride + position), so the graph is bit-identical for any rank count.
"""
from __future__ import annotations

import ctypes as C
import os

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


def exchange_ranges(src: torch.Tensor, send_rows, group=None, ws_name=None, recv_rows=None):
    """All-to-all of contiguous row ranges: the first send_rows[0] rows of `src` go to rank 0, the next
    send_rows[1] to rank 1, ...  Returns (received rows, rows received from each rank).  ws_name: take the
    receive buffer from ga_device's persistent workspace instead of the allocator.  recv_rows: the rows
    every rank will send here when the caller already knows them (skips the count exchange and its
    host synchronisation)."""
    world = dist.get_world_size(group)
    if recv_rows is None:
        send = torch.tensor(list(send_rows), dtype=torch.int64, device=src.device)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        recv_rows = [int(x) for x in recv.tolist()]
    shape = (sum(recv_rows),) + tuple(src.shape[1:])
    if ws_name is not None:
        import ga_device as gd
        dst = gd.workspace(ws_name, shape, src.dtype)
    else:
        dst = torch.empty(shape, dtype=src.dtype, device=src.device)
    dist.all_to_all_single(dst, src.contiguous(), output_split_sizes=recv_rows,
                           input_split_sizes=[int(x) for x in send_rows], group=group)
    assert len(recv_rows) == world
    return dst, recv_rows


def gather_rows(local: torch.Tensor, dst: int = 0, group=None, recv_rows=None, want_rows: bool = False):
    """Rows of every rank concatenated on rank `dst` (rank order); other ranks get an empty tensor.
    recv_rows / want_rows: reuse the per-rank row counts of a previous gather of equally long tensors."""
    world = dist.get_world_size(group)
    send_rows = [local.shape[0] if g == dst else 0 for g in range(world)]
    out, rows = exchange_ranges(local, send_rows, group, recv_rows=recv_rows)
    return (out, rows) if want_rows else out


_SIGN = -(1 << 63)


def all_reduce_min_u64(stamps: torch.Tensor, group=None):
    """In-place minimum of unsigned 64-bit stamps held in an int64 tensor (0xFF..FF = none):
    flipping the top bit makes signed order equal unsigned order."""
    stamps.bitwise_xor_(_SIGN)
    dist.all_reduce(stamps, op=dist.ReduceOp.MIN, group=group)
    stamps.bitwise_xor_(_SIGN)
    return stamps


def push_plan(matrix, rank: int):
    """matrix[s][g] = records source rank s sends to owner g (every rank holds the same matrix).  Returns
    (dst_start, seg_start, recv_total): dst_start[g] = row of rank g's receive arrays where THIS rank's
    segment begins (sources are laid out in rank order), seg_start[s] = row of this rank's receive arrays
    where source s's segment begins, recv_total = rows this rank receives."""
    world = len(matrix)
    dst_start = [sum(matrix[s][g] for s in range(rank)) for g in range(world)]
    seg_start = [sum(matrix[t][rank] for t in range(s)) for s in range(world)]
    return dst_start, seg_start, sum(matrix[s][rank] for s in range(world))


class _RawBuffer:
    """A device range outside torch's allocator, as far as ga_device needs one (data_ptr + device)."""

    def __init__(self, address: int, device):
        self._address, self.device = int(address), device

    def data_ptr(self) -> int:
        return self._address


class PeerRegion:
    """A device buffer of this rank that every other rank of the node has mapped (ga_peer_alloc / ga_peer_open:
    cudaMalloc + CUDA IPC, peer access over NVLink).  `ensure(nbytes)` is collective: every rank passes the same
    size, so all ranks decide alike whether to reallocate; reallocation is rare (the regions only grow)."""

    def __init__(self):
        self.nbytes = 0
        self.local = None            # this rank's buffer (raw address)
        self.mapped = []             # per rank: address of that rank's buffer in this process

    def close(self):
        import ga_native as gn
        L = gn.lib()
        rank = dist.get_rank()
        torch.cuda.synchronize()
        dist.barrier()                                   # nobody still uses a buffer about to go
        for g, address in enumerate(self.mapped):
            if g != rank and address:
                gn.check(L.ga_peer_close(C.c_void_p(address)))
        torch.cuda.synchronize()
        dist.barrier()                                   # every mapping is gone before the owners free
        if self.local:
            gn.check(L.ga_peer_free(C.c_void_p(self.local)))
        self.nbytes, self.local, self.mapped = 0, None, []

    def ensure(self, nbytes: int):
        """Collective.  A failure on any rank (allocation, CUDA IPC) raises GaError on EVERY rank, after all ranks
        have passed the same collectives -- nobody is left waiting in one."""
        import ga_native as gn
        if self.local and nbytes <= self.nbytes:
            return
        L = gn.lib()
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = torch.device("cuda", torch.cuda.current_device())
        if self.local:
            self.close()
        nbytes = (nbytes + nbytes // 16 + 65536) // 256 * 256
        handle = (C.c_uint8 * 64)()
        address = C.c_void_p()
        problem = ""
        if L.ga_peer_alloc(nbytes, C.byref(address), handle) != gn.GA_OK:
            torch.cuda.empty_cache()                     # the caching allocator may sit on the room
            if L.ga_peer_alloc(nbytes, C.byref(address), handle) != gn.GA_OK:
                problem = gn.last_error()
        mine = torch.tensor(list(handle) + [0 if problem else 1], dtype=torch.uint8, device=dev)
        everyone = torch.empty(world * 65, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(everyone, mine)
        everyone = everyone.cpu().numpy().reshape(world, 65)
        mapped = []
        if everyone[:, 64].all():
            for g in range(world):
                if g == rank:
                    mapped.append(address.value)
                    continue
                peer = C.c_void_p()
                raw = (C.c_uint8 * 64)(*[int(x) for x in everyone[g, :64]])
                if L.ga_peer_open(raw, C.byref(peer)) != gn.GA_OK:
                    problem = problem or gn.last_error()
                    mapped.append(0)
                else:
                    mapped.append(peer.value)
        else:
            problem = problem or "rank %d could not allocate" % int((everyone[:, 64] == 0).argmax())
        fine = torch.tensor([0 if problem else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(fine, op=dist.ReduceOp.MIN)
        if not int(fine.item()):
            for g, at in enumerate(mapped):              # undo what this rank did, then fail on every rank together
                if g != rank and at:
                    L.ga_peer_close(C.c_void_p(at))
            torch.cuda.synchronize()
            dist.barrier()
            if address.value:
                L.ga_peer_free(address)
            raise gn.GaError("peer memory: %s" % (problem or "another rank failed"))
        self.mapped, self.local, self.nbytes = mapped, address.value, nbytes


class PeerBuffers(PeerRegion):
    """The receive arrays of the record exchange (bases 16 B + meta 8 B per row) in one region."""

    def __init__(self):
        super().__init__()
        self.rows = 0

    def ensure(self, rows: int):
        if self.local and rows <= self.rows:
            return
        rows = rows + rows // 16 + 1024
        super().ensure(rows * 24)
        self.rows = rows

    def close(self):
        super().close()
        self.rows = 0

    def bases_at(self, g: int, row: int) -> int:
        return self.mapped[g] + 16 * row

    def meta_at(self, g: int, row: int) -> int:
        return self.mapped[g] + 16 * self.rows + 8 * row


_PEERS = {}
_PEER_OK = {}


def peers_available() -> bool:
    """Collective, once per process and device: can every rank map every other rank's memory (CUDA IPC between the
    processes of this node)?  If any rank cannot, all ranks take the NCCL route -- the exchange is then two
    all_to_all_single calls instead of one kernel over peer memory, the result is the same."""
    dev = torch.cuda.current_device()
    if dev not in _PEER_OK:
        import sys
        ok, probe, why = 1, PeerRegion(), ""
        try:
            probe.ensure(4096)
        except Exception as exc:        # noqa: BLE001 -- reported below, every rank must reach the all-reduce
            ok, why = 0, str(exc)
        flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", dev))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        _PEER_OK[dev] = bool(int(flag.item()))
        if ok:
            try:
                probe.close()
            except Exception:           # noqa: BLE001
                pass
        if not _PEER_OK[dev] and dist.get_rank() == 0:
            print("ga_multi: peer memory over CUDA IPC is not available (%s); records travel by NCCL all-to-all"
                  % (why or "another rank failed"), file=sys.stderr)
    return _PEER_OK[dev]


def release_peers():
    """Unmap and free the exchange buffers (collective)."""
    for peers in _PEERS.values():
        peers.close()
    _PEERS.clear()


def sharded_step(reads, k: int, threshold: int, timers=None, to_host: bool = False, feed=None):
    """One pass of the hot path over this rank's read shard; returns the graph on rank 0 (a
    BuiltGraph whose CSR stays on the device unless to_host) and None elsewhere."""
    import ga_native as gn
    import ga_device as gd
    if reads.alphabet.sym_bits > 2 and not reads.paired:
        raise NotImplementedError("multi-GPU build of unpaired reads covers alphabets of <= 4 symbols")
    gd.TIMERS = timers
    try:
        gd._mark("step begin")
        if not reads.paired and gd.superkmer_supported(reads, k, threshold) and USE_BUCKETS:
            return _sharded_step_buckets(reads, k, threshold, to_host, gn, gd, feed)
        # the table route all-reduces 8-bit pre-filter cells that each rank clamps at threshold + 1
        if threshold < 0 or (threshold + 1) * dist.get_world_size() > 255:
            raise NotImplementedError("multi-GPU pre-filter needs 0 <= threshold and (threshold+1)*ranks <= 255")
        if feed is not None:
            for _ in feed:      # the table route walks resident reads: drain the stream first
                pass
        return _sharded_step(reads, k, threshold, to_host, gn, gd)
    finally:
        gd.TIMERS = None


def raise_together(error, group=None):
    """Collective error check between phases: if any rank holds an exception, every rank raises (the failing
    rank its own, the others a RuntimeError naming it) instead of hanging in the next collective."""
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    flag = torch.tensor([0 if error is None else dist.get_rank(group) + 1], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if error is not None:
        raise error
    if int(flag.item()):
        raise RuntimeError("sharded build: rank %d failed (see its own traceback)" % (int(flag.item()) - 1))


USE_BUCKETS = True      # scripts/multi_check.py also runs the table route by clearing this
EXCHANGE = os.environ.get("GA_MULTI_EXCHANGE", "auto")     # "peer": the owners gather over NVLink while they count (no
                                                            # copy); "push": one send kernel over NVLink peer memory;
                                                            # "nccl": round 1 (dense copy + all_to_all_single);
                                                            # "auto": push on 2-3 ranks, peer from 4 on (C4: 157.0 vs
                                                            # 160.4 ms at N = 2, 67.6 vs 63.6 ms at N = 8)
PUSH = os.environ.get("GA_MULTI_PUSH", "gather")            # "gather": index pass, then ga_sk_push_records;
                                                            # "sorted": level-2 split + send in one pass (ga_sk_push_sorted)
PHASES = 1              # 2 cuts the exchange in two halves so that the second overlaps the counting of the first;
                        # measured on 2 and 8 B200s it gains nothing (77.2 vs 75.4 ms at N=8: the second bucket
                        # pass and its host round trip cost what the overlap saves), so one phase is the default


def _exchange_nccl(reads, k, l1_bits, l2_bits, n_buckets, feed, gn, gd):
    """Round-1 exchange: dense local copy of the records in bucket order, then all_to_all_single per record
    array.  Returns [(bases, meta, per-source hist [world, mine], seg_start, mine)] per phase."""
    dev = reads.words.device
    world, rank = dist.get_world_size(), dist.get_rank()
    bases, meta, offsets, hist, total = gd.sk_scatter_local(reads, k, l1_bits, l2_bits, feed)
    # ownership: the bucket ids are cut into PHASES halves, each half is split over the ranks, so that
    # the exchange of the second half runs on the NCCL stream while the first half is being counted
    phases = PHASES if n_buckets >= PHASES * world else 1
    per_phase = n_buckets // phases
    bounds = [[h * per_phase + g * per_phase // world for g in range(world + 1)] for h in range(phases)]
    flat = [b for row in bounds for b in row]
    cut = offsets[torch.tensor(flat, dtype=torch.int64, device=dev)].tolist()
    cut = [cut[h * (world + 1):(h + 1) * (world + 1)] for h in range(phases)]
    send_rows = [[int(cut[h][g + 1] - cut[h][g]) for g in range(world)] for h in range(phases)]
    mine = [bounds[h][rank + 1] - bounds[h][rank] for h in range(phases)]
    gd._mark("multi: cut")
    # the row counts of every phase in one small all-to-all (one host synchronisation)
    send_cnt = torch.tensor([[send_rows[h][g] for h in range(phases)] for g in range(world)], dtype=torch.int64,
                            device=dev)
    recv_cnt = torch.empty_like(send_cnt)
    dist.all_to_all_single(recv_cnt, send_cnt)
    recv_cnt = recv_cnt.tolist()
    recv_rows = [[int(recv_cnt[s][h]) for s in range(world)] for h in range(phases)]
    bases2 = bases.view(-1, 2)
    received = []
    with gd._timed("exchange_issue"):
        for h in range(phases):
            lo, hi = int(cut[h][0]), int(cut[h][world])
            got_bases = gd.workspace("sk_recv_bases%d" % h, (sum(recv_rows[h]), 2), torch.int64)
            got_meta = gd.workspace("sk_recv_meta%d" % h, (sum(recv_rows[h]),), torch.int64)
            got_hist = torch.empty(world * mine[h], dtype=torch.int64, device=dev)
            works = [dist.all_to_all_single(got_bases, bases2[lo:hi], output_split_sizes=recv_rows[h],
                                            input_split_sizes=send_rows[h], async_op=True),
                     dist.all_to_all_single(got_meta, meta[lo:hi], output_split_sizes=recv_rows[h],
                                            input_split_sizes=send_rows[h], async_op=True),
                     dist.all_to_all_single(got_hist, hist[bounds[h][0]:bounds[h][world]],
                                            output_split_sizes=[mine[h]] * world,
                                            input_split_sizes=[bounds[h][g + 1] - bounds[h][g] for g in range(world)],
                                            async_op=True)]
            starts = [0]
            for rows in recv_rows[h][:-1]:
                starts.append(starts[-1] + rows)
            received.append((works, got_bases.view(-1), got_meta, got_hist, starts, mine[h]))
    return received


def _exchange_push(reads, k, l1_bits, l2_bits, n_buckets, feed, gn, gd):
    """Sort + send in one kernel over NVLink peer memory (csrc/ga_peer.cu).  Same return shape as
    _exchange_nccl, one phase."""
    L = gn.lib()
    dev = reads.words.device
    world, rank = dist.get_world_size(), dist.get_rank()
    fused = PUSH == "sorted"
    if fused:
        rec, _, offsets, hist, total, _, cap1, cursors1, cursors2 = gd.sk_scatter_local(reads, k, l1_bits, l2_bits, feed,
                                                                                        dense=False, level2=False)
    else:
        rec, _, offsets, hist, total, index, cap1 = gd.sk_scatter_local(reads, k, l1_bits, l2_bits, feed, dense=False)
    bounds = [g * n_buckets // world for g in range(world + 1)]
    mine = bounds[rank + 1] - bounds[rank]
    # cut[g] = first bucket-sorted position of owner g's range, of every source: one small all-gather.  It doubles
    # as the barrier that keeps this step's stores out of buffers a peer's previous bucket pass still reads
    # (a rank enters it, in stream order, after its own previous step)
    my_cut = offsets[torch.tensor(bounds, dtype=torch.int64, device=dev)].contiguous()
    all_cut = torch.empty(world * (world + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_cut, my_cut)
    all_cut = all_cut.view(world, world + 1).tolist()
    matrix = [[int(all_cut[s][g + 1] - all_cut[s][g]) for g in range(world)] for s in range(world)]
    dst_start, seg_start, _ = push_plan(matrix, rank)
    gd._mark("multi: cut")
    peers = _PEERS.setdefault((dev.index, "received"), PeerBuffers())
    peers.ensure(max(sum(matrix[s][g] for s in range(world)) for g in range(world)))
    got_hist = torch.empty(world * mine, dtype=torch.int64, device=dev)
    work = dist.all_to_all_single(got_hist, hist, output_split_sizes=[mine] * world,
                                  input_split_sizes=[bounds[g + 1] - bounds[g] for g in range(world)], async_op=True)
    cut = (C.c_uint64 * (world + 1))(*[int(x) for x in all_cut[rank]])
    dst_bases = (C.c_void_p * world)(*[peers.bases_at(g, dst_start[g]) for g in range(world)])
    dst_meta = (C.c_void_p * world)(*[peers.meta_at(g, dst_start[g]) for g in range(world)])
    if os.environ.get("GA_PUSH_SKIP"):
        os.environ["GA_PUSH_SELF"] = str(rank)          # timing probe of the push kernel (results are then wrong)
    with gd._timed("sk_push", reads.windows_total(k)):
        if fused:
            gn.check(L.ga_sk_push_sorted(gn.ptr(rec), cap1, gn.ptr(cursors1), l1_bits, l2_bits, gn.ptr(cursors2), world,
                                         cut, dst_bases, dst_meta, gd._stream()))
        else:
            gn.check(L.ga_sk_push_records(gn.ptr(rec), cap1, gn.ptr(index), gn.ptr(offsets), l1_bits, l2_bits, world,
                                          cut, dst_bases, dst_meta, gd._stream()))
    # every rank's stores have landed before anybody counts: a one-word all-reduce, stream-ordered after the push
    done = torch.zeros(1, dtype=torch.int32, device=dev)
    barrier = dist.all_reduce(done, async_op=True)
    return [([work, barrier], _RawBuffer(peers.bases_at(rank, 0), dev), _RawBuffer(peers.meta_at(rank, 0), dev),
             got_hist, seg_start, mine)]


_GEOMETRY = {}          # (device, l1_bits) -> slots per level-1 bucket in force (the same on every rank, only grows)


_OUT_ROWS = {}          # device -> rows of the shared result buffers in force (the same on every rank, only grows)


def _count_in_place(reads, k, threshold, l1_bits, l2_bits, n_buckets, slots_wanted, n_occ, feed, gn, gd):
    """EXCHANGE = "peer": the exchange fused into the count.  Nothing is copied: every rank leaves its records in
    its own level-1 slots (peer-mapped regions), sorts a 32-bit index, and the owner of a bucket range gathers the
    records of its buckets from all ranks -- its own memory or NVLink -- inside the bucket kernel
    (ga_sk_count_build_from).  The results go where the graph is built: every rank appends its solid windows and
    candidate edge stamps to rank 0's buffers (mapped everywhere; one shared counter), so no gather follows.
    Returns (solid keys, n_solid, candidate edge stamps) of ALL ranks on rank 0, (None, 0, None) elsewhere."""
    dev = reads.words.device
    world, rank = dist.get_world_size(), dist.get_rank()
    n_l1 = 1 << l1_bits
    key = (dev.index, l1_bits)
    cap1 = max(_GEOMETRY.get(key, 0), slots_wanted)
    slots_region = _PEERS.setdefault((dev.index, "slots"), PeerRegion())
    index_region = _PEERS.setdefault((dev.index, "index"), PeerRegion())
    bounds = [g * n_buckets // world for g in range(world + 1)]
    mine = bounds[rank + 1] - bounds[rank]
    while True:
        _GEOMETRY[key] = cap1
        slots_region.ensure(n_l1 * cap1 * 32)
        index_region.ensure(n_l1 * cap1 * 4)
        regions = {"sk_l1_records": slots_region, "sk_index": index_region}
        needed = 0
        try:
            rec, _, offsets, hist, total, index, _ = gd.sk_scatter_local(
                reads, k, l1_bits, l2_bits, feed, dense=False, cap1=cap1,
                alloc=lambda name, n, dtype: _RawBuffer(regions[name].local, dev))
        except gd.ScatterOverflow as exc:
            needed = exc.needed
        feed = None
        # every rank must have fitted its records, or all of them cut again with larger buckets
        worst = torch.tensor([needed], dtype=torch.int64, device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        worst = int(worst.item())
        if worst == 0:
            break
        cap1 = max(cap1, worst)
    gd._mark("multi: cut")
    # the owner of a bucket range needs, from every source, where that source's entries of its buckets sit
    # (bucket-sorted positions, mine + 1 words) and how many records / windows each bucket holds there
    send_off = torch.cat([offsets[bounds[g]:bounds[g + 1] + 1] for g in range(world)])
    got_off = torch.empty(world * (mine + 1), dtype=torch.int64, device=dev)
    got_hist = torch.empty(world * mine, dtype=torch.int64, device=dev)
    sizes = [bounds[g + 1] - bounds[g] for g in range(world)]
    # these two collectives are also the barrier between "every rank has cut and indexed its records" and the gathers
    dist.all_to_all_single(got_off, send_off, output_split_sizes=[mine + 1] * world,
                           input_split_sizes=[n + 1 for n in sizes])
    dist.all_to_all_single(got_hist, hist, output_split_sizes=[mine] * world, input_split_sizes=sizes)
    gd._mark("multi: exchange 0")
    sources = gn.GaSkSources()
    for g in range(world):
        sources.records[g] = slots_region.mapped[g]
        sources.index[g] = index_region.mapped[g]
        sources.l1_capacity[g] = cap1
    sources.first_bucket, sources.n_sources = bounds[rank], world
    if mine:
        summed = got_hist.view(world, mine).sum(dim=0).contiguous()
        n_occ_mine = int((summed & 0xFFFFFFFF).sum().item())
    # shared result buffers on rank 0: [counter, padded to 64 B][keys: rows x 8 B][stamps: rows x 32 B]
    L = gn.lib()
    out_region = _PEERS.setdefault((dev.index, "solid"), PeerRegion())
    rows = max(_OUT_ROWS.get(dev.index, 0), 1 << 20, min(n_occ // (int(threshold) + 1), n_occ // 48) + 1024)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    while True:
        _OUT_ROWS[dev.index] = rows
        out_region.ensure(64 + rows * 40)             # collective; its barriers order the zeroing below for a retry
        root = out_region.mapped[0]
        if rank == 0:
            gn.check(L.ga_fill_bytes(C.c_void_p(root), 0, 64, gd._stream()))
        # nobody appends before the counter is zero: a one-word all-reduce, stream-ordered on every rank
        dist.all_reduce(total)
        error = None
        if mine:
            sources.solid_counter = root
            try:
                gd.sk_bucket_pass(None, None, got_off, world, summed, mine, k, threshold, max(n_occ_mine, 1),
                                  reads.status, l2_bits=l2_bits, sources=sources,
                                  out=(_RawBuffer(root + 64, dev), _RawBuffer(root + 64 + 8 * rows, dev), rows))
            except Exception as exc:        # noqa: BLE001 -- raised on every rank below, nobody waits in a collective
                error = exc
        # every rank's results have landed (all-reduce = barrier; it also carries "some rank failed"), then rank 0
        # reads the total and tells the others
        total.fill_(0 if error is None else 1)
        dist.all_reduce(total)
        if error is not None:
            raise error
        if int(total.item()):
            raise RuntimeError("sharded build: the bucket pass failed on another rank (see its traceback)")
        if rank == 0:
            gn.check(L.ga_copy_bytes(gn.ptr(total), C.c_void_p(root), 8, gd._stream()))
        dist.broadcast(total, 0)
        n_all = int(total.item())
        total.zero_()
        if n_all <= rows:
            break
        rows = n_all + n_all // 16               # did not fit (nothing was written beyond the rows): count again
    if rank != 0:
        return None, 0, None
    return _RawBuffer(root + 64, dev), n_all, _RawBuffer(root + 64 + 8 * rows, dev)


def _sharded_step_buckets(reads, k, threshold, to_host, gn, gd, feed=None):
    dev = reads.words.device
    world, rank = dist.get_world_size(), dist.get_rank()
    # occurrences of all ranks (the bucket geometry must be the same everywhere); this first collective of a step
    # is also what keeps a rank from cutting new records over slots a peer's previous bucket pass still gathers from
    mine_occ = torch.tensor([reads.windows_total(k), gd.sk_records_estimate(reads, k)], dtype=torch.int64, device=dev)
    every_occ = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(every_occ, mine_occ)
    every_occ = every_occ.view(world, 2).tolist()
    n_occ = sum(int(row[0]) for row in every_occ)
    most_records = max(int(row[1]) for row in every_occ)
    l1_bits, l2_bits = gd.sk_geometry(n_occ)
    n_buckets = 1 << (l1_bits + l2_bits)
    graph = gd.BuiltGraph(False, k - 1, reads.alphabet, 1)
    key_parts, stamp_parts = [], []
    received = []
    route = EXCHANGE if EXCHANGE != "auto" else ("peer" if world >= 4 else "push")
    if route != "nccl" and not (dist.get_backend() == "nccl" and peers_available()):
        route = "nccl"
    if route == "peer":
        # 1.-3. in one go: records stay where they were cut, the owners gather them while they count
        slots = gd.sk_l1_capacity(reads, k, l1_bits, most_records)        # the same number on every rank
        all_keys, n_all, all_stamps = _count_in_place(reads, k, threshold, l1_bits, l2_bits, n_buckets, slots, n_occ,
                                                      feed, gn, gd)
        gd._mark("multi: bucket pass")
        if rank != 0:
            return None
        if n_all == 0:
            return graph
        return gd.resolve_and_emit(graph, all_keys, n_all, all_stamps, k, reads.alphabet, reads.status, to_host)
    else:
        # 1. local records by bucket, 2. every bucket's records to its owner
        exchange = _exchange_push if route == "push" else _exchange_nccl
        received = exchange(reads, k, l1_bits, l2_bits, n_buckets, feed, gn, gd)
    for h, (works, got_bases, got_meta, got_hist, starts, mine) in enumerate(received):
        for wk in works:
            wk.wait()                  # orders the current stream after the transfer; the host does not block
        gd._mark("multi: exchange %d" % h)
        error = None
        try:
            if mine:
                # 3. one segment per source rank: positions from the per-source record counts
                per_source = got_hist.view(world, mine)
                seg_offsets = torch.zeros((world, mine + 1), dtype=torch.int64, device=dev)
                seg_offsets[:, 1:] = torch.cumsum(per_source >> 32, dim=1)
                seg_offsets += torch.tensor(starts, dtype=torch.int64, device=dev).view(world, 1)
                summed = per_source.sum(dim=0).contiguous()
                n_occ_mine = int((summed & 0xFFFFFFFF).sum().item())
                solid_keys, n_solid, edge_stamp = gd.sk_bucket_pass(
                    got_bases, got_meta, seg_offsets.contiguous(), world, summed, mine, k, threshold,
                    max(n_occ_mine, 1), reads.status)
                # the pass reuses its output workspace: keep this phase's (small) result
                key_parts.append(solid_keys[:n_solid].clone())
                stamp_parts.append(edge_stamp[:4 * n_solid].view(-1, 4).clone())
        except Exception as exc:            # noqa: BLE001 -- reported on every rank before the next collective
            error = exc
        raise_together(error)
    del received
    if key_parts:
        solid_keys, edge_stamp = torch.cat(key_parts), torch.cat(stamp_parts)
    else:
        solid_keys = torch.zeros((0, 1), dtype=torch.int64, device=dev)
        edge_stamp = torch.zeros((0, 4), dtype=torch.int64, device=dev)
    n_solid = solid_keys.shape[0]
    # 4. everything solid meets on rank 0
    gd._mark("multi: bucket pass")
    with gd._timed("gather"):
        all_keys, rows = gather_rows(solid_keys, want_rows=True)
        all_stamps = gather_rows(edge_stamp, recv_rows=rows)
    gd._mark("multi: gather")
    if rank != 0:
        return None
    n_all = all_keys.shape[0]
    if n_all == 0:
        return graph
    return gd.resolve_and_emit(graph, all_keys.contiguous(), n_all, all_stamps.contiguous().view(-1), k, reads.alphabet,
                               reads.status, to_host)


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
    if reads.paired:
        return _paired_tail(reads, k, to_host, gn, gd, graph, solid, solid_cap, solid_keys, n_solid, kw)
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


def _export_stamp_table(L, gd, gn, table, cap, queries=None):
    """Occupied slots of a query (queries=None) or query-edge table as (keys[, keys2], stamps) device lists."""
    dev = table.device
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    room = cap
    keys = torch.empty(room, dtype=torch.int64, device=dev)
    keys2 = torch.empty(room, dtype=torch.int64, device=dev) if queries is not None else None
    stamps = torch.empty(room, dtype=torch.int64, device=dev)
    gn.check(L.ga_stamp_table_export(gn.ptr(table), cap, gn.ptr(queries), gn.ptr(keys), gn.ptr(keys2),
                                     gn.ptr(stamps), room, gn.ptr(n_out), gd._stream()))
    n = int(n_out.item())
    return (keys[:n], stamps[:n]) if queries is None else (keys[:n], keys2[:n], stamps[:n])


def _paired_tail(reads, k, to_host, gn, gd, graph, solid, solid_cap, solid_keys, n_solid, kw):
    """Steps 5-6 for read pairs: every rank builds its query / query-edge tables from its own pairs against
    the replicated solid table, the tables travel to rank 0 as key lists and are folded with min(stamp)
    (ga_paired_merge); rank 0 then resolves the fuzzy node groups and emits the CSR exactly as one GPU would
    (debruijn_graph.py:269-347)."""
    L = gn.lib()
    rank = dist.get_rank()
    status = reads.status
    error = None
    try:
        queries, qedges, cap, dh = gd.paired_tables(reads, k, solid, solid_cap, n_solid, status)
        if gd._check_status(status) & (gn.ST_TABLE_FULL | gn.ST_BAD_SYMBOL):
            raise gn.GaError("table overflow or bad symbol in the sharded build")
        q_keys, q_stamps = _export_stamp_table(L, gd, gn, queries, cap)
        e_pk, e_sk, e_stamps = _export_stamp_table(L, gd, gn, qedges, cap, queries)
    except Exception as exc:            # noqa: BLE001  -- reported on every rank below
        error = exc
    raise_together(error)
    del queries, qedges
    q_keys, rows_q = gather_rows(q_keys, want_rows=True)
    q_stamps = gather_rows(q_stamps, recv_rows=rows_q)
    e_pk, rows_e = gather_rows(e_pk, want_rows=True)
    e_sk = gather_rows(e_sk, recv_rows=rows_e)
    e_stamps = gather_rows(e_stamps, recv_rows=rows_e)
    all_dh = gather_rows(dh.view(1, -1))
    if rank != 0:
        return None
    dev = solid.device
    # the two smallest occurrences per (symbol A, symbol B) over all ranks (stamps < 2^63: signed order is fine
    # once "none" = -1 is moved to the top)
    world = all_dh.shape[0]
    pairs = all_dh.view(world, -1, 2).permute(1, 0, 2).reshape(-1, 2 * world)
    pairs = torch.where(pairs < 0, torch.full_like(pairs, (1 << 63) - 1), pairs)
    pairs = torch.sort(pairs, dim=1).values[:, :2]
    dh_all = torch.where(pairs == (1 << 63) - 1, torch.full_like(pairs, -1), pairs).contiguous().view(-1)
    n_q, n_e = q_keys.shape[0], e_pk.shape[0]
    cap = max(1024, 3 * max(n_q, n_e))
    while True:
        status.zero_()
        queries = torch.full((cap, 2), -1, dtype=torch.int64, device=dev)
        qedges = torch.full((cap, 2), -1, dtype=torch.int64, device=dev)
        gn.check(L.ga_paired_merge(gn.ptr(q_keys.contiguous()), gn.ptr(q_stamps.contiguous()), n_q,
                                   gn.ptr(e_pk.contiguous()), gn.ptr(e_sk.contiguous()), gn.ptr(e_stamps.contiguous()),
                                   n_e, gn.ptr(queries), cap, gn.ptr(qedges), cap, gn.ptr(status), gd._stream()))
        if not gd._check_status(status) & gn.ST_STAMP_FULL:
            break
        cap *= 2
    return gd.emit_paired(graph, solid, solid_cap, solid_keys, n_solid, kw, k, reads.alphabet, queries, qedges, cap,
                          dh_all, to_host)


def sharded_host_step(ascii_pinned: torch.Tensor, n_reads: int, read_len: int, first_read: int, k: int,
                      threshold: int):
    """Host-buffer entry of the sharded build: this rank's reads as ASCII in pinned host memory ->
    (streamed H2D + pack overlapped with the scatter) -> exchange -> CSR arrays on rank 0's host."""
    import numpy as np
    import ga_device as gd
    alphabet = gd.Alphabet(np.zeros(0))
    stride = max(1, -(-int(read_len) // 32))
    words = torch.empty(max(1, n_reads * stride), dtype=torch.int64, device=gd._dev())
    reads = gd.DeviceReads.from_packed(words, n_reads, read_len, False, first_read=first_read, estride=read_len,
                                       alphabet=alphabet)
    return sharded_step(reads, k, threshold, to_host=True, feed=gd._stream_in(ascii_pinned, reads, read_len))
