"""Host-side logic of the multi-GPU path on the CPU: world_size-2 gloo processes exercise the
tensor-moving helpers of ga_multi (variable all-gather, hash-owner all-to-all, unsigned
min all-reduce).  The kernels themselves are covered by the -m gpu tests."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    sys.path[:0] = [os.path.join(ROOT, "genome-assembler_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ga_multi
    try:
        # variable all-gather: rank r contributes r+2 rows
        local = torch.arange((rank + 2) * 2, dtype=torch.int64).reshape(rank + 2, 2) + 100 * rank
        cat, sizes = ga_multi.all_gather_var(local)
        want = torch.cat([torch.arange((r + 2) * 2, dtype=torch.int64).reshape(r + 2, 2) + 100 * r for r in range(world)])
        assert sizes == [r + 2 for r in range(world)] and torch.equal(cat, want)
        # empty contribution from one rank
        cat, sizes = ga_multi.all_gather_var(local[:0] if rank == 0 else local)
        assert sizes[0] == 0 and cat.shape[0] == sum(sizes)

        # hash-owner exchange: keys 0..19 on each rank (+1000*rank), owner = key % world
        keys = torch.arange(20, dtype=torch.int64) + 1000 * rank
        counts = (torch.arange(20, dtype=torch.int32) + 1) * (rank + 1)
        owner = (keys % world).to(torch.int32)
        got_keys, got_counts = ga_multi.exchange_by_owner(owner, [keys.reshape(-1, 1), counts])
        assert torch.all(got_keys[:, 0] % world == rank)
        assert got_keys.shape[0] == got_counts.shape[0] == 20        # 10 from each of 2 ranks
        total = torch.tensor([int(got_counts.sum())], dtype=torch.int64)
        dist.all_reduce(total)
        assert int(total) == sum((i + 1) * (r + 1) for i in range(20) for r in range(world))
        # (key, count) rows stay paired
        for kk, cc in zip(got_keys[:, 0].tolist(), got_counts.tolist()):
            src = kk // 1000
            assert cc == ((kk % 1000) + 1) * (src + 1)

        # contiguous range exchange: rank r sends r+1 rows to rank 0 and 2 rows to rank 1
        rows = torch.arange((rank + 1 + 2) * 3, dtype=torch.int64).reshape(-1, 3) + 1000 * rank
        got, recv_rows = ga_multi.exchange_ranges(rows, [rank + 1, 2])
        if rank == 0:
            assert recv_rows == [1, 2] and got[:, 0].tolist() == [0, 1000, 1003]
        else:
            assert recv_rows == [2, 2] and got[:, 0].tolist() == [3, 6, 1006, 1009]
        # gather on rank 0 (rank 1 contributes nothing the second time)
        cat = ga_multi.gather_rows(rows[:rank + 1])
        assert cat.shape[0] == (3 if rank == 0 else 0)
        if rank == 0:
            assert cat[:, 0].tolist() == [0, 1000, 1003]
        cat = ga_multi.gather_rows(rows[:0] if rank == 1 else rows[:2])
        assert cat.shape[0] == (2 if rank == 0 else 0)

        # unsigned minimum with all-ones meaning "no stamp"
        stamps = torch.full((6,), -1, dtype=torch.int64)
        if rank == 0:
            stamps[0], stamps[1], stamps[3] = 7, 1 << 40, 5
        else:
            stamps[0], stamps[2], stamps[3] = 3, 9, 6
        ga_multi.all_reduce_min_u64(stamps)
        assert stamps.tolist() == [3, 1 << 40, 9, 5, -1, -1]

        # bookkeeping of the NVLink push exchange: every rank derives, from the all-gathered cut matrix, where its
        # segment starts in each owner's receive arrays; the owners' own view of the segments must agree
        my_cut = torch.tensor([0, 5 + rank, 5 + rank + 3 * (rank + 1)], dtype=torch.int64)   # rows for owner 0, owner 1
        all_cut = torch.empty(world * (world + 1), dtype=torch.int64)
        dist.all_gather_into_tensor(all_cut, my_cut)
        all_cut = all_cut.view(world, world + 1).tolist()
        matrix = [[all_cut[s][g + 1] - all_cut[s][g] for g in range(world)] for s in range(world)]
        assert matrix == [[5, 3], [6, 6]]
        dst_start, seg_start, recv_total = ga_multi.push_plan(matrix, rank)
        assert recv_total == (11 if rank == 0 else 9)
        told = torch.empty(world, dtype=torch.int64)           # told[s] = where source s says its segment starts here
        dist.all_to_all_single(told, torch.tensor(dst_start, dtype=torch.int64))
        assert told.tolist() == seg_start
        assert seg_start[0] == 0 and seg_start[1] == matrix[0][rank]
        # a failure on one rank between two collectives is raised on every rank (nobody is left waiting in the next
        # collective): the failing rank raises its own exception, the others name it
        ga_multi.raise_together(None)                              # nobody failed: returns
        try:
            ga_multi.raise_together(MemoryError("bucket pass failed") if rank == 1 else None)
            raised = None
        except MemoryError as exc:
            raised = "own:" + str(exc)
        except RuntimeError as exc:
            raised = "other:" + str(exc)
        assert raised == ("own:bucket pass failed" if rank == 1 else
                          "other:sharded build: rank 1 failed (see its own traceback)"), raised
        # both ranks are still in step afterwards
        probe = torch.tensor([rank + 1], dtype=torch.int64)
        dist.all_reduce(probe)
        assert int(probe) == 3
        results[rank] = "ok"
    except Exception as exc:      # noqa: BLE001
        results[rank] = repr(exc)
    finally:
        dist.destroy_process_group()


def test_multi_rank_helpers_gloo():
    world = 2
    port = 29600 + os.getpid() % 300
    with mp.get_context("spawn").Manager() as manager:      # not fork: the pytest process has threads by now
        results = manager.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert dict(results) == {0: "ok", 1: "ok"}
