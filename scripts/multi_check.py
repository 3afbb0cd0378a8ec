"""Run under torchrun on N GPUs: the sharded build must equal the single-GPU build bit for bit.
Rank 0 also builds the whole read set alone and compares every CSR array."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genome-assembler_b200")]
import numpy as np
import torch
import torch.distributed as dist
import ga_native as gn
import ga_device as gd
import ga_multi


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L = gn.lib()
    ok = True
    # unpaired: the bucketed route with both exchanges (one kernel over NVLink peer memory; dense copy + NCCL
    # all-to-all), a case large enough for several level-1 buckets, and the k > 32 table route
    for genome_size, n_reads, read_len, k, F, exchange in ((200000, 60000, 100, 31, 3, "peer"),
                                                           (3000000, 1500003, 150, 31, 3, "peer"),
                                                           (200000, 60000, 100, 31, 3, "push"),
                                                           (200000, 60000, 100, 31, 3, "push-gather"),
                                                           (200000, 60000, 100, 31, 3, "nccl"),
                                                           (3000000, 1500003, 150, 31, 3, "push"),
                                                           (3000000, 1500003, 150, 31, 3, "push-gather"),
                                                           (50000, 30011, 150, 41, 2, "push")):
        ga_multi.EXCHANGE = exchange.split("-")[0]
        ga_multi.PUSH = "gather" if exchange.endswith("gather") else "sorted"
        if exchange == "push":
            ga_multi.PUSH = "sorted"
        stride = (read_len + 31) // 32
        genome = torch.empty(genome_size, dtype=torch.uint8, device=dev)
        gn.check(L.ga_gen_genome(gn.ptr(genome), genome_size, 5, None))
        lo, hi = n_reads * rank // world, n_reads * (rank + 1) // world
        words = torch.empty(max(1, (hi - lo) * stride), dtype=torch.int64, device=dev)
        gn.check(L.ga_gen_reads(gn.ptr(genome), genome_size, lo, hi - lo, read_len, 5, 100, gn.ptr(words), stride, 0, 0, None))
        shard = gd.DeviceReads.from_packed(words, hi - lo, read_len, False, first_read=lo, estride=read_len)
        got = ga_multi.sharded_step(shard, k, F, to_host=True)
        if rank == 0:
            allw = torch.empty(n_reads * stride, dtype=torch.int64, device=dev)
            gn.check(L.ga_gen_reads(gn.ptr(genome), genome_size, 0, n_reads, read_len, 5, 100, gn.ptr(allw), stride, 0, 0, None))
            whole = gd.DeviceReads.from_packed(allw, n_reads, read_len, False, estride=read_len)
            want = gd.build_graph(gd.KmerCounts(k, whole), whole, F, to_host=True)
            same = all(np.array_equal(getattr(got, f), getattr(want, f))
                       for f in ("rowptr", "col", "indeg", "branching", "last_char", "keys_a"))
            print("multi_check world=%d k=%d reads=%d exchange=%s: nodes %d/%d edges %d/%d identical=%s" %
                  (world, k, n_reads, exchange, got.n_nodes, want.n_nodes, got.n_edges, want.n_edges, same), flush=True)
            ok = ok and same and got.n_nodes > 0
    # read pairs (table route: replicated solid set, per-rank query tables merged on rank 0)
    for genome_size, n_pairs, read_len, k, F in ((150000, 40000, 100, 29, 3), (30000, 9001, 100, 41, 2)):
        stride = (read_len + 31) // 32
        genome = torch.empty(genome_size, dtype=torch.uint8, device=dev)
        gn.check(L.ga_gen_genome(gn.ptr(genome), genome_size, 6, None))
        lo, hi = n_pairs * rank // world, n_pairs * (rank + 1) // world
        words = torch.empty(max(1, 2 * (hi - lo) * stride), dtype=torch.int64, device=dev)
        gn.check(L.ga_gen_reads(gn.ptr(genome), genome_size, 2 * lo, 2 * (hi - lo), read_len, 6, 100, gn.ptr(words),
                                stride, 1, 125, None))
        shard = gd.DeviceReads.from_packed(words, 2 * (hi - lo), read_len, True, first_read=lo, estride=read_len)
        got = ga_multi.sharded_step(shard, k, F, to_host=True)
        if rank == 0:
            allw = torch.empty(2 * n_pairs * stride, dtype=torch.int64, device=dev)
            gn.check(L.ga_gen_reads(gn.ptr(genome), genome_size, 0, 2 * n_pairs, read_len, 6, 100, gn.ptr(allw), stride,
                                    1, 125, None))
            whole = gd.DeviceReads.from_packed(allw, 2 * n_pairs, read_len, True, estride=read_len)
            want = gd.build_graph(gd.KmerCounts(k, whole), whole, F, to_host=True)
            same = all(np.array_equal(getattr(got, f), getattr(want, f))
                       for f in ("rowptr", "col", "indeg", "branching", "last_char", "keys_a", "keys_b"))
            same = same and got.num_edges_attr == want.num_edges_attr
            print("multi_check paired world=%d k=%d: nodes %d/%d edges %d/%d identical=%s" %
                  (world, k, got.n_nodes, want.n_nodes, got.n_edges, want.n_edges, same), flush=True)
            ok = ok and same and got.n_nodes > 0
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    ga_multi.release_peers()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
