"""Times ga_sk_count_build alone on a quarter-size C4 instance.  With a probe build of the library (GA_SK_DBG:
1 no second walk, 3 counting only, 7 window walk + hash only, 15 record loads only) it shows what each stage of
the bucket kernel costs; results with GA_SK_DBG != 0 are garbage."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genome-assembler_b200")]
import torch
import ga_native as gn
import ga_device as gd

L = gn.lib()
dev = torch.device("cuda", 0)
n, rl, k, F = int(os.environ.get("PROBE_READS", 25000000)), 150, 31, 3
G = n // 2
stride = (rl + 31) // 32
genome = torch.empty(G, dtype=torch.uint8, device=dev)
gn.check(L.ga_gen_genome(gn.ptr(genome), G, 4, None))
words = torch.empty(n * stride, dtype=torch.int64, device=dev)
gn.check(L.ga_gen_reads(gn.ptr(genome), G, 0, n, rl, 4, 100, gn.ptr(words), stride, 0, 0, None))
reads = gd.DeviceReads.from_packed(words, n, rl, False, estride=rl)
n_occ = reads.windows_total(k)
l1_bits, l2_bits = 8, 10            # 2^18 buckets: the windows per bucket of the full workload
n_buckets = 1 << (l1_bits + l2_bits)
rec, _, offsets, hist, total, index, cap1 = gd.sk_scatter_local(reads, k, l1_bits, l2_bits, dense=False)
out_cap = n_occ // 48 + 1024
solid_keys = torch.empty((out_cap, 1), dtype=torch.int64, device=dev)
edge_stamp = torch.empty(4 * out_cap, dtype=torch.int64, device=dev)
spill_list = torch.empty(1 << 16, dtype=torch.int64, device=dev)
status = reads.status
for spec in (sys.argv[1:] or ["0"]):
    os.environ["GA_SK_DBG"] = spec
    best = 1e9
    for _ in range(3):
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
        status.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gn.check(L.ga_sk_count_build(gn.ptr(rec), None, gn.ptr(offsets), 1, gn.ptr(hist), n_buckets, k, F,
                                     gd.SUPERKMER_TABLE_SLOTS, gd.SUPERKMER_MAX_SOLID, gn.ptr(solid_keys),
                                     gn.ptr(edge_stamp), out_cap, gn.ptr(counters), gn.ptr(spill_list), 1 << 16,
                                     gn.ptr(status), gn.ptr(index), cap1, l2_bits, None))
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    c = counters.cpu().tolist()
    print("dbg=%-3s %7.2f ms  solid=%d spilled=%d passes=%d failed=%d" %
          (spec, best, c[1], c[2], c[3] & 0xFFFFFFFF, c[3] >> 32), flush=True)
