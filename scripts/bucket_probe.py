"""Times ga_sk_count_build alone on a quarter-size C4 instance (same coverage, same windows per bucket) and
prints what the buckets hold: windows / records per bucket (percentiles), distinct windows and candidates per
pass.  GA_LIB picks another build of the library (A/B of kernel variants)."""
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
l1_bits, l2_bits = int(os.environ.get("PROBE_L1", "8")), 10            # 2^18 buckets: the windows per bucket of the full workload
n_buckets = 1 << (l1_bits + l2_bits)
rec, _, offsets, hist, total, index, cap1 = gd.sk_scatter_local(reads, k, l1_bits, l2_bits, dense=False)
win = (hist & 0xFFFFFFFF).float()
recs = (hist >> 32).float()
q = torch.tensor([0.01, 0.1, 0.5, 0.9, 0.99, 1.0], device=dev)
print("windows per bucket  mean %.0f  pct(1,10,50,90,99,100) %s" % (win.mean().item(), [int(v) for v in torch.quantile(win, q).tolist()]))
print("records per bucket  mean %.0f  pct %s" % (recs.mean().item(), [int(v) for v in torch.quantile(recs, q).tolist()]))
out_cap = n_occ // 48 + 1024
solid_keys = torch.empty((out_cap, 1), dtype=torch.int64, device=dev)
edge_stamp = torch.empty(4 * out_cap, dtype=torch.int64, device=dev)
spill_list = torch.empty(1 << 16, dtype=torch.int64, device=dev)
status = reads.status
for slots in [int(v) for v in (sys.argv[1:] or [str(gd.SUPERKMER_TABLE_SLOTS)])]:
    best = 1e9
    for _ in range(3):
        counters = torch.zeros(8, dtype=torch.int64, device=dev)
        status.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gn.check(L.ga_sk_count_build(gn.ptr(rec), None, gn.ptr(offsets), 1, gn.ptr(hist), n_buckets, k, F,
                                     slots, gd.SUPERKMER_MAX_SOLID, gn.ptr(solid_keys),
                                     gn.ptr(edge_stamp), out_cap, gn.ptr(counters), gn.ptr(spill_list), 1 << 16,
                                     gn.ptr(status), gn.ptr(index), cap1, l2_bits, None))
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    c = counters.cpu().tolist()
    passes = max(c[3] & 0xFFFFFFFF, 1)
    print("slots=%-5d %7.2f ms  solid=%d spilled=%d passes=%d failed=%d (table %d, queue %d, solid %d) distinct/pass=%.0f notes/pass=%.0f" %
          (slots, best, c[1], c[2], passes, c[3] >> 32, c[6] & 0xFFFFFFFF, c[7], c[6] >> 32, c[4] / passes, c[5] / passes), flush=True)
