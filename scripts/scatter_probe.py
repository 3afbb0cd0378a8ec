"""Times ga_sk_scatter_reads alone on a quarter-size C4 instance (level-1 geometry of the full workload) for each
kernel variant (GA_SK_SCATTER).  The same script, with debug switches in the kernel that dropped the histogram RED,
either store or the cursor atomic, produced the numbers quoted in csrc/ga_superkmer.cu and DESIGN.md."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genome-assembler_b200")]
import torch
import ga_native as gn
import ga_device as gd

L = gn.lib()
dev = torch.device("cuda", 0)
n, rl, k = int(os.environ.get("PROBE_READS", 25000000)), 150, 31
G = n // 2
stride = (rl + 31) // 32
genome = torch.empty(G, dtype=torch.uint8, device=dev)
gn.check(L.ga_gen_genome(gn.ptr(genome), G, 4, None))
words = torch.empty(n * stride, dtype=torch.int64, device=dev)
gn.check(L.ga_gen_reads(gn.ptr(genome), G, 0, n, rl, 4, 100, gn.ptr(words), stride, 0, 0, None))
reads = gd.DeviceReads.from_packed(words, n, rl, False, estride=rl)
n_occ = reads.windows_total(k)
l1_bits, l2_bits = 10, 10
n_l1, n_buckets = 1 << l1_bits, 1 << (l1_bits + l2_bits)
cap1 = int(n_occ * 2.0 / 17 / n_l1 * 1.5) + 4096
rec = torch.empty(n_l1 * cap1 * 4, dtype=torch.int64, device=dev)
cstride = L.ga_sk_cursor_stride()
status = reads.status


def run(variant, dbg):
    os.environ["GA_SK_SCATTER"] = variant
    best = 1e9
    for _ in range(3):
        cursors1 = torch.zeros(n_l1 * cstride, dtype=torch.int64, device=dev)
        hist = torch.zeros(n_buckets, dtype=torch.int64, device=dev)
        status.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gn.check(L.ga_sk_scatter_reads(C.byref(reads.struct_range(0, n)), k, l1_bits, l2_bits, gn.ptr(rec),
                                       cap1, gn.ptr(cursors1), gn.ptr(hist), gn.ptr(status), None))
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    recs = int(cursors1.sum().item())
    print("%-10s dbg=%-2d %7.2f ms  records=%d status=%d" % (variant, dbg, best, recs, int(status[0].item())), flush=True)


for spec in (sys.argv[1:] or ["warp", "lane", "lane128"]):
    variant, _, dbg = spec.partition(":")
    os.environ["GA_SK_DBG"] = dbg or "0"          # only a probe build of the library looks at it
    run(variant, int(dbg or 0))
