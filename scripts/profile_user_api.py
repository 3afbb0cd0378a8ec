"""Host-side profile of the call a reference user makes (stdin bytes -> read_input -> graph class -> contigs) on one of
the small workloads: cProfile over a few warm passes, top entries by cumulative and by own time.
    python scripts/profile_user_api.py c2|c3 [passes]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genome-assembler_b200")]
import numpy as np
import torch
import bench
import ga_native as gn
import ga_device as gd
import assemble as cli
import debruijn_graph as dg

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
genome_size, n, read_len, paired, k, F, _ = bench.WORKLOADS[which]
mates = 2 if paired else 1
L = gn.lib()
dev = torch.device("cuda", 0)
genome = torch.empty(genome_size, dtype=torch.uint8, device=dev)
gn.check(L.ga_gen_genome(gn.ptr(genome), genome_size, bench.SEED, None))
stride = (read_len + 31) // 32
words = torch.empty(n * mates * stride, dtype=torch.int64, device=dev)
gn.check(L.ga_gen_reads(gn.ptr(genome), genome_size, 0, n * mates, read_len, bench.SEED, 100, gn.ptr(words), stride,
                        int(paired), 125, None))
reads = gd.DeviceReads.from_packed(words, n * mates, read_len, paired, estride=read_len)
ascii_dev = torch.empty(n * mates * read_len, dtype=torch.uint8, device=dev)
gn.check(L.ga_unpack_reads(gn.ptr(words), n * mates, read_len, stride, 2, gn.ptr(reads.alphabet.inv_dev),
                           gn.ptr(ascii_dev), None))
rows = ascii_dev.cpu().numpy().reshape(n, mates * read_len)
if paired:
    tail = np.frombuffer(b"|125\n", dtype=np.uint8)
    lines = np.concatenate([rows[:, :read_len], np.full((n, 1), ord("|"), dtype=np.uint8), rows[:, read_len:],
                            np.broadcast_to(tail, (n, tail.size))], axis=1)
else:
    lines = np.concatenate([rows, np.full((n, 1), ord("\n"), dtype=np.uint8)], axis=1)
text = (b"%d\n" % n) + lines.tobytes()
del rows, lines, ascii_dev, reads, words
cls = dg.PairedDeBruijnGraph if paired else dg.DeBruijnGraph


def one():
    parsed, _, _, _ = cli.IOHandler.read_input(io.BytesIO(text))
    g = cls(parsed, k=k, hamming_dist=F)
    return g.enumerate_contigs()


for _ in range(2):
    one()
gd._TRACE = 1            # host wall time between the stage marks, with a device sync at each (ga_device._mark)
gd._last[0] = time.perf_counter()
t0 = time.perf_counter()
one()
print("one traced pass: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
gd._TRACE = 0
prof = cProfile.Profile()
prof.enable()
for _ in range(passes):
    one()
prof.disable()
for order in ("cumulative", "tottime"):
    out = io.StringIO()
    pstats.Stats(prof, stream=out).sort_stats(order).print_stats(28)
    print(out.getvalue()[:6000])
