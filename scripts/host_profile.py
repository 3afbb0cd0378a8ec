"""Where does the host time of one device_step go?  (cProfile + wall clock, on the GPU box)"""
import cProfile, pstats, sys, os, time, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genome-assembler_b200")]
import torch
import ga_native as gn, ga_device as gd
from bench import WORKLOADS, SEED
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
genome_size, n_reads, read_len, paired, k, F, desc = WORKLOADS[wl]
dev = torch.device("cuda", 0); L = gn.lib(); mates = 2 if paired else 1; stride = (read_len + 31) // 32
genome = torch.empty(genome_size, dtype=torch.uint8, device=dev)
gn.check(L.ga_gen_genome(gn.ptr(genome), genome_size, SEED, None))
words = torch.empty(n_reads * mates * stride, dtype=torch.int64, device=dev)
gn.check(L.ga_gen_reads(gn.ptr(genome), genome_size, 0, n_reads * mates, read_len, SEED, 100, gn.ptr(words), stride, int(paired), 125, None))
reads = gd.DeviceReads.from_packed(words, n_reads * mates, read_len, paired, estride=read_len)
for _ in range(3):
    gd.device_step(reads, k, F)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    gd.device_step(reads, k, F)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    gd.device_step(reads, k, F)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:6000])
