"""Debug: candidate edge stamps of the bucketed path vs the stamps of the table path, per (key, symbol)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genome-assembler_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import numpy as np, torch
import ga_native as gn, ga_device as gd
from helpers import GOLDEN, reads_for
gold = GOLDEN["cases"]["nd-unpaired"]
reads = reads_for(gold["recipe"]); k, F = gold["k"], gold["F"]
dr = gd.DeviceReads(reads, False)
L = gn.lib()
# table path
counts = gd.KmerCounts(k, dr)
keys, n = gd._solid_keys(counts, F)
cap = int(1.7 * n) + 64
solid = torch.empty(cap * 16, dtype=torch.uint8, device="cuda")
L.ga_table_clear(gn.ptr(solid), cap, 1, None)
L.ga_table_insert_ids(gn.ptr(keys), n, 1, 0, gn.ptr(solid), cap, gn.ptr(dr.status), None)
ns = torch.full((n,), -1, dtype=torch.int64, device="cuda"); es = torch.full((4 * n,), -1, dtype=torch.int64, device="cuda")
gd.build_dna4(dr, k, solid, cap, keys, n, 1, ns, es, dr.status)
torch.cuda.synchronize()
ka = keys[:n, 0].cpu().numpy().view(np.uint64); ea = es.cpu().numpy().view(np.uint64).reshape(n, 4)
oa = np.argsort(ka); ka, ea = ka[oa], ea[oa]
# bucketed path
gd.SUPERKMER_MIN_OCC = 0
for name, val in [a.split("=") for a in sys.argv[1:]]:
    setattr(gd, name, int(val))
kb, nb, eb = gd.superkmer_stamps(dr, k, F)
torch.cuda.synchronize()
kb = kb[:nb, 0].cpu().numpy().view(np.uint64); eb = eb[:4 * nb].cpu().numpy().view(np.uint64).reshape(nb, 4)
ob = np.argsort(kb); kb, eb = kb[ob], eb[ob]
print("solid", n, nb, "same keys", np.array_equal(ka, kb))
NONE = np.uint64(0xFFFFFFFFFFFFFFFF)
have = ea != NONE
print("edges table path", int(have.sum()), "candidates", int((eb != NONE).sum()))
bad = have & (eb != ea)
print("mismatching stamps", int(bad.sum()), "of which candidate missing", int((bad & (eb == NONE)).sum()),
      "candidate larger", int((bad & (eb != NONE) & (eb > ea)).sum()), "candidate smaller", int((bad & (eb < ea)).sum()))
idx = np.argwhere(bad)[:10]
for i, c in idx:
    print(hex(int(ka[i])), c, int(ea[i, c]), int(eb[i, c]) if eb[i, c] != NONE else None, "read", int(ea[i, c]) // 100, "pos", int(ea[i, c]) % 100)
