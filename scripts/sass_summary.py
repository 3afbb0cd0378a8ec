#!/usr/bin/env python3
"""Opcode histogram of the hot kernels from the built objects (cuobjdump -sass), written to profiles/<round>/sass_summary.txt.
    python scripts/sass_summary.py > profiles/r02/sass_summary.txt
Shows which memory / atomic / TMA instructions each kernel really uses (e.g. that no ATOMS.CAST.SPIN loop is left in the
bucket kernel, that the push kernel prefetches with UBLKPF.L2, that records move with 256-bit LDG/STG)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "genome-assembler_b200", "build")
HOT = ["sk_scatter_reads_lane_kernel", "sk_index_buckets_sorted_kernel", "sk_bucket_kernel", "sk_push_records_kernel",
       "sk_resolve_kernel", "sk_scatter_buckets_kernel", "count_kernel", "build_paired_kernel", "sketch_update"]
WATCH = re.compile(r"^(ATOMS|ATOMG|ATOM|RED|LDG|STG|LDS|STS|LDSM|UBLK|UTMA|SYNCS|MATCH|REDUX|VOTE|SHFL|BAR|MEMBAR|CCTL|FENCE|LDGSTS|UTC|TCGEN)")


def main():
    for obj in sorted(os.listdir(BUILD)):
        if not obj.endswith(".o") or "_" in obj.replace("ga_", "", 1):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        name, ops = None, None
        kernels = []
        for line in sass.split("\n"):
            m = re.search(r"Function : (\S+)", line)
            if m:
                name, ops = m.group(1), collections.Counter()
                kernels.append((name, ops))
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if m and ops is not None:
                ops[m.group(1)] += 1
        for name, ops in kernels:
            if not any(h in name for h in HOT):
                continue
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            hot = next(h for h in HOT if h in demangled)
            demangled = demangled[demangled.index(hot):].split("(")[0]
            total = sum(ops.values())
            print("== %s :: %s  (%d SASS instructions)" % (obj, demangled[-90:], total))
            watched = sorted(((op, n) for op, n in ops.items() if WATCH.match(op)), key=lambda t: -t[1])
            print("   memory/sync: " + ", ".join("%s %d" % t for t in watched))
            top = ", ".join("%s %d" % t for t in ops.most_common(12))
            print("   most frequent: " + top)
            spin = [op for op in ops if "SPIN" in op]
            print("   CAS-spin emulation (ATOMS.CAST.SPIN): %s" % (", ".join(spin) if spin else "none"))
            print()


if __name__ == "__main__":
    main()
