#!/usr/bin/env python3
"""Registers, stack (spills), static shared memory and constant bank size of every kernel in the built objects
(cuobjdump --dump-resource-usage), written to profiles/<round>/resource_summary.txt.
    python scripts/resource_summary.py > profiles/r02/resource_summary.txt
What to read off: STACK / LOCAL = 0 means no spills and no local arrays; REG x threads per CTA x CTAs per SM must fit the
65 536 registers of an SM (sk_bucket: 64 x 512 x 2 = 65 536, i.e. the launch bound is exactly met); the dynamic
shared memory (the bucket kernel's 110 KB pool) is asked for at launch and not listed here."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "genome-assembler_b200", "build")


def main():
    print("%-16s %-74s %4s %6s %7s %6s %6s" % ("object", "kernel", "REG", "STACK", "SHARED", "LOCAL", "CONST0"))
    with_stack = []
    for obj in sorted(os.listdir(BUILD)):
        if not obj.endswith(".o") or "_" in obj.replace("ga_", "", 1):
            continue
        out = subprocess.run(["cuobjdump", "--dump-resource-usage", os.path.join(BUILD, obj)],
                             capture_output=True, text=True).stdout
        rows = []
        name = None
        for line in out.split("\n"):
            m = re.match(r"\s*Function (\S+):", line)
            if m:
                name = m.group(1)
                continue
            m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+) CONSTANT\[0\]:(\d+)", line)
            if m and name:
                demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
                demangled = re.sub(r"\(anonymous namespace\)::", "", demangled)
                demangled = re.sub(r"^void ", "", demangled).split("(")[0]
                rows.append((demangled,) + tuple(int(x) for x in m.groups()))
                name = None
        for row in sorted(rows):
            print("%-16s %-74s %4d %6d %7d %6d %6d" % ((obj, row[0][-74:]) + row[1:]))
            if row[2] or row[4]:
                with_stack.append("%s (%d B)" % (row[0], row[2] + row[4]))
    print()
    print("Kernels with a stack frame or local memory (everything else keeps all of its state in registers):")
    print("  " + (", ".join(with_stack) if with_stack else "none"))


if __name__ == "__main__":
    main()
