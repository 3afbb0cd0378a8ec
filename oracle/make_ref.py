#!/usr/bin/env python3
"""Recipe for oracle/_ref/: the UNMODIFIED reference (pure Python, six files) copied byte for byte from
/root/reference into the git-ignored directory oracle/_ref/, so that it can be executed on the GPU box
(where /root/reference does not exist) as the CPU arm of bench.py (`--impl reference`, `cpu_baseline`).

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing under genome-assembler_b200/ imports, loads or executes it; the
sources never enter the repository history (oracle/_ref/ is listed in .gitignore, not in .gpurunignore).
There is nothing to compile: the reference has no native code.  Run by __graft_entry__.build() whenever
/root/reference is present (the build container); on the GPU box the copied files are used as they are.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("GA_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("assemble.py", "countminsketch.py", "debruijn_graph.py", "debruijn_node.py", "debug_graph.py",
         "generate_reads.py")


def make() -> bool:
    """True when oracle/_ref/ holds the reference afterwards."""
    if os.path.isdir(SRC):
        os.makedirs(DST, exist_ok=True)
        for name in FILES:
            src, dst = os.path.join(SRC, name), os.path.join(DST, name)
            if not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
                shutil.copyfile(src, dst)
    return all(os.path.exists(os.path.join(DST, name)) for name in FILES)


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref: %s" % ("ready" if ok else "reference not available here"))
    sys.exit(0 if ok else 1)
