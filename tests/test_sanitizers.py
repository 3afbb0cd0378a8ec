"""Race and memory checks of the threaded host code, built as plain C++ with ThreadSanitizer and with AddressSanitizer +
UBSan; any report fails the test.
  * csrc/ga_traverse.cu: piecewise contig traversal on several host threads (marks written by one thread next to
    fields read by another), on random chain graphs, against the serial sweep;
  * csrc/ga_parse.cu: the raw-byte read parser on several host threads, against its own serial scan."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "genome-assembler_b200", "csrc")
HERE = os.path.join(ROOT, "tests", "sanitize")
CASES = {"traverse": ("ga_traverse.cu", "traverse_main.cpp", 12), "parse": ("ga_parse.cu", "parse_main.cpp", 6)}


@pytest.mark.parametrize("what", sorted(CASES))
@pytest.mark.parametrize("flags,needle", [("thread", "ThreadSanitizer"), ("address,undefined", "Sanitizer")])
def test_host_threads_under_sanitizers(what, flags, needle, tmp_path):
    source, driver, lines = CASES[what]
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / (what + "_san"))
    # the file is host-only code: built from a copy that sits next to tests/sanitize/ga_common.cuh, which stands in
    # for the CUDA header of the same name (`#include "ga_common.cuh"` looks beside the including file first)
    for name in ("ga_common.cuh", driver):
        shutil.copy(os.path.join(HERE, name), tmp_path / name)
    shutil.copy(os.path.join(CSRC, source), tmp_path / "unit.cpp")
    build = subprocess.run([gxx, "-O1", "-g", "-std=c++17", "-fsanitize=" + flags, "-I", os.path.join(ROOT, "include"),
                            "unit.cpp", driver, "-o", exe, "-lpthread"],
                           capture_output=True, text=True, timeout=600, cwd=tmp_path)
    if build.returncode != 0 and "sanitize" in build.stderr and "cannot find" in build.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0", ASAN_OPTIONS="detect_leaks=1"))
    out = run.stdout + run.stderr
    assert run.returncode == 0 and needle not in out and "runtime error" not in out, out[-3000:]
    assert out.count(" same") == lines and "DIFFERENT" not in out
