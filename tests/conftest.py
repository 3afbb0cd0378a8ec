"""pytest configuration: markers and import paths.

* ``gpu`` marks tests that need a B200 (run with ``-m gpu`` on the GPU box).
* The product lives in the flat directory ``genome-assembler_b200/`` (same module names as
  the upstream scripts, so it is a drop-in); the oracle lives in ``oracle/`` and is
  imported by tests only.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("", "genome-assembler_b200", os.path.join("tests", "golden"), "tests"):
    path = os.path.join(ROOT, sub)
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
