"""Test configuration: registers the ``gpu`` marker and puts the repo root and the drop-in
package directory (``br2-simulator_b200/``, which holds the importable ``br2`` package) on
``sys.path``."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "br2-simulator_b200")
for path in (ROOT, PKG):
    if path not in sys.path:
        sys.path.insert(0, path)
os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(ROOT, ".numba_cache"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "tests", "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU check")
