"""pytest configuration: markers + import paths.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI symbol check, gloo sharding (CPU only).
`-m gpu`       : CUDA path vs oracle / goldens through the C-ABI (needs a B200).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "time-opt-ilqr_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
