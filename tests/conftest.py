import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # The built libraries are git-ignored: in a fresh checkout build them once (nvcc cross-compiles without a GPU).
    lib = os.path.join(ROOT, "mixedprecisionblockqr_b200", "libmpqr.so")
    orc = os.path.join(ROOT, "oracle", "libmpqr_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            import __graft_entry__
            __graft_entry__.build()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# reference test shapes, Cuda/qr.cu:1762-1783 (m, n, r)
REF_SHAPES = [
    (6, 4, 2), (6, 4, 1), (6, 4, 3), (12, 8, 4), (12, 8, 5), (12, 8, 6), (12, 8, 2), (12, 8, 8),
    (12, 8, 3), (24, 16, 8), (24, 16, 12), (60, 40, 8), (60, 40, 16), (80, 80, 16), (97, 90, 16),
    (100, 80, 16), (128, 80, 16), (129, 80, 16), (240, 160, 16), (600, 400, 16),
]
