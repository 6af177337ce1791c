import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without CUDA (or without the built library) skips the gpu-marked tests instead
    of failing them one by one.  An explicit `-m gpu` run is NOT softened: there a missing GPU / library must fail."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    import torch

    lib = os.path.join(ROOT, "oprl_b200", "liboprl_b200.so")
    if torch.cuda.is_available() and os.path.exists(lib):
        return
    why = "no CUDA device" if not torch.cuda.is_available() else "liboprl_b200.so is not built"
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
