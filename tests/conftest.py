import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a B200: skip them (instead of failing on the first CUDA call) when this host has no Blackwell GPU,
    so a plain `pytest tests` is green on a CPU box and the parity tests run by themselves on the GPU box."""
    try:
        import torch
        ok = torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:  # noqa: BLE001
        ok = False
    if ok:
        return
    skip = pytest.mark.skip(reason="needs a B200 (compute capability 10.x); run under gpurun")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
