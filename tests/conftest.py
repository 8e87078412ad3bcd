import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle


def _have_gpu():
    """Is there a CUDA device?  Asked of torch, NOT of the engine library: on a GPU box a library that fails to build or
    load must fail the gpu tests loudly, not skip them."""
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` run on a machine without a device skips the `gpu` tests instead of failing them."""
    have = None
    for item in items:
        if item.get_closest_marker("gpu") is not None:
            if have is None:
                have = _have_gpu()
            if not have:
                item.add_marker(pytest.mark.skip(reason="needs a B200"))
