import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "gpu_next: needs a B200 and was written AFTER the round's GPU budget was spent, so it has never "
                                       "run on one; deliberately outside `-m gpu` until it has (tools/next_gpu_call.sh runs "
                                       "`-m \"gpu or gpu_next\"` first thing next round, then the marker becomes `gpu`)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()
    return oracle


def _have_gpu():
    try:
        from miosqp_b200 import engine
        return engine.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu_next` tests are not `gpu`, so `-m "not gpu"` selects them: skip them where there is no device."""
    have = None
    for item in items:
        if item.get_closest_marker("gpu_next") is not None:
            if have is None:
                have = _have_gpu()
            if not have:
                item.add_marker(pytest.mark.skip(reason="gpu_next: needs a B200 (never run on one yet)"))
