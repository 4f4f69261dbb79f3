import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return load


def set_opt(monkeypatch, name, value):
    """Set a debug switch of libsnvc_b200 for the rest of the test: the library reads its switches from the environment
    only when it is loaded, so a test sets them through snvc_set_option (and the environment, for spawned ranks)."""
    from snvc_b200 import _lib
    monkeypatch.setenv(name, str(value))
    _lib.set_option(name, value)


@pytest.fixture(autouse=True)
def _reset_library_options():
    yield
    try:
        from snvc_b200 import _lib
        if _lib._lib is not None:
            _lib.set_option(None)
    except Exception:
        pass
