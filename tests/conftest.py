import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _cuda_extension_present():
    """The C-ABI library is git-ignored: a fresh checkout that runs the tests before ``__graft_entry__.build()`` gets it
    built here (nvcc cross-compiles sm_100a without a GPU).  Never rebuilt when present — the GPU box receives the .so."""
    from ipp_rl_b200 import build as b

    if not os.path.exists(b.LIB):
        b.build_extension(force=True)
    yield


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
