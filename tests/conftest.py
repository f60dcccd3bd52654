import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def native_libs():
    """Build the in-tree native artefacts when they are missing (nvcc cross-compiles without a GPU)."""
    from enclone_ranger_b200 import _abi
    if not (os.path.exists(_abi.LIB_JOIN) and os.path.exists(_abi.LIB_SYNTH) and os.path.exists(os.path.join(ROOT, 'oracle', 'liboracle.so'))):
        import __graft_entry__
        __graft_entry__.build()
    return True


@pytest.fixture(scope="session")
def gpu_ctx():
    import enclone_ranger_b200 as E
    ctx = E.JoinContext()
    yield ctx
    ctx.close()
