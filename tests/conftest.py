import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def lib_built():
    """The CUDA shared library, built in-tree with nvcc if stale (cross-compiles without a GPU)."""
    from althea_b200 import build
    return build.build()


def _make_ctx(parity):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from althea_b200 import engine
    torch.cuda.init()
    return engine.Context(0, parity_math=parity)


@pytest.fixture(scope="session")
def ctx_fast(lib_built):
    c = _make_ctx(False)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx_parity(lib_built):
    c = _make_ctx(True)
    yield c
    c.close()
