import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_api():
    from oracle_lib import oracle_api as f
    return f()


@pytest.fixture()
def oir(oracle_api):
    """A fresh oracle Ir (CPU restatement of the reference)."""
    from vkjit_b200.ir import Ir
    ir = Ir(_api=oracle_api)
    yield ir
    ir.close()


@pytest.fixture(scope="session")
def cuda_backend():
    """Initialises the product backend on cuda:0; fails loudly when it cannot."""
    import vkjit_b200
    vkjit_b200.init(-1)
    return vkjit_b200


@pytest.fixture()
def cir(cuda_backend):
    """A fresh product Ir (CUDA path through the C ABI)."""
    ir = cuda_backend.Ir()
    yield ir
    ir.close()
