import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build every library once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from voxelyze_b200 import capi
    return capi.load_oracle()


@pytest.fixture(scope="session")
def reference(built):
    from voxelyze_b200 import capi
    if not os.path.exists(capi.REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return capi.load_reference()


@pytest.fixture(scope="session")
def product(built):
    from voxelyze_b200 import capi
    return capi.load_product()       # raises if the CUDA library is missing: no fallback
