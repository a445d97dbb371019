import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_ffi
    return oracle_ffi.Oracle()


@pytest.fixture(scope="session")
def reference():
    import oracle_ffi
    if not oracle_ffi.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle_ffi.Reference("fix")


@pytest.fixture(scope="session")
def reference_asc():
    import oracle_ffi
    if not oracle_ffi.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle_ffi.Reference("asc")
