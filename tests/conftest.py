import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import _oracle

    return _oracle.load_oracle()


@pytest.fixture(scope="session")
def reflib():
    import _oracle

    if not _oracle.ref_available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    return _oracle.load_ref()
