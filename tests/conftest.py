import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding

    return binding.Oracle()


@pytest.fixture(scope="session")
def reference_stable():
    from oracle import binding

    if not binding.Reference.available("stable"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return binding.Reference("stable")


@pytest.fixture(scope="session")
def reference_verbatim():
    from oracle import binding

    if not binding.Reference.available("verbatim"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return binding.Reference("verbatim")
