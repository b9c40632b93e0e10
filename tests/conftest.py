import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.pyoracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref); present in the build container and shipped prebuilt to the GPU box."""
    from oracle.pyoracle import Ref, build
    if not Ref.available() and os.path.isdir("/root/reference/src"):
        build(("ref",))
    if not Ref.available():
        pytest.skip("oracle/_ref/libmcmc_ref.so not built (no /root/reference here)")
    return Ref()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
