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
    from tests import oracle_util
    return oracle_util.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libxaac_ref.so). Built by `make ref` where /root/reference exists;
    the prebuilt .so travels to the GPU box. Tests that need it skip when it is absent."""
    from tests import oracle_util
    r = oracle_util.Ref.try_load()
    if r is None:
        pytest.skip("oracle/_ref/libxaac_ref.so not built (needs /root/reference; run `make ref`)")
    return r


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import libxaac_b200
    c = libxaac_b200.Context(0)
    yield c
    c.close()
