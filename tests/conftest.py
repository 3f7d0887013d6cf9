import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library. Fails (does not skip) if it is missing on a GPU box."""
    from athena_b200 import capi
    return capi.load()


@pytest.fixture(scope="session")
def emul_lib():
    """TEST-ONLY host build of the operator bodies (tests/emul)."""
    from emul_lib import load_emul
    return load_emul()
