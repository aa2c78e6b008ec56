import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
_TESTS = os.path.join(ROOT, "tests")
if _TESTS not in sys.path:  # helper modules next to the tests (ransac_lapack)
    sys.path.insert(0, _TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def aps():
    """The product package (directory name is not an identifier, so it is loaded by path)."""
    import __graft_entry__ as ge

    pkg = ge.load_package()
    mode = os.environ.get("APS_TEST_PAIR_EPILOGUE")  # GPU runs only: every pairwise test through the other epilogue
    if mode is not None:
        pkg._lib.default_context().set_pairwise_epilogue(int(mode))
    return pkg
