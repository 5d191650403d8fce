import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement (oracle/srcnn_oracle.c)."""
    from oracle.loader import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref():
    """The reference's own kernels compiled by g++ (oracle/_ref); skip when not built."""
    from oracle.loader import Oracle, have, build
    if not have("reference"):
        try:
            build("reference")
        except Exception:
            pass
    if not have("reference"):
        pytest.skip("oracle/_ref/libsrcnn_ref.so not built (needs /root/reference)")
    return Oracle("reference")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))
