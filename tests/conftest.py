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
    """The CPU oracle (test infrastructure): oracle/abcdez_oracle.c via ctypes."""
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def A():
    """The product package (abcdez.jl_b200/ loaded as abcdez_b200)."""
    import abcdez_b200
    return abcdez_b200


@pytest.fixture(scope="session")
def gpu_ctx(A):
    import ctypes
    try:
        ctx = A.default_context()
    except A.ABCdeZError as e:
        pytest.fail(f"CUDA context could not be created on a gpu-marked test: {e}")
    return ctx
