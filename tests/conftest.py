"""Test configuration.

`-m "not gpu"`: oracle against the golden fixtures and (when oracle/_ref was built here) against the unmodified
reference, host logic, C-ABI symbol/struct checks — no compute call needs a GPU.
`-m gpu`: parity of the CUDA path (through the C ABI) against the oracle and the fixtures. GPU tests FAIL, not skip,
when libvgl_b200.so is missing or cannot initialise a device: there is no fallback path to hide behind.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the reference segfaults with one OpenMP thread; PR goldens were produced with T = 8
os.environ.setdefault("OMP_NUM_THREADS", "8")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()  # builds liboracle.so on first use
    return O


@pytest.fixture(scope="session")
def vgl():
    import vectorgraphlibrary_b200 as V
    if not os.path.exists(V.LIB_PATH):
        from vectorgraphlibrary_b200 import build
        build.build()
    V.lib()
    return V


@pytest.fixture(scope="session")
def ctx(vgl):
    c = vgl.Context(0)  # raises VglbError when no device: GPU tests then fail loudly
    yield c
    c.close()


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["rmat_s8_ef4", "kron_s10_ef16", "ru_s10_ef32", "rmat_s11_ef8"]


@pytest.fixture(scope="session", params=GOLDEN_CASES)
def golden(request):
    import numpy as np
    return request.param, np.load(os.path.join(GOLDEN_DIR, request.param + ".npz"))
