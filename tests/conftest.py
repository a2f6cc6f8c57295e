import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "merzbild.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): a C++ restatement of the reference's algorithms."""
    from oracle import oracle as o

    o.lib()
    return o


@pytest.fixture(scope="session")
def mb():
    """The product: ctypes mirror of the reference API over libmerzbild_b200.so (CUDA, sm_100a)."""
    import merzbild_b200 as m

    if not os.path.exists(m.LIB_PATH):  # fresh checkout: build the library from source like __graft_entry__.build() does (nvcc, sm_100a)
        import subprocess

        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "merzbild.jl_b200", "csrc"), "all"])
    return m
