import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def gpu_lib():
    """The shipped library.  Fails (does not skip) when it is missing: the GPU
    tests must never pass on anything but the CUDA path."""
    from flake_b200 import api
    return api.load_library()


@pytest.fixture(scope="session")
def emu_lib():
    """Kernel sources compiled for the fiber emulator (tests/cuda_emu): logic only."""
    from flake_b200 import api, build
    return api.load_library(build.build_emu())
