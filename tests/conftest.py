import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import helpers as H
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
    return H.oracle()


@pytest.fixture(scope="session")
def product_lib():
    """libfermi_b200.so, (re)built for sm_100a if stale."""
    from fermi_b200.build import build_library
    build_library()
    from fermi_b200._lib import lib
    return lib()


@pytest.fixture(scope="session")
def emu():
    """Host build of the device-side core (tests/emu/emu.cpp) -- checker infrastructure only."""
    import emu_binding
    return emu_binding.load()


def golden_cases():
    return ["reads10x", "noisy", "genome"]
