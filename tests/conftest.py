import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def assets_dir(tmp_path_factory):
    """<tmp>/id1/... with palette, colormap and every synthetic map the tests use."""
    from rcutil import make_assets
    base = str(tmp_path_factory.mktemp("rc_assets"))
    make_assets(base)
    return base


@pytest.fixture(scope="session")
def sim_lib():
    """Host build of the C-ABI with the CPU loop backend (test infrastructure)."""
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    return os.path.join(d, "_build", "libref_cuda_sim.so")


@pytest.fixture(scope="session")
def oracle_available():
    from oracle import refsoft
    if not refsoft.available() and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    if not refsoft.available(True) and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "big"])
    return refsoft.available()
