"""Test configuration.  `-m "not gpu"` runs everywhere (oracle vs the reference's golden vectors, the
host logic, the C-ABI export check, and the CUDA sources executed by the fiber emulator in
tests/hostsim); `-m gpu` runs the parity tests proper through the real library on a B200."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SIM_SO = os.path.join(ROOT, "tests", "hostsim", "librsigpu_sim.so")
GPU_SO = os.path.join(ROOT, "rsicnv_b200", "librsigpu.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "librsi_oracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(target):
    subprocess.run(["make", "-s", target], cwd=ROOT, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    if not os.path.exists(ORACLE_SO):
        _make("oracle")
    from bind import Lib
    return Lib("oracle")


@pytest.fixture(scope="session")
def ref():
    from bind import Lib, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference; make ref)")
    return Lib("ref")


@pytest.fixture(scope="session")
def sim_lib():
    if not os.path.exists(SIM_SO):
        _make("sim")
    return SIM_SO


@pytest.fixture(scope="session")
def gpu_lib():
    """the real library; GPU tests FAIL (not skip) when it is missing: there is no fallback"""
    assert os.path.exists(GPU_SO), "rsicnv_b200/librsigpu.so missing: run `make lib` / __graft_entry__.build()"
    return GPU_SO
