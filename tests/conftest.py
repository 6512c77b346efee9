import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

EMU_PATH = os.path.join(ROOT, "tests", "_emu", "libcmbl_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def emu(pkg):
    """Host emulator build of the CUDA kernel sources (g++ -DCMBL_EMU): lets the CPU suite check kernel index logic
    against the oracle.  Test infrastructure only — the package never loads it by itself."""
    csrc = os.path.join(ROOT, "cmblensing.jl_b200", "csrc")
    subprocess.run(["make", "-C", csrc, "-j", str(os.cpu_count() or 2), "emu"], check=True, stdout=subprocess.DEVNULL)
    return pkg._lib.Library(EMU_PATH)


class _BackEnd:
    """Where a kernel test runs: `lib` is passed to ProjLambert (None = the product library), `device` is the torch device."""
    def __init__(self, name, lib, device):
        self.name, self.lib, self.device = name, lib, device

    def library(self, pkg):
        return self.lib if self.lib is not None else pkg.load()


@pytest.fixture(scope="session", params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def be(request, pkg):
    """Back end of tests/test_kernels.py: the host emulator (CPU suite) or the sm_100a library on cuda:0 (-m gpu)."""
    if request.param == "emu":
        return _BackEnd("emu", request.getfixturevalue("emu"), "cpu")
    request.getfixturevalue("cuda_pkg")
    return _BackEnd("cuda", None, "cuda:0")


@pytest.fixture(scope="session")
def cuda_pkg(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    pkg.load()          # raises loudly if libcmbl_b200.so is missing: there is no fallback
    return pkg
