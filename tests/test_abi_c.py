"""tests/abi_smoke.c compiled by gcc (C99) against include/cmbl_b200.h and linked with libcmbl_b200.so: the boundary is usable from plain C.
Without a GPU the program checks the error contract; with one (-m gpu) it runs plan, FFT, LenseFlow, dot and get_max_lensing_step."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    so_dir = os.path.join(ROOT, "cmblensing.jl_b200")
    assert os.path.exists(os.path.join(so_dir, "libcmbl_b200.so")), "build the library first (__graft_entry__.build())"
    import torch  # noqa: F401  (only to locate the CUDA runtime that ships with it)
    import nvidia.cuda_runtime as rt
    cudart = os.path.join(os.path.dirname(rt.__file__), "lib")
    exe = os.path.join(tmp, "abi_smoke")
    lib = [f for f in os.listdir(cudart) if f.startswith("libcudart.so")][0]
    cmd = [shutil.which("gcc") or "gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe,
           "-L" + so_dir, "-lcmbl_b200", os.path.join(cudart, lib), "-lm", "-Wl,-rpath," + so_dir, "-Wl,-rpath," + cudart]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_abi_from_c_without_gpu(tmp_path):
    exe = _build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "error contract ok" in r.stdout


@pytest.mark.gpu
def test_abi_from_c_on_gpu(tmp_path, cuda_pkg):
    exe = _build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi smoke ok" in r.stdout and "kernel path 3" in r.stdout
