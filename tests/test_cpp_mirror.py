"""The C++ host mirror (include/rapier_b200.hpp) compiles against the C ABI, links the product library,
fails loudly without a GPU, and on the GPU box reproduces the expected pyramid."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "pyramid_example.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "pyramid_example")
LIBDIR = os.path.join(ROOT, "rapier_b200", "csrc")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", EXE, SRC, "-L" + LIBDIR, "-lrapier_b200", "-Wl,-rpath," + LIBDIR])


def test_cpp_mirror_builds_and_refuses_cpu(built):
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE, "5"], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)          # RB_ERR_NO_DEVICE: no CPU fallback
    assert "no CPU fallback" in r.stderr or "CUDA" in r.stderr


@pytest.mark.gpu
def test_cpp_mirror_runs_on_gpu(built):
    _build()
    r = subprocess.run([EXE, "200"], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "manifolds=290" in r.stdout
