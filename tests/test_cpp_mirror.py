"""The Eigen-free C++ mirror of the reference interface (include/jrlqp_b200.hpp): compiles against the C-ABI,
fails loudly without a GPU, and reproduces the reference's GI-paper test on a B200."""
import os
import subprocess

import pytest
import torch

import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    S.load_library()
    exe = str(tmp_path / "paper_example")
    libdir = os.path.join(ROOT, "jrl-qp_b200", "_build")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "paper_example.cpp"), "-L" + libdir, "-ljrlqp_b200",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_paper_example(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout and "status 0 iterations 1" in r.stdout
