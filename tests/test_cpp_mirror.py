"""Builds tests/cpp/test_host_mirror.cpp against include/cheq_b200.hpp + libcheq_b200.so and runs it:
the C++ host-side mirror of cheq's API (the reference is compiled Rust; no Rust toolchain here)."""
import subprocess
from pathlib import Path

import pytest

from conftest import ROOT


def _build(tmp_path_factory):
    from cheq_b200 import build
    lib = build.build()
    exe = tmp_path_factory.mktemp("cpp") / "test_host_mirror"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "test_host_mirror.cpp"),
           "-o", str(exe), "-L", str(lib.parent), "-lcheq_b200", f"-Wl,-rpath,{lib.parent}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    return _build(tmp_path_factory)


def test_cpp_host_mirror_cpu(exe):
    r = subprocess.run([str(exe), "cpu", str(ROOT / "cheq_b200" / "data" / "qeq_params.toml")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_gpu(exe):
    r = subprocess.run([str(exe), "gpu", str(ROOT / "cheq_b200" / "data" / "qeq_params.toml")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
