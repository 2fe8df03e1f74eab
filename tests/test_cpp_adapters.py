"""The C++ host mirrors (include/fast_gicp_b200/fast_gicp.hpp, include/elevation_b200/gpu_process.hpp) compile with a plain
g++ against the C-ABI and link to libmrslam_b200.so; on the GPU box the small program also runs a registration."""
import os
import subprocess

import pytest

from conftest import ROOT


def _build(tmp_path):
    exe = str(tmp_path / "test_adapters")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_adapters.cpp"),
           "-o", exe, "-L", os.path.join(ROOT, "mr_slam_b200"), "-lmrslam_b200", "-Wl,-rpath," + os.path.join(ROOT, "mr_slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_adapters_compile_and_link(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "NO_GPU" in r.stdout or "ADAPTERS_OK" in r.stdout


@pytest.mark.gpu
def test_adapters_run_on_gpu(tmp_path, gpu):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ADAPTERS_OK" in r.stdout, r.stdout + r.stderr
