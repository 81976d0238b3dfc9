"""The C++ facade (include/ohmb200/GpuMap.hpp) compiles against the C ABI with the reference's class/method names;
on the GPU box the example program drives RayMapper::integrateRays the way the reference tests do."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "facade_example.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "facade_example")
    libdir = os.path.join(ROOT, "ohm_b200")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
                           "-L", libdir, "-lohmb200", f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


def test_facade_compiles_and_links(tmp_path):
    from ohm_b200 import _lib
    _lib.load()
    exe = _build(tmp_path)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_facade_example_runs(gpu, tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "regions" in out.stdout
