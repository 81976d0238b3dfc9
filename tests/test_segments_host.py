"""Exactness of the region-binned decomposition, proven on the host (no GPU): the same enumerateSegments /
resumeSegment code the CUDA kernels use is compiled for the CPU by nvcc and compared, visit by visit and bit by bit
(voxel keys and enter/exit ranges), with the sequential walk over ~400k rays: long and short random rays, the
LineWalkTests lattice (exact voxel-boundary ties), axis-aligned and degenerate rays, six map geometries (odd region
dimensions included) and every start/end exclusion flag."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_segments_reproduce_the_sequential_walk(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "segments_host_test")
    subprocess.check_call([
        nvcc, "-std=c++17", "-O1", "-Xcompiler", "-ffp-contract=off", "--fmad=false", "-diag-suppress", "20013",
        "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "ohm_b200", "csrc"),
        "-o", exe, os.path.join(ROOT, "tests", "cpp", "segments_host_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "failures 0" in out.stdout
