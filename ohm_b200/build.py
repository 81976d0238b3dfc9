"""Builds libohmb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libohmb200.so")
SOURCES = [os.path.join(HERE, "csrc", "ohmb200.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))
                  if f.endswith((".cuh", ".inl", ".h", ".cu"))] + [
    os.path.join(ROOT, "include", "ohmb200.h")]

# --fmad=false: voxel sequences must match the CPU mapper bit for bit (see ohmb200_device.cuh).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIB] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
