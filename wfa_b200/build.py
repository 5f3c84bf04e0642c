"""Builds the in-tree native libraries (nvcc cross-compiles sm_100a without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwfacuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]


def _stale(target, sources):
    return (not os.path.exists(target)) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in sources)


def build_cuda(force=False, verbose=False):
    src = [os.path.join(CSRC, f) for f in ("wfacuda.cu", "wfa_kernels.cuh", "wfa_lane.cuh", "wfa_slim.cuh", "wfa_render.cuh")] + \
          [os.path.join(os.path.dirname(HERE), "include", "wfacuda.h")]
    if force or _stale(LIB, src):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, src[0]]
        subprocess.check_call(cmd)
    return LIB


def build_all(force=False):
    from . import datagen
    datagen.build(force)
    return build_cuda(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(LIB)
