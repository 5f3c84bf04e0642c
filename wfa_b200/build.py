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
    src = [os.path.join(CSRC, f) for f in ("wfacuda.cu", "wfa_kernels.cuh", "wfa_lane.cuh", "wfa_slim.cuh", "wfa_wide.cuh", "wfa_render.cuh")] + \
          [os.path.join(os.path.dirname(HERE), "include", "wfacuda.h")]
    if force or _stale(LIB, src):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, src[0]]
        subprocess.check_call(cmd)
    return LIB


def build_host_bench(force=False):
    """bench_api: the reference's call shape (AlignBatch on per-pair byte strings) through the C++
    mirror of the Go API -- what bench.py reports as e2e.api_value."""
    exe = os.path.join(HERE, "host", "bench_api")
    src = [os.path.join(HERE, "host", "bench_api.cpp"), os.path.join(HERE, "host", "wfa.hpp"), LIB]
    if force or _stale(exe, src):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-o", exe, src[0], "-I" + os.path.join(HERE, "host"),
                               "-L" + HERE, "-lwfacuda", "-lwfagen", "-Wl,-rpath,$ORIGIN/.."])
    return exe


def build_all(force=False):
    from . import datagen
    datagen.build(force)
    lib = build_cuda(force)
    build_host_bench(force)
    return lib


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(LIB)
