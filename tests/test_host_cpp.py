"""The C++ host mirror of the Go API (wfa_b200/host/wfa.hpp): compiles and links
against libwfacuda.so on CPU; on a GPU it runs the reference-style test."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host")


def _build(built_lib):
    libdir = os.path.dirname(built_lib)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "test_host.cpp"),
                           "-L" + libdir, "-lwfacuda", "-Wl,-rpath," + libdir])


def test_host_cpp_compiles_and_links(built_lib):
    _build(built_lib)
    # without a GPU New() must fail loudly (exit code 2 + message), never fall back
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([EXE], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_host_cpp_runs(built_lib):
    _build(built_lib)
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0 and "host API ok" in r.stdout, (r.stdout, r.stderr)
