"""The C-ABI library builds here (nvcc cross-compiles sm_100a), loads, and exports
every symbol include/wfacuda.h declares.  No compute call is made without a GPU;
the product must fail loudly instead of falling back to a CPU path."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "wfacuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(wfacuda_[a-z_]+)\s*\(", hdr)))


def test_header_symbols_exported(built_lib):
    from wfa_b200 import api
    names = _declared()
    assert set(names) == set(api.EXPORTS), (names, api.EXPORTS)
    lib = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), n


def test_header_is_plain_c(tmp_path):
    """include/wfacuda.h is what a cgo / FFI binding includes: it must compile as C99 on its own."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "wfacuda.h"\nint main(void) { wfacuda_result r; wfacuda_wavefront w; wfacuda_config c; (void)r; (void)w; (void)c; return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                           "-I" + os.path.join(ROOT, "include"), str(src)])


def test_built_for_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_no_gpu_means_loud_failure(built_lib):
    import torch
    from wfa_b200 import api
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert api.device_count() == 0
    with pytest.raises(api.WfaError, match="no CPU fallback"):
        api.New()


def test_product_never_touches_oracle():
    """wfa_b200/ must not import, link or load anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wfa_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".hpp", ".go")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
    deps = subprocess.run(["ldd", os.path.join(ROOT, "wfa_b200", "libwfacuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_invalid_config_rejected(built_lib):
    """create() validates before touching CUDA, so this is checkable without a GPU."""
    from wfa_b200 import api
    for pen in ((0, 6, 2), (4, 6, 0)):
        with pytest.raises(api.WfaError, match="must be > 0"):
            api.New(api.Penalties(*pen), api.Options(True))
