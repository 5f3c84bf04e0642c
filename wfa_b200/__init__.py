"""wfa_b200 -- B200-native wavefront alignment behind the shenwei356/wfa API.

The hot path (next / extend / reduce / backtrace) is hand-written CUDA for
sm_100a in csrc/, shipped as libwfacuda.so with the C ABI of include/wfacuda.h.
`wfa_b200.api` mirrors the reference's Go API on top of that C ABI; there is no
CPU fallback -- importing the API without the built library raises.
"""
from . import datagen  # noqa: F401

__all__ = ["datagen"]
