"""urnn_b200 -- B200-native kernels for the U-RNN ConvGRU encoder-decoder time step.

Layout: csrc/ (CUDA + C ABI, built into urnn_b200/liburnn_b200.so), urnn_b200/ (ctypes binding, torch-facing
ops, sequence runner), src/lib/model/networks/ (mirror of the reference's module interface).
"""
from . import _capi  # noqa: F401
from .ops import set_default_math, get_default_math  # noqa: F401

__all__ = ["set_default_math", "get_default_math"]
