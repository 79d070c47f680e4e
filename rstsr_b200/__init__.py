"""rstsr_b200 -- a B200-native `DeviceCuda` backend for the rstsr tensor toolkit.

The product is `lib/librstsr_cuda.so` (hand-written CUDA for sm_100a behind the C ABI of
include/rstsr_cuda.h); this package is the thin host-side mirror of the reference's device-trait interface.
Importing it requires the built library: there is no CPU fallback.
"""
from . import _ffi
from ._ffi import RstsrCudaError
from .device import (bfloat16, Comm, CudaRaw, DeviceCuda, Layout, broadcast_layout, broadcast_shapes, layout_for_array_copy, layout_for_binary_op,
                     layout_for_reduce, layout_reshapeable)
from .tensor import (COL_MAJOR, ROW_MAJOR, Tensor, allclose, isclose, arange, asarray, atleast_1d, atleast_2d, concat, concatenate,
                     diag, empty, eye, full, hstack, meshgrid, ones, stack, unstack, vecdot, vstack, zeros)

_ffi.lib()  # fail loudly at import time if the extension is missing

__all__ = ["bfloat16", "DeviceCuda", "CudaRaw", "Layout", "Tensor", "Comm", "RstsrCudaError", "ROW_MAJOR", "COL_MAJOR", "asarray",
           "arange", "zeros", "ones", "eye", "full", "empty", "vecdot", "allclose", "isclose", "concat", "concatenate", "stack", "hstack", "vstack", "unstack", "diag", "meshgrid", "atleast_1d",
           "atleast_2d", "broadcast_layout", "broadcast_shapes", "layout_for_array_copy", "layout_for_binary_op",
           "layout_for_reduce", "layout_reshapeable"]
