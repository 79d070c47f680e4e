"""ctypes binding of librstsr_cuda.so -- the same stub a Rust `extern "C"` block would declare
(see INTEGRATION.md).  No torch types cross this boundary: device pointers are plain integers.

The library is REQUIRED: there is no CPU fallback.  If the shared object is missing, or a compute entry point
is called on a machine without a CUDA device, the call fails loudly (RstsrCudaError / OSError).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint8, c_uint64, c_void_p

RC_MAX_NDIM = 16
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librstsr_cuda.so")

# status codes (include/rstsr_cuda.h)
STATUS_NAMES = {0: "Ok", 1: "ValueOutOfRange", 2: "InvalidValue", 3: "InvalidLayout", 4: "RuntimeError",
                5: "DeviceMismatch", 6: "UnImplemented", 7: "MemoryError", 8: "DeviceError", 9: "IndexError"}

# dtype codes
BOOL, I8, I16, I32, I64, U8, U16, U32, U64, F32, F64, F16, BF16, C32, C64 = range(15)
ROW_MAJOR, COL_MAJOR = 0, 1
ITER_C, ITER_F, ITER_A, ITER_K = range(4)

BINOPS = dict(add=0, sub=1, mul=2, div=3, rem=4, bitor=5, bitand=6, bitxor=7, shl=8, shr=9, maximum=10, minimum=11,
              floor_divide=12, pow=13, atan2=14, copysign=15, hypot=16, logaddexp=17, nextafter=18,
              eq=32, ne=33, lt=34, le=35, gt=36, ge=37)
UNOPS = dict(neg=0, not_=1, abs=2, square=3, sign=4, sqrt=5, exp=6, expm1=7, log=8, log2=9, log10=10, sin=11, cos=12,
             tan=13, asin=14, acos=15, atan=16, sinh=17, cosh=18, tanh=19, asinh=20, acosh=21, atanh=22, floor=23,
             ceil=24, round=25, trunc=26, reciprocal=27, conj=28, real=29, imag=30, isnan=48, isinf=49, isfinite=50,
             signbit=51,
             inv=27)  # OpInvAPI and OpReciprocalAPI are both b.recip() (auto_impl/op_binary_common.rs:25,29)
REDOPS = dict(sum=0, prod=1, max=2, min=3, mean=4, var=5, std=6, l2_norm=7, argmin=8, argmax=9, all=10, any=11,
              count_nonzero=12)


UPLO = {"U": 0, "L": 1}                              # rc_uplo / FlagUpLo
SYMM = {"Sy": 0, "He": 1, "Ay": 2, "Ah": 3, "N": 4}  # rc_symm / FlagSymm


class RstsrCudaError(RuntimeError):
    """Mirror of rstsr's `Error` (rstsr-common/src/error.rs:66-99): `.kind` is the RSTSRError variant name."""

    def __init__(self, status: int, message: str):
        self.status = status
        self.kind = STATUS_NAMES.get(status, f"status {status}")
        super().__init__(f"{self.kind}: {message}")


class CLayout(ctypes.Structure):
    """rc_layout"""
    _fields_ = [("ndim", c_int32), ("shape", c_int64 * RC_MAX_NDIM), ("stride", c_int64 * RC_MAX_NDIM),
                ("offset", c_int64)]


# every exported symbol of include/rstsr_cuda.h: name -> (restype, argtypes)
_P = c_void_p
_L = POINTER(CLayout)
SIGNATURES = {
    "rc_last_error": (c_char_p, []),
    "rc_version": (c_char_p, []),
    "rc_device_count": (c_int, [POINTER(c_int)]),
    "rc_device_create": (c_int, [c_int, c_int, POINTER(_P)]),
    "rc_device_create_on_stream": (c_int, [c_int, c_int, _P, POINTER(_P)]),
    "rc_device_destroy": (c_int, [_P]),
    "rc_device_default_order": (c_int, [_P, POINTER(c_int)]),
    "rc_device_set_default_order": (c_int, [_P, c_int]),
    "rc_device_same_device": (c_int, [_P, _P, POINTER(c_int)]),
    "rc_device_ordinal": (c_int, [_P, POINTER(c_int)]),
    "rc_device_stream": (c_int, [_P, POINTER(_P)]),
    "rc_device_synchronize": (c_int, [_P]),
    "rc_device_launch_count": (c_int, [_P, POINTER(c_uint64)]),
    "rc_malloc": (c_int, [_P, c_size_t, POINTER(_P)]),
    "rc_free": (c_int, [_P, _P]),
    "rc_memcpy_h2d": (c_int, [_P, _P, _P, c_size_t]),
    "rc_memcpy_d2h": (c_int, [_P, _P, _P, c_size_t]),
    "rc_memcpy_d2d": (c_int, [_P, _P, _P, c_size_t]),
    "rc_memset": (c_int, [_P, _P, c_int, c_size_t]),
    "rc_memcpy_h2d_async": (c_int, [_P, _P, _P, c_size_t]),
    "rc_memcpy_d2h_async": (c_int, [_P, _P, _P, c_size_t]),
    "rc_memcpy2d_d2h_async": (c_int, [_P, _P, c_size_t, _P, c_size_t, c_size_t, c_size_t]),
    "rc_memcpy2d_h2d_async": (c_int, [_P, _P, c_size_t, _P, c_size_t, c_size_t, c_size_t]),
    "rc_device_wait": (c_int, [_P, _P]),
    "rc_memcpy_peer": (c_int, [_P, _P, _P, _P, c_size_t]),
    "rc_get_index": (c_int, [_P, c_int, _P, c_int64, _P]),
    "rc_set_index": (c_int, [_P, c_int, _P, c_int64, _P]),
    "rc_host_alloc": (c_int, [c_size_t, POINTER(_P)]),
    "rc_host_free": (c_int, [_P]),
    "rc_dtype_size": (c_size_t, [c_int]),
    "rc_layout_check": (c_int, [_L]),
    "rc_layout_bounds_index": (c_int, [_L, POINTER(c_int64), POINTER(c_int64)]),
    "rc_layout_c_contig": (c_int, [_L, POINTER(c_int)]),
    "rc_layout_f_contig": (c_int, [_L, POINTER(c_int)]),
    "rc_layout_new_contig": (c_int, [POINTER(c_int64), c_int, c_int, c_int64, _L]),
    "rc_layout_broadcast": (c_int, [_L, _L, c_int, _L, _L]),
    "rc_layout_for_binary_op": (c_int, [_L, _L, c_int, _L]),
    "rc_layout_for_array_copy": (c_int, [_L, c_int, c_int, _L]),
    "rc_layout_for_reduce": (c_int, [_L, POINTER(c_int64), c_int, _L]),
    "rc_layout_reshapeable": (c_int, [_L, POINTER(c_int64), c_int, c_int, POINTER(c_int), _L]),
    "rc_layout_equal": (c_int, [_L, _L, POINTER(c_int)]),
    "rc_assign": (c_int, [_P, c_int, _P, _L, c_int, _P, _L]),
    "rc_assign_arbitary": (c_int, [_P, c_int, _P, _L, c_int, _P, _L]),
    "rc_fill": (c_int, [_P, c_int, _P, _L, c_int, _P]),
    "rc_arange": (c_int, [_P, c_int, _P, _P, _P, POINTER(_P), POINTER(c_int64)]),
    "rc_linspace": (c_int, [_P, c_int, _P, _P, c_int64, c_int, POINTER(_P)]),
    "rc_tril": (c_int, [_P, c_int, _P, _L, c_int64]),
    "rc_triu": (c_int, [_P, c_int, _P, _L, c_int64]),
    "rc_op_mutc_refa_refb": (c_int, [_P, c_int, c_int, _P, _L, _P, _L, _P, _L]),
    "rc_op_mutc_refa_numb": (c_int, [_P, c_int, c_int, _P, _L, _P, _L, _P]),
    "rc_op_mutc_numa_refb": (c_int, [_P, c_int, c_int, _P, _L, _P, _P, _L]),
    "rc_op_muta_refb": (c_int, [_P, c_int, c_int, _P, _L, _P, _L, c_int]),
    "rc_op_muta_numb": (c_int, [_P, c_int, c_int, _P, _L, _P, c_int]),
    "rc_unary_muta_refb": (c_int, [_P, c_int, c_int, _P, _L, _P, _L]),
    "rc_unary_muta": (c_int, [_P, c_int, c_int, _P, _L]),
    "rc_binop_out_dtype": (c_int, [c_int, c_int, POINTER(c_int)]),
    "rc_unop_out_dtype": (c_int, [c_int, c_int, POINTER(c_int)]),
    "rc_redop_out_dtype": (c_int, [c_int, c_int, POINTER(c_int)]),
    "rc_reduce_all": (c_int, [_P, c_int, c_int, _P, _L, _P]),
    "rc_reduce_all_device": (c_int, [_P, c_int, c_int, _P, _L, _P]),
    "rc_reduce_axes": (c_int, [_P, c_int, c_int, _P, _L, POINTER(c_int64), c_int, POINTER(_P), _L]),
    "rc_reduce_axes_into": (c_int, [_P, c_int, c_int, _P, _L, POINTER(c_int64), c_int, _P, _L]),
    "rc_vecdot": (c_int, [_P, c_int, _P, _L, _P, _L, _P, _L, POINTER(c_int64), POINTER(c_int64), c_int]),
    "rc_allclose_all": (c_int, [_P, c_int, _P, _L, _P, _L, c_double, c_double, c_int, POINTER(c_int)]),
    "rc_index_select": (c_int, [_P, c_int, _P, _L, _P, _L, c_int, POINTER(c_int64), c_int64]),
    "rc_pack_tri": (c_int, [_P, c_int, _P, _L, _P, _L, c_int]),
    "rc_unpack_tri": (c_int, [_P, c_int, _P, _L, _P, _L, c_int, c_int]),
    "rc_comm_get_unique_id": (c_int, [POINTER(c_uint8)]),
    "rc_comm_init_rank": (c_int, [_P, c_int, c_int, POINTER(c_uint8), POINTER(_P)]),
    "rc_comm_destroy": (c_int, [_P]),
    "rc_comm_all_reduce": (c_int, [_P, c_int, c_int, _P, c_size_t]),
    "rc_reduce_all_sharded": (c_int, [_P, _P, c_int, c_int, _P, _L, c_int64, _P]),
    "rc_comm_set_peer_window": (c_int, [_P, c_int]),
    "rc_comm_info": (c_int, [_P, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "rc_reduce_axes_sharded": (c_int, [_P, _P, c_int, c_int, _P, _L, POINTER(c_int64), c_int, c_int64, _P, _L]),
    "rc_device_numa_node": (c_int, [_P, POINTER(c_int)]),
    "rc_host_alloc_on_node": (c_int, [c_size_t, c_int, POINTER(_P), POINTER(c_int)]),
    "rc_assign_arbitary_order": (c_int, [_P, c_int, c_int, _P, _L, c_int, _P, _L]),
    "rc_dtype_promote": (c_int, [c_int, c_int, POINTER(c_int)]),
    "rc_binop_out_dtype_ex": (c_int, [c_int, c_int, c_int, POINTER(c_int)]),
    "rc_op_mutc_refa_refb_ex": (c_int, [_P, c_int, c_int, _P, _L, c_int, _P, _L, c_int, _P, _L]),
    "rc_op_mutc_refa_numb_ex": (c_int, [_P, c_int, c_int, _P, _L, c_int, _P, _L, c_int, _P]),
    "rc_op_mutc_numa_refb_ex": (c_int, [_P, c_int, c_int, _P, _L, c_int, _P, c_int, _P, _L]),
    "rc_reduce_unraveled_arg_all": (c_int, [_P, c_int, c_int, _P, _L, POINTER(c_int64)]),
    "rc_reduce_unraveled_arg_axes": (c_int, [_P, c_int, c_int, _P, _L, POINTER(c_int64), c_int, POINTER(_P), _L]),
    "rc_isclose": (c_int, [_P, c_int, _P, _L, _P, _L, _P, _L, c_double, c_double, c_int]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load librstsr_cuda.so (once) and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C rstsr_b200/csrc`.  There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().rc_last_error()
        raise RstsrCudaError(status, msg.decode() if msg else "")


__all__ = ["lib", "check", "CLayout", "RstsrCudaError", "SIGNATURES", "LIB_PATH", "byref"]
