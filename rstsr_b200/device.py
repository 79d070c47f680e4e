"""`DeviceCuda`: host-side mirror of the reference's device-trait surface over the C ABI.

Method names and argument order follow the reference traits (paths inside RESTGroup/rstsr v0.7.10) so the parity
tests read like the reference's own:

    DeviceBaseAPI / DeviceStorageAPI / DeviceCreation*API   rstsr-core/src/storage/{device,creation}.rs
    OpAssignAPI, OpAssignArbitaryAPI                        rstsr-core/src/operators/assignment.rs:5-53
    Op{Add..Shr}API, Op*AssignAPI, OpL/RConsume*API          rstsr-core/src/operators/ops/op_{ternary,binary}_arithmetic.rs
    unary / binary-function traits                          rstsr-core/src/operators/ops/op_{binary,ternary}_common.rs
    Op{Sum,Prod,Max,Min,Mean}API                            rstsr-core/src/operators/reduction.rs:3-33

Everything numerical happens in librstsr_cuda.so (hand-written CUDA for sm_100a); this file only marshals
arguments.  It never touches `oracle/` and has no CPU fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from ._ffi import CLayout, RstsrCudaError, byref, check

_NP_TO_DT = {np.dtype(np.bool_): _ffi.BOOL, np.dtype(np.int8): _ffi.I8, np.dtype(np.int16): _ffi.I16,
             np.dtype(np.int32): _ffi.I32, np.dtype(np.int64): _ffi.I64, np.dtype(np.uint8): _ffi.U8,
             np.dtype(np.uint16): _ffi.U16, np.dtype(np.uint32): _ffi.U32, np.dtype(np.uint64): _ffi.U64,
             np.dtype(np.float32): _ffi.F32, np.dtype(np.float64): _ffi.F64,
             # round 2: half and complex element types (half::f16, half::bf16, num::Complex<f32|f64>)
             np.dtype(np.float16): _ffi.F16, np.dtype(np.complex64): _ffi.C32, np.dtype(np.complex128): _ffi.C64}
try:  # bfloat16 has no NumPy dtype of its own; ml_dtypes provides one with the same f32-compute-round-once arithmetic
    import ml_dtypes as _ml_dtypes
    _NP_TO_DT[np.dtype(_ml_dtypes.bfloat16)] = _ffi.BF16
    bfloat16 = _ml_dtypes.bfloat16
except ImportError:  # pragma: no cover
    bfloat16 = None
_DT_TO_NP = {v: k for k, v in _NP_TO_DT.items()}


def dtype_code(dt) -> int:
    return _NP_TO_DT[np.dtype(dt)]


def dtype_np(code: int) -> np.dtype:
    return _DT_TO_NP[code]


@dataclass(frozen=True)
class Layout:
    """Layout<IxD> (rstsr-common/src/layout/layoutbase.rs:15-23): element strides, element offset."""
    shape: Tuple[int, ...]
    stride: Tuple[int, ...]
    offset: int = 0

    def __post_init__(self):
        object.__setattr__(self, "shape", tuple(int(x) for x in self.shape))
        object.__setattr__(self, "stride", tuple(int(x) for x in self.stride))
        object.__setattr__(self, "offset", int(self.offset))

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        n = 1
        for d in self.shape:
            n *= d
        return n

    def to_c(self) -> CLayout:
        if len(self.shape) != len(self.stride) or len(self.shape) > _ffi.RC_MAX_NDIM:
            raise RstsrCudaError(3, "invalid layout rank")
        c = CLayout()
        c.ndim = len(self.shape)
        for i, (d, s) in enumerate(zip(self.shape, self.stride)):
            c.shape[i] = d
            c.stride[i] = s
        c.offset = self.offset
        return c

    @staticmethod
    def from_c(c: CLayout) -> "Layout":
        n = c.ndim
        return Layout(tuple(c.shape[i] for i in range(n)), tuple(c.stride[i] for i in range(n)), c.offset)

    # ---- host-side layout algebra, evaluated by the library (reference-identical by contract) ----
    def bounds_index(self) -> Tuple[int, int]:
        lo, hi = ctypes.c_int64(), ctypes.c_int64()
        check(_ffi.lib().rc_layout_bounds_index(byref(self.to_c()), byref(lo), byref(hi)))
        return lo.value, hi.value

    def check(self) -> "Layout":
        check(_ffi.lib().rc_layout_check(byref(self.to_c())))
        return self

    def c_contig(self) -> bool:
        out = ctypes.c_int()
        check(_ffi.lib().rc_layout_c_contig(byref(self.to_c()), byref(out)))
        return bool(out.value)

    def f_contig(self) -> bool:
        out = ctypes.c_int()
        check(_ffi.lib().rc_layout_f_contig(byref(self.to_c()), byref(out)))
        return bool(out.value)

    def same_as(self, other: "Layout") -> bool:
        out = ctypes.c_int()
        check(_ffi.lib().rc_layout_equal(byref(self.to_c()), byref(other.to_c()), byref(out)))
        return bool(out.value)

    @staticmethod
    def contig(shape: Sequence[int], order: int, offset: int = 0) -> "Layout":
        arr = (ctypes.c_int64 * max(len(shape), 1))(*[int(x) for x in shape])
        out = CLayout()
        check(_ffi.lib().rc_layout_new_contig(arr, len(shape), order, offset, byref(out)))
        return Layout.from_c(out)


def broadcast_layout(la: Layout, lb: Layout, order: int) -> Tuple[Layout, Layout]:
    oa, ob = CLayout(), CLayout()
    check(_ffi.lib().rc_layout_broadcast(byref(la.to_c()), byref(lb.to_c()), order, byref(oa), byref(ob)))
    return Layout.from_c(oa), Layout.from_c(ob)


def broadcast_shapes(shapes: Sequence[Sequence[int]], order: int = _ffi.ROW_MAJOR) -> Tuple[int, ...]:
    """rt::broadcast_shapes (rstsr-core/src/tensor/manipulation/broadcast.rs; rule of rstsr-common/src/layout/
    broadcast.rs:21-68): fold the shapes pairwise through the library's broadcast (right-aligned for RowMajor,
    left-aligned for ColMajor)."""
    out = Layout.contig((), order)
    for shape in shapes:
        out, _ = broadcast_layout(out, Layout.contig(tuple(int(d) for d in shape), order), order)
        out = Layout.contig(out.shape, order)
    return out.shape


def layout_for_binary_op(la: Layout, lb: Layout, order: int) -> Layout:
    out = CLayout()
    check(_ffi.lib().rc_layout_for_binary_op(byref(la.to_c()), byref(lb.to_c()), order, byref(out)))
    return Layout.from_c(out)


def layout_for_array_copy(la: Layout, iter_order: int = _ffi.ITER_K, default_order: int = _ffi.ROW_MAJOR) -> Layout:
    out = CLayout()
    check(_ffi.lib().rc_layout_for_array_copy(byref(la.to_c()), iter_order, default_order, byref(out)))
    return Layout.from_c(out)


def layout_for_reduce(la: Layout, axes: Sequence[int]) -> Layout:
    arr = (ctypes.c_int64 * max(len(axes), 1))(*[int(x) for x in axes])
    out = CLayout()
    check(_ffi.lib().rc_layout_for_reduce(byref(la.to_c()), arr, len(axes), byref(out)))
    return Layout.from_c(out)


def layout_reshapeable(la: Layout, shape: Sequence[int], order: int) -> Optional[Layout]:
    arr = (ctypes.c_int64 * max(len(shape), 1))(*[int(x) for x in shape])
    ok = ctypes.c_int()
    out = CLayout()
    check(_ffi.lib().rc_layout_reshapeable(byref(la.to_c()), arr, len(shape), order, byref(ok), byref(out)))
    return Layout.from_c(out) if ok.value else None


class CudaRaw:
    """Device buffer: the `Raw` of DeviceRawAPI (Vec<T> on the CPU devices).  Drop = free."""

    def __init__(self, device: "DeviceCuda", ptr: int, length: int, dtype, owned: bool = True):
        self.device = device
        self.ptr = int(ptr)
        self.len = int(length)
        self.dtype = np.dtype(dtype)
        self._owned = owned

    @property
    def nbytes(self) -> int:
        return self.len * self.dtype.itemsize

    def free(self):
        if self._owned and self.ptr and self.device._handle:
            _ffi.lib().rc_free(self.device._handle, self.ptr)
        self.ptr = 0
        self._owned = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def clone(self) -> "CudaRaw":
        out = self.device.uninit_impl(self.dtype, self.len)
        check(_ffi.lib().rc_memcpy_d2d(self.device._handle, out.ptr, self.ptr, self.nbytes))
        return out


class DeviceCuda:
    """A CUDA device: {ordinal, default_order, stream} (cf. DeviceFaer: {pool, default_order},
    rstsr-core/src/device_faer/device.rs:5-60)."""

    def __init__(self, ordinal: int = 0, default_order: int = _ffi.ROW_MAJOR, stream: Optional[int] = None):
        self._handle = None
        h = ctypes.c_void_p()
        if stream is None:
            check(_ffi.lib().rc_device_create(ordinal, default_order, byref(h)))
        else:
            check(_ffi.lib().rc_device_create_on_stream(ordinal, default_order, ctypes.c_void_p(stream), byref(h)))
        self._handle = h
        self.ordinal = ordinal

    @staticmethod
    def device_count() -> int:
        n = ctypes.c_int(0)
        check(_ffi.lib().rc_device_count(byref(n)))
        return n.value

    def close(self):
        if self._handle:
            _ffi.lib().rc_device_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- DeviceBaseAPI ----
    def default_order(self) -> int:
        out = ctypes.c_int()
        check(_ffi.lib().rc_device_default_order(self._handle, byref(out)))
        return out.value

    def set_default_order(self, order: int):
        check(_ffi.lib().rc_device_set_default_order(self._handle, order))

    def same_device(self, other: "DeviceCuda") -> bool:
        out = ctypes.c_int()
        check(_ffi.lib().rc_device_same_device(self._handle, other._handle, byref(out)))
        return bool(out.value)

    def synchronize(self):
        check(_ffi.lib().rc_device_synchronize(self._handle))

    def launch_count(self) -> int:
        out = ctypes.c_uint64()
        check(_ffi.lib().rc_device_launch_count(self._handle, byref(out)))
        return out.value

    def stream(self) -> int:
        out = ctypes.c_void_p()
        check(_ffi.lib().rc_device_stream(self._handle, byref(out)))
        return out.value or 0

    # ---- DeviceCreationAnyAPI / NumAPI / DeviceStorageAPI ----
    def uninit_impl(self, dtype, length: int) -> CudaRaw:
        p = ctypes.c_void_p()
        dt = np.dtype(dtype)
        check(_ffi.lib().rc_malloc(self._handle, int(length) * dt.itemsize, byref(p)))
        return CudaRaw(self, p.value, length, dt)

    empty_impl = uninit_impl

    def zeros_impl(self, dtype, length: int) -> CudaRaw:
        raw = self.uninit_impl(dtype, length)
        check(_ffi.lib().rc_memset(self._handle, raw.ptr, 0, raw.nbytes))
        return raw

    def full_impl(self, dtype, length: int, value) -> CudaRaw:
        raw = self.uninit_impl(dtype, length)
        self.fill(raw, Layout((length,), (1,), 0), value)
        return raw

    def ones_impl(self, dtype, length: int) -> CudaRaw:
        return self.full_impl(dtype, length, 1)

    # ---- DeviceCreationArangeAPI / linspace / TriAPI (storage/creation.rs:41-62) ----
    def arange_impl(self, start, end, step, dtype) -> CudaRaw:
        dt = np.dtype(dtype)
        s, e, st = (np.array([v], dtype=dt) for v in (start, end, step))
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        check(_ffi.lib().rc_arange(self._handle, dtype_code(dt), s.ctypes.data, e.ctypes.data, st.ctypes.data, byref(p),
                                   byref(n)))
        return CudaRaw(self, p.value, n.value, dt)

    def linspace_impl(self, start, end, n: int, endpoint: bool, dtype) -> CudaRaw:
        dt = np.dtype(dtype)
        s, e = np.array([start], dtype=dt), np.array([end], dtype=dt)
        p = ctypes.c_void_p()
        check(_ffi.lib().rc_linspace(self._handle, dtype_code(dt), s.ctypes.data, e.ctypes.data, int(n),
                                     1 if endpoint else 0, byref(p)))
        return CudaRaw(self, p.value, int(n), dt)

    def tril_impl(self, raw: CudaRaw, layout: Layout, k: int = 0):
        check(_ffi.lib().rc_tril(self._handle, dtype_code(raw.dtype), raw.ptr, byref(layout.to_c()), int(k)))

    def triu_impl(self, raw: CudaRaw, layout: Layout, k: int = 0):
        check(_ffi.lib().rc_triu(self._handle, dtype_code(raw.dtype), raw.ptr, byref(layout.to_c()), int(k)))

    def outof_cpu_vec(self, vec: np.ndarray) -> CudaRaw:
        vec = np.ascontiguousarray(vec).reshape(-1)
        raw = self.uninit_impl(vec.dtype, vec.size)
        check(_ffi.lib().rc_memcpy_h2d(self._handle, raw.ptr, vec.ctypes.data, vec.nbytes))
        self.synchronize()  # `vec` may be a temporary
        return raw

    from_cpu_vec = outof_cpu_vec

    def to_cpu_vec(self, raw: CudaRaw) -> np.ndarray:
        out = np.empty(raw.len, dtype=raw.dtype)
        check(_ffi.lib().rc_memcpy_d2h(self._handle, out.ctypes.data, raw.ptr, raw.nbytes))
        return out

    # ---- stream-ordered staging (pinned host memory); see rc_memcpy_*_async in the header ----
    def wait(self, other: "DeviceCuda"):
        """Work enqueued on this handle from now on runs after what `other` has enqueued so far."""
        check(_ffi.lib().rc_device_wait(self._handle, other._handle))

    def change_device(self, raw: CudaRaw, target: "DeviceCuda") -> CudaRaw:
        """DeviceChangeAPI::change_device (rstsr-core/src/storage/conversion.rs:3-21) from this handle to `target`
        (another stream of the same GPU, or another GPU: peer copy).  Returns new storage owned by `target`."""
        if raw.device is not self and not raw.device.same_device(self):
            raise _ffi.RstsrCudaError(5, "storage does not live on this device")
        out = target.uninit_impl(raw.dtype, raw.len)
        nbytes = raw.len * np.dtype(raw.dtype).itemsize
        check(_ffi.lib().rc_memcpy_peer(target._handle, out.ptr, self._handle, raw.ptr, nbytes))
        return out

    def h2d_async(self, raw: CudaRaw, host_ptr: int, nbytes: int, dst_byte_offset: int = 0):
        check(_ffi.lib().rc_memcpy_h2d_async(self._handle, raw.ptr + dst_byte_offset, host_ptr, nbytes))

    def d2h_async(self, host_ptr: int, raw: CudaRaw, nbytes: int, src_byte_offset: int = 0):
        check(_ffi.lib().rc_memcpy_d2h_async(self._handle, host_ptr, raw.ptr + src_byte_offset, nbytes))

    def d2h_2d_async(self, host_ptr: int, host_pitch: int, raw: CudaRaw, src_byte_offset: int, src_pitch: int,
                     width_bytes: int, height: int):
        check(_ffi.lib().rc_memcpy2d_d2h_async(self._handle, host_ptr, host_pitch, raw.ptr + src_byte_offset, src_pitch,
                                               width_bytes, height))

    def numa_node(self) -> int:
        """NUMA node of the GPU's PCIe root (-1 if the platform does not say)."""
        out = ctypes.c_int(-1)
        check(_ffi.lib().rc_device_numa_node(self._handle, byref(out)))
        return out.value

    @staticmethod
    def host_alloc(nbytes: int, node: int = -1) -> Tuple[int, bool]:
        """Pinned staging buffer, optionally bound to a NUMA node: (address, bound?).  Free with host_free."""
        p, bound = ctypes.c_void_p(), ctypes.c_int(0)
        check(_ffi.lib().rc_host_alloc_on_node(int(nbytes), int(node), byref(p), byref(bound)))
        return p.value, bool(bound.value)

    @staticmethod
    def host_free(ptr: int):
        check(_ffi.lib().rc_host_free(ctypes.c_void_p(ptr)))

    def wrap(self, ptr: int, length: int, dtype) -> CudaRaw:
        """Borrow device memory owned by someone else (e.g. a torch tensor's data_ptr())."""
        return CudaRaw(self, ptr, length, dtype, owned=False)

    def get_index(self, raw: CudaRaw, index: int):
        out = np.empty(1, dtype=raw.dtype)
        check(_ffi.lib().rc_get_index(self._handle, dtype_code(raw.dtype), raw.ptr, index, out.ctypes.data))
        return out[0]

    def set_index(self, raw: CudaRaw, index: int, value):
        v = np.array([value], dtype=raw.dtype)
        check(_ffi.lib().rc_set_index(self._handle, dtype_code(raw.dtype), raw.ptr, index, v.ctypes.data))

    # ---- OpAssignAPI / OpAssignArbitaryAPI ----
    def assign(self, c: CudaRaw, lc: Layout, a: CudaRaw, la: Layout):
        check(_ffi.lib().rc_assign(self._handle, dtype_code(c.dtype), c.ptr, byref(lc.to_c()), dtype_code(a.dtype), a.ptr,
                                   byref(la.to_c())))

    assign_uninit = assign

    def assign_arbitary(self, c: CudaRaw, lc: Layout, a: CudaRaw, la: Layout, order: Optional[int] = None):
        """`order`: pairing order of the flattened elements; None = the handle's default order.  Passing it explicitly
        (rc_assign_arbitary_order) keeps a handle that several threads share untouched."""
        if order is None:
            check(_ffi.lib().rc_assign_arbitary(self._handle, dtype_code(c.dtype), c.ptr, byref(lc.to_c()),
                                                dtype_code(a.dtype), a.ptr, byref(la.to_c())))
        else:
            check(_ffi.lib().rc_assign_arbitary_order(self._handle, int(order), dtype_code(c.dtype), c.ptr,
                                                      byref(lc.to_c()), dtype_code(a.dtype), a.ptr, byref(la.to_c())))

    assign_arbitary_uninit = assign_arbitary

    def fill(self, c: CudaRaw, lc: Layout, value):
        v = np.array([value])
        if v.dtype not in _NP_TO_DT or c.dtype.kind == "c" or c.dtype.itemsize == 2 and c.dtype.kind not in "iu":
            v = v.astype(c.dtype)  # half / complex targets: the host scalar arrives in the target type
        check(_ffi.lib().rc_fill(self._handle, dtype_code(c.dtype), c.ptr, byref(lc.to_c()), dtype_code(v.dtype),
                                 v.ctypes.data))

    # ---- elementwise ----
    @staticmethod
    def _scalar(value, dtype) -> np.ndarray:
        return np.array([value]).astype(dtype)

    def op_mutc_refa_refb(self, op: str, c: CudaRaw, lc: Layout, a: CudaRaw, la: Layout, b: CudaRaw, lb: Layout):
        """c = a op b.  Operands of one dtype run the single fused kernel; mixed dtypes follow the reference's promotion
        rules (rc_op_mutc_refa_refb_ex: DTypePromoteAPI / DTypeIntoFloatAPI).  c's dtype must be the op's output type
        (the reference fixes it through `TOut`): anything else is a DTypeMismatch instead of reinterpreted bytes."""
        want = self.binop_out_dtype_ex(op, a.dtype, b.dtype)
        if c.dtype != want:
            raise RstsrCudaError(6, f"DTypeMismatch: {op}({a.dtype}, {b.dtype}) writes {want}, output storage is {c.dtype}")
        # rc_op_mutc_refa_refb_ex runs the single fused kernel of rc_op_mutc_refa_refb when no operand needs a cast
        check(_ffi.lib().rc_op_mutc_refa_refb_ex(self._handle, _ffi.BINOPS[op], dtype_code(c.dtype), c.ptr,
                                                 byref(lc.to_c()), dtype_code(a.dtype), a.ptr, byref(la.to_c()),
                                                 dtype_code(b.dtype), b.ptr, byref(lb.to_c())))

    def op_mutc_refa_numb(self, op: str, c: CudaRaw, lc: Layout, a: CudaRaw, la: Layout, b, b_dtype=None):
        """b is a host scalar of a's dtype, or of `b_dtype` (then the pair is promoted like two tensors)."""
        bt = a.dtype if b_dtype is None else np.dtype(b_dtype)
        want = self.binop_out_dtype_ex(op, a.dtype, bt)
        if c.dtype != want:
            raise RstsrCudaError(6, f"DTypeMismatch: {op}({a.dtype}, {bt}) writes {want}, output storage is {c.dtype}")
        s = self._scalar(b, bt)
        check(_ffi.lib().rc_op_mutc_refa_numb_ex(self._handle, _ffi.BINOPS[op], dtype_code(c.dtype), c.ptr,
                                                 byref(lc.to_c()), dtype_code(a.dtype), a.ptr, byref(la.to_c()),
                                                 dtype_code(bt), s.ctypes.data))

    def op_mutc_numa_refb(self, op: str, c: CudaRaw, lc: Layout, a, b: CudaRaw, lb: Layout, a_dtype=None):
        at = b.dtype if a_dtype is None else np.dtype(a_dtype)
        want = self.binop_out_dtype_ex(op, at, b.dtype)
        if c.dtype != want:
            raise RstsrCudaError(6, f"DTypeMismatch: {op}({at}, {b.dtype}) writes {want}, output storage is {c.dtype}")
        s = self._scalar(a, at)
        check(_ffi.lib().rc_op_mutc_numa_refb_ex(self._handle, _ffi.BINOPS[op], dtype_code(c.dtype), c.ptr,
                                                 byref(lc.to_c()), dtype_code(at), s.ctypes.data, dtype_code(b.dtype),
                                                 b.ptr, byref(lb.to_c())))

    def op_muta_refb(self, op: str, a: CudaRaw, la: Layout, b: CudaRaw, lb: Layout, reverse: bool = False):
        """a = a o b (Op*AssignAPI / OpLConsume*API); reverse: a = b o a (OpRConsume*API)."""
        if a.dtype != b.dtype:  # `TA: AddAssign<TB>` etc.: the reference has no mixed-type in-place op
            raise RstsrCudaError(6, f"DTypeMismatch: in-place {op} needs operands of one dtype ({a.dtype} vs {b.dtype})")
        check(_ffi.lib().rc_op_muta_refb(self._handle, _ffi.BINOPS[op], dtype_code(a.dtype), a.ptr, byref(la.to_c()),
                                         b.ptr, byref(lb.to_c()), 1 if reverse else 0))

    def op_muta_numb(self, op: str, a: CudaRaw, la: Layout, b, reverse: bool = False):
        s = self._scalar(b, a.dtype)
        check(_ffi.lib().rc_op_muta_numb(self._handle, _ffi.BINOPS[op], dtype_code(a.dtype), a.ptr, byref(la.to_c()),
                                         s.ctypes.data, 1 if reverse else 0))

    def unary_muta_refb(self, op: str, a: CudaRaw, la: Layout, b: CudaRaw, lb: Layout):
        check(_ffi.lib().rc_unary_muta_refb(self._handle, _ffi.UNOPS[op], dtype_code(b.dtype), a.ptr, byref(la.to_c()),
                                            b.ptr, byref(lb.to_c())))

    def unary_muta(self, op: str, a: CudaRaw, la: Layout):
        check(_ffi.lib().rc_unary_muta(self._handle, _ffi.UNOPS[op], dtype_code(a.dtype), a.ptr, byref(la.to_c())))

    @staticmethod
    def binop_out_dtype(op: str, dtype) -> np.dtype:
        out = ctypes.c_int()
        check(_ffi.lib().rc_binop_out_dtype(_ffi.BINOPS[op], dtype_code(dtype), byref(out)))
        return dtype_np(out.value)

    @staticmethod
    def binop_out_dtype_ex(op: str, ta, tb) -> np.dtype:
        """TOut of the op for operand types (ta, tb): rc_binop_out_dtype_ex (promotion rules of the reference)."""
        out = ctypes.c_int()
        check(_ffi.lib().rc_binop_out_dtype_ex(_ffi.BINOPS[op], dtype_code(ta), dtype_code(tb), byref(out)))
        return dtype_np(out.value)

    @staticmethod
    def promote_types(ta, tb) -> np.dtype:
        """<TA as DTypePromoteAPI<TB>>::Res (rstsr-dtype-traits/src/promotion.rs)."""
        out = ctypes.c_int()
        check(_ffi.lib().rc_dtype_promote(dtype_code(ta), dtype_code(tb), byref(out)))
        return dtype_np(out.value)

    def isclose(self, c: CudaRaw, lc: Layout, a: CudaRaw, la: Layout, b: CudaRaw, lb: Layout, rtol: float = 1.0e-5,
                atol: float = 1.0e-8, equal_nan: bool = False):
        """OpIsCloseAPI::op_mutc_refa_refb (rstsr-core/src/operators/ops/op_ternary_common.rs:67-82), TE = f64."""
        if a.dtype != b.dtype:
            raise RstsrCudaError(6, "DTypeMismatch: isclose takes operands of one dtype")
        if c.dtype != np.dtype(np.bool_):
            raise RstsrCudaError(6, "DTypeMismatch: isclose writes bool")
        check(_ffi.lib().rc_isclose(self._handle, dtype_code(a.dtype), c.ptr, byref(lc.to_c()), a.ptr, byref(la.to_c()),
                                    b.ptr, byref(lb.to_c()), float(rtol), float(atol), int(bool(equal_nan))))

    @staticmethod
    def unop_out_dtype(op: str, dtype) -> np.dtype:
        out = ctypes.c_int()
        check(_ffi.lib().rc_unop_out_dtype(_ffi.UNOPS[op], dtype_code(dtype), byref(out)))
        return dtype_np(out.value)

    # ---- reductions ----
    @staticmethod
    def redop_out_dtype(op: str, dtype) -> np.dtype:
        out = ctypes.c_int()
        check(_ffi.lib().rc_redop_out_dtype(_ffi.REDOPS[op], dtype_code(dtype), byref(out)))
        return dtype_np(out.value)

    def reduce_all(self, op: str, a: CudaRaw, la: Layout):
        out = np.empty(1, dtype=self.redop_out_dtype(op, a.dtype))
        check(_ffi.lib().rc_reduce_all(self._handle, _ffi.REDOPS[op], dtype_code(a.dtype), a.ptr, byref(la.to_c()),
                                       out.ctypes.data))
        return out[0]

    def reduce_all_device(self, op: str, a: CudaRaw, la: Layout, out: CudaRaw):
        check(_ffi.lib().rc_reduce_all_device(self._handle, _ffi.REDOPS[op], dtype_code(a.dtype), a.ptr,
                                              byref(la.to_c()), out.ptr))

    def reduce_axes(self, op: str, a: CudaRaw, la: Layout, axes: Sequence[int]) -> Tuple[CudaRaw, Layout]:
        arr = (ctypes.c_int64 * max(len(axes), 1))(*[int(x) for x in axes])
        p = ctypes.c_void_p()
        lo = CLayout()
        check(_ffi.lib().rc_reduce_axes(self._handle, _ffi.REDOPS[op], dtype_code(a.dtype), a.ptr, byref(la.to_c()), arr,
                                        len(axes), byref(p), byref(lo)))
        layout = Layout.from_c(lo)
        return CudaRaw(self, p.value, max(layout.size, 1), self.redop_out_dtype(op, a.dtype)), layout

    def reduce_axes_into(self, op: str, a: CudaRaw, la: Layout, axes: Sequence[int], out: CudaRaw, lo: Layout):
        if out.dtype != self.redop_out_dtype(op, a.dtype):
            raise RstsrCudaError(6, f"DTypeMismatch: {op} of {a.dtype} writes {self.redop_out_dtype(op, a.dtype)}, "
                                    f"output storage is {out.dtype}")
        arr = (ctypes.c_int64 * max(len(axes), 1))(*[int(x) for x in axes])
        check(_ffi.lib().rc_reduce_axes_into(self._handle, _ffi.REDOPS[op], dtype_code(a.dtype), a.ptr, byref(la.to_c()),
                                             arr, len(axes), out.ptr, byref(lo.to_c())))

    def unraveled_arg_all(self, op: str, a: CudaRaw, la: Layout) -> Tuple[int, ...]:
        """OpUnraveledArgMin/MaxAPI::unraveled_arg*_all (operators/reduction.rs:35-55): index tuple within `la`."""
        out = (ctypes.c_int64 * max(la.ndim, 1))()
        check(_ffi.lib().rc_reduce_unraveled_arg_all(self._handle, _ffi.REDOPS[op], dtype_code(a.dtype), a.ptr,
                                                     byref(la.to_c()), out))
        return tuple(int(out[i]) for i in range(la.ndim))

    def unraveled_arg_axes(self, op: str, a: CudaRaw, la: Layout, axes: Sequence[int]) -> Tuple[CudaRaw, Layout]:
        """unraveled_arg*_axes: u64 tuples [lo.size][naxes] (position within the reduced axes, in the order given) and
        the layout `lo` of the output elements."""
        arr = (ctypes.c_int64 * max(len(axes), 1))(*[int(x) for x in axes])
        p = ctypes.c_void_p()
        lo = CLayout()
        check(_ffi.lib().rc_reduce_unraveled_arg_axes(self._handle, _ffi.REDOPS[op], dtype_code(a.dtype), a.ptr,
                                                      byref(la.to_c()), arr, len(axes), byref(p), byref(lo)))
        layout = Layout.from_c(lo)
        n = max(layout.bounds_index()[1], 1) * max(len(axes), 1)
        return CudaRaw(self, p.value, n, np.uint64), layout

    # ---- binary reductions ----
    def vecdot(self, c: CudaRaw, lc: Layout, a: CudaRaw, la: Layout, b: CudaRaw, lb: Layout, axes_a: Sequence[int],
               axes_b: Sequence[int]):
        """DeviceVecdotAPI::vecdot (rstsr-core/src/device_cpu_serial/linalg/vecdot.rs:16-28)."""
        if not (a.dtype == b.dtype == c.dtype):
            raise _ffi.RstsrCudaError(6, "vecdot: a, b and c must share one dtype")
        if len(axes_a) != len(axes_b):
            raise _ffi.RstsrCudaError(2, "axes_a and axes_b should have the same length")
        n = len(axes_a)
        aa = (ctypes.c_int64 * max(n, 1))(*[int(x) for x in axes_a])
        ab = (ctypes.c_int64 * max(n, 1))(*[int(x) for x in axes_b])
        check(_ffi.lib().rc_vecdot(self._handle, dtype_code(a.dtype), c.ptr, byref(lc.to_c()), a.ptr, byref(la.to_c()),
                                   b.ptr, byref(lb.to_c()), aa, ab, n))

    def allclose_all(self, a: CudaRaw, la: Layout, b: CudaRaw, lb: Layout, rtol: float = 1.0e-5, atol: float = 1.0e-8,
                     equal_nan: bool = False) -> bool:
        """OpAllCloseAPI::allclose_all (rstsr-core/src/device_cpu_serial/reduction.rs:660-683); la / lb already
        broadcast to one shape; defaults are IsCloseArgs::from(None) (rstsr-dtype-traits/src/isclose.rs:139-147)."""
        if a.dtype != b.dtype:
            raise _ffi.RstsrCudaError(6, "allclose_all: a and b must share one dtype")
        r = ctypes.c_int(0)
        check(_ffi.lib().rc_allclose_all(self._handle, dtype_code(a.dtype), a.ptr, byref(la.to_c()), b.ptr,
                                         byref(lb.to_c()), float(rtol), float(atol), int(bool(equal_nan)), byref(r)))
        return bool(r.value)

    # ---- index-driven movement ----
    def index_select(self, c: CudaRaw, lc: Layout, a: CudaRaw, la: Layout, axis: int, indices: Sequence[int]):
        """DeviceIndexSelectAPI::index_select (rstsr-core/src/device_cpu_serial/adv_indexing.rs:9-19); `indices`
        are non-negative host integers (usize in the trait)."""
        if a.dtype != c.dtype:
            raise _ffi.RstsrCudaError(6, "index_select: a and c must share one dtype")
        idx = np.ascontiguousarray(indices, dtype=np.int64).reshape(-1)
        arr = idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
        check(_ffi.lib().rc_index_select(self._handle, dtype_code(a.dtype), c.ptr, byref(lc.to_c()), a.ptr,
                                         byref(la.to_c()), int(axis), arr, idx.size))

    def pack_tri(self, a: CudaRaw, la: Layout, b: CudaRaw, lb: Layout, uplo: str):
        """OpPackTriAPI::pack_tri (rstsr-core/src/device_cpu_serial/operators/op_tri.rs:8-27): a = packed out."""
        if a.dtype != b.dtype:
            raise _ffi.RstsrCudaError(6, "pack_tri: a and b must share one dtype")
        check(_ffi.lib().rc_pack_tri(self._handle, dtype_code(a.dtype), a.ptr, byref(la.to_c()), b.ptr, byref(lb.to_c()),
                                     _ffi.UPLO[uplo]))

    def unpack_tri(self, a: CudaRaw, la: Layout, b: CudaRaw, lb: Layout, uplo: str, symm: str):
        """OpUnpackTriAPI::unpack_tri (rstsr-core/src/device_cpu_serial/operators/op_tri.rs:34-52): a = full out."""
        if a.dtype != b.dtype:
            raise _ffi.RstsrCudaError(6, "unpack_tri: a and b must share one dtype")
        check(_ffi.lib().rc_unpack_tri(self._handle, dtype_code(a.dtype), a.ptr, byref(la.to_c()), b.ptr,
                                       byref(lb.to_c()), _ffi.UPLO[uplo], _ffi.SYMM[symm]))

    # trait-named conveniences: sum_all / sum_axes / ... (operators/reduction.rs:26-32)
    def sum_all(self, a, la): return self.reduce_all("sum", a, la)
    def prod_all(self, a, la): return self.reduce_all("prod", a, la)
    def max_all(self, a, la): return self.reduce_all("max", a, la)
    def min_all(self, a, la): return self.reduce_all("min", a, la)
    def mean_all(self, a, la): return self.reduce_all("mean", a, la)
    def sum_axes(self, a, la, axes): return self.reduce_axes("sum", a, la, axes)
    def prod_axes(self, a, la, axes): return self.reduce_axes("prod", a, la, axes)
    def max_axes(self, a, la, axes): return self.reduce_axes("max", a, la, axes)
    def min_axes(self, a, la, axes): return self.reduce_axes("min", a, la, axes)
    def mean_axes(self, a, la, axes): return self.reduce_axes("mean", a, la, axes)


class Comm:
    """NCCL communicator for cross-shard reductions: one rank per process / GPU."""

    def __init__(self, device: DeviceCuda, nranks: int, rank: int, unique_id: bytes):
        self.device = device
        self.nranks, self.rank = nranks, rank
        h = ctypes.c_void_p()
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        check(_ffi.lib().rc_comm_init_rank(device._handle, nranks, rank, buf, byref(h)))
        self._handle = h

    @staticmethod
    def unique_id() -> bytes:
        buf = (ctypes.c_uint8 * 128)()
        check(_ffi.lib().rc_comm_get_unique_id(buf))
        return bytes(buf)

    def all_reduce(self, op: str, buf: CudaRaw, count: Optional[int] = None):
        check(_ffi.lib().rc_comm_all_reduce(self._handle, _ffi.REDOPS[op], dtype_code(buf.dtype), buf.ptr,
                                            buf.len if count is None else count))

    def info(self) -> Tuple[int, int, bool]:
        """(nranks, rank, peer_window): peer_window = results up to 256 KiB are combined by the one-kernel NVLink
        exchange instead of ncclAllReduce."""
        n, r, p = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(_ffi.lib().rc_comm_info(self._handle, byref(n), byref(r), byref(p)))
        return n.value, r.value, bool(p.value)

    def set_peer_window(self, enable: bool):
        """Every rank must make the same call at the same point (see rc_comm_set_peer_window)."""
        check(_ffi.lib().rc_comm_set_peer_window(self._handle, 1 if enable else 0))

    def reduce_axes_sharded(self, op: str, a: CudaRaw, la: Layout, axes: Sequence[int], n_reduced_global: int,
                            out: CudaRaw, lo: Layout):
        """`*_axes` of a tensor whose SHARDED axis is reduced: `la` is this rank's shard, (out, lo) the full output."""
        if out.dtype != a.dtype:
            raise RstsrCudaError(6, "DTypeMismatch: sharded reductions write the element type")
        arr = (ctypes.c_int64 * max(len(axes), 1))(*[int(x) for x in axes])
        check(_ffi.lib().rc_reduce_axes_sharded(self.device._handle, self._handle, _ffi.REDOPS[op], dtype_code(a.dtype),
                                                a.ptr, byref(la.to_c()), arr, len(axes), int(n_reduced_global), out.ptr,
                                                byref(lo.to_c())))

    def reduce_all_sharded(self, op: str, a: CudaRaw, la: Layout, n_global: int):
        out = np.empty(1, dtype=a.dtype)
        check(_ffi.lib().rc_reduce_all_sharded(self.device._handle, self._handle, _ffi.REDOPS[op], dtype_code(a.dtype),
                                               a.ptr, byref(la.to_c()), n_global, out.ctypes.data))
        return out[0]

    def close(self):
        if self._handle:
            _ffi.lib().rc_comm_destroy(self._handle)
            self._handle = None
