"""A thin `Tensor` over DeviceCuda -- just enough of rstsr's L4 tensor API to drive the device path the way the
reference's callers do (`&a + &b`, `a.to_contig(order)`, `a.sum_axes(ax)`, ...), so tests read like the reference's.

Each method states the reference caller it mirrors; all of them do what the reference does on the host
(broadcast, choose the output layout, allocate) and then make exactly ONE device call.  Host-side decisions are
evaluated by the C++ layout algebra inside librstsr_cuda.so (via rstsr_b200.device helpers), never by `oracle/`.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np

from . import _ffi
from .device import (CudaRaw, DeviceCuda, Layout, broadcast_layout, layout_for_array_copy, layout_for_binary_op,
                     layout_reshapeable)

ROW_MAJOR, COL_MAJOR = _ffi.ROW_MAJOR, _ffi.COL_MAJOR
_FUNC_OPS = {"maximum", "minimum", "floor_divide", "pow", "atan2", "copysign", "hypot", "logaddexp", "nextafter",
             "eq", "ne", "lt", "le", "gt", "ge"}


class Tensor:
    """(storage, layout) pair: TensorBase<Storage<R, T, B>, D> (rstsr-core/src/tensorbase.rs:5-25)."""

    def __init__(self, raw: CudaRaw, layout: Layout, owned: bool = True):
        self.raw = raw
        self.layout = layout
        self.owned = owned  # Tensor (owns its buffer) vs TensorView

    # ---- basic properties ----
    @property
    def device(self) -> DeviceCuda:
        return self.raw.device

    @property
    def dtype(self) -> np.dtype:
        return self.raw.dtype

    @property
    def shape(self) -> Tuple[int, ...]:
        return self.layout.shape

    @property
    def stride(self) -> Tuple[int, ...]:
        return self.layout.stride

    @property
    def ndim(self) -> int:
        return self.layout.ndim

    @property
    def size(self) -> int:
        return self.layout.size

    def view(self) -> "Tensor":
        return Tensor(self.raw, self.layout, owned=False)

    def _with(self, layout: Layout) -> "Tensor":
        return Tensor(self.raw, layout, owned=False)

    # ---- host transfer ----
    def to_numpy(self) -> np.ndarray:
        """Download the whole storage and materialise this view (C-contiguous numpy array)."""
        host = self.device.to_cpu_vec(self.raw)
        l = self.layout
        if l.size == 0:
            return np.zeros(l.shape, dtype=self.dtype)
        item = host.dtype.itemsize
        v = np.lib.stride_tricks.as_strided(host[l.offset:], shape=l.shape, strides=tuple(s * item for s in l.stride),
                                            writeable=False)
        return np.array(v)

    def to_vec(self) -> np.ndarray:
        """to_vec (1-D only; rstsr-core/src/tensor/ownership_conversion.rs:177-187)."""
        if self.ndim != 1:
            raise _ffi.RstsrCudaError(3, "to_vec is only defined for 1-D tensors")
        return self.to_numpy()

    def to_scalar(self):
        if self.size != 1:
            raise _ffi.RstsrCudaError(3, "to_scalar needs exactly one element")
        return self.device.get_index(self.raw, self.layout.offset)

    # ---- views: no data movement (rstsr-core/src/tensor/manipulation/{transpose,...}.rs) ----
    def transpose(self, axes: Optional[Sequence[int]] = None) -> "Tensor":
        n = self.ndim
        if axes is None:
            axes = list(range(n))[::-1]
        ax = [a + n if a < 0 else a for a in axes]
        if sorted(ax) != list(range(n)):
            raise _ffi.RstsrCudaError(3, "invalid permutation")
        l = self.layout
        return self._with(Layout(tuple(l.shape[a] for a in ax), tuple(l.stride[a] for a in ax), l.offset))

    into_transpose = transpose
    permute_dims = transpose

    def swapaxes(self, a1: int, a2: int) -> "Tensor":
        a1, a2 = _check_axis(a1, self.ndim), _check_axis(a2, self.ndim)  # out of range: InvalidValue, as in the reference
        ax = list(range(self.ndim))
        ax[a1], ax[a2] = ax[a2], ax[a1]
        return self.transpose(ax)

    into_swapaxes = swapaxes

    def reverse_axes(self) -> "Tensor":
        return self.transpose(None)

    def flip(self, axes: Union[None, int, Sequence[int]] = None) -> "Tensor":
        """rt::flip (manipulation/flip.rs:3-20): reverse the given axes (all of them for None); a view."""
        if isinstance(axes, (int, np.integer)):
            ax = [int(axes) % self.ndim] if -self.ndim <= axes < self.ndim else _normalize_axes([int(axes)], self.ndim)
        elif axes is None:
            ax = list(range(self.ndim))
        else:
            ax = _normalize_axes(axes, self.ndim)
        return self[tuple(slice(None, None, -1) if i in ax else slice(None) for i in range(self.ndim))]

    def expand_dims(self, axes: Union[int, Sequence[int]]) -> "Tensor":
        """rt::expand_dims / unsqueeze (manipulation/expand_dims.rs:3-20): the axes index the RESULT (ndim + len(axes)
        dimensions), duplicates are an error, insertion runs in ascending order.  An inserted unit axis gets stride 1
        (the reference's dim_insert copies a neighbour's stride; a unit axis is never stepped along)."""
        l = self.layout
        if isinstance(axes, (int, np.integer)):
            axis = int(axes)
            if not -(l.ndim + 1) <= axis <= l.ndim:
                raise _ffi.RstsrCudaError(2, f"axis {axis} out of bounds for inserting into ndim {l.ndim}")
            ins = [axis + l.ndim + 1 if axis < 0 else axis]
        else:
            ins = sorted(_normalize_axes(axes, l.ndim + len(list(axes))))
        shape, stride = list(l.shape), list(l.stride)
        for axis in ins:
            shape.insert(axis, 1)
            stride.insert(axis, 1)
        return self._with(Layout(tuple(shape), tuple(stride), l.offset))

    unsqueeze = expand_dims

    def squeeze(self, axes: Union[None, int, Sequence[int]] = None) -> "Tensor":
        """rt::squeeze (manipulation/squeeze.rs:3-38): drop the given unit axes (every unit axis for None); an axis
        of another extent and repeated axes are InvalidValue errors (dim_eliminate, indexer.rs:256-267)."""
        l = self.layout
        if axes is None:
            drop = [i for i in range(l.ndim) if l.shape[i] == 1]
        else:
            raw = [int(axes)] if isinstance(axes, (int, np.integer)) else [int(a) for a in axes]
            drop = [a + l.ndim if a < 0 else a for a in raw]
            if any(a < 0 for a in drop):
                raise _ffi.RstsrCudaError(2, "Some negative index is too small.")
            if len(set(drop)) != len(drop):
                raise _ffi.RstsrCudaError(2, "Same axes is not allowed here.")
            for a in drop:
                if a >= l.ndim:
                    raise _ffi.RstsrCudaError(2, f"axis {a} out of bounds for ndim {l.ndim}")
                if l.shape[a] != 1:
                    raise _ffi.RstsrCudaError(2, "Dimension to be eliminated is not 1.")
        keep = [i for i in range(l.ndim) if i not in drop]
        return self._with(Layout(tuple(l.shape[i] for i in keep), tuple(l.stride[i] for i in keep), l.offset))

    def moveaxis(self, source: Union[int, Sequence[int]], destination: Union[int, Sequence[int]]) -> "Tensor":
        """rt::moveaxis (manipulation/moveaxis.rs:3-36): NumPy's rule -- the axes not named keep their order, each
        source axis is inserted at its destination, destinations taken in ascending order."""
        src = _normalize_axes([source] if isinstance(source, (int, np.integer)) else source, self.ndim)
        dst = _normalize_axes([destination] if isinstance(destination, (int, np.integer)) else destination, self.ndim)
        if len(src) != len(dst):
            raise _ffi.RstsrCudaError(2, "`source` and `destination` arguments must have the same number of elements")
        order = [i for i in range(self.ndim) if i not in src]
        for d, s_ in sorted(zip(dst, src)):
            order.insert(d, s_)
        return self.transpose(order)

    def broadcast_to(self, shape: Sequence[int]) -> "Tensor":
        target = Layout.contig(shape, self.device.default_order())
        la_b, _ = broadcast_layout(self.layout, target, self.device.default_order())
        if la_b.shape != tuple(shape):
            raise _ffi.RstsrCudaError(3, "Broadcasting failed.")
        return self._with(la_b)

    def __getitem__(self, key) -> "Tensor":
        """`.i(...)` with ints, slices, None and Ellipsis, NumPy slice semantics for in-range bounds."""
        if not isinstance(key, tuple):
            key = (key,)
        n_real = sum(1 for k in key if k is not None and k is not Ellipsis)
        if Ellipsis in key:
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - n_real) + key[i + 1:]
        shape, stride, offset = [], [], self.layout.offset
        axis = 0
        for k in key:
            if k is None:
                shape.append(1)
                stride.append(1)
                continue
            d, s = self.layout.shape[axis], self.layout.stride[axis]
            if isinstance(k, slice):
                start, stop, step = k.indices(d)
                length = len(range(start, stop, step))
                if length > 0:
                    offset += start * s
                shape.append(length)
                stride.append(s * step)
            else:
                k = int(k)
                if k < 0:
                    k += d
                if not 0 <= k < d:
                    raise _ffi.RstsrCudaError(9, "index out of bounds")
                offset += k * s
            axis += 1
        while axis < self.ndim:
            shape.append(self.layout.shape[axis])
            stride.append(self.layout.stride[axis])
            axis += 1
        return self._with(Layout(tuple(shape), tuple(stride), offset))

    i = __getitem__

    # ---- copies ----
    def to_layout(self, layout: Layout) -> "Tensor":
        """change_layout_f (rstsr-core/src/tensor/manipulation/to_layout.rs:8-38)."""
        if layout.size != self.layout.size:
            raise _ffi.RstsrCudaError(3, "size mismatch")
        if layout.same_as(self.layout):
            return self.view()
        dev = self.device
        raw = dev.uninit_impl(self.dtype, layout.bounds_index()[1])
        dev.assign_arbitary_uninit(raw, layout, self.raw, self.layout)
        return Tensor(raw, layout)

    def to_contig(self, order: Optional[int] = None) -> "Tensor":
        """to_contig / change_contig_f (manipulation/to_contig.rs:8-23,81-103)."""
        order = self.device.default_order() if order is None else order
        return self.to_layout(Layout.contig(self.shape, order))

    def to_prefer(self, order: Optional[int] = None) -> "Tensor":
        """to_prefer / change_prefer_f (manipulation/to_contig.rs:229-243): a view if the layout already is c- (f-)
        preferred, otherwise the contiguous copy."""
        order = self.device.default_order() if order is None else order
        if self._prefer(order == COL_MAJOR):
            return self.view()
        return self.to_contig(order)

    def to_owned(self) -> "Tensor":
        """asarray((&t, K)) (rstsr-core/src/tensor/asarray.rs:321-341): same-shape copy into a K-order layout."""
        dev = self.device
        lc = layout_for_array_copy(self.layout, _ffi.ITER_K, dev.default_order())
        raw = dev.uninit_impl(self.dtype, lc.bounds_index()[1])
        dev.assign_uninit(raw, lc, self.raw, self.layout)
        return Tensor(raw, lc)

    def to_device(self, target: DeviceCuda) -> "Tensor":
        """tensor.into_device / to_device (DeviceChangeAPI, rstsr-core/src/storage/conversion.rs:3-21): the whole raw
        buffer moves, the layout is kept as is."""
        return Tensor(self.device.change_device(self.raw, target), self.layout)

    def astype(self, dtype) -> "Tensor":
        dev = self.device
        lc = layout_for_array_copy(self.layout, _ffi.ITER_K, dev.default_order())
        raw = dev.uninit_impl(dtype, lc.bounds_index()[1])
        dev.assign_uninit(raw, lc, self.raw, self.layout)
        return Tensor(raw, lc)

    def reshape(self, shape: Sequence[int], order: Optional[int] = None) -> "Tensor":
        """change_shape_with_args_f (manipulation/reshape.rs:113-166): view if possible, else copy."""
        dev = self.device
        order = dev.default_order() if order is None else order
        shape = list(shape) if not isinstance(shape, int) else [shape]
        if -1 in shape:
            rest = 1
            for v in shape:
                if v != -1:
                    rest *= v
            shape[shape.index(-1)] = self.size // rest if rest else 0
        view = layout_reshapeable(self.layout, shape, order)
        if view is not None:
            return self._with(view)
        target = Layout.contig(shape, order)
        raw = dev.uninit_impl(self.dtype, max(target.size, 1))
        # reshape.rs:155-162: pairing order = the requested order, passed explicitly (the handle may be shared)
        dev.assign_arbitary_uninit(raw, target, self.raw, self.layout, order)
        return Tensor(raw, target)

    into_shape = reshape

    def assign(self, other: Union["Tensor", int, float]):
        """a.assign(&b) / a.fill(v) (rstsr-core/src/tensor/assignment.rs:26-52,111-124)."""
        dev = self.device
        if not isinstance(other, Tensor):
            dev.fill(self.raw, self.layout, other)
            return self
        la, lb = broadcast_layout(self.layout, other.layout, dev.default_order())
        if la.shape != self.layout.shape:
            raise _ffi.RstsrCudaError(3, "cannot broadcast to the assigned tensor")
        dev.assign(self.raw, la, other.raw, lb)
        return self

    fill = assign

    # ---- elementwise ----
    def _binary(self, op: str, other, reverse: bool = False) -> "Tensor":
        dev = self.device
        order = dev.default_order()
        out_dtype = dev.binop_out_dtype(op, self.dtype)
        if isinstance(other, Tensor):
            a, b = (other, self) if reverse else (self, other)
            out_dtype = dev.binop_out_dtype_ex(op, a.dtype, b.dtype)  # promotion when the operand types differ
            if not a.device.same_device(b.device):
                raise _ffi.RstsrCudaError(5, "DeviceMismatch")
            la_b, lb_b = broadcast_layout(a.layout, b.layout, order)
            if op in _FUNC_OPS:  # op_binary_common.rs:79-93
                l1 = layout_for_array_copy(la_b, _ffi.ITER_K, order)
                l2 = layout_for_array_copy(lb_b, _ffi.ITER_K, order)
                lc = l1 if l1.same_as(l2) else Layout.contig(la_b.shape, order)
            else:  # op_binary_arithmetic.rs:170-213
                lc = layout_for_binary_op(la_b, lb_b, order)
            raw = dev.uninit_impl(out_dtype, lc.bounds_index()[1])
            dev.op_mutc_refa_refb(op, raw, lc, a.raw, la_b, b.raw, lb_b)
            return Tensor(raw, lc)
        # scalar operand (op_binary_arithmetic.rs:676-677, 746): layout_for_array_copy(K).  The scalar has the tensor's
        # element type, except a float scalar against an integer tensor, which is an f64 operand (promoted pair).
        st = self.dtype
        if isinstance(other, (float, np.floating)) and self.dtype.kind in "iub":
            st = np.dtype(np.float64)
        out_dtype = dev.binop_out_dtype_ex(op, *((st, self.dtype) if reverse else (self.dtype, st)))
        lc = layout_for_array_copy(self.layout, _ffi.ITER_K, order)
        raw = dev.uninit_impl(out_dtype, lc.bounds_index()[1])
        if reverse:
            dev.op_mutc_numa_refb(op, raw, lc, other, self.raw, self.layout, a_dtype=st)
        else:
            dev.op_mutc_refa_numb(op, raw, lc, self.raw, self.layout, other, b_dtype=st)
        return Tensor(raw, lc)

    def consume(self, op: str, other: "Tensor", reverse: bool = False) -> "Tensor":
        """`a op &b` with an OWNED `a` (reverse: `&b op a`): the result reuses `a`'s buffer when `b` broadcasts to
        `a`'s own layout, through OpLConsume*API / OpRConsume*API (tensor/operators/op_binary_arithmetic.rs:280-354);
        otherwise falls back to the allocating path."""
        dev = self.device
        if self.owned and op not in _FUNC_OPS and 0 not in [s for d, s in zip(self.shape, self.stride) if d > 1]:
            try:
                la_b, lb_b = broadcast_layout(self.layout, other.layout, dev.default_order())
            except _ffi.RstsrCudaError:
                la_b = None
            if la_b is not None and la_b.ndim == self.ndim and la_b.same_as(self.layout):
                dev.op_muta_refb(op, self.raw, la_b, other.raw, lb_b, reverse=reverse)
                return Tensor(self.raw, la_b)
        return other._binary(op, self) if reverse else self._binary(op, other)

    def _inplace(self, op: str, other) -> "Tensor":
        dev = self.device
        if isinstance(other, Tensor):
            la, lb = broadcast_layout(self.layout, other.layout, dev.default_order())
            if la.shape != self.layout.shape or la.stride != self.layout.stride:
                raise _ffi.RstsrCudaError(3, "cannot broadcast to the in-place operand")
            dev.op_muta_refb(op, self.raw, la, other.raw, lb)
        else:
            dev.op_muta_numb(op, self.raw, self.layout, other)
        return self

    def __add__(self, o): return self._binary("add", o)
    def __radd__(self, o): return self._binary("add", o, reverse=True)
    def __sub__(self, o): return self._binary("sub", o)
    def __rsub__(self, o): return self._binary("sub", o, reverse=True)
    def __mul__(self, o): return self._binary("mul", o)
    def __rmul__(self, o): return self._binary("mul", o, reverse=True)
    def __truediv__(self, o): return self._binary("div", o)
    def __rtruediv__(self, o): return self._binary("div", o, reverse=True)
    def __mod__(self, o): return self._binary("rem", o)
    def __or__(self, o): return self._binary("bitor", o)
    def __and__(self, o): return self._binary("bitand", o)
    def __xor__(self, o): return self._binary("bitxor", o)
    def __lshift__(self, o): return self._binary("shl", o)
    def __rshift__(self, o): return self._binary("shr", o)
    def __iadd__(self, o): return self._inplace("add", o)
    def __isub__(self, o): return self._inplace("sub", o)
    def __imul__(self, o): return self._inplace("mul", o)
    def __itruediv__(self, o): return self._inplace("div", o)

    def binary(self, op: str, other) -> "Tensor":
        return self._binary(op, other)

    def _unary(self, op: str) -> "Tensor":
        dev = self.device
        la = layout_for_array_copy(self.layout, _ffi.ITER_K, dev.default_order())  # op_unary_common.rs:150-153
        raw = dev.uninit_impl(dev.unop_out_dtype(op, self.dtype), la.bounds_index()[1])
        dev.unary_muta_refb(op, raw, la, self.raw, self.layout)
        return Tensor(raw, la)

    def unary(self, op: str) -> "Tensor":
        return self._unary(op)

    def unary_inplace(self, op: str) -> "Tensor":
        self.device.unary_muta(op, self.raw, self.layout)
        return self

    def __neg__(self): return self._unary("neg")
    def __invert__(self): return self._unary("not_")
    def __abs__(self): return self._unary("abs")

    # ---- reductions (rstsr-core/src/tensor/reduction.rs:3-133) ----
    def _reduce(self, op: str, axes=None):
        dev = self.device
        if axes is None:
            return dev.reduce_all(op, self.raw, self.layout)
        if isinstance(axes, int):
            axes = [axes]
        raw, lo = dev.reduce_axes(op, self.raw, self.layout, list(axes))
        return Tensor(raw, lo)

    def sum_all(self): return self._reduce("sum")
    def prod_all(self): return self._reduce("prod")
    def max_all(self): return self._reduce("max")
    def min_all(self): return self._reduce("min")
    def mean_all(self): return self._reduce("mean")
    def sum_axes(self, axes): return self._reduce("sum", axes)
    def prod_axes(self, axes): return self._reduce("prod", axes)
    def max_axes(self, axes): return self._reduce("max", axes)
    def min_axes(self, axes): return self._reduce("min", axes)
    def mean_axes(self, axes): return self._reduce("mean", axes)
    sum, prod, max, min, mean = sum_all, prod_all, max_all, min_all, mean_all

    # the "next" reductions (tensor/reduction.rs:118-133): same call shape, other monoids
    def var_all(self): return self._reduce("var")
    def std_all(self): return self._reduce("std")
    def l2_norm_all(self): return self._reduce("l2_norm")
    def argmin_all(self): return self._reduce("argmin")
    def argmax_all(self): return self._reduce("argmax")
    def all_all(self): return bool(self._reduce("all"))
    def any_all(self): return bool(self._reduce("any"))
    def count_nonzero_all(self): return self._reduce("count_nonzero")
    def var_axes(self, axes): return self._reduce("var", axes)
    def std_axes(self, axes): return self._reduce("std", axes)
    def l2_norm_axes(self, axes): return self._reduce("l2_norm", axes)
    def argmin_axes(self, axes): return self._reduce("argmin", axes)
    def argmax_axes(self, axes): return self._reduce("argmax", axes)
    def all_axes(self, axes): return self._reduce("all", axes)
    def any_axes(self, axes): return self._reduce("any", axes)
    def count_nonzero_axes(self, axes): return self._reduce("count_nonzero", axes)

    def unraveled_argmin_all(self) -> Tuple[int, ...]:
        """OpUnraveledArgMinAPI::unraveled_argmin_all (operators/reduction.rs:38-55): the multi-index of argmin_all's
        row-major flat index.  (The `_axes` variant returns a tensor of index VECTORS -- not a POD element type.)"""
        return tuple(int(i) for i in np.unravel_index(int(self.argmin_all()), self.shape))

    def unraveled_argmax_all(self) -> Tuple[int, ...]:
        return tuple(int(i) for i in np.unravel_index(int(self.argmax_all()), self.shape))

    # ---- index-driven movement ----
    def index_select(self, axis: int, indices: Sequence[int]) -> "Tensor":
        """tensor.index_select(axis, indices) (rstsr-core/src/tensor/adv_indexing.rs:10-48): negative indices count
        from the end; the output is contiguous in the device's default order."""
        dev = self.device
        if not -self.ndim <= axis < self.ndim:
            raise _ffi.RstsrCudaError(2, "axis out of bounds")
        axis = axis + self.ndim if axis < 0 else axis
        n = self.shape[axis]
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        idx = np.where(idx < 0, idx + n, idx)
        if idx.size and (idx.min() < 0 or idx.max() >= n):
            raise _ffi.RstsrCudaError(9, f"Invalid index that exceeds shape length at axis {axis}.")
        shape = list(self.shape)
        shape[axis] = int(idx.size)
        lo = Layout.contig(shape, dev.default_order())
        raw = dev.uninit_impl(self.dtype, max(lo.size, 1))
        dev.index_select(raw, lo, self.raw, self.layout, axis, idx)
        return Tensor(raw, lo)

    def take(self, indices: Sequence[int], axis: int) -> "Tensor":
        return self.index_select(axis, indices)

    def _prefer(self, col: bool) -> bool:
        """Layout::f_prefer / c_prefer (rstsr-common/src/layout/layoutbase.rs:89-142)."""
        if self.ndim == 0 or self.layout.size == 0:
            return True
        dims = list(zip(self.stride, self.shape))
        last = 0
        for s, d in (dims if col else reversed(dims)):
            if d != 1:
                if s < last or (last == 0 and s != 1):
                    return False
                last = s
            elif last == 0:
                last = 1
        return True

    def _tri_out_layout(self, shape) -> Layout:
        c, f = self._prefer(False), self._prefer(True)
        if c and not f:
            return Layout.contig(shape, ROW_MAJOR)
        if f and not c:
            return Layout.contig(shape, COL_MAJOR)
        return Layout.contig(shape, self.device.default_order())

    def pack_tri(self, uplo: str) -> "Tensor":
        """tensor.pack_tri(uplo) (rstsr-core/src/tensor/operators/op_tri.rs:13-71): row-major devices pack the LAST two
        axes into one of n(n+1)/2, col-major devices the FIRST two; the output follows the input's c/f preference."""
        dev = self.device
        if self.ndim < 2:
            raise _ffi.RstsrCudaError(3, "pack_tri needs at least two axes")
        if dev.default_order() == ROW_MAJOR:
            n, m, rest = self.shape[-2], self.shape[-1], list(self.shape[:-2])
            if n != m:
                raise _ffi.RstsrCudaError(3, "Last two dimensions should be the same for pack_tri.")
            shape = rest + [n * (n + 1) // 2]
        else:
            n, m, rest = self.shape[0], self.shape[1], list(self.shape[2:])
            if n != m:
                raise _ffi.RstsrCudaError(3, "First two dimensions should be the same for pack_tri.")
            shape = [n * (n + 1) // 2] + rest
        la = self._tri_out_layout(shape)
        raw = dev.uninit_impl(self.dtype, max(la.bounds_index()[1], 1))
        dev.pack_tri(raw, la, self.raw, self.layout, uplo)
        return Tensor(raw, la)

    def pack_tril(self): return self.pack_tri("L")
    def pack_triu(self): return self.pack_tri("U")

    def unpack_tri(self, uplo: str, symm: str) -> "Tensor":
        """tensor.unpack_tri(uplo, symm) (rstsr-core/src/tensor/operators/op_tri.rs:106-163)."""
        dev = self.device
        if self.ndim < 1:
            raise _ffi.RstsrCudaError(3, "unpack_tri needs at least one axis")
        row = dev.default_order() == ROW_MAJOR
        n_tp = self.shape[-1] if row else self.shape[0]
        n = int(np.floor(np.sqrt(np.float64(2 * n_tp))))
        if n * (n + 1) // 2 != n_tp:
            raise _ffi.RstsrCudaError(3, ("Last" if row else "First") + " dimension should be triangular number for unpack_tri.")
        shape = list(self.shape[:-1]) + [n, n] if row else [n, n] + list(self.shape[1:])
        la = self._tri_out_layout(shape)
        raw = dev.uninit_impl(self.dtype, max(la.bounds_index()[1], 1))
        dev.unpack_tri(raw, la, self.raw, self.layout, uplo, symm)
        return Tensor(raw, la)

    def unpack_tril(self, symm: str): return self.unpack_tri("L", symm)
    def unpack_triu(self, symm: str): return self.unpack_tri("U", symm)


# ---- binary reductions ----
def _kept_layout(l: Layout, axes: Sequence[int]) -> Layout:
    keep = [i for i in range(l.ndim) if i not in axes]
    return Layout(tuple(l.shape[i] for i in keep), tuple(l.stride[i] for i in keep), l.offset)


def vecdot(a: Tensor, b: Tensor, axis: Union[None, int, Tuple[Sequence[int], Sequence[int]]] = None) -> Tensor:
    """rt::vecdot(a, b, axis) (rstsr-core/src/tensor/linalg/vecdot.rs:160-243): contracts `axis` (default -1, negative
    counted from the end of EACH operand, non-negative must be < min(ndim)) or the axes pair (axes_a, axes_b); the
    remaining axes broadcast in the device's default order; output layout by get_layout_for_binary_op."""
    dev = a.device
    if not dev.same_device(b.device):
        raise _ffi.RstsrCudaError(5, "DeviceMismatch")
    nmin = min(a.ndim, b.ndim)
    if axis is None:
        axis = -1
    if isinstance(axis, (int, np.integer)):
        axis = int(axis)
        if axis < 0:
            if not (-nmin <= axis <= -1):
                raise _ffi.RstsrCudaError(2, "axis should be [-N, -1] where N is min(a.ndim, b.ndim)")
            axes_a, axes_b = [axis + a.ndim], [axis + b.ndim]
        else:
            if not (0 <= axis < nmin):
                raise _ffi.RstsrCudaError(2, "axis should be [0, N) where N is min(a.ndim, b.ndim)")
            axes_a, axes_b = [axis], [axis]
    else:
        axes_a = [int(x) + a.ndim if int(x) < 0 else int(x) for x in axis[0]]
        axes_b = [int(x) + b.ndim if int(x) < 0 else int(x) for x in axis[1]]
        if len(axes_a) != len(axes_b):
            raise _ffi.RstsrCudaError(2, "axes_a and axes_b should have the same length")
        for ax, nd in ((axes_a, a.ndim), (axes_b, b.ndim)):
            if any(not (0 <= x < nd) for x in ax) or len(set(ax)) != len(ax):
                raise _ffi.RstsrCudaError(2, "axes out of bounds or repeated")
    if [a.shape[i] for i in axes_a] != [b.shape[i] for i in axes_b]:
        raise _ffi.RstsrCudaError(3, "the dimensions of a and b along the contracted axis should be the same")
    order = dev.default_order()
    lam_b, lbm_b = broadcast_layout(_kept_layout(a.layout, axes_a), _kept_layout(b.layout, axes_b), order)
    lc = layout_for_binary_op(lam_b, lbm_b, order)
    raw = dev.uninit_impl(a.dtype, max(lc.bounds_index()[1], 1))
    dev.vecdot(raw, lc, a.raw, a.layout, b.raw, b.layout, axes_a, axes_b)
    return Tensor(raw, lc)


def allclose(a: Tensor, b: Tensor, rtol: float = 1.0e-5, atol: float = 1.0e-8, equal_nan: bool = False) -> bool:
    """rt::allclose(a, b, args) (rstsr-core/src/tensor/reduction.rs:324-351): broadcast, then the device's
    allclose_all.  Note the reference's isclose treats inf vs inf as NOT close (|inf - inf| is NaN)."""
    dev = a.device
    if not dev.same_device(b.device):
        raise _ffi.RstsrCudaError(5, "DeviceMismatch")
    la_b, lb_b = broadcast_layout(a.layout, b.layout, dev.default_order())
    return dev.allclose_all(a.raw, la_b, b.raw, lb_b, rtol, atol, equal_nan)


def isclose(a: Tensor, b: Tensor, rtol: float = 1.0e-5, atol: float = 1.0e-8, equal_nan: bool = False) -> Tensor:
    """rt::isclose(a, b, args): elementwise OpIsCloseAPI (rstsr-core/src/operators/ops/op_ternary_common.rs:59-102);
    output layout by the rule of the other binary-function ops (tensor/operators/op_binary_common.rs:79-93)."""
    dev = a.device
    if not dev.same_device(b.device):
        raise _ffi.RstsrCudaError(5, "DeviceMismatch")
    order = dev.default_order()
    la_b, lb_b = broadcast_layout(a.layout, b.layout, order)
    l1 = layout_for_array_copy(la_b, _ffi.ITER_K, order)
    l2 = layout_for_array_copy(lb_b, _ffi.ITER_K, order)
    lc = l1 if l1.same_as(l2) else Layout.contig(la_b.shape, order)
    raw = dev.uninit_impl(np.bool_, max(lc.bounds_index()[1], 1))
    dev.isclose(raw, lc, a.raw, la_b, b.raw, lb_b, rtol, atol, equal_nan)
    return Tensor(raw, lc)


# ---- creation from tensors: compositions of OpAssignAPI (rstsr-core/src/tensor/creation_from_tensor.rs) ----
def _normalize_axes(axes: Sequence[int], ndim: int) -> list:
    """normalize_axes_index(allow_duplicate = false) (rstsr-common/src/axis_index.rs:379-412), order kept."""
    out = []
    for a in axes:
        a = int(a)
        if not -ndim <= a < ndim:
            raise _ffi.RstsrCudaError(2, f"axis {a} out of bounds for ndim {ndim}")
        out.append(a + ndim if a < 0 else a)
    if len(set(out)) != len(out):
        raise _ffi.RstsrCudaError(2, "Duplicate axes are not allowed.")
    return out


def _check_axis(axis: int, ndim: int) -> int:
    if not -ndim <= axis < ndim:
        raise _ffi.RstsrCudaError(2, f"axis {axis} out of bounds for ndim {ndim}")
    return axis + ndim if axis < 0 else axis


def concat(tensors: Sequence[Tensor], axis: int = 0) -> Tensor:
    """rt::concat((tensors, axis)) (creation_from_tensor.rs:321-386): a default-order contiguous result, one strided
    `assign` per input into its slab."""
    tensors = list(tensors)
    if not tensors:
        raise _ffi.RstsrCudaError(2, "concat requires at least one tensor.")
    dev, ndim = tensors[0].device, tensors[0].ndim
    if ndim == 0:
        raise _ffi.RstsrCudaError(3, "All tensors must have ndim > 0 in concat.")
    for t in tensors:
        if t.ndim != ndim:
            raise _ffi.RstsrCudaError(3, "All tensors must have the same ndim.")
        if not t.device.same_device(dev):
            raise _ffi.RstsrCudaError(5, "All tensors must be on the same device.")
        if t.dtype != tensors[0].dtype:
            raise _ffi.RstsrCudaError(6, "concat: all tensors must share one dtype")
    axis = _check_axis(axis, ndim)
    other = [d for i, d in enumerate(tensors[0].shape) if i != axis]
    total = 0
    for t in tensors:
        if [d for i, d in enumerate(t.shape) if i != axis] != other:
            raise _ffi.RstsrCudaError(3, "All tensors must have the same shape except for the concatenation axis.")
        total += t.shape[axis]
    shape = other[:axis] + [total] + other[axis:]
    result = empty(shape, dev, dtype=tensors[0].dtype)
    offset = 0
    for t in tensors:
        n = t.shape[axis]
        key = tuple(slice(offset, offset + n) if i == axis else slice(None) for i in range(ndim))
        slab = result[key]
        dev.assign(result.raw, slab.layout, t.raw, t.layout)
        offset += n
    return result


concatenate = concat


def stack(tensors: Sequence[Tensor], axis: int = 0) -> Tensor:
    """rt::stack((tensors, axis)) (creation_from_tensor.rs:643-684): expand_dims(axis) on every input, then concat."""
    tensors = list(tensors)
    if not tensors:
        raise _ffi.RstsrCudaError(2, "stack requires at least one tensor.")
    ndim = tensors[0].ndim
    for t in tensors:
        if t.shape != tensors[0].shape:
            raise _ffi.RstsrCudaError(3, "All tensors must have the same shape.")
    if not -(ndim + 1) <= axis <= ndim:
        raise _ffi.RstsrCudaError(2, f"axis {axis} out of bounds for inserting into ndim {ndim}")
    axis = axis + ndim + 1 if axis < 0 else axis
    return concat([t.expand_dims(axis) for t in tensors], axis)


def atleast_1d(t: Tensor) -> Tensor:
    return t.expand_dims(0) if t.ndim == 0 else t.view()


def atleast_2d(t: Tensor) -> Tensor:
    if t.ndim == 0:
        return t.expand_dims(0).expand_dims(1)
    return t.expand_dims(0) if t.ndim == 1 else t.view()


def hstack(tensors: Sequence[Tensor]) -> Tensor:
    """creation_from_tensor.rs:505-526: atleast_1d, then concat along axis 0 (1-D inputs) or 1."""
    if not len(tensors):
        raise _ffi.RstsrCudaError(2, "hstack requires at least one tensor.")
    ts = [atleast_1d(t) for t in tensors]
    return concat(ts, 0 if ts[0].ndim == 1 else 1)


def vstack(tensors: Sequence[Tensor]) -> Tensor:
    """creation_from_tensor.rs:580-594: atleast_2d, then concat along axis 0."""
    if not len(tensors):
        raise _ffi.RstsrCudaError(2, "vstack requires at least one tensor.")
    return concat([atleast_2d(t) for t in tensors], 0)


def unstack(t: Tensor, axis: int = 0) -> list:
    """creation_from_tensor.rs:783-812: views, one per index along `axis` (no copy)."""
    if t.ndim == 0:
        raise _ffi.RstsrCudaError(3, "unstack requires a tensor with ndim > 0.")
    axis = _check_axis(axis, t.ndim)
    return [t[tuple(i if k == axis else slice(None) for k in range(t.ndim))] for i in range(t.shape[axis])]


def meshgrid(tensors: Sequence[Tensor], indexing: str = "xy", copy: bool = True) -> list:
    """rt::meshgrid((tensors, indexing, copy)) (creation_from_tensor.rs:131-207): each 1-D input is reshaped to
    (1, .., -1, .., 1), broadcast against the others (stride-0 views) and, with `copy`, materialised contiguously in
    the device's default order -- one strided `assign` from a broadcast source per output."""
    if indexing not in ("ij", "xy"):
        raise _ffi.RstsrCudaError(2, "indexing must be 'ij' or 'xy'.")
    tensors = list(tensors)
    if not tensors:
        return []
    for t in tensors:
        if t.ndim != 1:
            raise _ffi.RstsrCudaError(3, "meshgrid only support 1-D tensor.")
        if not t.device.same_device(tensors[0].device):
            raise _ffi.RstsrCudaError(5, "All tensors must be on the same device.")
    if len(tensors) == 1:
        return [tensors[0].to_contig()] if copy else [tensors[0].view()]
    nd = len(tensors)
    shape = [t.shape[0] for t in tensors]
    if indexing == "xy":
        shape[0], shape[1] = shape[1], shape[0]
    outs = []
    for i, t in enumerate(tensors):
        ax = (1 if i == 0 else 0 if i == 1 else i) if indexing == "xy" else i
        stride = [0] * nd
        stride[ax] = t.stride[0]
        v = t._with(Layout(tuple(shape), tuple(stride), t.layout.offset))
        outs.append(v.to_contig(t.device.default_order()) if copy else v)
    return outs


def _diagonal_layout(l: Layout, offset: int) -> Layout:
    """Layout::diagonal(offset, 0, 1) of a 2-D layout (rstsr-common/src/layout/layoutbase.rs:322-384)."""
    d1, d2 = l.shape
    t1, t2 = l.stride
    if -d1 + 1 <= offset < 0:
        off, n = l.offset + t1 * (-offset), min(d1 + offset, d2)
    elif 0 <= offset < d1:
        off, n = l.offset + t2 * offset, max(min(d2 - offset, d1), 0)
    else:
        off, n = l.offset, 0
    return Layout((n,), (t1 + t2,), off)


def diag(t: Tensor, offset: int = 0) -> Tensor:
    """rt::diag((tensor, offset)) (creation_from_tensor.rs:48-82): 1-D -> matrix with the vector on diagonal `offset`
    (zeros elsewhere); 2-D -> the diagonal as a new 1-D tensor."""
    dev = t.device
    if t.ndim == 1:
        n = t.size + abs(offset)
        result = full([n, n], 0, dev, dtype=t.dtype)
        dev.assign(result.raw, _diagonal_layout(result.layout, offset), t.raw, t.layout)
        return result
    if t.ndim == 2:
        ld = _diagonal_layout(t.layout, offset)
        result = empty([ld.shape[0]], dev, dtype=t.dtype)
        if ld.shape[0]:
            dev.assign(result.raw, result.layout, t.raw, ld)
        return result
    raise _ffi.RstsrCudaError(3, "diag only support 1-D or 2-D tensor.")


# ---- creation (rstsr-core/src/tensor/{asarray,creation}.rs) ----
def asarray(data, device: DeviceCuda, layout: Optional[Layout] = None, dtype=None) -> Tensor:
    """asarray((vec, layout, &device)): upload a flat vector and view it through `layout`; a numpy array is
    uploaded as is (its C-order flattening) with the device's default-order contiguous layout of its shape."""
    arr = np.asarray(data, dtype=dtype)
    shape = arr.shape
    raw = device.outof_cpu_vec(arr.reshape(-1))
    if layout is None:
        if arr.ndim <= 1 or device.default_order() == ROW_MAJOR:
            layout = Layout.contig(shape, ROW_MAJOR)
        else:
            # keep the values at the same logical indices: the upload is C-ordered
            layout = Layout.contig(shape, ROW_MAJOR)
    return Tensor(raw, layout)


def arange(n: int, device: DeviceCuda, dtype=np.int64) -> Tensor:
    return asarray(np.arange(n, dtype=dtype), device)


def zeros(shape: Sequence[int], device: DeviceCuda, dtype=np.float64) -> Tensor:
    l = Layout.contig(shape, device.default_order())
    return Tensor(device.zeros_impl(dtype, max(l.size, 1)), l)


def full(shape: Sequence[int], value, device: DeviceCuda, dtype=np.float64) -> Tensor:
    l = Layout.contig(shape, device.default_order())
    return Tensor(device.full_impl(dtype, max(l.size, 1), value), l)


def ones(shape: Sequence[int], device: DeviceCuda, dtype=np.float64) -> Tensor:
    l = Layout.contig(shape, device.default_order())
    return Tensor(device.ones_impl(dtype, max(l.size, 1)), l)


def eye(n_rows: int, device: DeviceCuda, n_cols: Optional[int] = None, k: int = 0, order: Optional[int] = None,
        dtype=np.float64) -> Tensor:
    """rt::eye((n_rows, n_cols, k, order, &device)) (rstsr-core/src/tensor/creation.rs:409-427): zeros_impl, then `fill`
    of the k-th diagonal view with one.  As in the reference, ColMajor builds the [n_cols, n_rows].f() layout."""
    n_cols = n_rows if n_cols is None else n_cols
    order = device.default_order() if order is None else order
    layout = Layout.contig([n_rows, n_cols], ROW_MAJOR) if order == ROW_MAJOR else Layout.contig([n_cols, n_rows], COL_MAJOR)
    raw = device.zeros_impl(dtype, max(layout.size, 1))
    ld = _diagonal_layout(layout, k)
    if ld.shape[0]:
        device.fill(raw, ld, 1)
    return Tensor(raw, layout)


def empty(shape: Sequence[int], device: DeviceCuda, dtype=np.float64) -> Tensor:
    l = Layout.contig(shape, device.default_order())
    return Tensor(device.uninit_impl(dtype, max(l.size, 1)), l)
