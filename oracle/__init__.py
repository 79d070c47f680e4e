"""CPU ORACLE (test infrastructure, NOT product code).

`oracle` restates, on the CPU, what the reference's DeviceCpuSerial / DeviceFaer compute for the DeviceCuda hot
path: layout algebra in `oracle.layout` (plain Python) and the data loops in `rstsr_oracle.c` (plain C, loaded
here through ctypes).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it; nothing under `rstsr_b200/` does.

Raw storage is a 1-D numpy array (the reference's `Vec<T>`); views are `oracle.layout.Layout`s over it, exactly
the (raw, layout) pairs the reference's device traits receive.

Parity status: the reference is Rust and cannot be built in this image, so the oracle is pinned on the
reference's own known-answer tests (tests/test_oracle_golden.py lists each vector with its file:line).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile
from typing import Optional, Sequence, Tuple

import numpy as np

from . import layout as L
from .layout import COL_MAJOR, ROW_MAJOR, Layout, LayoutError

_HERE = os.path.dirname(os.path.abspath(__file__))

BINOPS = dict(add=0, sub=1, mul=2, div=3, rem=4, bitor=5, bitand=6, bitxor=7, shl=8, shr=9, maximum=10, minimum=11,
              floor_divide=12, pow=13, atan2=14, copysign=15, hypot=16, logaddexp=17, nextafter=18,
              eq=32, ne=33, lt=34, le=35, gt=36, ge=37)
UNOPS = dict(neg=0, not_=1, abs=2, square=3, sign=4, sqrt=5, exp=6, expm1=7, log=8, log2=9, log10=10, sin=11, cos=12,
             tan=13, asin=14, acos=15, atan=16, sinh=17, cosh=18, tanh=19, asinh=20, acosh=21, atanh=22, floor=23,
             ceil=24, round=25, trunc=26, reciprocal=27, conj=28, real=29, imag=30, isnan=48, isinf=49, isfinite=50,
             signbit=51)
REDOPS = dict(sum=0, prod=1, max=2, min=3, mean=4)
DTYPE_CODE = {np.dtype(np.bool_): 0, np.dtype(np.int8): 1, np.dtype(np.int16): 2, np.dtype(np.int32): 3,
              np.dtype(np.int64): 4, np.dtype(np.uint8): 5, np.dtype(np.uint16): 6, np.dtype(np.uint32): 7,
              np.dtype(np.uint64): 8, np.dtype(np.float32): 9, np.dtype(np.float64): 10}
_SUF = {1: "i8", 2: "i16", 3: "i32", 4: "i64", 5: "u8", 6: "u16", 7: "u32", 8: "u64", 9: "f32", 10: "f64"}


class _CLayout(ctypes.Structure):
    _fields_ = [("ndim", ctypes.c_int32), ("shape", ctypes.c_int64 * 16), ("stride", ctypes.c_int64 * 16),
                ("offset", ctypes.c_int64)]


def _cl(l: Layout) -> _CLayout:
    c = _CLayout()
    c.ndim = l.ndim
    for i in range(l.ndim):
        c.shape[i] = l.shape[i]
        c.stride[i] = l.stride[i]
    c.offset = l.offset
    return c


_lib = None


def build(native: bool = False, out_dir: Optional[str] = None) -> str:
    """Compile rstsr_oracle.c with gcc.  native=True adds -march=native (for the timed CPU baseline on the box
    it runs on); the default build is portable (x86-64-v2) because the .so travels to the GPU box."""
    out_dir = out_dir or _HERE
    name = "liboracle_native.so" if native else "liboracle.so"
    out = os.path.join(out_dir, name)
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [gcc, "-O3", "-march=native" if native else "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared",
           "-fno-fast-math", "-ffp-contract=off", "-o", out, os.path.join(_HERE, "rstsr_oracle.c"), "-lm"]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


def load(native: bool = False):
    """Load the oracle library (building it if needed)."""
    global _lib
    if _lib is not None and not native:
        return _lib
    if native:
        try:
            path = build(native=True, out_dir=tempfile.mkdtemp(prefix="rstsr_oracle_"))
        except Exception:
            path = None
        if path:
            return ctypes.CDLL(path)
        return load(False)
    path = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rstsr_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        try:
            build(False)
        except Exception:
            if not os.path.exists(path):
                raise
    _lib = ctypes.CDLL(path)
    return _lib


def num_threads() -> int:
    lib = load()
    lib.orc_num_threads.restype = ctypes.c_int
    return int(lib.orc_num_threads())


def _code(a: np.ndarray) -> int:
    return DTYPE_CODE[a.dtype]


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _check_raw(a: np.ndarray, l: Layout):
    assert a.ndim == 1 and a.flags.c_contiguous, "raw storage must be a flat contiguous numpy array"
    lo, hi = L.bounds_index(l)
    if l.size:
        assert hi <= a.size, f"layout reaches {hi} but storage has {a.size} elements"


# ---------------------------------------------------------------------------------------------
# device-level ops (what the reference's device traits do)
# ---------------------------------------------------------------------------------------------
def op_mutc_refa_refb(op: str, c, lc: Layout, a, la: Optional[Layout], b, lb: Optional[Layout], lib=None):
    """c = a o b.  `a` / `b` may be python scalars (op_mutc_numa_refb / op_mutc_refa_numb).
    Follows cpu_serial/op_with_func.rs:10-199 with the closures of auto_impl/op_ternary_{arithmetic,common}.rs."""
    lib = lib or load()
    code = BINOPS[op]
    arr = a if isinstance(a, np.ndarray) else b
    dt = arr.dtype
    t = DTYPE_CODE[np.dtype(np.uint8) if dt == np.bool_ else dt]
    fn = getattr(lib, f"orc_binary_{_SUF[t]}")
    sa = sb = None
    pa = pb = None
    if isinstance(a, np.ndarray):
        _check_raw(a, la)
        pa = _ptr(a)
    else:
        sa = np.array([a], dtype=dt)
    if isinstance(b, np.ndarray):
        _check_raw(b, lb)
        pb = _ptr(b)
    else:
        sb = np.array([b], dtype=dt)
    _check_raw(c, lc)
    cla = _cl(la) if la is not None else _cl(lc)
    clb = _cl(lb) if lb is not None else _cl(lc)
    fn(ctypes.c_int(code), _ptr(c), ctypes.byref(_cl(lc)), pa, ctypes.byref(cla), pb, ctypes.byref(clb), _ptr(sa), _ptr(sb))


def op_muta_refb_unary(op: str, c, lc: Layout, a, la: Layout):
    """c = f(a) (auto_impl/op_binary_common.rs, op_binary_arithmetic.rs:94-113)."""
    lib = load()
    dt = a.dtype
    t = DTYPE_CODE[np.dtype(np.uint8) if dt == np.bool_ else dt]
    _check_raw(c, lc)
    _check_raw(a, la)
    if dt == np.bool_ and op == "not_":
        # logical not on bool
        tmp = np.zeros_like(a, dtype=np.uint8)
        getattr(lib, "orc_unary_u8")(ctypes.c_int(UNOPS["sign"]), _ptr(tmp), ctypes.byref(_cl(la)), _ptr(a.view(np.uint8)),
                                      ctypes.byref(_cl(la)))
        view_c = c.view(np.uint8)
        one = np.array([1], dtype=np.uint8)
        getattr(lib, "orc_binary_u8")(ctypes.c_int(BINOPS["bitxor"]), _ptr(view_c), ctypes.byref(_cl(lc)), _ptr(tmp),
                                       ctypes.byref(_cl(la)), None, ctypes.byref(_cl(lc)), None, _ptr(one))
        return
    getattr(lib, f"orc_unary_{_SUF[t]}")(ctypes.c_int(UNOPS[op]), _ptr(c), ctypes.byref(_cl(lc)), _ptr(a),
                                          ctypes.byref(_cl(la)))


def assign(c, lc: Layout, a, la: Layout):
    """OpAssignAPI::assign with cast (cpu_serial/assignment.rs:91-119)."""
    lib = load()
    if lc.shape != la.shape:
        raise LayoutError("InvalidLayout", "All shape of layout in this function must be the same.")
    _check_raw(c, lc)
    _check_raw(a, la)
    lib.orc_assign(ctypes.c_int(_code(c)), _ptr(c), ctypes.byref(_cl(lc)), ctypes.c_int(_code(a)), _ptr(a),
                   ctypes.byref(_cl(la)))


def assign_arbitary(c, lc: Layout, a, la: Layout, order: str, parallel: bool = False, lib=None):
    """OpAssignArbitaryAPI::assign_arbitary (cpu_serial/assignment.rs:28-67, cpu_rayon/assignment.rs:41-93)."""
    lib = lib or load()
    if lc.size != la.size:
        raise LayoutError("InvalidLayout", "size mismatch")
    _check_raw(c, lc)
    _check_raw(a, la)
    if order == ROW_MAJOR:
        contig = L.c_contig(lc) and L.c_contig(la)
        it = "C"
    else:
        contig = L.f_contig(lc) and L.f_contig(la)
        it = "F"
    tc, ta = L.translate_to_col_major_unary(lc, it), L.translate_to_col_major_unary(la, it)
    lib.orc_assign_arbitary(ctypes.c_int(_code(c)), _ptr(c), ctypes.byref(_cl(tc)), ctypes.c_int(_code(a)), _ptr(a),
                            ctypes.byref(_cl(ta)), ctypes.c_int(1 if contig else 0), ctypes.c_int(1 if parallel else 0))


def fill(c, lc: Layout, value):
    lib = load()
    _check_raw(c, lc)
    dt = c.dtype
    t = DTYPE_CODE[np.dtype(np.uint8) if dt == np.bool_ else dt]
    v = np.array([value]).astype(dt)
    getattr(lib, f"orc_fill_{_SUF[t]}")(_ptr(c), ctypes.byref(_cl(lc)), _ptr(v))


def reduce_all(op: str, a, la: Layout, device: str = "serial", lib=None):
    """`*_all` (cpu_serial/reduction.rs:144-178; device='rayon': cpu_rayon/reduction.rs:20-106)."""
    lib = lib or load()
    if op in ("max", "min") and la.size == 0:
        raise LayoutError("InvalidValue", f"zero-size array is not supported for {op}")
    _check_raw(a, la)
    t = _code(a)
    lay = L.translate_to_col_major_unary(la, "K")
    outer, size_contig = L.translate_to_col_major_with_contig([lay])
    use = outer[0] if size_contig >= 32 else lay
    out = np.zeros(1, dtype=a.dtype)
    name = "orc_reduce_all_par_" if device == "rayon" else "orc_reduce_all_"
    if device == "rayon" and la.size < 1024:
        name = "orc_reduce_all_"
    getattr(lib, name + _SUF[t])(ctypes.c_int(REDOPS[op]), _ptr(a), ctypes.byref(_cl(use)), ctypes.c_int64(size_contig),
                                 ctypes.c_int64(la.size), _ptr(out))
    return out[0]


def reduce_axes(op: str, a, la: Layout, axes: Sequence[int], device: str = "serial", lib=None) -> Tuple[np.ndarray, Layout]:
    """`*_axes` -> (raw output, output layout).  Follows reduce_axes_cpu_serial (cpu_serial/reduction.rs:180-368);
    device='rayon' follows reduce_axes_cpu_rayon (cpu_rayon/reduction.rs:109-328) incl. its 1024-element serial
    cut-over, 64-column chunks and the extra `init (+) x` of rayon's reduce.

    Stride-0 (broadcast) reduced axes: the reference sizes the repeat count from the KEPT layout
    (`lm.shape()[i]`, cpu_rayon/reduction.rs:162) -- a suspected defect.  The oracle uses the
    reduced layout's extents, i.e. the mathematically intended value; parity is not pinned on that corner."""
    lib = lib or load()
    if op in ("max", "min") and la.size == 0:
        raise LayoutError("InvalidValue", f"zero-size array is not supported for {op}")
    _check_raw(a, la)
    t = _code(a)
    suf = _SUF[t]
    parallel = device == "rayon" and la.size >= 1024
    ls, lm = L.dim_split_axes(la, axes)
    offset = la.offset
    lo = L.layout_for_array_copy(lm, "K")
    out = np.zeros(max(lo.size, 0), dtype=a.dtype)
    if lo.size == 0:
        return out, lo
    _, as0, asc, asd = L.get_axes_composition(ls)
    _, am0, amc, amd = L.get_axes_composition(lm)

    def prod(xs):
        r = 1
        for x in xs:
            r *= x
        return r

    size_s0 = prod(ls.shape[i] for i in as0)
    size_sc = prod(ls.shape[i] for i in asc)
    size_m0 = prod(lm.shape[i] for i in am0)
    size_mc = prod(lm.shape[i] for i in amc)
    n_mean = ls.size
    code = ctypes.c_int(REDOPS[op])
    par = ctypes.c_int(1 if parallel else 0)
    off = ctypes.c_int64(offset)

    def sub(l: Layout, ax):
        return L.dim_split_axes(l, ax)[0]

    if size_sc > 1:
        amcd = amc + amd
        getattr(lib, "orc_reduce_axes_a_" + suf)(code, _ptr(a), _ptr(out), ctypes.byref(_cl(sub(lm, amcd))),
                                                  ctypes.byref(_cl(sub(lo, amcd))), ctypes.byref(_cl(sub(ls, asd))),
                                                  ctypes.c_int64(size_sc), ctypes.c_int64(size_s0), off,
                                                  ctypes.c_int64(n_mean), par)
    elif size_mc > 1:
        ascd = asc + asd
        loc = sub(lo, amc)
        assert L.f_contig(loc), "the contiguous part of input must be the same applied to output"
        getattr(lib, "orc_reduce_axes_b_" + suf)(code, _ptr(a), _ptr(out), ctypes.byref(_cl(sub(lm, amd))),
                                                  ctypes.byref(_cl(sub(lo, amd))), ctypes.byref(_cl(sub(ls, ascd))),
                                                  ctypes.c_int64(size_mc), ctypes.c_int64(size_s0), off,
                                                  ctypes.c_int64(n_mean), par)
    else:
        getattr(lib, "orc_reduce_axes_c_" + suf)(code, _ptr(a), _ptr(out), ctypes.byref(_cl(sub(lm, amd))),
                                                  ctypes.byref(_cl(sub(lo, amd))), ctypes.byref(_cl(sub(ls, asd))),
                                                  ctypes.c_int64(size_s0), off, ctypes.c_int64(n_mean))
    if size_m0 > 1:
        # replicate along stride-0 kept axes.  The reference subtracts the INPUT offset from OUTPUT indices here
        # (cpu_rayon/reduction.rs:291-298); the output layout's offset is 0, which is what is used.
        amcd = amc + amd
        lib.orc_reduce_axes_bcast_fixup(ctypes.c_int(t), _ptr(out), ctypes.byref(_cl(sub(lo, am0))),
                                        ctypes.byref(_cl(sub(lo, amcd))), ctypes.c_int64(0))
    return out, lo


# ---------------------------------------------------------------------------------------------
# tensor-level flows (L4 callers): broadcast + output layout + one device call
# ---------------------------------------------------------------------------------------------
def tensor_binary(op: str, a, la: Layout, b, lb: Layout, order: str = ROW_MAJOR):
    """`&a o &b` (rstsr-core/src/tensor/operators/op_binary_arithmetic.rs:170-213) -> (raw c, lc)."""
    la_b, lb_b = L.broadcast_layout(la, lb, order)
    lc = L.get_layout_for_binary_op(la_b, lb_b, order)
    out_dtype = np.bool_ if BINOPS[op] >= 32 else a.dtype
    c = np.zeros(L.bounds_index(lc)[1], dtype=out_dtype)
    cc = c.view(np.uint8) if out_dtype == np.bool_ else c
    op_mutc_refa_refb(op, cc, lc, a, la_b, b, lb_b)
    return c, lc


def tensor_to_contig(a, la: Layout, target_order: str, device_order: str = ROW_MAJOR):
    """to_contig / change_layout (tensor/manipulation/to_contig.rs:8-23, to_layout.rs:8-38) -> (raw, layout, copied)."""
    target = L.contig_layout(la.shape, target_order)
    if target.same_as(la):
        return a, la, False
    c = np.zeros(L.bounds_index(target)[1], dtype=a.dtype)
    assign_arbitary(c, target, a, la, device_order)
    return c, target, True


def tensor_reshape(a, la: Layout, shape: Sequence[int], order: str = ROW_MAJOR):
    """reshape (tensor/manipulation/reshape.rs:113-166) -> (raw, layout, copied)."""
    shape = L.reshape_substitute_negatives(shape, la.size)
    view = L.layout_reshapeable(la, shape, order)
    if view is not None:
        return a, view, False
    target = L.contig_layout(shape, order)
    c = np.zeros(max(target.size, 1), dtype=a.dtype)
    assign_arbitary(c, target, a, la, order)
    return c, target, True


def to_numpy(raw: np.ndarray, l: Layout) -> np.ndarray:
    """Materialise a (raw, layout) view as a C-contiguous numpy array (for comparisons)."""
    if l.size == 0:
        return np.zeros(l.shape, dtype=raw.dtype)
    item = raw.dtype.itemsize
    lo, _ = L.bounds_index(l)
    view = np.lib.stride_tricks.as_strided(raw[l.offset:] if l.offset <= raw.size else raw, shape=l.shape,
                                           strides=tuple(s * item for s in l.stride), writeable=False)
    return np.array(view)


# ---------------------------------------------------------------------------------------------
# "next" reductions: var / std / l2_norm / argmin / argmax / all / any / count_nonzero.
# Closures of rstsr-core/src/feature_rayon/auto_impl/reduction.rs:207-715; arg semantics of
# rstsr-native-impl/src/cpu_serial/reduction.rs:421-582 (row-major first occurrence, NaN never accepted unless it
# is the first element).  Evaluated with numpy / plain Python on the materialised view: values, not loop order,
# are what these pin (float results are compared with a tolerance anyway).
# ---------------------------------------------------------------------------------------------
def _arg_fold(values, is_max: bool) -> int:
    """reduce_all_unraveled_arg_cpu_serial's fold over a row-major sequence -> row-major index."""
    acc_i, acc_v = None, None
    for i, y in enumerate(values):
        if acc_i is None:
            acc_i, acc_v = i, y  # f_comp(None, y) = Some(true)
            continue
        better = (y > acc_v) if is_max else (y < acc_v)
        if better:
            acc_i, acc_v = i, y
        # equal values: the smaller index stays (iteration is in increasing index order)
    return acc_i


def reduce_ext(op: str, a, la: Layout, axes=None):
    """-> (raw output, output layout) for axes, or a scalar for axes=None."""
    if op in ("argmin", "argmax") and la.size == 0:
        raise LayoutError("InvalidLayout", "empty sequence is not allowed for reduce_arg.")
    view = to_numpy(a, la)
    nd = la.ndim
    if axes is None:
        ax = list(range(nd))
    else:
        ax = L.normalize_axes(axes, nd)
    kept = [i for i in range(nd) if i not in ax]
    # reduced axes in the ORDER GIVEN, flattened row-major (layout_axes of dim_split_axes keeps that order)
    moved = np.transpose(view, kept + ax)
    kshape = tuple(la.shape[i] for i in kept)
    n_red = 1
    for i in ax:
        n_red *= la.shape[i]
    flat = moved.reshape(kshape + (n_red,))
    dt = a.dtype
    if op in ("var", "std"):
        s = flat.sum(-1, dtype=dt)
        q = (flat * flat).sum(-1, dtype=dt)
        n = dt.type(n_red)
        mean = s / n
        res = q / n - mean * mean
        if op == "std":
            res = np.sqrt(res)
        res = res.astype(dt)
    elif op == "l2_norm":
        res = np.sqrt((flat * flat).sum(-1, dtype=dt)).astype(dt)
    elif op in ("argmin", "argmax"):
        res = np.zeros(kshape, dtype=np.uint64)
        it = np.ndindex(*kshape) if kshape else [()]
        for idx in it:
            res[idx] = _arg_fold(list(flat[idx]), op == "argmax")
    elif op == "count_nonzero":
        res = (flat != 0).sum(-1).astype(np.uint64)
    elif op == "all":
        res = flat.astype(bool).all(-1)
    elif op == "any":
        res = flat.astype(bool).any(-1)
    else:
        raise ValueError(op)
    if axes is None:
        return res.reshape(())[()]
    _, lm = L.dim_split_axes(la, ax)
    lo = L.layout_for_array_copy(lm, "K")
    out = np.zeros(max(lo.size, 1), dtype=res.dtype)
    if lo.size:
        item = out.dtype.itemsize
        dst = np.lib.stride_tricks.as_strided(out, shape=lo.shape, strides=tuple(s * item for s in lo.stride))
        dst[...] = res
    return out, lo


# ---- binary reductions (test infrastructure, like everything in this package) ----
def tensor_vecdot(a, la: Layout, b, lb: Layout, axes_a: Sequence[int], axes_b: Sequence[int], order: str = ROW_MAJOR):
    """rt::vecdot (rstsr-core/src/tensor/linalg/vecdot.rs:177-243) + vecdot_naive_cpu_serial
    (rstsr-native-impl/src/cpu_serial/vecdot.rs:6-168): -> (raw c, layout of c).  Sums in the element type
    (integers wrap); the summation ORDER of the reference depends on its contiguity regime, so float parity is
    a tolerance, not bit equality."""
    axes_a = L.normalize_axes(axes_a, la.ndim)
    axes_b = L.normalize_axes(axes_b, lb.ndim)
    las, lam = L.dim_split_axes(la, axes_a)
    lbs, lbm = L.dim_split_axes(lb, axes_b)
    if tuple(las.shape) != tuple(lbs.shape):
        raise LayoutError("InvalidLayout", "the dimensions of a and b along the contracted axis should be the same")
    lam_b, lbm_b = L.broadcast_layout(lam, lbm, order)
    lc = L.get_layout_for_binary_op(lam_b, lbm_b, order)
    n = len(axes_a)
    item_a, item_b = a.dtype.itemsize, b.dtype.itemsize
    full_a = np.lib.stride_tricks.as_strided(a[la.offset:], shape=tuple(lam_b.shape) + tuple(las.shape),
                                             strides=tuple(s * item_a for s in tuple(lam_b.stride) + tuple(las.stride)))
    full_b = np.lib.stride_tricks.as_strided(b[lb.offset:], shape=tuple(lbm_b.shape) + tuple(lbs.shape),
                                             strides=tuple(s * item_b for s in tuple(lbm_b.stride) + tuple(lbs.stride)))
    red = tuple(range(len(lam_b.shape), len(lam_b.shape) + n))
    with np.errstate(over="ignore"):
        res = (full_a * full_b).sum(axis=red, dtype=a.dtype) if n else (full_a * full_b).astype(a.dtype)
    out = np.zeros(max(L.bounds_index(lc)[1] if lc.size else 1, 1), dtype=a.dtype)
    if lc.size:
        dst = np.lib.stride_tricks.as_strided(out[lc.offset:], shape=lc.shape, strides=tuple(s * item_a for s in lc.stride))
        dst[...] = res
    return out, lc


def isclose_scalar(a, b, rtol: float, atol: float, equal_nan: bool) -> bool:
    """rstsr-dtype-traits/src/isclose.rs:92-106 with TE = f64 for two scalars of one numpy dtype."""
    dt = np.asarray(a).dtype
    with np.errstate(all="ignore"):
        if dt.kind == "f":
            diff = np.float64(np.abs(dt.type(a) - dt.type(b)))
            abs_b = np.float64(np.abs(dt.type(b)))
        elif dt.kind == "i":
            ai, bi = int(a), int(b)
            bits = dt.itemsize * 8
            wrap = lambda v: ((v + (1 << (bits - 1))) % (1 << bits)) - (1 << (bits - 1))
            diff = np.float64(wrap(ai - bi) if ai >= bi else wrap(bi - ai))
            abs_b = np.float64(wrap(-bi) if bi < 0 else bi)
        else:
            ai, bi = int(a), int(b)
            diff = np.float64(ai - bi if ai >= bi else bi - ai)
            abs_b = np.float64(bi)
        comp = bool(diff <= np.float64(atol) + np.float64(rtol) * abs_b)
    nan_check = bool(equal_nan) and dt.kind == "f" and bool(np.isnan(a)) and bool(np.isnan(b))
    return comp or nan_check


def tensor_allclose(a, la: Layout, b, lb: Layout, rtol: float = 1.0e-5, atol: float = 1.0e-8, equal_nan: bool = False,
                    order: str = ROW_MAJOR) -> bool:
    """rt::allclose (rstsr-core/src/tensor/reduction.rs:324-351) -> allclose_all
    (rstsr-core/src/device_cpu_serial/reduction.rs:660-683), vectorised restatement of isclose_scalar."""
    la_b, lb_b = L.broadcast_layout(la, lb, order)
    if la_b.size == 0 or lb_b.size == 0:
        raise LayoutError("InvalidValue", "zero-size array is not supported for allclose")
    va, vb = to_numpy(a, la_b), to_numpy(b, lb_b)
    dt = a.dtype
    with np.errstate(all="ignore"):
        if dt.kind == "f":
            diff = np.abs(va - vb).astype(np.float64)
            abs_b = np.abs(vb).astype(np.float64)
        elif dt.kind == "i":
            ua, ub = va.astype(np.dtype(f"u{dt.itemsize}")), vb.astype(np.dtype(f"u{dt.itemsize}"))
            diff = np.where(va >= vb, ua - ub, ub - ua).astype(dt).astype(np.float64)
            abs_b = np.where(vb < 0, (np.zeros_like(ub) - ub).astype(dt), vb).astype(np.float64)
        else:
            diff = np.where(va >= vb, va - vb, vb - va).astype(np.float64)
            abs_b = vb.astype(np.float64)
        ok = diff <= np.float64(atol) + np.float64(rtol) * abs_b
        if equal_nan and dt.kind == "f":
            ok = ok | (np.isnan(va) & np.isnan(vb))
    return bool(ok.all())


# ---- index-driven movement (test infrastructure) ----
def _wview(raw: np.ndarray, l: Layout) -> np.ndarray:
    """writable strided view of `raw` through layout `l`"""
    item = raw.dtype.itemsize
    return np.lib.stride_tricks.as_strided(raw[l.offset:], shape=l.shape, strides=tuple(s * item for s in l.stride))


def index_select(c, lc: Layout, a, la: Layout, axis: int, indices: Sequence[int]):
    """index_select_cpu_serial (rstsr-native-impl/src/cpu_serial/adv_indexing.rs:3-90): c[.., i, ..] = a[.., idx[i], ..]."""
    if lc.ndim != la.ndim:
        raise LayoutError("InvalidLayout", "Input and output ndim should same.")
    if lc.shape[axis] != len(indices):
        raise LayoutError("InvalidLayout", "Invalid index length.")
    if len(indices) and not (0 <= max(indices) < la.shape[axis]):
        raise LayoutError("IndexError", "Index out of range.")
    if lc.size == 0:
        return
    _wview(c, lc)[...] = np.take(to_numpy(a, la), np.asarray(indices, dtype=np.int64), axis=axis)


def tensor_index_select(a, la: Layout, axis: int, indices: Sequence[int], order: str = ROW_MAJOR):
    """index_select_f (rstsr-core/src/tensor/adv_indexing.rs:10-48) -> (raw, layout)."""
    axis = axis + la.ndim if axis < 0 else axis
    n = la.shape[axis]
    idx = []
    for i in indices:
        i = n + i if i < 0 else i
        if not 0 <= i < n:
            raise LayoutError("IndexError", "Invalid index that exceeds shape length at axis")
        idx.append(i)
    shape = list(la.shape)
    shape[axis] = len(idx)
    lo = L.c_contig_layout(shape) if order == ROW_MAJOR else L.f_contig_layout(shape)
    out = np.zeros(max(lo.size, 1), dtype=a.dtype)
    index_select(out, lo, a, la, axis, idx)
    return out, lo


def _tri_pairs(n: int, uplo: str):
    """(i, j) of the packed triangle in packing order: row by row (cpu_serial/op_tri.rs:25-71)."""
    if uplo == "L":
        pairs = [(i, j) for i in range(n) for j in range(i + 1)]
    else:
        pairs = [(i, j) for i in range(n) for j in range(i, n)]
    I = np.array([p[0] for p in pairs], dtype=np.int64)
    J = np.array([p[1] for p in pairs], dtype=np.int64)
    return I, J


def pack_tri(a, la: Layout, b, lb: Layout, uplo: str, order: str = ROW_MAJOR):
    """OpPackTriAPI::pack_tri for DeviceCpuSerial (rstsr-core/src/device_cpu_serial/operators/op_tri.rs:8-27) +
    pack_tri_cpu_serial (rstsr-native-impl/src/cpu_serial/op_tri.rs:73-148): a (packed) <- b (full)."""
    if order != ROW_MAJOR:
        la, lb = la.reverse_axes(), lb.reverse_axes()
        uplo = "L" if uplo == "U" else "U"
    n = lb.shape[-1]
    if la.size == 0:
        return
    I, J = _tri_pairs(n, uplo)
    _wview(a, la)[...] = to_numpy(b, lb)[..., I, J]


def unpack_tri(a, la: Layout, b, lb: Layout, uplo: str, symm: str, order: str = ROW_MAJOR):
    """OpUnpackTriAPI::unpack_tri (device_cpu_serial/operators/op_tri.rs:34-52) + unpack_tri_cpu_serial
    (rstsr-native-impl/src/cpu_serial/op_tri.rs:152-522): a (full) <- b (packed).  symm N writes one triangle only."""
    if order != ROW_MAJOR:
        la, lb = la.reverse_axes(), lb.reverse_axes()
        uplo = "L" if uplo == "U" else "U"
    n = la.shape[-1]
    if la.size == 0:
        return
    I, J = _tri_pairs(n, uplo)
    va, vb = _wview(a, la), to_numpy(b, lb)
    off = I != J
    if symm in ("Sy", "He", "N"):
        va[..., I, J] = vb
        if symm != "N":
            va[..., J[off], I[off]] = vb[..., off]
    elif symm in ("Ay", "Ah"):
        va[..., I[off], J[off]] = vb[..., off]
        va[..., J[off], I[off]] = -vb[..., off]
        d = np.arange(n)
        va[..., d, d] = 0
    else:
        raise ValueError(symm)
