"""TEST INFRASTRUCTURE (CPU oracle) -- operand-type promotion of the binary-function traits.

Restates, for checking only:
  * DTypePromoteAPI  rstsr-dtype-traits/src/promotion.rs:23-57 (trait), :123-181 (bool x T -> T),
                     :186-273 (`impl_promotion_asable!(T1, T2, .., Res)` table; isize = i64, usize = u64)
  * DTypeIntoFloatAPI promotion.rs:62-118 (integers -> f64, floats stay)
  * the per-op rules of rstsr-core/src/feature_rayon/auto_impl/op_ternary_common.rs:
        :6-73    atan2 copysign hypot nextafter logaddexp : promote_pair, into_float, f   -> TOut = float type
        :75-139  maximum minimum floor_divide == != > >= < <= : promote_pair, f           -> TOut = Res / bool
        :141-189 pow : `TA: num::Pow<TB>`, TOut = TA::Output
  * powi: Rust f32::powi / f64::powi lower to compiler-rt's __powisf2 / __powidf2 (square-and-multiply, one reciprocal
    at the end for negative exponents) -- restated so values can be compared bit for bit.
The table below is written out rule by rule (it is pinned against the reference's own macro lines by
tests/golden/promotion_table.json, generated with scripts/gen_promotion_golden.py).
"""
from __future__ import annotations

import numpy as np

_BITS = {"i8": 8, "i16": 16, "i32": 32, "i64": 64, "u8": 8, "u16": 16, "u32": 32, "u64": 64}
NAMES = ["bool", "i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64", "f32", "f64"]
NP = {"bool": np.bool_, "i8": np.int8, "i16": np.int16, "i32": np.int32, "i64": np.int64, "u8": np.uint8,
      "u16": np.uint16, "u32": np.uint32, "u64": np.uint64, "f32": np.float32, "f64": np.float64}
NAME_OF = {np.dtype(v): k for k, v in NP.items()}


def name_of(dt) -> str:
    return NAME_OF[np.dtype(dt)]


EXT_NAMES = ["f16", "bf16", "c32", "c64"]
NP_EXT = {"f16": np.float16, "c32": np.complex64, "c64": np.complex128}  # bf16: ml_dtypes.bfloat16 where installed


def promote(a: str, b: str) -> str:
    """<a as DTypePromoteAPI<b>>::Res"""
    if a == b:
        return a                                    # promotion.rs:41-55
    if a == "bool":
        return b                                    # :123-141 (every T incl. f16 / bf16 / c32 / c64: :195-200)
    if b == "bool":
        return a                                    # :143-157
    if a in EXT_NAMES or b in EXT_NAMES:
        ca, cb = a in ("c32", "c64"), b in ("c32", "c64")
        if ca and cb:
            return "c64"                            # c32 x c64 (:516-545)
        if (ca or cb) and a not in ("f16", "bf16") and b not in ("f16", "bf16"):
            c, p = (a, b) if ca else (b, a)
            if c == "c64":
                return "c64"                        # Complex<f64> x any primitive (:412-423, :493-505)
            # Complex<f32> keeps f32 components for what f32 holds exactly (:406-410), else widens (:425-431)
            return "c32" if p in ("i8", "i16", "u8", "u16", "f32") else "c64"
        raise TypeError(f"DTypePromoteAPI<{b}> is not implemented for {a}")  # the half types pair with bool only
    fa, fb = a[0] == "f", b[0] == "f"
    if fa and fb:
        return "f64"                                # f32 x f64 (:266, :275)
    if fa or fb:
        f, i = (a, b) if fa else (b, a)
        if f == "f64":
            return "f64"
        return "f32" if _BITS[i] <= 16 else "f64"   # f32 x i8/i16/u8/u16 -> f32, wider ints -> f64 (:258-265)
    sa, sb = a[0] == "i", b[0] == "i"
    ba, bb = _BITS[a], _BITS[b]
    if sa == sb:
        return ("i" if sa else "u") + str(max(ba, bb))
    bs, bu = (ba, bb) if sa else (bb, ba)
    if bs > bu:
        return "i" + str(bs)                        # i16 x u8 -> i16
    if bu == 64:
        return "f64"                                # i64 x u64 -> f64
    return "i" + str(2 * bu)                        # i8 x u8 -> i16, i32 x u32 -> i64


def into_float(t: str) -> str:
    if t in ("f32", "f64") or t in EXT_NAMES:
        return t                                    # promotion.rs:85-101
    if t == "bool":
        raise TypeError("DTypeIntoFloatAPI is not implemented for bool")
    return "f64"


FLOAT_FUNCS = {"atan2", "copysign", "hypot", "nextafter", "logaddexp"}
COMPARES = {"eq", "ne", "lt", "le", "gt", "ge"}


def pow_kind(ta: str, tb: str) -> str:
    if ta[0] == "f":
        if tb == ta:
            return "same"
        if tb in ("i8", "u8", "i16", "u16", "i32"):
            return "powi"
    elif ta[0] in "iu" and tb[0] == "u":
        return "ipow"
    raise TypeError(f"num::Pow<{tb}> is not implemented for {ta}")


def op_types(op: str, ta: str, tb: str):
    """(compute type both operands are brought to, output type)"""
    if op == "pow":
        if ta == tb and ta in EXT_NAMES:
            return ta, ta                           # num_traits::Float::powf / Complex::powc of one type
        pow_kind(ta, tb)
        return ta, ta
    r = promote(ta, tb)
    if op in FLOAT_FUNCS:
        k = into_float(r)
        return k, k
    return r, ("bool" if op in COMPARES else r)


def powi(a: np.ndarray, n: np.ndarray) -> np.ndarray:
    """__powisf2 / __powidf2 elementwise, in a's precision."""
    a = np.array(a, copy=True)
    out = np.empty_like(a)
    one = a.dtype.type(1)
    flat_a, flat_n, flat_o = a.reshape(-1), np.broadcast_to(n, a.shape).reshape(-1), out.reshape(-1)
    for i in range(flat_a.size):
        x, b = flat_a[i], int(np.int32(flat_n[i]))
        recip = b < 0
        r = one
        while True:
            if b & 1:
                r = a.dtype.type(r * x)
            b = int(b / 2)  # C division: toward zero
            if b == 0:
                break
            x = a.dtype.type(x * x)
        flat_o[i] = one / r if recip else r
    return out


def ipow(a: np.ndarray, e: np.ndarray) -> np.ndarray:
    """release-mode {integer}::pow: wrapping power."""
    bits = a.dtype.itemsize * 8
    mask = (1 << bits) - 1
    flat_a = a.reshape(-1)
    flat_e = np.broadcast_to(e, a.shape).reshape(-1)
    out = np.empty_like(flat_a)
    for i in range(flat_a.size):
        v = pow(int(flat_a[i]) & mask, int(flat_e[i]) & 0xFFFFFFFF, 1 << bits)
        if a.dtype.kind == "i" and v >= 1 << (bits - 1):
            v -= 1 << bits
        out[i] = v
    return out.reshape(a.shape)


def cast(x: np.ndarray, t: str) -> np.ndarray:
    """The promoting `as` casts are value-preserving (or round-to-nearest for i64/u64 -> f64): numpy's astype agrees."""
    if np.dtype(NP[t]) == x.dtype:
        return x
    if x.dtype == np.bool_:
        return x.astype(np.uint8).astype(NP[t])
    return x.astype(NP[t])
