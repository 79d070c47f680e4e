"""GPU parity for the round-2 element types: half::f16, half::bf16, num::Complex<f32>, num::Complex<f64>
(DeviceComplexFloatAPI, rstsr-core/src/operators/combined_trait.rs:6-55).

Oracle = NumPy (float16, complex64 / complex128) and ml_dtypes.bfloat16, whose arithmetic is the `half` crate's: both
operands to f32, the op, ONE rounding back.  Complex products / quotients are restated from num-complex's `impl Mul` /
`impl Div` with separate real NumPy operations (NumPy's own complex loops may fuse or rescale).  Bit-exact for data
movement, casts, + - * /, comparisons, conj / real / imag / square; libm-style functions within a half-precision ulp resp.
1e-5 / 1e-13 for complex; reductions against an f64 / complex128 sum with the tolerance stated at the check."""
import numpy as np
import pytest

import rstsr_b200 as rt
from oracle import layout as L

from helpers import P, seed_of, upload

pytestmark = pytest.mark.gpu

BF16 = rt.bfloat16
HALF = [np.float16] + ([BF16] if BF16 is not None else [])
CPLX = [np.complex64, np.complex128]
ALL = HALF + CPLX


def _name(dt):
    return np.dtype(dt).name


def _data(rng, n, dt, positive=False):
    dt = np.dtype(dt)
    if dt.kind == "c":
        r = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        return r.astype(dt)
    x = rng.standard_normal(n).astype(np.float32)
    if positive:
        x = np.abs(x) + np.float32(0.25)
    return x.astype(dt)


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64,
                   16: np.dtype([("lo", np.uint64), ("hi", np.uint64)])}[a.dtype.itemsize])


def _same_bits(a, b):
    return np.array_equal(_bits(a), _bits(b))


@pytest.mark.parametrize("dt", ALL, ids=_name)
def test_data_movement_is_bit_exact(dev, dt):
    rng = np.random.default_rng(seed_of(("extmove", _name(dt))))
    a = _data(rng, 70 * 130, dt)
    t = rt.asarray(a, dev).reshape([70, 130])
    assert _same_bits(t.to_numpy(), a.reshape(70, 130))
    assert _same_bits(t.transpose([1, 0]).to_contig(rt.ROW_MAJOR).to_numpy(), a.reshape(70, 130).T)
    assert _same_bits(t[3:60:2, ::-1].to_contig(rt.COL_MAJOR).to_numpy(), a.reshape(70, 130)[3:60:2, ::-1])
    b = _data(rng, 12 * 66 * 68, dt).reshape(12, 66, 68)
    tb = rt.asarray(b.reshape(-1), dev).reshape([12, 66, 68])
    assert _same_bits(tb.transpose([2, 0, 1]).to_contig(rt.ROW_MAJOR).to_numpy(), b.transpose(2, 0, 1))
    idx = [int(i) for i in rng.integers(0, 70, 33)]
    assert _same_bits(t.index_select(0, idx).to_numpy(), a.reshape(70, 130)[idx])
    val = (1.5 - 2j) if np.dtype(dt).kind == "c" else 1.5
    assert _same_bits(rt.full([5, 7], val, dev, dtype=dt).to_numpy(), np.full((5, 7), val, dtype=dt))
    assert _same_bits(rt.ones([9], dev, dtype=dt).to_numpy(), np.ones(9, dtype=dt))
    assert _same_bits(rt.zeros([9], dev, dtype=dt).to_numpy(), np.zeros(9, dtype=dt))


def test_casts(dev):
    rng = np.random.default_rng(seed_of("extcast"))
    x64 = rng.standard_normal(5000) * np.exp(rng.uniform(-12, 12, 5000))
    x64[:6] = [0.0, -0.0, np.inf, -np.inf, np.nan, 65504.0]
    x32 = x64.astype(np.float32)
    pairs = [(np.float32, np.float16), (np.float64, np.float16), (np.float16, np.float32), (np.float16, np.float64)]
    if BF16 is not None:
        pairs += [(np.float32, BF16), (BF16, np.float32), (BF16, np.float64), (np.float16, BF16), (BF16, np.float16)]
    for src, dst in pairs:
        a = (x64 if np.dtype(src) == np.float64 else x32).astype(src)
        got = rt.asarray(a, dev).astype(dst).to_numpy()
        want = a.astype(np.float32).astype(dst) if np.dtype(src).itemsize == 2 else a.astype(dst)
        assert _same_bits(got, want) or np.array_equal(got, want, equal_nan=True), (src, dst)
    if BF16 is not None:  # f64 -> bf16 is ONE rounding (bf16::from_f64); ml_dtypes rounds through f32: allow the last bit
        got = rt.asarray(x64, dev).astype(BF16).to_numpy().astype(np.float64)
        want = x64.astype(np.float32).astype(BF16).astype(np.float64)
        fin = np.isfinite(want)
        assert np.all(np.abs(got[fin] - want[fin]) <= np.abs(want[fin]) * 2.0 ** -7)
        assert np.array_equal(np.isnan(got), np.isnan(want))
    for hdt in HALF:
        b = rng.integers(0, 2, 300).astype(np.bool_)
        assert _same_bits(rt.asarray(b, dev).astype(hdt).to_numpy(), b.astype(np.float32).astype(hdt))
        h = _data(rng, 300, hdt)
        h[:3] = np.array([0.0, -0.0, np.nan], dtype=np.float32).astype(hdt)
        assert np.array_equal(rt.asarray(h, dev).astype(np.bool_).to_numpy(), h.astype(np.float32) != 0)
    # real -> complex (a as R, 0) and complex <-> complex (componentwise `as`)
    for src, dst in ((np.float32, np.complex64), (np.float64, np.complex128), (np.float32, np.complex128), (np.float64, np.complex64),
                     (np.int32, np.complex128), (np.int64, np.complex128)):
        a = (rng.standard_normal(400) * 100).astype(src)
        got = rt.asarray(a, dev).astype(dst).to_numpy()
        rdt = np.float32 if np.dtype(dst) == np.complex64 else np.float64
        assert _same_bits(got.real.copy(), a.astype(rdt)) and not got.imag.any(), (src, dst)
    z = _data(rng, 400, np.complex128)
    assert _same_bits(rt.asarray(z, dev).astype(np.complex64).to_numpy(), z.astype(np.complex64))
    z32 = z.astype(np.complex64)
    assert _same_bits(rt.asarray(z32, dev).astype(np.complex128).to_numpy(), z32.astype(np.complex128))
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.asarray(z, dev).astype(np.float64)   # the reference has no complex -> real DTypeCastAPI either
    assert e.value.kind == "UnImplemented"


def _complex_of(re, im, dt):
    out = np.empty(np.broadcast(re, im).shape, dtype=dt)  # componentwise: `re + 1j * im` would lose the sign of a zero
    out.real, out.imag = re, im
    return out


def _cmul(a, b):
    rdt = a.real.dtype
    re = (a.real * b.real).astype(rdt) - (a.imag * b.imag).astype(rdt)
    im = (a.real * b.imag).astype(rdt) + (a.imag * b.real).astype(rdt)
    return _complex_of(re, im, a.dtype)


def _cdiv(a, b):
    rdt = a.real.dtype
    n = (b.real * b.real).astype(rdt) + (b.imag * b.imag).astype(rdt)
    re = ((a.real * b.real).astype(rdt) + (a.imag * b.imag).astype(rdt)) / n
    im = ((a.imag * b.real).astype(rdt) - (a.real * b.imag).astype(rdt)) / n
    return _complex_of(re.astype(rdt), im.astype(rdt), a.dtype)


@pytest.mark.parametrize("dt", ALL, ids=_name)
def test_arithmetic_bit_exact(dev, dt):
    rng = np.random.default_rng(seed_of(("extarith", _name(dt))))
    a, b = _data(rng, 48 * 80, dt).reshape(48, 80), _data(rng, 80, dt)
    ta, tb = rt.asarray(a.reshape(-1), dev).reshape([48, 80]), rt.asarray(b, dev)
    cplx = np.dtype(dt).kind == "c"
    f32 = lambda x: x.astype(np.float32)
    want = {"add": (a + b) if cplx else (f32(a) + f32(b)).astype(dt), "sub": (a - b) if cplx else (f32(a) - f32(b)).astype(dt),
            "mul": _cmul(a, np.broadcast_to(b, a.shape)) if cplx else (f32(a) * f32(b)).astype(dt),
            "div": _cdiv(a, np.broadcast_to(b, a.shape)) if cplx else (f32(a) / f32(b)).astype(dt)}
    for op, w in want.items():
        got = ta.binary(op, tb)
        assert got.dtype == np.dtype(dt)
        assert _same_bits(got.to_numpy(), w), (dt, op)
    # transposed operand (tile kernel with 2- / 8- / 16-byte elements), scalar operand, in place
    sq = _data(rng, 96 * 96, dt).reshape(96, 96)
    tsq = rt.asarray(sq.reshape(-1), dev).reshape([96, 96])
    w = (sq + sq.T) if cplx else (f32(sq) + f32(sq.T)).astype(dt)
    assert _same_bits((tsq + tsq.transpose([1, 0])).to_numpy(), w)
    s = np.array([0.75 - 0.5j if cplx else 0.75]).astype(dt)[0]
    w = _cmul(a, np.full_like(a, s)) if cplx else (f32(a) * np.float32(s)).astype(dt)
    assert _same_bits((ta * s).to_numpy(), w)
    tc = rt.asarray(a.reshape(-1).copy(), dev).reshape([48, 80])
    tc -= tb
    assert _same_bits(tc.to_numpy(), want["sub"])
    assert _same_bits((-ta).to_numpy(), -a)
    # comparisons
    eq = ta.binary("eq", ta).to_numpy()
    assert eq.dtype == np.bool_ and eq.all()
    assert np.array_equal(ta.binary("ne", tb).to_numpy(), a != b)
    if not cplx:
        for op, fn in (("lt", np.less), ("le", np.less_equal), ("gt", np.greater), ("ge", np.greater_equal)):
            assert np.array_equal(ta.binary(op, tb).to_numpy(), fn(f32(a), f32(b))), (dt, op)
        assert _same_bits(ta.binary("maximum", tb).to_numpy(), np.fmax(f32(a), f32(b)).astype(dt))
        assert _same_bits(ta.binary("minimum", tb).to_numpy(), np.fmin(f32(a), f32(b)).astype(dt))
    else:
        with pytest.raises(rt.RstsrCudaError) as e:
            ta.binary("lt", tb)
        assert e.value.kind == "UnImplemented"


@pytest.mark.parametrize("dt", HALF, ids=_name)
def test_half_math_functions(dev, dt):
    rng = np.random.default_rng(seed_of(("halfmath", _name(dt))))
    x = _data(rng, 3000, dt, positive=True)
    t = rt.asarray(x, dev)
    ulp = 2.0 ** -9 if np.dtype(dt) == np.float16 else 2.0 ** -6
    x32 = x.astype(np.float32)
    for op, fn in (("sqrt", np.sqrt), ("exp", np.exp), ("log", np.log), ("sin", np.sin), ("cos", np.cos), ("tanh", np.tanh),
                   ("reciprocal", lambda v: np.float32(1) / v), ("floor", np.floor), ("abs", np.abs), ("square", np.square)):
        got = t.unary(op).to_numpy().astype(np.float64)
        want = fn(x32).astype(dt).astype(np.float64)
        assert np.allclose(got, want, rtol=ulp, atol=1e-7), (dt, op)
    assert np.array_equal(t.unary("isnan").to_numpy(), np.isnan(x32))
    pw = t.binary("pow", rt.asarray(_data(rng, 3000, dt), dev)).to_numpy().astype(np.float64)
    assert np.all(np.isfinite(pw))


@pytest.mark.parametrize("dt", CPLX, ids=_name)
def test_complex_unary(dev, dt):
    rng = np.random.default_rng(seed_of(("cplxun", _name(dt))))
    z = _data(rng, 2500, dt)
    t = rt.asarray(z, dev)
    rdt = z.real.dtype
    tol = 2e-6 if rdt == np.float32 else 1e-14
    a = t.unary("abs")
    assert a.dtype == rdt and np.allclose(a.to_numpy(), np.hypot(z.real, z.imag), rtol=tol, atol=0)
    assert _same_bits(t.unary("real").to_numpy(), z.real.copy()) and _same_bits(t.unary("imag").to_numpy(), z.imag.copy())
    assert _same_bits(t.unary("conj").to_numpy(), np.conj(z))
    assert _same_bits(t.unary("square").to_numpy(), _cmul(z, z))
    n = (z.real * z.real).astype(rdt) + (z.imag * z.imag).astype(rdt)
    assert _same_bits(t.unary("reciprocal").to_numpy(), ((z.real / n).astype(rdt) + 1j * (-z.imag / n).astype(rdt)).astype(dt))
    loose = 3e-5 if rdt == np.float32 else 1e-12
    for op, fn in (("exp", np.exp), ("log", np.log), ("sqrt", np.sqrt), ("sin", np.sin), ("cos", np.cos), ("sinh", np.sinh),
                   ("cosh", np.cosh), ("tanh", np.tanh)):
        got = t.unary(op).to_numpy().astype(np.complex128)
        want = fn(z.astype(np.complex128))
        assert np.allclose(got, want, rtol=loose, atol=loose), (dt, op)
    with pytest.raises(rt.RstsrCudaError):
        t.unary("floor")


@pytest.mark.parametrize("dt", ALL, ids=_name)
def test_reductions(dev, dt):
    rng = np.random.default_rng(seed_of(("extred", _name(dt))))
    cplx = np.dtype(dt).kind == "c"
    a = _data(rng, 300 * 257, dt).reshape(300, 257)
    t = rt.asarray(a.reshape(-1), dev).reshape([300, 257])
    wide = a.astype(np.complex128 if cplx else np.float64)
    # one rounding to the element type on top of an f32 / element-precision accumulation
    eps = {"float16": 2.0 ** -9, "bfloat16": 2.0 ** -6, "complex64": 1e-5, "complex128": 1e-12}[_name(dt)]
    for axes in ([0], [1], [0, 1]):
        got = t.sum_axes(axes).to_numpy().astype(wide.dtype)
        want = wide.sum(axis=tuple(axes))
        l1 = np.abs(wide).sum(axis=tuple(axes))
        assert np.all(np.abs(got - want) <= eps * np.maximum(l1, np.abs(want)) + 1e-30), (dt, axes, "sum")
        got = t.mean_axes(axes).to_numpy().astype(wide.dtype)
        n = np.prod([a.shape[i] for i in axes])
        assert np.all(np.abs(got - want / n) <= eps * np.maximum(l1 / n, np.abs(want / n)) + 1e-30), (dt, axes, "mean")
    s = t.sum_all()
    assert abs(complex(s) - complex(wide.sum())) <= eps * max(np.abs(wide).sum(), 1.0)
    sm = (1 + 0.01 * _data(rng, 40, dt).astype(wide.dtype)).astype(dt)
    p = complex(rt.asarray(sm, dev).prod_all())
    pw = complex(np.prod(sm.astype(wide.dtype)))
    assert abs(p - pw) <= 40 * eps * abs(pw)
    if not cplx:
        for axes in ([0], [1]):
            assert _same_bits(t.max_axes(axes).to_numpy(), a.astype(np.float32).max(axis=axes[0]).astype(dt))
            assert _same_bits(t.min_axes(axes).to_numpy(), a.astype(np.float32).min(axis=axes[0]).astype(dt))
        assert float(t.max_all()) == float(a.astype(np.float32).max())
    else:
        with pytest.raises(rt.RstsrCudaError) as e:
            t.max_all()
        assert e.value.kind == "UnImplemented"


@pytest.mark.parametrize("dt", ALL, ids=_name)
def test_var_std_l2_and_arg_reductions(dev, dt):
    """var / std / l2_norm (REAL output for complex: TOut = T::Real, auto_impl/reduction.rs:207-354); for the half types
    also argmin / argmax / count_nonzero.  Values against f64 / complex128 NumPy with the element type's rounding."""
    rng = np.random.default_rng(seed_of(("extvar", _name(dt))))
    cplx = np.dtype(dt).kind == "c"
    a = _data(rng, 120 * 96, dt).reshape(120, 96)
    t = rt.asarray(a.reshape(-1), dev).reshape([120, 96])
    wide = a.astype(np.complex128 if cplx else np.float64)
    out_dt = a.real.dtype if cplx else np.dtype(dt)
    eps = {"float16": 2.0 ** -8, "bfloat16": 2.0 ** -5, "complex64": 2e-5, "complex128": 1e-12}[_name(dt)]
    for axes in ([0], [1], [0, 1]):
        ax = tuple(axes)
        n = np.prod([a.shape[i] for i in axes])
        want_var = (np.abs(wide) ** 2).sum(axis=ax) / n - np.abs(wide.sum(axis=ax) / n) ** 2
        got = t.var_axes(axes)
        assert got.dtype == out_dt
        assert np.allclose(got.to_numpy().astype(np.float64), want_var, rtol=4 * eps, atol=4 * eps), (dt, axes, "var")
        assert np.allclose(t.std_axes(axes).to_numpy().astype(np.float64), np.sqrt(want_var), rtol=4 * eps, atol=4 * eps), (dt, axes, "std")
        want_l2 = np.sqrt((np.abs(wide) ** 2).sum(axis=ax))
        got = t.l2_norm_axes(axes)
        assert got.dtype == out_dt and np.allclose(got.to_numpy().astype(np.float64), want_l2, rtol=2 * eps, atol=0), (dt, axes, "l2")
    assert abs(float(t.l2_norm_all()) - float(np.sqrt((np.abs(wide) ** 2).sum()))) <= 2 * eps * float(np.sqrt((np.abs(wide) ** 2).sum()))
    if not cplx:
        a32 = a.astype(np.float32)
        assert np.array_equal(t.argmax_axes([1]).to_numpy(), a32.argmax(axis=1).astype(np.uint64))
        assert np.array_equal(t.argmin_axes([0]).to_numpy(), a32.argmin(axis=0).astype(np.uint64))
        assert int(t.argmax_all()) == int(a32.argmax())
        z = a.copy()
        z[::3] = np.zeros(1, dtype=dt)[0]
        assert int(rt.asarray(z.reshape(-1), dev).reshape([120, 96]).count_nonzero_all()) == int(np.count_nonzero(z.astype(np.float32)))
    else:
        with pytest.raises(rt.RstsrCudaError) as e:
            t.argmax_all()
        assert e.value.kind == "UnImplemented"


def test_promotion_with_complex_and_half_operands(dev):
    """Mixed operand types through the reference's table (promotion.rs:195-200, :368-545): both operands are brought to
    promote(ta, tb) -- primitive -> Complex<R> is (v as R, 0) -- and the one-type kernel runs: bit-exact against the same
    two steps in NumPy."""
    rng = np.random.default_rng(seed_of("extpromote"))
    n = 3000
    z32, z64 = _data(rng, n, np.complex64), _data(rng, n, np.complex128)
    prim = {np.float32: rng.standard_normal(n).astype(np.float32), np.float64: rng.standard_normal(n),
            np.int8: rng.integers(-100, 100, n).astype(np.int8), np.uint16: rng.integers(1, 60000, n).astype(np.uint16),
            np.int32: rng.integers(-10**6, 10**6, n).astype(np.int32), np.int64: rng.integers(-10**12, 10**12, n),
            np.uint64: rng.integers(1, 2**63, n).astype(np.uint64), np.bool_: rng.integers(0, 2, n).astype(np.bool_)}
    up = lambda a: rt.asarray(a, dev)
    for zdt, z in ((np.complex64, z32), (np.complex128, z64)):
        for pdt, p in prim.items():
            k = rt.DeviceCuda.promote_types(zdt, pdt)
            small = np.dtype(pdt).kind == "b" or np.dtype(pdt) in (np.dtype(np.int8), np.dtype(np.uint16), np.dtype(np.float32))
            assert k == np.dtype(zdt if (np.dtype(zdt) == np.complex128 or small) else np.complex128), (zdt, pdt)
            rdt = np.float32 if k == np.complex64 else np.float64
            zk, pk = z.astype(k), _complex_of(p.astype(rdt), np.zeros(n, rdt), k)
            got = up(z) + up(p)
            assert got.dtype == k and _same_bits(got.to_numpy(), zk + pk), (zdt, pdt, "add")
            got = up(p) - up(z)
            assert got.dtype == k and _same_bits(got.to_numpy(), pk - zk), (zdt, pdt, "sub")
            assert _same_bits((up(z) * up(p)).to_numpy(), _cmul(zk, pk)), (zdt, pdt, "mul")
            assert np.array_equal(up(z).binary("ne", up(p)).to_numpy(), zk != pk)
    assert _same_bits((up(z32) / up(z64)).to_numpy(), _cdiv(z32.astype(np.complex128), z64))
    # broadcast operand and a transposed view on the promoted side
    col = prim[np.int32][:50]
    m = z32[:50 * 60].reshape(50, 60)
    got = rt.asarray(m.reshape(-1), dev).reshape([50, 60]).transpose([1, 0]) + up(col)
    assert got.dtype == np.complex128 and _same_bits(got.to_numpy(), m.T.astype(np.complex128) + col.astype(np.float64))
    for hdt in HALF:  # bool x half -> half; half x anything else is not in the reference's table
        h, b = _data(rng, n, hdt), prim[np.bool_]
        got = up(h) * up(b)
        assert got.dtype == np.dtype(hdt) and _same_bits(got.to_numpy(), (h.astype(np.float32) * b.astype(np.float32)).astype(hdt))
        with pytest.raises(rt.RstsrCudaError) as e:
            up(h) + up(prim[np.float32])
        assert e.value.kind == "UnImplemented"
    # staged casts: primitive -> complex (v as R, 0), integer <-> half through f64 / f32
    for pdt, p in prim.items():
        for zdt, rdt in ((np.complex64, np.float32), (np.complex128, np.float64)):
            got = up(p).astype(zdt).to_numpy()
            assert _same_bits(got.real.copy(), p.astype(rdt)) and not got.imag.any(), (pdt, zdt)
    i = rng.integers(-70000, 70000, n).astype(np.int32)
    with np.errstate(over="ignore"):
        want = i.astype(np.float64).astype(np.float16)
    assert _same_bits(up(i).astype(np.float16).to_numpy(), want)
    h = (rng.standard_normal(n) * 300).astype(np.float16)
    assert np.array_equal(up(h).astype(np.int32).to_numpy(), h.astype(np.float32).astype(np.int32))


@pytest.mark.parametrize("dt", CPLX, ids=_name)
def test_complex_inverse_functions_and_predicates(dev, dt):
    """tan, the inverse trigonometric / hyperbolic functions, log2 / log10 and is_nan / is_infinite / is_finite of complex
    numbers (ComplexFloat rows of auto_impl/op_binary_common.rs:10-35, :97-102).  libm-style: 1e-5 / 1e-12 relative
    against NumPy's C99 functions (same principal branches as num-complex)."""
    rng = np.random.default_rng(seed_of(("cinv", _name(dt))))
    z = _data(rng, 4000, dt)
    t = rt.asarray(z, dev)
    tol = 2e-5 if np.dtype(dt) == np.complex64 else 1e-12
    fns = {"tan": np.tan, "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan, "asinh": np.arcsinh, "acosh": np.arccosh,
           "atanh": np.arctanh, "log2": np.log2, "log10": np.log10}
    for op, fn in fns.items():
        got = t.unary(op)
        assert got.dtype == np.dtype(dt)
        want = fn(z.astype(np.complex128))
        err = np.abs(got.to_numpy().astype(np.complex128) - want) / np.maximum(np.abs(want), 1.0)
        assert err.max() <= tol, (dt, op, err.max())
    w = z.copy()
    w[:6] = [complex(np.nan, 1), complex(1, np.nan), complex(np.inf, 0), complex(0, -np.inf), complex(np.inf, np.nan), 0]
    tw = rt.asarray(w, dev)
    nan = np.isnan(w.real) | np.isnan(w.imag)
    inf = ~nan & (np.isinf(w.real) | np.isinf(w.imag))
    assert np.array_equal(tw.unary("isnan").to_numpy(), nan)
    assert np.array_equal(tw.unary("isinf").to_numpy(), inf)
    assert np.array_equal(tw.unary("isfinite").to_numpy(), np.isfinite(w.real) & np.isfinite(w.imag))


@pytest.mark.parametrize("dt", ALL, ids=_name)
def test_vecdot_and_allclose(dev, dt):
    """vecdot = sum conj(a) * b (cpu_serial/vecdot.rs:96-157) and allclose_all with TE = f64 (isclose.rs:92-106) for the
    half and complex types: against complex128 / f64 sums, tolerance = rounding of the element type."""
    rng = np.random.default_rng(seed_of(("extdot", _name(dt))))
    a, b = _data(rng, 64 * 300, dt).reshape(64, 300), _data(rng, 64 * 300, dt).reshape(64, 300)
    ta, tb = rt.asarray(a.reshape(-1), dev).reshape([64, 300]), rt.asarray(b.reshape(-1), dev).reshape([64, 300])
    cplx = np.dtype(dt).kind == "c"
    wide = np.complex128 if cplx else np.float64
    tol = {"float16": 2e-3, "bfloat16": 2e-2, "complex64": 2e-5, "complex128": 1e-12}[_name(dt)]
    for axis, npaxis in ((-1, 1), (0, 0)):
        got = rt.vecdot(ta, tb, axis=axis)
        assert got.dtype == np.dtype(dt)
        want = np.sum(np.conj(a.astype(wide)) * b.astype(wide), axis=npaxis)
        scale = np.sum(np.abs(a.astype(wide)) * np.abs(b.astype(wide)), axis=npaxis)
        assert np.all(np.abs(got.to_numpy().astype(wide) - want) <= tol * scale), (dt, axis)
    assert rt.allclose(ta, ta)
    c = a.copy()
    c[17, 123] = c[17, 123] * dt(1.5) + dt(1)
    tc = rt.asarray(c.reshape(-1), dev).reshape([64, 300])
    assert not rt.allclose(ta, tc)
    assert rt.allclose(ta, tc, rtol=10.0, atol=10.0)
    n = a.copy()
    n[3, 4] = np.nan
    tn = rt.asarray(n.reshape(-1), dev).reshape([64, 300])
    assert not rt.allclose(tn, tn)
    assert rt.allclose(tn, tn, equal_nan=True)


def test_creation_of_half_and_complex(dev):
    """linspace in the element type's own arithmetic (start + T::from(i) * step, cpu_rayon/creation.rs:108-131), the serial
    arange recurrence of the 8- / 16-bit types (cpu_serial/creation.rs:7-19) and tril on 16-byte elements."""
    # complex: (end - start) / Complex::from(n - 1) is the textbook quotient, i * step the textbook product
    for dt, rdt in ((np.complex64, np.float32), (np.complex128, np.float64)):
        s, e, n = dt(1.0 + 2.0j), dt(3.5 + 4.7j), 10          # the reference's own linspace_impl test values
        got = dev.to_cpu_vec(dev.linspace_impl(s, e, n, True, dt))
        step = _cdiv(np.array([e - s], dtype=dt), np.array([n - 1 + 0j], dtype=dt))
        want = np.array([s + _cmul(np.array([i + 0j], dtype=dt), step)[0] for i in range(n)], dtype=dt)
        assert _same_bits(got, want), dt
        assert got[0] == s and abs(got[-1] - e) <= 4 * np.finfo(rdt).eps * abs(e)
        got = dev.to_cpu_vec(dev.linspace_impl(s, e, 7, False, dt))
        assert len(got) == 7 and abs(got[-1] - (s + (e - s) * 6 / 7)) <= 8 * np.finfo(rdt).eps * abs(e)
    for hdt in HALF:
        f = lambda x: np.float32(x)
        s, e, n = hdt(-2.0), hdt(3.0), 33
        got = dev.to_cpu_vec(dev.linspace_impl(s, e, n, True, hdt))
        step = hdt(f(hdt(f(e) - f(s))) / f(hdt(f(n - 1))))
        want = np.array([hdt(f(s) + f(hdt(f(hdt(f(i))) * f(step)))) for i in range(n)], dtype=hdt)
        assert _same_bits(got, want), hdt
        # arange: current = current + step rounded to the half type at every step
        got = dev.to_cpu_vec(dev.arange_impl(hdt(0.0), hdt(2.0), hdt(0.1), hdt))
        want, cur = [], hdt(0.0)
        while f(cur) < f(hdt(2.0)):
            want.append(cur)
            cur = hdt(f(cur) + f(hdt(0.1)))
        assert _same_bits(got, np.array(want, dtype=hdt)), hdt
    for idt in (np.int8, np.uint8, np.int16, np.uint16):
        got = dev.to_cpu_vec(dev.arange_impl(3, 100, 7, idt))
        assert got.dtype == np.dtype(idt) and np.array_equal(got, np.arange(3, 100, 7, dtype=idt))
    assert len(dev.to_cpu_vec(dev.arange_impl(5, 5, 1, np.int16))) == 0
    z = _data(np.random.default_rng(5), 6 * 9, np.complex128).reshape(6, 9)
    raw = upload(dev, z.reshape(-1).copy())
    dev.tril_impl(raw, rt.Layout((6, 9), (9, 1)), 1)
    assert _same_bits(dev.to_cpu_vec(raw).reshape(6, 9), np.tril(z, 1))


@pytest.mark.parametrize("dt", ALL, ids=_name)
def test_elementwise_isclose_and_complex_sign(dev, dt):
    """OpIsCloseAPI on half / complex operands (TE = f64; |a - b| and |b| formed in the element type) and ext_sign of
    complex numbers: z / |z|, zero for zero (ext_num.rs:268-280)."""
    rng = np.random.default_rng(seed_of(("extclose", _name(dt))))
    a = _data(rng, 3000, dt)
    b = a.copy()
    b[::3] = (b[::3].astype(np.complex128 if np.dtype(dt).kind == "c" else np.float64) * 1.01).astype(dt)
    b[5] = np.nan
    a[5] = np.nan
    rtol, atol = 5e-3, 1e-3
    wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    if np.dtype(dt).kind == "c":
        diff = np.abs((a - b)).astype(np.float64)
    else:
        diff = np.abs((a.astype(np.float32) - b.astype(np.float32)).astype(dt).astype(np.float64))
    with np.errstate(invalid="ignore"):
        want = diff <= atol + rtol * np.abs(b.astype(wide))
    for equal_nan in (False, True):
        w = want.copy()
        w[5] = equal_nan
        got = rt.isclose(rt.asarray(a, dev), rt.asarray(b, dev), rtol=rtol, atol=atol, equal_nan=equal_nan).to_numpy()
        # entries within one rounding of the threshold may fall either way (hypot / the half rounding of a - b)
        edge = np.abs(diff - (atol + rtol * np.abs(b.astype(wide)))) <= 1e-3 * (atol + rtol * np.abs(b.astype(wide)))
        edge[5] = False
        assert np.array_equal(got[~edge], w[~edge]), (dt, equal_nan)
    if np.dtype(dt).kind == "c":
        z = a.copy()
        z[5] = 0
        z[6] = complex(0.0, -2.0)
        got = rt.asarray(z, dev).unary("sign").to_numpy()
        n = np.abs(z)
        want = np.where(n == 0, 0, z / np.where(n == 0, 1, n)).astype(dt)
        assert got[5] == 0 and np.allclose(got, want, rtol=4 * np.finfo(z.real.dtype).eps, atol=0)
