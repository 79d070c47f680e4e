"""GPU parity: elementwise ops through the C ABI vs the CPU oracle on the same seeded inputs.
Bit-exact for integer and IEEE (+ - * / %, min/max, comparisons, copysign, nextafter, sqrt, floor...) results;
libm-style functions (exp, sin, pow, ...) within 4 ulp-ish relative tolerance (CUDA libm vs glibc)."""
import numpy as np
import pytest

import oracle
import rstsr_b200 as rt
from oracle import layout as L

from helpers import O, P, rand_data, random_view, same, seed_of, upload, view_np

pytestmark = pytest.mark.gpu

EXACT_BIN = ["add", "sub", "mul", "div", "rem", "maximum", "minimum", "floor_divide"]
BIT_BIN = ["bitor", "bitand", "bitxor", "shl", "shr"]
CMP_BIN = ["eq", "ne", "lt", "le", "gt", "ge"]
FLOAT_EXACT_BIN = ["copysign", "nextafter"]
FLOAT_TOL_BIN = ["pow", "atan2", "hypot", "logaddexp"]
INT_DTYPES = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64]
FLT_DTYPES = [np.float32, np.float64]


def _operands(rng, dtype, shape_a, shape_b=None):
    """Two random views whose shapes broadcast (row-major rule): b drops/unit-izes random axes of a."""
    la, na = random_view(rng)
    return la, na


def _bcast_pair(rng, order):
    """Random (la, lb) that broadcast under `order`, with their buffer sizes."""
    la, na = random_view(rng, max_ndim=4, max_extent=6)
    # derive b's shape from a's: drop leading axes (row-major) / trailing (col-major), set some extents to 1
    shape = list(la.shape)
    k = int(rng.integers(0, len(shape) + 1))
    shape_b = shape[k:] if order == L.ROW_MAJOR else shape[:len(shape) - k]
    shape_b = [1 if rng.random() < 0.3 else d for d in shape_b]
    base = L.c_contig_layout(shape_b) if rng.random() < 0.5 else L.f_contig_layout(shape_b)
    perm = [int(p) for p in rng.permutation(len(shape_b))]
    # permute storage order but keep the logical shape: build strides for a permuted storage
    stor = L.c_contig_layout([shape_b[p] for p in perm])
    inv = [perm.index(i) for i in range(len(shape_b))]
    lb = stor.transpose(inv) if shape_b else base
    for ax in range(lb.ndim):
        if rng.random() < 0.2:
            lb = lb.narrow(ax, slice(None, None, -1))
    nb = 1
    for d in shape_b:
        nb *= d
    return la, na, lb, max(nb, 1)


def _nonzero(x):
    x = x.copy()
    x[x == 0] = 1
    return x


def run_binary(dev, op, dtype, la, a, lb, b, order):
    """`&a o &b` on the device and in the oracle; returns (got, want, lc_got, lc_want)."""
    ta = rt.Tensor(upload(dev, a), P(la))
    tb = rt.Tensor(upload(dev, b), P(lb))
    tc = ta.binary(op, tb)
    la_b, lb_b = L.broadcast_layout(la, lb, order)
    lc = O(tc.layout)  # the product's layout choice is checked separately (function ops use another rule)
    c = np.zeros(max(L.bounds_index(lc)[1], 1), dtype=np.bool_ if op in CMP_BIN else dtype)
    cc = c.view(np.uint8) if c.dtype == np.bool_ else c
    aa = a.view(np.uint8) if a.dtype == np.bool_ else a
    bb = b.view(np.uint8) if b.dtype == np.bool_ else b
    oracle.op_mutc_refa_refb(op, cc, lc, aa, la_b, bb, lb_b)
    return tc.to_numpy(), view_np(c, lc), tc.layout


@pytest.mark.parametrize("dtype", INT_DTYPES + FLT_DTYPES)
@pytest.mark.parametrize("op", EXACT_BIN)
def test_binary_exact(dev, op, dtype):
    rng = np.random.default_rng(seed_of((op, np.dtype(dtype).name)))
    for _ in range(12):
        la, na, lb, nb = _bcast_pair(rng, L.ROW_MAJOR)
        a, b = rand_data(rng, na, dtype), rand_data(rng, nb, dtype)
        if op in ("div", "rem", "floor_divide") and np.dtype(dtype).kind in "iu":
            b = _nonzero(b)  # Rust panics on integer division by zero: not a defined result
        got, want, lc = run_binary(dev, op, dtype, la, a, lb, b, L.ROW_MAJOR)
        if op in EXACT_BIN[:5]:
            la_b, lb_b = L.broadcast_layout(la, lb, L.ROW_MAJOR)
            assert same(lc, L.get_layout_for_binary_op(la_b, lb_b, L.ROW_MAJOR))
        assert got.dtype == want.dtype and got.shape == want.shape
        assert np.array_equal(got, want, equal_nan=True), (op, dtype, la, lb)


@pytest.mark.parametrize("dtype", INT_DTYPES)
@pytest.mark.parametrize("op", BIT_BIN)
def test_binary_bit_ops(dev, op, dtype):
    rng = np.random.default_rng(seed_of((op, np.dtype(dtype).name)))
    for _ in range(8):
        la, na, lb, nb = _bcast_pair(rng, L.ROW_MAJOR)
        a, b = rand_data(rng, na, dtype), rand_data(rng, nb, dtype)
        if op in ("shl", "shr"):
            b = (np.abs(b.astype(np.int64)) % (8 * np.dtype(dtype).itemsize)).astype(dtype)
        got, want, _ = run_binary(dev, op, dtype, la, a, lb, b, L.ROW_MAJOR)
        assert np.array_equal(got, want), (op, dtype, la, lb)


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.uint8, np.uint64, np.float32, np.float64, np.bool_])
@pytest.mark.parametrize("op", CMP_BIN)
def test_binary_comparisons_give_bool(dev, op, dtype):
    rng = np.random.default_rng(seed_of((op, np.dtype(dtype).name)))
    for _ in range(6):
        la, na, lb, nb = _bcast_pair(rng, L.ROW_MAJOR)
        a, b = rand_data(rng, na, dtype), rand_data(rng, nb, dtype)
        if np.dtype(dtype).kind == "f":
            a[::3] = np.round(a[::3])
            b[::3] = np.round(b[::3])
            a[::7] = np.nan
        got, want, _ = run_binary(dev, op, dtype, la, a, lb, b, L.ROW_MAJOR)
        assert got.dtype == np.bool_
        assert np.array_equal(got, want), (op, dtype)


@pytest.mark.parametrize("dtype", FLT_DTYPES)
@pytest.mark.parametrize("op", FLOAT_EXACT_BIN + FLOAT_TOL_BIN)
def test_binary_float_functions(dev, op, dtype):
    rng = np.random.default_rng(seed_of((op, np.dtype(dtype).name)))
    for _ in range(6):
        la, na, lb, nb = _bcast_pair(rng, L.ROW_MAJOR)
        a, b = rand_data(rng, na, dtype), rand_data(rng, nb, dtype)
        if op == "pow":
            a = np.abs(a) + dtype(0.1)
        got, want, _ = run_binary(dev, op, dtype, la, a, lb, b, L.ROW_MAJOR)
        if op in FLOAT_EXACT_BIN:
            assert np.array_equal(got, want, equal_nan=True)
        else:
            tol = 2e-6 if dtype == np.float32 else 4e-15
            assert np.allclose(got, want, rtol=tol, atol=tol, equal_nan=True), (op, dtype)


@pytest.mark.parametrize("order", [(rt.ROW_MAJOR, L.ROW_MAJOR), (rt.COL_MAJOR, L.COL_MAJOR)])
def test_add_output_layout_follows_the_reference(dev, dev_col, order):
    """get_layout_for_binary_op parity incl. the col-major device (op_binary_arithmetic.rs:1072-1143)."""
    d = dev if order[0] == rt.ROW_MAJOR else dev_col
    rng = np.random.default_rng(7 + order[0])
    for _ in range(25):
        la, na, lb, nb = _bcast_pair(rng, order[1])
        a, b = rand_data(rng, na, np.float64), rand_data(rng, nb, np.float64)
        tc = rt.Tensor(upload(d, a), P(la)) + rt.Tensor(upload(d, b), P(lb))
        c, lc = oracle.tensor_binary("add", a, la, b, lb, order[1])
        assert same(tc.layout, lc), (la, lb, tc.layout, lc)
        assert np.array_equal(tc.to_numpy(), view_np(c, lc))


def test_reference_add_kats_on_device(dev, dev_col):
    """op_binary_arithmetic.rs:992-1143 through the Tensor mirror."""
    lin = np.linspace
    a = rt.asarray(lin(1, 6, 6), dev).reshape([2, 3])
    b = rt.asarray(lin(2, 6, 3), dev)
    assert (a + b).to_numpy().reshape(-1).tolist() == [3., 6., 9., 6., 9., 12.]
    a = rt.asarray(lin(1, 6, 6), dev).reshape([1, 2, 3])
    b = rt.asarray(lin(1, 10, 10), dev).reshape([5, 1, 2, 1])
    assert (a + b).to_numpy().reshape(-1).tolist() == [2., 3., 4., 6., 7., 8., 4., 5., 6., 8., 9., 10., 6., 7., 8., 10.,
                                                      11., 12., 8., 9., 10., 12., 13., 14., 10., 11., 12., 14., 15., 16.]
    a = rt.asarray(lin(1, 9, 9), dev).reshape([3, 3])
    b = rt.asarray(lin(2, 18, 9), dev).reshape([3, 3]).reverse_axes()
    assert (a + b).to_numpy().reshape(-1).tolist() == [3., 10., 17., 8., 15., 22., 13., 20., 27.]
    a, b = rt.asarray(lin(1, 5, 5), dev), rt.asarray(lin(2, 10, 5), dev)
    assert (a.flip(0) + b).to_numpy().tolist() == [7., 8., 9., 10., 11.]
    assert (a + b.flip(0)).to_numpy().tolist() == [11., 10., 9., 8., 7.]
    assert (a - b).to_numpy().tolist() == [-1., -2., -3., -4., -5.]
    assert (a * b).to_numpy().tolist() == [2., 8., 18., 32., 50.]
    # col-major device: raw buffer order is compared (op_binary_arithmetic.rs:1086-1102)
    a = rt.asarray(lin(1, 6, 6), dev_col).reshape([3, 2])
    b = rt.asarray(lin(2, 6, 3), dev_col)
    c = a + b
    assert dev_col.to_cpu_vec(c.raw)[:6].tolist() == [3., 6., 9., 6., 9., 12.]
    a = rt.asarray(lin(1, 6, 6), dev_col).reshape([3, 2, 1])
    b = rt.asarray(lin(1, 10, 10), dev_col).reshape([1, 2, 1, 5])
    c = a + b
    assert dev_col.to_cpu_vec(c.raw)[:30].tolist() == [2., 3., 4., 6., 7., 8., 4., 5., 6., 8., 9., 10., 6., 7., 8., 10., 11.,
                                                       12., 8., 9., 10., 12., 13., 14., 10., 11., 12., 14., 15., 16.]


@pytest.mark.parametrize("dtype", [np.int32, np.uint64, np.float32, np.float64])
def test_scalar_operands_and_inplace(dev, dtype):
    rng = np.random.default_rng(11)
    for _ in range(10):
        la, na = random_view(rng)
        a = rand_data(rng, na, dtype)
        s = dtype(3)
        for op in ("add", "sub", "mul", "div"):
            ta = rt.Tensor(upload(dev, a), P(la))
            lk = L.layout_for_array_copy(la, "K")
            for reverse in (False, True):
                got_t = ta._binary(op, s, reverse=reverse)
                assert same(got_t.layout, lk)
                c = np.zeros(max(lk.size, 1), dtype=dtype)
                if reverse:
                    oracle.op_mutc_refa_refb(op, c, lk, s, None, a, la)
                else:
                    oracle.op_mutc_refa_refb(op, c, lk, a, la, s, None)
                assert np.array_equal(got_t.to_numpy(), view_np(c, lk), equal_nan=True), (op, reverse, la)
            # a op= s  and  a = s op a  in place (Op*AssignAPI / OpRConsume*API)
            for reverse in (False, True):
                raw = upload(dev, a)
                dev.op_muta_numb(op, raw, P(la), s, reverse=reverse)
                ref = a.copy()
                if reverse:
                    oracle.op_mutc_refa_refb(op, ref, la, s, None, a, la)
                else:
                    oracle.op_mutc_refa_refb(op, ref, la, a, la, s, None)
                assert np.array_equal(dev.to_cpu_vec(raw), ref, equal_nan=True), (op, reverse, la)
        # a op= b with b broadcast to a
        la2, na2, lb2, nb2 = _bcast_pair(rng, L.ROW_MAJOR)
        a2, b2 = rand_data(rng, na2, dtype), _nonzero(rand_data(rng, nb2, dtype))
        la_b, lb_b = L.broadcast_layout(la2, lb2, L.ROW_MAJOR)
        if la_b.shape == la2.shape:
            for op in ("add", "mul", "sub"):
                for reverse in (False, True):
                    raw = upload(dev, a2)
                    dev.op_muta_refb(op, raw, P(la_b), upload(dev, b2), P(lb_b), reverse=reverse)
                    ref = a2.copy()
                    if reverse:
                        oracle.op_mutc_refa_refb(op, ref, la_b, b2, lb_b, a2, la_b)
                    else:
                        oracle.op_mutc_refa_refb(op, ref, la_b, a2, la_b, b2, lb_b)
                    assert np.array_equal(dev.to_cpu_vec(raw), ref, equal_nan=True)


UNARY_EXACT = ["neg", "abs", "square", "sign", "sqrt", "floor", "ceil", "round", "trunc", "reciprocal", "conj", "real",
               "imag"]
UNARY_TOL = ["exp", "expm1", "log", "log2", "log10", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh",
             "asinh", "acosh", "atanh"]
UNARY_PRED = ["isnan", "isinf", "isfinite", "signbit"]


@pytest.mark.parametrize("dtype", FLT_DTYPES)
@pytest.mark.parametrize("op", UNARY_EXACT + UNARY_TOL + UNARY_PRED)
def test_unary_float(dev, op, dtype):
    rng = np.random.default_rng(seed_of((op, np.dtype(dtype).name)))
    for _ in range(5):
        la, na = random_view(rng)
        a = rand_data(rng, na, dtype)
        if op in ("log", "log2", "log10", "sqrt", "acosh"):
            a = np.abs(a) + dtype(1.0)
        if op in ("asin", "acos", "atanh"):
            a = np.tanh(a).astype(dtype) * dtype(0.99)
        if op in UNARY_PRED:
            a[::5] = np.nan
            a[1::7] = np.inf
            a[2::9] = -0.0
        t = rt.Tensor(upload(dev, a), P(la)).unary(op)
        lk = L.layout_for_array_copy(la, "K")
        assert same(t.layout, lk)
        out_dtype = np.uint8 if op in UNARY_PRED else dtype
        c = np.zeros(max(lk.size, 1), dtype=out_dtype)
        oracle.op_muta_refb_unary(op, c, lk, a, la)
        got, want = t.to_numpy(), view_np(c, lk)
        if op in UNARY_PRED:
            assert got.dtype == np.bool_ and np.array_equal(got.astype(np.uint8), want)
        elif op in UNARY_EXACT:
            assert np.array_equal(got, want, equal_nan=True), op
        else:
            tol = 2e-6 if dtype == np.float32 else 4e-15
            assert np.allclose(got, want, rtol=tol, atol=tol, equal_nan=True), op
        if op in ("neg", "abs", "square", "sqrt"):  # in-place form
            raw = upload(dev, a)
            dev.unary_muta(op, raw, P(la))
            ref = a.copy()
            oracle.op_muta_refb_unary(op, ref, la, a, la)
            assert np.array_equal(dev.to_cpu_vec(raw), ref, equal_nan=True)


@pytest.mark.parametrize("dtype", [np.int8, np.int32, np.int64, np.uint8, np.uint32, np.uint64])
@pytest.mark.parametrize("op", ["neg", "not_", "abs", "square", "sign"])
def test_unary_int(dev, op, dtype):
    if op == "neg" and np.dtype(dtype).kind == "u":
        pytest.skip("Neg is not implemented for unsigned integers in Rust")
    rng = np.random.default_rng(5)
    for _ in range(5):
        la, na = random_view(rng)
        a = rand_data(rng, na, dtype)
        t = rt.Tensor(upload(dev, a), P(la)).unary(op)
        lk = L.layout_for_array_copy(la, "K")
        c = np.zeros(max(lk.size, 1), dtype=dtype)
        oracle.op_muta_refb_unary(op, c, lk, a, la)
        assert np.array_equal(t.to_numpy(), view_np(c, lk)), (op, dtype)


def test_not_on_bool(dev):
    a = np.array([True, False, True, True])
    t = ~rt.asarray(a, dev)
    assert t.to_numpy().tolist() == [False, True, False, False]


# ---- shapes chosen to hit every kernel variant (flat ND=1 / ND=0, rows, tile, scalar, >8 dims, 2^31 split) ----
KERNEL_SHAPES = [
    ("flat 1-D vector", (1 << 16,), None, None),
    ("flat 1-D odd length (scalar tail)", ((1 << 16) + 3,), None, None),
    ("rows kernel + row broadcast", (300, 4096), None, (0, 1)),
    ("rows kernel, no broadcast (pitched)", (64, 2048), "pitched", None),
    ("column broadcast", (128, 2048), None, (1, 0)),
    ("3-D generic vector", (5, 7, 64), None, (64, 0, 1)),
    ("tile kernel a + b^T", (96, 160), None, "T"),
    ("tile kernel batched", (3, 70, 90), None, "T3"),
    ("short rows", (4096, 6), None, (0, 1)),
    ("rect tile a + b^T, narrow X=17", (1000, 17), None, "T"),
    ("rect tile a + b^T, narrow X=40", (700, 40), None, "T"),
    ("rect tile a + b^T, narrow Y=24", (24, 900), None, "T"),
    ("rect tile batched, narrow X", (3, 310, 33), None, "T3"),
    ("rect tile batched, narrow Y", (3, 18, 333), None, "T3"),
    ("rect tile a + b^T, narrow X=5", (1000, 5), None, "T"),
    ("rect tile a + b^T, narrow Y=31", (31, 257), None, "T"),
]


@pytest.mark.parametrize("name,shape,amode,bmode", KERNEL_SHAPES, ids=[k[0] for k in KERNEL_SHAPES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.uint8])
def test_kernel_variants(dev, name, shape, amode, bmode, dtype):
    rng = np.random.default_rng(seed_of(name))
    n = int(np.prod(shape))
    la = L.c_contig_layout(shape)
    if amode == "pitched":
        la = L.Layout(shape, (shape[1] + 8, 1), 0)
        a = rand_data(rng, shape[0] * (shape[1] + 8), dtype)
    else:
        a = rand_data(rng, n, dtype)
    if bmode is None:
        lb, b = L.c_contig_layout(shape), rand_data(rng, n, dtype)
    elif bmode == "T":
        lb, b = L.c_contig_layout(shape[::-1]).reverse_axes(), rand_data(rng, n, dtype)
    elif bmode == "T3":
        lb = L.c_contig_layout((shape[0], shape[2], shape[1])).swapaxes(1, 2)
        b = rand_data(rng, n, dtype)
    else:
        lb = L.Layout(shape, bmode, 0)
        b = rand_data(rng, L.bounds_index(lb)[1], dtype)
    for op in ("add", "mul"):
        tc = rt.Tensor(upload(dev, a), P(la)).binary(op, rt.Tensor(upload(dev, b), P(lb)))
        c, lc = oracle.tensor_binary(op, a, la, b, lb)
        assert same(tc.layout, lc)
        assert np.array_equal(tc.to_numpy(), view_np(c, lc), equal_nan=True), (name, op)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32])
@pytest.mark.parametrize("shape", [(37, 4096), (300, 2048), (5, 1024 + 16)])
def test_outer_sum_of_two_broadcast_operands(dev, dtype, shape):
    """(n,1) + (1,m): one operand is a splat along the output's contiguous axis that changes from row to row, the other
    is broadcast over the rows (orbital-energy denominators e_i + e_a are built this way); flat vector kernel."""
    rng = np.random.default_rng(seed_of("outer", shape, np.dtype(dtype).name))
    n, m = shape
    col, row = rand_data(rng, n, dtype), rand_data(rng, m, dtype)
    lcol, lrow = L.Layout(shape, (1, 0), 0), L.Layout(shape, (0, 1), 0)
    for op in ("add", "sub", "mul"):
        for (x, lx, y, ly) in ((col, lcol, row, lrow), (row, lrow, col, lcol)):
            tc = rt.Tensor(upload(dev, x), P(lx)).binary(op, rt.Tensor(upload(dev, y), P(ly)))
            c, lc = oracle.tensor_binary(op, x, lx, y, ly)
            assert same(tc.layout, lc)
            assert np.array_equal(tc.to_numpy(), view_np(c, lc)), (op, shape)


def test_misaligned_views_fall_back_to_scalar_path(dev):
    rng = np.random.default_rng(3)
    a = rand_data(rng, 5000, np.float64)
    b = rand_data(rng, 5000, np.float64)
    la = L.Layout((4097,), (1,), 3)  # odd offset: not 16-byte aligned
    lb = L.Layout((4097,), (1,), 1)
    tc = rt.Tensor(upload(dev, a), P(la)) + rt.Tensor(upload(dev, b), P(lb))
    assert np.array_equal(tc.to_numpy(), a[3:4100] + b[1:4098])


def test_many_dims_are_split_on_the_host(dev):
    rng = np.random.default_rng(4)
    shape = (2,) * 10  # 10 non-mergeable dims after striding every axis by 2
    big = rand_data(rng, 4 ** 10, np.float64)
    la = L.Layout(shape, tuple(2 * 4 ** (9 - i) for i in range(10)), 0)
    b = rand_data(rng, 2 ** 10, np.float64)
    lb = L.c_contig_layout(shape)
    tc = rt.Tensor(upload(dev, big), P(la)) - rt.Tensor(upload(dev, b), P(lb))
    c, lc = oracle.tensor_binary("sub", big, la, b, lb)
    assert np.array_equal(tc.to_numpy(), view_np(c, lc))


def test_zero_size_and_scalar_shapes(dev):
    a = rt.asarray(np.zeros(0), dev).reshape([0, 3])
    b = rt.asarray(np.ones(3), dev)
    c = a + b
    assert c.shape == (0, 3) and c.to_numpy().shape == (0, 3)
    s = rt.Tensor(upload(dev, np.array([2.5])), rt.Layout((), (), 0))
    t = s + rt.asarray(np.array([1.0, 2.0]), dev)
    assert t.to_numpy().tolist() == [3.5, 4.5]


def test_errors_mirror_the_reference(dev):
    a = rt.asarray(np.arange(6.0), dev).reshape([2, 3])
    b = rt.asarray(np.arange(4.0), dev)
    with pytest.raises(rt.RstsrCudaError) as e:
        a + b
    assert e.value.kind == "InvalidLayout" and "Broadcasting failed" in str(e.value)
    # shape mismatch at the device level: "All shape of layout in this function must be the same."
    with pytest.raises(rt.RstsrCudaError) as e:
        dev.op_mutc_refa_refb("add", a.raw, a.layout, a.raw, a.layout, b.raw, b.layout)
    assert e.value.kind == "InvalidLayout"
    # unsupported dtype/op combination is UnImplemented, never a silent fallback
    i = rt.asarray(np.arange(4, dtype=np.int32), dev)
    with pytest.raises(rt.RstsrCudaError) as e:
        i.unary("sin")
    assert e.value.kind == "UnImplemented"


def test_reference_consume_kats_on_device(dev):
    """test_add_consume_row_major / test_sub_consume (op_binary_arithmetic.rs:1166-1317): an owned operand's
    buffer is reused when the other operand broadcasts to its layout, otherwise a new one is allocated."""
    lin = np.linspace
    a, b = rt.asarray(lin(1, 5, 5), dev), rt.asarray(lin(2, 10, 5), dev)
    c = a.consume("add", b)                                   # a + &b, same shape
    assert c.raw.ptr == a.raw.ptr and c.to_numpy().tolist() == [3., 6., 9., 12., 15.]
    a = rt.asarray(lin(1, 10, 10), dev).reshape([2, 5])
    a.owned = True
    c = a.consume("add", b)                                   # a + &b, b broadcasts to a
    assert c.raw.ptr == a.raw.ptr
    assert c.to_numpy().reshape(-1).tolist() == [3., 6., 9., 12., 15., 8., 11., 14., 17., 20.]
    a = rt.asarray(lin(2, 10, 5), dev)
    b2 = rt.asarray(lin(1, 10, 10), dev).reshape([2, 5])
    c = a.consume("add", b2)                                  # a + &b, a would have to grow: new buffer
    assert c.raw.ptr != a.raw.ptr
    assert c.to_numpy().reshape(-1).tolist() == [3., 6., 9., 12., 15., 8., 11., 14., 17., 20.]
    a, b = rt.asarray(lin(1, 5, 5), dev), rt.asarray(lin(2, 10, 5), dev)
    c = b.consume("sub", a, reverse=True)                     # &a - b reuses b (OpRConsumeSubAPI: b = a - b)
    assert c.raw.ptr == b.raw.ptr and c.to_numpy().tolist() == [-1., -2., -3., -4., -5.]
    a, b = rt.asarray(lin(1, 5, 5), dev), rt.asarray(lin(2, 10, 5), dev)
    c = a.consume("sub", b.view())                            # a - &b reuses a
    assert c.raw.ptr == a.raw.ptr and c.to_numpy().tolist() == [-1., -2., -3., -4., -5.]
