"""GPU parity of the boundary completions: mixed operand types (rc_op_mutc_*_ex) against the oracle's
promote -> into_float -> f restatement, elementwise isclose, operands broadcast along a common axis, explicit
pairing order of assign_arbitary, and the dtype guards of the host mirror.  Bit-exact except libm functions."""
import numpy as np
import pytest

import oracle
import rstsr_b200 as rt
from oracle import layout as L
from oracle import promotion as PR

from helpers import O, P, rand_data, seed_of, upload, view_np

pytestmark = pytest.mark.gpu

PAIRS = [(np.int32, np.float64), (np.float32, np.float64), (np.int64, np.int64), (np.int8, np.uint8),
         (np.uint64, np.int64), (np.int32, np.float32), (np.uint16, np.float32), (np.bool_, np.int32),
         (np.float64, np.uint8), (np.int16, np.int64)]
EXACT_OPS = ["add", "sub", "mul", "maximum", "minimum", "eq", "ne", "lt", "le", "gt", "ge", "copysign", "nextafter"]
TOL_OPS = ["atan2", "hypot", "logaddexp"]


def _oracle_mixed(op, a, la, b, lb, lc):
    """promote_pair + into_float + f per element: cast both operands to the compute type, then the same-type oracle."""
    k, out = PR.op_types(op, PR.name_of(a.dtype), PR.name_of(b.dtype))
    ak = PR.cast(a, k) if isinstance(a, np.ndarray) else a
    bk = PR.cast(b, k) if isinstance(b, np.ndarray) else b
    c = np.zeros(max(L.bounds_index(lc)[1], 1), dtype=PR.NP[out])
    cc = c.view(np.uint8) if c.dtype == np.bool_ else c
    ak = ak.view(np.uint8) if ak.dtype == np.bool_ else ak
    bk = bk.view(np.uint8) if bk.dtype == np.bool_ else bk
    oracle.op_mutc_refa_refb(op, cc, lc, ak, la, bk, lb)
    return c


@pytest.mark.parametrize("pair", PAIRS, ids=lambda p: f"{np.dtype(p[0]).name}-{np.dtype(p[1]).name}")
@pytest.mark.parametrize("op", EXACT_OPS + TOL_OPS)
def test_mixed_binary(dev, op, pair):
    ta, tb = pair
    na, nb = PR.name_of(ta), PR.name_of(tb)
    try:
        k, out = PR.op_types(op, na, nb)
    except TypeError:
        pytest.skip("the reference has no impl for this pair")
    if k == "bool" and op not in ("eq", "ne", "lt", "le", "gt", "ge"):
        pytest.skip("bool arithmetic")
    rng = np.random.default_rng(seed_of((op, na, nb)))
    for shape_a, shape_b in (((37, 50), (37, 50)), ((6, 1, 40), (5, 40)), ((33,), (4, 33)), ((8, 9), (1,))):
        a, b = rand_data(rng, int(np.prod(shape_a)), ta), rand_data(rng, int(np.prod(shape_b)), tb)
        la, lb = L.c_contig_layout(list(shape_a)), L.c_contig_layout(list(shape_b))
        if len(shape_b) == 2 and shape_b[0] != 1:  # a transposed operand too
            lb = L.f_contig_layout(list(shape_b))
        x = rt.Tensor(upload(dev, a), P(la))
        y = rt.Tensor(upload(dev, b), P(lb))
        z = x.binary(op, y)
        assert z.dtype == np.dtype(PR.NP[out]), (op, na, nb, z.dtype)
        la_b, lb_b = L.broadcast_layout(la, lb, L.ROW_MAJOR)
        want = view_np(_oracle_mixed(op, a, la_b, b, lb_b, O(z.layout)), O(z.layout))
        got = z.to_numpy()
        if op in TOL_OPS:
            tol = 2e-6 if out == "f32" else 4e-15  # CUDA libm vs glibc, as in test_gpu_elementwise.py
            assert np.allclose(got, want, rtol=tol, atol=tol, equal_nan=True), (op, na, nb)
        else:
            assert np.array_equal(got, want, equal_nan=(got.dtype.kind == "f")), (op, na, nb, shape_a, shape_b)


@pytest.mark.parametrize("pair", PAIRS[:5], ids=lambda p: f"{np.dtype(p[0]).name}-{np.dtype(p[1]).name}")
def test_mixed_scalar_operands(dev, pair):
    """op_mutc_refa_numb / op_mutc_numa_refb with a scalar of another type (promoted pair)."""
    ta, tb = pair
    rng = np.random.default_rng(seed_of(("numb", np.dtype(ta).name, np.dtype(tb).name)))
    a = rand_data(rng, 300, ta)
    la = L.c_contig_layout([12, 25])
    sb = rand_data(rng, 1, tb)[0]
    for op in ("maximum", "lt", "hypot"):
        k, out = PR.op_types(op, PR.name_of(ta), PR.name_of(tb))
        raw_a = upload(dev, a)
        lc = rt.layout_for_array_copy(P(la), 3, rt.ROW_MAJOR)
        for reverse in (False, True):
            k2, out2 = PR.op_types(op, *((PR.name_of(tb), PR.name_of(ta)) if reverse else (PR.name_of(ta), PR.name_of(tb))))
            c = dev.uninit_impl(PR.NP[out2], 300)
            if reverse:
                dev.op_mutc_numa_refb(op, c, lc, sb, raw_a, P(la), a_dtype=tb)
            else:
                dev.op_mutc_refa_numb(op, c, lc, raw_a, P(la), sb, b_dtype=tb)
            got = dev.to_cpu_vec(c)
            ak = PR.cast(a, k2)
            bk = PR.cast(np.array([sb]), k2)[0]
            cw = np.zeros(300, dtype=PR.NP[out2])
            cc = cw.view(np.uint8) if cw.dtype == np.bool_ else cw
            if reverse:
                oracle.op_mutc_refa_refb(op, cc, O(lc), bk, None, ak, la)
            else:
                oracle.op_mutc_refa_refb(op, cc, O(lc), ak, la, bk, None)
            if op == "hypot":
                tol = 4e-15 if out2 == "f64" else 2e-6
                assert np.allclose(got, cw, rtol=tol, atol=tol)
            else:
                assert np.array_equal(got, cw), (op, reverse)


def test_pow_mixed_exponent(dev):
    rng = np.random.default_rng(7)
    # float ^ small signed / unsigned integers: powi, bit-identical to __powi?f2
    for tf in (np.float32, np.float64):
        for te in (np.int8, np.uint8, np.int16, np.uint16, np.int32):
            a = (rng.random(500) * 3.0 - 1.5).astype(tf)
            e = rng.integers(-6 if np.dtype(te).kind == "i" else 0, 7, 500).astype(te)
            x = rt.Tensor(upload(dev, a), rt.Layout((20, 25), (25, 1)))
            y = rt.Tensor(upload(dev, e), rt.Layout((20, 25), (1, 20)))  # transposed exponent
            z = x.binary("pow", y)
            assert z.dtype == np.dtype(tf)
            want = PR.powi(a.reshape(20, 25), e.reshape(25, 20).T.astype(np.int32))
            assert np.array_equal(z.to_numpy(), want, equal_nan=True), (tf, te)
    # integer ^ unsigned: wrapping power
    for ti in (np.int8, np.int32, np.int64, np.uint16, np.uint64):
        for te in (np.uint8, np.uint32, np.uint64):
            a = rng.integers(-5 if np.dtype(ti).kind == "i" else 0, 6, 400).astype(ti)
            e = rng.integers(0, 40, 400).astype(te)
            z = rt.Tensor(upload(dev, a), rt.Layout((400,), (1,))).binary("pow", rt.Tensor(upload(dev, e), rt.Layout((400,), (1,))))
            assert z.dtype == np.dtype(ti)
            assert np.array_equal(z.to_numpy(), PR.ipow(a, e)), (ti, te)
    # scalar exponent
    a = (rng.random(64) + 0.5)
    raw = upload(dev, a)
    c = dev.uninit_impl(np.float64, 64)
    l = rt.Layout((64,), (1,))
    dev.op_mutc_refa_numb("pow", c, l, raw, l, -3, b_dtype=np.int32)
    assert np.array_equal(dev.to_cpu_vec(c), PR.powi(a, np.full(64, -3)))
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.Tensor(upload(dev, a), l).binary("pow", rt.Tensor(upload(dev, a.astype(np.float32)), l))
    assert e.value.kind == "UnImplemented"


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.uint64])
def test_isclose_elementwise(dev, dtype):
    rng = np.random.default_rng(seed_of(("isclose", np.dtype(dtype).name)))
    a = rand_data(rng, 4000, dtype)
    b = a.copy()
    idx = rng.integers(0, 4000, 600)
    if np.dtype(dtype).kind == "f":
        b[idx] = b[idx] * (1 + 3e-5 * rng.standard_normal(600).astype(dtype))
        a[5], b[5] = np.nan, np.nan
        a[6], b[6] = np.inf, np.inf   # the reference: |inf - inf| is NaN -> NOT close
        a[7] = np.nan
    else:
        b[idx] += rng.integers(-1, 2, 600).astype(dtype)
    la, lb = L.c_contig_layout([40, 100]), L.c_contig_layout([100, 40]).transpose([1, 0])
    for rtol, atol, eq_nan in ((1e-5, 1e-8, False), (1e-5, 1e-8, True), (0.0, 1.0, False)):
        z = rt.isclose(rt.Tensor(upload(dev, a), P(la)), rt.Tensor(upload(dev, b), P(lb)), rtol, atol, eq_nan)
        assert z.dtype == np.bool_
        av, bv = view_np(a, la), view_np(b, lb)
        want = np.array([[oracle.isclose_scalar(av[i, j], bv[i, j], rtol, atol, eq_nan) for j in range(100)] for i in range(40)])
        assert np.array_equal(z.to_numpy(), want), (dtype, rtol, atol, eq_nan)


def test_common_broadcast_axes(dev):
    """x.broadcast_to(s) + y.broadcast_to(s): get_layout_for_binary_op keeps stride 0 on the axis both operands
    broadcast, and the op must run (the reference iterates the stride-0 output)."""
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(3), rng.standard_normal(3)
    la = L.Layout((4, 3), (0, 1), 0)
    x = rt.Tensor(upload(dev, a), P(la))
    y = rt.Tensor(upload(dev, b), P(la))
    for op in ("add", "maximum", "lt"):
        z = x.binary(op, y)
        lc_ref = L.get_layout_for_binary_op(la, la, L.ROW_MAJOR) if op == "add" else None
        if lc_ref is not None:
            assert (z.layout.shape, z.layout.stride) == (lc_ref.shape, lc_ref.stride)
        want = {"add": a + b, "maximum": np.maximum(a, b), "lt": a < b}[op]
        assert np.array_equal(z.to_numpy(), np.broadcast_to(want, (4, 3)))
    # vecdot with a kept axis broadcast in both operands
    m = rng.standard_normal(5)
    lm = L.Layout((6, 5), (0, 1), 0)
    v = rt.vecdot(rt.Tensor(upload(dev, m), P(lm)), rt.Tensor(upload(dev, m), P(lm)), -1)
    assert np.allclose(v.to_numpy(), np.full(6, (m * m).sum()), rtol=1e-13)
    # an input that varies along an axis where the OUTPUT is broadcast stays an error
    c = dev.uninit_impl(np.float64, 3)
    with pytest.raises(rt.RstsrCudaError) as e:
        dev.op_mutc_refa_refb("add", c, P(la), upload(dev, rng.standard_normal(12)), rt.Layout((4, 3), (3, 1)), y.raw, P(la))
    assert e.value.kind == "InvalidLayout"


def test_reshape_order_does_not_touch_the_handle(dev):
    a = np.arange(24, dtype=np.float64)
    t = rt.Tensor(upload(dev, a), rt.Layout((2, 3, 4), (12, 4, 1))).transpose([2, 0, 1])
    before = dev.default_order()
    r = t.reshape([6, 4], order=rt.COL_MAJOR)
    assert dev.default_order() == before
    want = np.reshape(a.reshape(2, 3, 4).transpose(2, 0, 1), (6, 4), order="F")
    assert np.array_equal(r.to_numpy(), want)


def test_dtype_guards(dev):
    a = upload(dev, np.zeros(8, dtype=np.float32))
    b = upload(dev, np.zeros(8, dtype=np.float64))
    l = rt.Layout((8,), (1,))
    with pytest.raises(rt.RstsrCudaError) as e:
        dev.op_muta_refb("add", a, l, b, l)
    assert "DTypeMismatch" in str(e.value)
    with pytest.raises(rt.RstsrCudaError):
        dev.op_mutc_refa_refb("add", a, l, a, l, b, l)        # f32 + f64 writes f64, not f32
    out = dev.uninit_impl(np.float32, 1)
    with pytest.raises(rt.RstsrCudaError):
        dev.reduce_axes_into("sum", b, l, [0], out, rt.Layout((), ()))
    # promoted add through the tensor layer: f32 + f64 -> f64, values exact
    x = np.arange(8, dtype=np.float32) / 3
    y = np.arange(8, dtype=np.float64) / 7
    z = rt.Tensor(upload(dev, x), l) + rt.Tensor(upload(dev, y), l)
    assert z.dtype == np.float64 and np.array_equal(z.to_numpy(), x.astype(np.float64) + y)


def test_same_device_needs_same_stream(dev):
    other = rt.DeviceCuda(0, rt.ROW_MAJOR)
    try:
        assert dev.same_device(dev) and not dev.same_device(other)
        l = rt.Layout((4,), (1,))
        x = rt.Tensor(upload(dev, np.ones(4)), l)
        y = rt.Tensor(upload(other, np.ones(4)), l)
        with pytest.raises(rt.RstsrCudaError) as e:
            _ = x + y
        assert e.value.kind == "DeviceMismatch"
        moved = rt.Tensor(other.change_device(y.raw, dev), l)
        assert np.array_equal((x + moved).to_numpy(), np.full(4, 2.0))
    finally:
        other.close()


@pytest.mark.parametrize("op", ["argmin", "argmax"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32])
def test_unraveled_arg(dev, op, dtype):
    """OpUnraveledArgMin/MaxAPI (operators/reduction.rs:35-55): the index TUPLE of the first extreme element -- within
    the tensor for `_all`, within the reduced axes (in the order given) for `_axes`
    (reduce_axes_unraveled_arg_cpu_serial, cpu_serial/reduction.rs:479-529)."""
    rng = np.random.default_rng(seed_of(("unravel", op, np.dtype(dtype).name)))
    shape = (7, 9, 11)
    a = rand_data(rng, int(np.prod(shape)), dtype)
    a[rng.integers(0, a.size, 40)] = a.max() if op == "argmax" else a.min()   # ties: the first occurrence wins
    for la in (L.c_contig_layout(list(shape)), L.c_contig_layout(list(shape)).transpose([2, 0, 1]),
               L.f_contig_layout(list(shape)).narrow(1, slice(None, None, -1))):
        raw = upload(dev, a)
        flat = int(oracle.reduce_ext(op, a, la))
        assert dev.unraveled_arg_all(op, raw, P(la)) == tuple(int(i) for i in np.unravel_index(flat, la.shape))
        for axes in ([0], [2, 0], [1, 2], [-1], [0, 1, 2]):
            ref, lref = oracle.reduce_ext(op, a, la, axes)
            tuples, lo = dev.unraveled_arg_axes(op, raw, P(la), axes)
            assert (lo.shape, lo.stride, lo.offset) == (tuple(lref.shape), tuple(lref.stride), lref.offset)
            got = dev.to_cpu_vec(tuples).reshape(-1, len(axes))
            red_shape = tuple(la.shape[ax % 3] for ax in axes)
            want_flat = np.asarray(ref).reshape(-1)          # u64 positions, memory order of lref
            want = np.stack(np.unravel_index(want_flat.astype(np.int64), red_shape), axis=-1)
            assert np.array_equal(got[:want.shape[0]], want), (op, dtype, axes)
    with pytest.raises(rt.RstsrCudaError) as e:
        dev.unraveled_arg_all(op, upload(dev, np.zeros(1, dtype=dtype)), rt.Layout((0, 3), (3, 1)))
    assert e.value.kind == "InvalidLayout"


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int16])
def test_outer_ops_take_both_operand_orders(dev, dtype):
    """c[i, j] = u[i] op v[j] and v[j] op u[i] (ew_outer_kernel, SPL = 1 / 2): non-commutative ops pin the operand order;
    odd row counts exercise the last partial CTA, row strides != 1 the scalar walk."""
    rng = np.random.default_rng(seed_of(("outer", np.dtype(dtype).name)))
    for n_rows, n_cols, ustep in ((37, 2048, 1), (8, 4096, 3), (1001, 1024, 1), (5, 64, 2)):
        u = rand_data(rng, n_rows * ustep, dtype)
        v = rand_data(rng, n_cols, dtype)
        if np.dtype(dtype).kind != "f":
            v[v == 0] = 1
            u[u == 0] = 1
        lu = L.Layout((n_rows, 1), (ustep, 1), 0)   # column vector, every ustep-th element
        lv = L.c_contig_layout([n_cols])
        tu, tv = rt.Tensor(upload(dev, u), P(lu)), rt.Tensor(upload(dev, v), P(lv))
        uu = u[::ustep][:n_rows].reshape(n_rows, 1)
        for op, fn in (("sub", np.subtract), ("add", np.add), ("mul", np.multiply)):
            got = tu.binary(op, tv).to_numpy()
            assert np.array_equal(got, fn(uu, v[None, :]).astype(dtype)), (dtype, op, "u op v")
            got = tv.binary(op, tu).to_numpy()
            assert np.array_equal(got, fn(v[None, :], uu).astype(dtype)), (dtype, op, "v op u")
        if np.dtype(dtype).kind == "f":
            assert np.array_equal(tu.binary("div", tv).to_numpy(), uu / v[None, :])
            assert np.array_equal(tv.binary("div", tu).to_numpy(), v[None, :] / uu)
            assert np.array_equal(tu.binary("lt", tv).to_numpy(), uu < v[None, :])


FUSED_PAIRS = [(np.float32, np.float64), (np.int32, np.float64), (np.int64, np.float64), (np.int32, np.int64)]


@pytest.mark.parametrize("pair", FUSED_PAIRS, ids=lambda p: f"{np.dtype(p[0]).name}-{np.dtype(p[1]).name}")
def test_fused_promotion_pairs(dev, pair):
    """+ - * / on the pairs whose widening cast is fused into the kernel (rc_ew_mixed.cu), in both operand orders and
    through every kernel family -- 16/32-byte packs, rows with a broadcast operand, transposed operand (tile), short
    transposed axis (rect tile), outer product, scalar operand on either side, odd sizes on the scalar path: bit-exact
    against `a as K op b as K` (Rust's wrapping integer arithmetic for K = i64; / on integers only where b != 0)."""
    tn, tw = pair
    K = np.promote_types(tn, tw)
    assert rt.DeviceCuda.promote_types(tn, tw) == K
    rng = np.random.default_rng(seed_of(("fused", np.dtype(tn).name, np.dtype(tw).name)))

    def data(n, dt):
        if np.dtype(dt).kind == "f":
            return (rng.standard_normal(n) * 100).astype(dt)
        x = rng.integers(-2**20, 2**20, n).astype(dt)
        x[x == 0] = 7
        return x

    ops = {"add": np.add, "sub": np.subtract, "mul": np.multiply}
    if K.kind == "f":
        ops["div"] = np.divide

    def check(x, y, xa, ya, what):
        for op, fn in ops.items():
            z = x.binary(op, y)
            assert z.dtype == K, (op, what)
            want = fn(np.asarray(xa).astype(K), np.asarray(ya).astype(K))
            assert np.array_equal(z.to_numpy(), want), (op, what, pair)

    for first_narrow in (True, False):
        ta, tb = (tn, tw) if first_narrow else (tw, tn)
        # (1) flat packs + odd tail; (2) rows with a broadcast row; (3) transposed operand; (4) short transposed axis; (5) outer
        a, b = data(256 * 1024 + 3, ta), data(256 * 1024 + 3, tb)
        check(rt.asarray(a, dev), rt.asarray(b, dev), a, b, "flat")
        a, b = data(384 * 1024, ta).reshape(384, 1024), data(1024, tb)
        check(rt.asarray(a.reshape(-1), dev).reshape([384, 1024]), rt.asarray(b, dev), a, b, "rows")
        a, b = data(512 * 640, ta).reshape(512, 640), data(640 * 512, tb).reshape(640, 512)
        check(rt.asarray(a.reshape(-1), dev).reshape([512, 640]), rt.asarray(b.reshape(-1), dev).reshape([640, 512]).transpose([1, 0]),
              a, b.T, "tile")
        a, b = data(4096 * 6, ta).reshape(4096, 6), data(6 * 4096, tb).reshape(6, 4096)
        check(rt.asarray(a.reshape(-1), dev).reshape([4096, 6]), rt.asarray(b.reshape(-1), dev).reshape([6, 4096]).transpose([1, 0]),
              a, b.T, "rect")
        a, b = data(700, ta), data(900, tb)
        check(rt.asarray(a, dev).reshape([700, 1]), rt.asarray(b, dev).reshape([1, 900]), a.reshape(700, 1), b.reshape(1, 900), "outer")
    # scalar operands through the C ABI entry points (the tensor-level mirror only mixes float scalars with int tensors)
    a = data(5000, tn)
    s = np.asarray(data(1, tw))[0]
    la = rt.Layout((5000,), (1,))
    for op, fn in ops.items():
        out = dev.uninit_impl(K, 5000)
        dev.op_mutc_refa_numb(op, out, la, upload(dev, a), la, s, b_dtype=np.dtype(tw))
        assert np.array_equal(dev.to_cpu_vec(out), fn(a.astype(K), K.type(s))), (op, "numb")
        dev.op_mutc_numa_refb(op, out, la, s, upload(dev, a), la, a_dtype=np.dtype(tw))
        assert np.array_equal(dev.to_cpu_vec(out), fn(K.type(s), a.astype(K))), (op, "numa")
