"""GPU parity for the binary reductions: vecdot and allclose through the C ABI vs the oracle.
KATs: rstsr-core/src/tensor/linalg/vecdot.rs:41-93 (doc tests, also rstsr-core/tests/doc_draft/linalg/test_vecdot.rs)
and rstsr-dtype-traits/src/isclose.rs:152-181."""
import numpy as np
import pytest

import oracle
import rstsr_b200 as rt
from oracle import layout as L

from helpers import P, rand_data, random_view, same, seed_of, upload, view_np

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def check_vecdot(dev, order, a, la, b, lb, axes_a, axes_b):
    want, lc = oracle.tensor_vecdot(a, la, b, lb, axes_a, axes_b, order)
    ta, tb = rt.Tensor(upload(dev, a), P(la)), rt.Tensor(upload(dev, b), P(lb))
    got = rt.vecdot(ta, tb, (list(axes_a), list(axes_b)))
    assert same(got.layout, lc), (got.layout, lc)
    g, w = got.to_numpy(), view_np(want, lc)
    if a.dtype.kind == "f":
        # |delta| <= tol * sum_i |a_i b_i| per output (the reference's summation order is regime dependent)
        mag_raw, lm = oracle.tensor_vecdot(np.abs(a).astype(np.float64), la, np.abs(b).astype(np.float64), lb, axes_a, axes_b, order)
        mag = view_np(mag_raw, lm)
        assert np.all(np.abs(g.astype(np.float64) - w.astype(np.float64)) <= TOL[a.dtype] * mag + 1e-300), (g, w)
    else:
        assert np.array_equal(g, w)
    return got


def test_reference_kats(dev):
    a = rt.asarray(np.array([1, 2, 3]), dev)
    b = rt.asarray(np.array([4, 5, 6]), dev)
    r = rt.vecdot(a, b)
    assert r.ndim == 0 and int(r.to_numpy()) == 32
    a = rt.asarray(np.array([1, 2, 3, 4]), dev).reshape([2, 2])
    b = rt.asarray(np.array([5, 6, 7, 8]), dev).reshape([2, 2])
    assert rt.vecdot(a, b).to_numpy().tolist() == [17, 53]
    a = rt.asarray(np.array([0., 5., 0., 0., 0., 10., 0., 6., 8.]), dev).reshape([3, 3])
    b = rt.asarray(np.array([0., 0.6, 0.8]), dev)
    assert np.allclose(rt.vecdot(a, b).to_numpy(), [3., 8., 10.], rtol=1e-15)
    # isclose.rs tests through allclose on 1-element tensors
    x, y = rt.asarray(np.array([1.00001]), dev), rt.asarray(np.array([1.00002]), dev)
    assert rt.allclose(x, y) is True
    assert rt.allclose(x, y, rtol=1e-6, atol=1e-9) is False
    assert rt.allclose(rt.asarray(np.array([100], dtype=np.uint64), dev), rt.asarray(np.array([102], dtype=np.uint64), dev)) is False


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32, np.uint64, np.uint32])
@pytest.mark.parametrize("n", [1, 7, 64, 1000, 100003, 1 << 20])
def test_vecdot_1d(dev, dtype, n):
    rng = np.random.default_rng(seed_of("vd1", n, np.dtype(dtype).name))
    a, b = rand_data(rng, n, dtype), rand_data(rng, n, dtype)
    l = L.c_contig_layout([n])
    check_vecdot(dev, "row", a, l, b, l, [0], [0])
    if n > 8:  # misaligned, strided and reversed operands
        la = l.narrow(0, slice(1, None))
        lb = l.narrow(0, slice(None, -1))
        check_vecdot(dev, "row", a, la, b, lb, [0], [0])
        check_vecdot(dev, "row", a, l.narrow(0, slice(None, None, -1)), b, l, [0], [0])
        m = n // 2
        check_vecdot(dev, "row", a, l.narrow(0, slice(0, 2 * m, 2)), b, l.narrow(0, slice(0, m)), [0], [0])


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32])
@pytest.mark.parametrize("shape", [(5, 8), (64, 96), (257, 1024), (2048, 72), (3, 4, 5), (16, 33, 40), (8, 1, 64)])
def test_vecdot_each_axis(dev, dev_col, dtype, shape):
    rng = np.random.default_rng(seed_of("vdax", shape, np.dtype(dtype).name))
    n = int(np.prod(shape))
    a, b = rand_data(rng, n, dtype), rand_data(rng, n, dtype)
    for d, order, mk in ((dev, "row", L.c_contig_layout), (dev_col, "col", L.f_contig_layout)):
        l = mk(list(shape))
        for ax in range(len(shape)):
            check_vecdot(d, order, a, l, b, l, [ax], [ax])
        # transposed second operand pairs different axes
        if len(shape) == 2 and shape[0] != shape[1]:
            lt = mk([shape[1], shape[0]])
            bt = rand_data(rng, n, dtype)
            with pytest.raises(rt.RstsrCudaError):
                rt.vecdot(rt.Tensor(upload(d, a), P(l)), rt.Tensor(upload(d, bt), P(lt)), ([1], [1]))
        if len(shape) == 3:
            check_vecdot(d, order, a, l, b, l, [0, 2], [0, 2])
            check_vecdot(d, order, a, l, b, l, [2, 0], [2, 0])
            check_vecdot(d, order, a, l, b, l, [0, 1, 2], [0, 1, 2])


def test_vecdot_broadcast_and_axes_pairs(dev, dev_col):
    rng = np.random.default_rng(seed_of("vdbc"))
    for dtype in (np.float64, np.int64, np.float32):
        a = rand_data(rng, 6 * 40 * 5, dtype)
        b = rand_data(rng, 40 * 7, dtype)
        # a: (6, 40, 5) contract axis 1; b: (7, 40) contract axis 1 -> kept (6, 5) vs (7,): row-major broadcast fails
        with pytest.raises(oracle.LayoutError):
            oracle.tensor_vecdot(a, L.c_contig_layout([6, 40, 5]), b, L.c_contig_layout([7, 40]), [1], [1], "row")
        ta = rt.asarray(a, dev).reshape([6, 40, 5])
        tb = rt.asarray(b, dev).reshape([7, 40])
        with pytest.raises(rt.RstsrCudaError):
            rt.vecdot(ta, tb, ([1], [1]))
        # kept (6, 5) vs (5,) broadcasts in row-major; kept (6, 5) vs (6,) in col-major
        b2 = rand_data(rng, 40 * 5, dtype)
        check_vecdot(dev, "row", a, L.c_contig_layout([6, 40, 5]), b2, L.c_contig_layout([40, 5]), [1], [0])
        b3 = rand_data(rng, 40 * 6, dtype)
        check_vecdot(dev_col, "col", a, L.f_contig_layout([6, 40, 5]), b3, L.f_contig_layout([6, 40]), [1], [1])
        # b is a vector contracted against every row / every column
        v = rand_data(rng, 40, dtype)
        check_vecdot(dev, "row", a, L.c_contig_layout([6, 40, 5]), v, L.c_contig_layout([40]), [1], [0])
        check_vecdot(dev, "row", a, L.c_contig_layout([30, 40]), v, L.c_contig_layout([40]), [-1], [-1])
        check_vecdot(dev, "row", a, L.c_contig_layout([40, 30]), v, L.c_contig_layout([40]), [0], [0])
        # stride-0 operand
        one = rand_data(rng, 1, dtype)
        check_vecdot(dev, "row", a, L.c_contig_layout([30, 40]), one, L.Layout((30, 40), (0, 0), 0), [1], [1])


@pytest.mark.parametrize("seed", range(40))
def test_vecdot_random_views(dev, dev_col, seed):
    rng = np.random.default_rng(seed_of("vdrand", seed))
    dtype = [np.float64, np.float32, np.int64, np.int32, np.uint32, np.uint64][seed % 6]
    la, na = random_view(rng, max_ndim=4, max_extent=9)
    if la.ndim == 0:
        return
    # b: an independent view with the SAME shape but different strides
    perm = [int(p) for p in rng.permutation(la.ndim)]
    inv = [perm.index(i) for i in range(la.ndim)]
    lb = L.c_contig_layout([la.shape[p] for p in perm]).transpose(inv)
    if rng.random() < 0.5:
        lb = lb.narrow(int(rng.integers(0, la.ndim)), slice(None, None, -1))
    a, b = rand_data(rng, na, dtype), rand_data(rng, max(lb.size, 1), dtype)
    k = int(rng.integers(1, la.ndim + 1))
    axes = [int(x) for x in rng.permutation(la.ndim)[:k]]
    d, order = (dev, "row") if seed % 2 == 0 else (dev_col, "col")
    check_vecdot(d, order, a, la, b, lb, axes, axes)


def test_vecdot_empty_and_errors(dev):
    a = rt.zeros([4, 0], dev)
    r = rt.vecdot(a, a)  # empty contraction -> zeros (fold from Zero::zero())
    assert r.shape == (4,) and r.to_numpy().tolist() == [0.0] * 4
    e = rt.zeros([0, 5], dev)
    r = rt.vecdot(e, e)
    assert r.shape == (0,)
    x = rt.zeros([3, 4], dev)
    with pytest.raises(rt.RstsrCudaError) as ei:
        rt.vecdot(x, rt.zeros([3, 5], dev))
    assert ei.value.kind == "InvalidLayout"
    with pytest.raises(rt.RstsrCudaError) as ei:
        rt.vecdot(x, x, 2)
    assert ei.value.kind == "InvalidValue"
    with pytest.raises(rt.RstsrCudaError) as ei:
        rt.vecdot(x, x, -3)
    assert ei.value.kind == "InvalidValue"
    # device-level: c must be the broadcast of the kept axes
    c = rt.zeros([5], dev)
    with pytest.raises(rt.RstsrCudaError) as ei:
        dev.vecdot(c.raw, c.layout, x.raw, x.layout, x.raw, x.layout, [1], [1])
    assert ei.value.kind == "InvalidLayout"
    with pytest.raises(rt.RstsrCudaError):
        dev.vecdot(c.raw, c.layout, x.raw, x.layout, rt.zeros([3, 4], dev, dtype=np.float32).raw, x.layout, [1], [1])


def test_vecdot_into_strided_output(dev):
    rng = np.random.default_rng(seed_of("vdinto"))
    a, b = rng.standard_normal((33, 70)), rng.standard_normal((33, 70))
    ta, tb = rt.asarray(a, dev), rt.asarray(b, dev)
    out = rt.full([2, 33], -7.0, dev)
    oc = out[1, ::-1]  # reversed strided row of a larger buffer
    dev.vecdot(oc.raw, oc.layout, ta.raw, ta.layout, tb.raw, tb.layout, [1], [1])
    o = out.to_numpy()
    assert np.all(o[0] == -7.0)
    assert np.allclose(o[1, ::-1], (a * b).sum(1), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32, np.uint64, np.uint32])
def test_allclose_matches_oracle(dev, dev_col, dtype):
    rng = np.random.default_rng(seed_of("ac", np.dtype(dtype).name))
    for shape in ([1], [9], [1000], [37, 41], [5, 6, 7], [1 << 20]):
        n = int(np.prod(shape))
        a = rand_data(rng, n, dtype)
        for d, order, mk in ((dev, "row", L.c_contig_layout), (dev_col, "col", L.f_contig_layout)):
            l = mk(shape)
            cases = [a.copy()]
            for _ in range(3):  # one element off, at a random place, by a just-too-large / just-small-enough amount
                b = a.copy()
                i = int(rng.integers(0, n))
                if np.dtype(dtype).kind == "f":
                    b[i] = b[i] * (1 + rng.choice([2e-5, 0.9e-5, -3e-5])) + rng.choice([0, 1e-9])
                else:
                    b[i] = b[i] + rng.choice([0, 1, 2]).astype(dtype)
                cases.append(b)
            for b in cases:
                want = oracle.tensor_allclose(a, l, b, l, order=order)
                got = rt.allclose(rt.Tensor(upload(d, a), P(l)), rt.Tensor(upload(d, b), P(l)))
                assert got is want
                lt = l.transpose(list(range(l.ndim))[::-1])  # same pairing through a transposed view of both
                assert rt.allclose(rt.Tensor(upload(d, a), P(lt)), rt.Tensor(upload(d, b), P(lt))) is want


def test_allclose_edge_values(dev):
    nan, inf = np.nan, np.inf
    a = np.array([1.0, nan, 3.0, -0.0])
    b = np.array([1.0, nan, 3.0, 0.0])
    l = L.c_contig_layout([4])
    ta, tb = rt.asarray(a, dev), rt.asarray(b, dev)
    assert rt.allclose(ta, tb) is False and oracle.tensor_allclose(a, l, b, l) is False
    assert rt.allclose(ta, tb, equal_nan=True) is True and oracle.tensor_allclose(a, l, b, l, equal_nan=True) is True
    # reference quirk: |inf - inf| = NaN is not <= anything -> not close (NumPy would say close)
    i = rt.asarray(np.array([1.0, inf]), dev)
    assert rt.allclose(i, i) is False
    assert oracle.isclose_scalar(np.float64(inf), np.float64(inf), 1e-5, 1e-8, False) is False
    # rtol scales with |b| only (asymmetric)
    x, y = rt.asarray(np.array([100.0]), dev), rt.asarray(np.array([100.002]), dev)
    assert rt.allclose(x, y, rtol=1.9999e-5, atol=0.0) is oracle.tensor_allclose(np.array([100.0]), L.c_contig_layout([1]),
                                                                                  np.array([100.002]), L.c_contig_layout([1]),
                                                                                  rtol=1.9999e-5, atol=0.0)
    # signed integers: wrapping abs_diff at the extremes (release-mode Rust)
    lo, hi = np.iinfo(np.int32).min, np.iinfo(np.int32).max
    p, q = np.array([lo, hi, -5], dtype=np.int32), np.array([hi, lo, -5], dtype=np.int32)
    l3 = L.c_contig_layout([3])
    assert rt.allclose(rt.asarray(p, dev), rt.asarray(q, dev)) is oracle.tensor_allclose(p, l3, q, l3)
    # f32 difference is rounded in f32 before the f64 comparison
    f, g = np.array([1.0, 16777216.0], dtype=np.float32), np.array([1.0000001, 16777217.0], dtype=np.float32)
    l2 = L.c_contig_layout([2])
    assert rt.allclose(rt.asarray(f, dev), rt.asarray(g, dev), rtol=0.0, atol=1e-7) is oracle.tensor_allclose(f, l2, g, l2, 0.0, 1e-7)


def test_allclose_broadcast_and_errors(dev, dev_col):
    rng = np.random.default_rng(seed_of("acb"))
    row = rng.standard_normal(50)
    m = np.tile(row, (20, 1))
    assert rt.allclose(rt.asarray(m, dev), rt.asarray(row, dev)) is True
    m2 = m.copy()
    m2[13, 7] += 1e-3
    assert rt.allclose(rt.asarray(m2, dev), rt.asarray(row, dev)) is False
    # col-major devices align shapes on the left: (20, 50) vs (50,) does not broadcast, (20, 50) vs (20,) does
    with pytest.raises(rt.RstsrCudaError):
        rt.allclose(rt.asarray(m, dev_col), rt.asarray(row, dev_col))
    col = rng.standard_normal(20)
    mc = np.tile(col[:, None], (1, 50))
    assert rt.allclose(rt.asarray(mc, dev_col), rt.asarray(col, dev_col)) is True
    with pytest.raises(rt.RstsrCudaError) as ei:
        rt.allclose(rt.zeros([0, 3], dev), rt.zeros([0, 3], dev))
    assert ei.value.kind == "InvalidValue" and "zero-size array is not supported for allclose" in str(ei.value)
    with pytest.raises(rt.RstsrCudaError):
        rt.allclose(rt.zeros([3], dev), rt.zeros([3], dev, dtype=np.float32))
