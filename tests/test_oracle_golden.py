"""The oracle against the reference's own known-answer tests.  Each case cites the reference test
it restates (paths inside RESTGroup/rstsr v0.7.10).  No GPU, no product code."""
import numpy as np
import pytest

import oracle
from oracle import layout as L
from oracle.layout import COL_MAJOR, ROW_MAJOR, Layout


def lay(shape, stride, offset=0):
    return Layout(tuple(shape), tuple(stride), offset)


# ---- rstsr-common/src/layout/rearrangement.rs:465-496 (test_greedy_layout) ----
def test_greedy_layout_kats():
    l = L.c_contig_layout([2, 3, 4])
    assert L.greedy_layout(l, False)[0].same_as(L.f_contig_layout([4, 3, 2]))
    assert L.greedy_layout(l, True)[0].same_as(L.f_contig_layout([4, 3, 2]))
    l = L.f_contig_layout([2, 3, 4])
    assert L.greedy_layout(l, False)[0].same_as(l)
    assert L.greedy_layout(l, True)[0].same_as(l)
    l = lay([5, 1, 2, 1, 3, 6], [1000, 10, 10, 40, 0, 100])
    g, _ = L.greedy_layout(l, False)
    assert (g.shape, g.stride) == ((2, 6, 5, 1, 1, 1), (10, 100, 1000, 0, 0, 0))
    g, _ = L.greedy_layout(l, True)
    assert (g.shape, g.stride) == ((1, 1, 3, 2, 6, 5), (10, 40, 0, 10, 100, 1000))
    l = L.f_contig_layout([2, 3, 4]).narrow(1, slice(None, None, -1)).swapaxes(-1, -2)
    assert L.greedy_layout(l, True)[0].same_as(L.f_contig_layout([2, 3, 4]))
    assert L.greedy_layout(l, False)[0].same_as(L.f_contig_layout([2, 3, 4]).narrow(1, slice(None, None, -1)))


# ---- rstsr-common/src/layout/broadcast.rs:281-416 ----
@pytest.mark.parametrize("s1,s2,shape,t1,t2", [
    ([8, 1, 6, 1], [7, 1, 5], (8, 7, 6, 5), "PUPU", "EPUP"),
    ([5, 4], [1], (5, 4), "PP", "EU"),
    ([5, 4], [4], (5, 4), "PP", "EP"),
    ([15, 3, 5], [15, 1, 5], (15, 3, 5), "PPP", "PUP"),
    ([15, 3, 5], [3, 5], (15, 3, 5), "PPP", "EPP"),
    ([15, 3, 5], [3, 1], (15, 3, 5), "PPP", "EPU"),
    ([1, 1, 2], [1, 2], (1, 1, 2), "PPP", "EPP"),
    ([1, 2], [1, 1, 2], (1, 1, 2), "EPP", "PPP"),
])
def test_broadcast_shape_kats(s1, s2, shape, t1, t2):
    names = {"P": "preserve", "U": "upcast", "E": "expand"}
    got = L.broadcast_shape(s1, s2, ROW_MAJOR)
    assert got[0] == shape
    assert got[1] == [names[c] for c in t1]
    assert got[2] == [names[c] for c in t2]


@pytest.mark.parametrize("s1,s2", [([3], [4]), ([2, 1], [8, 4, 3]), ([15, 3, 5], [15, 3])])
def test_broadcast_shape_fail(s1, s2):
    with pytest.raises(L.LayoutError):
        L.broadcast_shape(s1, s2, ROW_MAJOR)


def test_broadcast_layout_kat():
    l1, l2 = L.broadcast_layout(L.c_contig_layout([8, 1, 6, 3, 1]), L.f_contig_layout([7, 1, 3, 5]), ROW_MAJOR)
    assert l1.shape == l2.shape == (8, 7, 6, 3, 5)
    assert l1.stride == (18, 0, 3, 1, 0)
    assert l2.stride == (0, 1, 0, 7, 21)


# ---- rstsr-common/src/layout/layoutbase.rs:696-950 ----
def test_layout_new_kats():
    L.check_layout(lay([3, 2, 6], [3, -300, 15], 917))
    with pytest.raises(L.LayoutError):
        L.check_layout(lay([3, 2, 6], [3, -300, 15], 0))
    with pytest.raises(L.LayoutError):
        L.check_layout(lay([3, 2, 6], [3, 4, 7], 1000))
    L.check_layout(lay([3, 2, 6], [3, -300, 0], 1000))
    L.check_layout(lay([], [], 1000))
    L.check_layout(lay([3, 1, 5], [1, 0, 15], 1))
    L.check_layout(lay([3, 0, 5], [-1, -2, -3], 1))


def test_contig_kats():
    assert L.f_contig(lay([3, 5, 7], [1, 3, 15]))
    assert not L.f_contig(lay([3, 5, 7], [1, 4, 20]))
    assert L.f_contig(lay([], []))
    assert L.f_contig(lay([2, 0, 4], [1, 10, 100]))
    assert L.f_contig(lay([2, 1, 4], [1, 1, 2]))
    assert L.c_contig(lay([3, 5, 7], [35, 7, 1]))
    assert not L.c_contig(lay([3, 5, 7], [36, 7, 1]))
    assert L.c_contig(lay([], []))
    assert L.c_contig(lay([2, 0, 4], [1, 10, 100]))
    assert L.c_contig(lay([2, 1, 4], [4, 1, 1]))


def test_bounds_index_kats():
    assert L.bounds_index(lay([3, 2, 6], [3, -180, 15], 782)) == (602, 864)
    with pytest.raises(L.LayoutError):
        L.bounds_index(lay([3, 2, 6], [3, -180, 15], 15))
    assert L.bounds_index(lay([], [], 10)) == (10, 11)


def test_transpose_kats():
    l = lay([3, 2, 6], [3, -180, 15], 782)
    t = l.transpose([2, 0, 1])
    assert (t.shape, t.stride) == ((6, 3, 2), (15, 3, -180))
    t = l.transpose([-1, 0, 1])
    assert (t.shape, t.stride) == ((6, 3, 2), (15, 3, -180))
    with pytest.raises(L.LayoutError):
        l.transpose([-2, 0, 1])
    with pytest.raises(L.LayoutError):
        l.transpose([1, 0])
    r = l.reverse_axes()
    assert (r.shape, r.stride) == ((6, 2, 3), (15, -180, 3))
    s = l.swapaxes(-1, -2)
    assert (s.shape, s.stride) == ((3, 6, 2), (3, 15, -180))


# ---- rstsr-common/src/layout/iterator.rs:1043-1125 (offset sequences for C / F / K orders) ----
def test_iter_order_kats():
    l = lay([3, 2, 6], [3, -180, 15], 782)
    c = [782, 797, 812, 827, 842, 857, 602, 617, 632, 647, 662, 677, 785, 800, 815, 830, 845, 860, 605, 620, 635,
         650, 665, 680, 788, 803, 818, 833, 848, 863, 608, 623, 638, 653, 668, 683]
    f = [782, 785, 788, 602, 605, 608, 797, 800, 803, 617, 620, 623, 812, 815, 818, 632, 635, 638, 827, 830, 833,
         647, 650, 653, 842, 845, 848, 662, 665, 668, 857, 860, 863, 677, 680, 683]
    k = [602, 605, 608, 617, 620, 623, 632, 635, 638, 647, 650, 653, 662, 665, 668, 677, 680, 683, 782, 785, 788,
         797, 800, 803, 812, 815, 818, 827, 830, 833, 842, 845, 848, 857, 860, 863]
    assert list(L.iter_offsets_col_major(L.translate_to_col_major_unary(l, "C"))) == c
    assert list(L.iter_offsets_col_major(L.translate_to_col_major_unary(l, "F"))) == f
    assert list(L.iter_offsets_col_major(L.translate_to_col_major_unary(l, "K"))) == k
    assert list(L.iter_offsets_col_major(L.f_contig_layout([3, 0, 5]))) == []


# ---- rstsr-core/src/tensor/reduction.rs:417-613 ----
def _sliced_view():
    l = L.c_contig_layout([12, 15, 18]).swapaxes(-1, -2)
    return l.narrow(0, slice(2, -3)).narrow(1, slice(1, -4, 2)).narrow(2, slice(-1, 3, -2))


def _sliced_view_col():
    # col-major device: into_shape([12,15,18]) of arange is F-contiguous (tensor/reduction.rs:453-485)
    l = L.f_contig_layout([12, 15, 18]).swapaxes(-1, -2)
    return l.narrow(0, slice(2, -3)).narrow(1, slice(1, -4, 2)).narrow(2, slice(-1, 3, -2))


@pytest.mark.parametrize("device", ["serial", "rayon"])
def test_sum_all_kats(device):
    a = np.arange(3240, dtype=np.uint64)
    assert oracle.reduce_all("sum", np.arange(24, dtype=np.uint64), L.c_contig_layout([24]), device) == 276
    assert oracle.reduce_all("sum", a, _sliced_view(), device) == 446586
    assert oracle.reduce_all("sum", a, _sliced_view_col(), device) == 403662


@pytest.mark.parametrize("device", ["serial", "rayon"])
def test_sum_axes_kats(device):
    a = np.arange(3240, dtype=np.uint64)
    l = L.c_contig_layout([4, 6, 15, 9]).transpose([2, 0, 3, 1])
    out, lo = oracle.reduce_axes("sum", a, l, [0, -2], device)
    s = oracle.to_numpy(out, lo)
    assert (s[0, 1], s[1, 2], s[3, 5]) == (27270, 154845, 428220)
    l = L.f_contig_layout([4, 6, 15, 9]).transpose([2, 0, 3, 1])  # col-major device
    out, lo = oracle.reduce_axes("sum", a, l, [0, -2], device)
    s = oracle.to_numpy(out, lo)
    assert (s[0, 1], s[1, 2], s[3, 5]) == (217620, 218295, 220185)


@pytest.mark.parametrize("device", ["serial", "rayon"])
def test_min_kats(device):
    v = np.array([8, 4, 2, 9, 3, 7, 2, 8, 1, 6, 10, 5], dtype=np.int64)
    l = L.c_contig_layout([4, 3])
    out, lo = oracle.reduce_axes("min", v, l, [0], device)
    assert list(oracle.to_numpy(out, lo)) == [2, 3, 1]
    out, lo = oracle.reduce_axes("min", v, l, [1], device)
    assert list(oracle.to_numpy(out, lo)) == [2, 3, 1, 5]
    assert oracle.reduce_all("min", v, l, device) == 1


@pytest.mark.parametrize("device", ["serial", "rayon"])
def test_mean_kats(device):
    a = np.arange(24.0)
    l = L.c_contig_layout([2, 3, 4])
    assert oracle.reduce_all("mean", a, l, device) == 11.5
    out, lo = oracle.reduce_axes("mean", a, l, [0, 2], device)
    assert list(oracle.to_numpy(out, lo)) == [7.5, 11.5, 15.5]
    v = l.narrow(0, slice(None, None, -1)).narrow(2, slice(None, None, -2))
    out, lo = oracle.reduce_axes("mean", a, v, [-1, 1], device)
    assert list(oracle.to_numpy(out, lo)) == [18.0, 6.0]
    # col-major device (tensor/reduction.rs:587-613)
    l = L.f_contig_layout([2, 3, 4])
    out, lo = oracle.reduce_axes("mean", a, l, [0, 2], device)
    assert list(oracle.to_numpy(out, lo)) == [9.5, 11.5, 13.5]
    v = l.narrow(0, slice(None, None, -1)).narrow(2, slice(None, None, -2))
    out, lo = oracle.reduce_axes("mean", a, v, [-1, 1], device)
    assert list(oracle.to_numpy(out, lo)) == [15.0, 14.0]


def test_max_min_zero_size_is_error():
    # auto_impl/reduction.rs:51-53,95-97
    with pytest.raises(L.LayoutError):
        oracle.reduce_all("max", np.zeros(0), L.c_contig_layout([0]))
    with pytest.raises(L.LayoutError):
        oracle.reduce_axes("min", np.zeros(0), L.c_contig_layout([0, 3]), [0])


def test_max_ignores_nan_and_starts_from_finite_min():
    # ext_real.rs:70-87: f64::max skips NaN, init = f64::MIN
    a = np.array([np.nan, -np.inf, np.nan])
    assert oracle.reduce_all("max", a, L.c_contig_layout([3])) == np.finfo(np.float64).min
    a = np.array([np.nan, 2.0, np.nan, -1.0])
    assert oracle.reduce_all("max", a, L.c_contig_layout([4])) == 2.0
    assert oracle.reduce_all("min", a, L.c_contig_layout([4])) == -1.0


# ---- rstsr-core/src/tensor/operators/op_binary_arithmetic.rs:992-1143 ----
def _lin(a, b, n):
    return np.linspace(a, b, n)


def _c_order(raw, l):
    return oracle.to_numpy(raw, l).reshape(-1)


def test_add_row_major_kats():
    a, b = _lin(1, 5, 5), _lin(2, 10, 5)
    c, lc = oracle.tensor_binary("add", a, L.c_contig_layout([5]), b, L.c_contig_layout([5]))
    assert list(_c_order(c, lc)) == [3., 6., 9., 12., 15.]
    c, lc = oracle.tensor_binary("add", _lin(1, 6, 6), L.c_contig_layout([2, 3]), _lin(2, 6, 3), L.c_contig_layout([3]))
    assert list(_c_order(c, lc)) == [3., 6., 9., 6., 9., 12.]
    c, lc = oracle.tensor_binary("add", _lin(1, 6, 6), L.c_contig_layout([1, 2, 3]), _lin(1, 10, 10),
                                 L.c_contig_layout([5, 1, 2, 1]))
    assert list(_c_order(c, lc)) == [2., 3., 4., 6., 7., 8., 4., 5., 6., 8., 9., 10., 6., 7., 8., 10., 11., 12., 8., 9.,
                                     10., 12., 13., 14., 10., 11., 12., 14., 15., 16.]
    c, lc = oracle.tensor_binary("add", _lin(1, 9, 9), L.c_contig_layout([3, 3]), _lin(2, 18, 9),
                                 L.c_contig_layout([3, 3]).reverse_axes())
    assert list(_c_order(c, lc)) == [3., 10., 17., 8., 15., 22., 13., 20., 27.]
    flip = L.c_contig_layout([5]).narrow(0, slice(None, None, -1))
    c, lc = oracle.tensor_binary("add", a, flip, b, L.c_contig_layout([5]))
    assert list(_c_order(c, lc)) == [7., 8., 9., 10., 11.]
    c, lc = oracle.tensor_binary("add", a, L.c_contig_layout([5]), b, flip)
    assert list(_c_order(c, lc)) == [11., 10., 9., 8., 7.]


def test_add_col_major_kats():
    # op_binary_arithmetic.rs:1072-1143: results compared on the RAW buffer of c
    c, lc = oracle.tensor_binary("add", _lin(1, 6, 6), L.f_contig_layout([3, 2]), _lin(2, 6, 3), L.f_contig_layout([3]),
                                 COL_MAJOR)
    assert list(c) == [3., 6., 9., 6., 9., 12.]
    c, lc = oracle.tensor_binary("add", _lin(1, 6, 6), L.f_contig_layout([3, 2, 1]), _lin(1, 10, 10),
                                 L.f_contig_layout([1, 2, 1, 5]), COL_MAJOR)
    assert list(c) == [2., 3., 4., 6., 7., 8., 4., 5., 6., 8., 9., 10., 6., 7., 8., 10., 11., 12., 8., 9., 10., 12., 13.,
                       14., 10., 11., 12., 14., 15., 16.]
    c, lc = oracle.tensor_binary("add", _lin(1, 9, 9), L.f_contig_layout([3, 3]), _lin(2, 18, 9),
                                 L.f_contig_layout([3, 3]).reverse_axes(), COL_MAJOR)
    assert list(c) == [3., 10., 17., 8., 15., 22., 13., 20., 27.]


def test_sub_mul_kats():
    a, b = _lin(1, 5, 5), _lin(2, 10, 5)
    c, lc = oracle.tensor_binary("sub", a, L.c_contig_layout([5]), b, L.c_contig_layout([5]))
    assert list(c) == [-1., -2., -3., -4., -5.]
    c, lc = oracle.tensor_binary("mul", a, L.c_contig_layout([5]), b, L.c_contig_layout([5]))
    assert list(c) == [2., 8., 18., 32., 50.]


# ---- rstsr-core/src/tensor/assignment.rs:150-178 (assign with i32 -> f32 cast, fill) ----
def test_assign_cast_and_fill_kats():
    a = np.zeros(15, dtype=np.float32)
    b = np.arange(15, dtype=np.int32)
    oracle.assign(a, L.c_contig_layout([3, 5]), b, L.c_contig_layout([3, 5]))
    assert list(a) == [float(i) for i in range(15)]
    # broadcast assign of a row: layouts already broadcast by the caller (tensor/assignment.rs:26-52)
    a = np.zeros(15, dtype=np.float64)
    row = np.arange(5, dtype=np.float64)
    la, lb = L.broadcast_layout_to_first(L.c_contig_layout([3, 5]), L.c_contig_layout([5]), ROW_MAJOR)
    oracle.assign(a, la, row, lb)
    assert list(a) == [0, 1, 2, 3, 4] * 3
    oracle.fill(a, L.c_contig_layout([3, 5]), 1.5)
    assert (a == 1.5).all()


# ---- rstsr-core/tests/core_func/manipulation/test_to_contig.rs:56-116 ----
def test_to_contig_kats():
    src = np.arange(12, dtype=np.int64)
    t = L.c_contig_layout([3, 4]).reverse_axes()
    assert not L.c_contig(t) and L.f_contig(t)
    r, lr, copied = oracle.tensor_to_contig(src, t, ROW_MAJOR)
    assert copied and L.c_contig(lr) and lr.shape == (4, 3)
    assert oracle.to_numpy(r, lr).tolist() == [[0, 4, 8], [1, 5, 9], [2, 6, 10], [3, 7, 11]]
    r, lr, copied = oracle.tensor_to_contig(src, t, COL_MAJOR)
    assert not copied and L.f_contig(lr)
    a = np.arange(24, dtype=np.int64)
    s = L.c_contig_layout([4, 6]).narrow(0, slice(None, None, 2)).narrow(1, slice(None, None, 2))
    assert (s.shape, s.stride) == ((2, 3), (12, 2))
    r, lr, copied = oracle.tensor_to_contig(a, s, ROW_MAJOR)
    assert copied and lr.stride == (3, 1)
    assert oracle.to_numpy(r, lr).tolist() == [[0, 2, 4], [12, 14, 16]]
    # already contiguous -> view; the opposite order copies (test_to_contig.rs:12-54)
    c3 = L.c_contig_layout([2, 3, 4])
    assert oracle.tensor_to_contig(a, c3, ROW_MAJOR)[2] is False
    r, lr, copied = oracle.tensor_to_contig(a, c3, COL_MAJOR)
    assert copied and L.f_contig(lr)
    assert np.array_equal(oracle.to_numpy(r, lr), a.reshape(2, 3, 4))
    # a C-contiguous slice with a non-zero offset must still copy (tests/test_issues/issue_77.rs)
    sl = L.c_contig_layout([4, 6]).narrow(0, slice(1, None))
    r, lr, copied = oracle.tensor_to_contig(a, sl, ROW_MAJOR)
    assert copied and lr.offset == 0 and np.array_equal(oracle.to_numpy(r, lr), a.reshape(4, 6)[1:])


# ---- reshape incl. F order (rstsr-core/tests/core_func/manipulation/test_reshape.rs:13-70) ----
def test_reshape_kats():
    a = np.arange(24, dtype=np.int64)
    r, l, copied = oracle.tensor_reshape(a, L.c_contig_layout([24]), [2, 3, 4], ROW_MAJOR)
    assert not copied and l.same_as(L.c_contig_layout([2, 3, 4]))
    r, l, copied = oracle.tensor_reshape(a, L.c_contig_layout([2, 3, 4]), [6, -1], ROW_MAJOR)
    assert not copied and l.shape == (6, 4)
    t = L.c_contig_layout([4, 6]).reverse_axes()  # (6,4) strides (1,6)
    r, l, copied = oracle.tensor_reshape(a, t, [24], ROW_MAJOR)
    assert copied
    assert np.array_equal(oracle.to_numpy(r, l), a.reshape(4, 6).T.reshape(-1))
    r, l, copied = oracle.tensor_reshape(a, t, [24], COL_MAJOR)  # F-order flattening of an F-contiguous view: a view
    assert not copied and np.array_equal(oracle.to_numpy(r, l), a)
    r, l, copied = oracle.tensor_reshape(a, L.c_contig_layout([4, 6]), [3, 8], COL_MAJOR)
    assert copied
    assert np.array_equal(oracle.to_numpy(r, l), a.reshape(4, 6).reshape((3, 8), order="F"))
    with pytest.raises(L.LayoutError):
        oracle.tensor_reshape(a, L.c_contig_layout([24]), [5, 5], ROW_MAJOR)


# ---- unrolled_reduce association (cpu_serial/reduction.rs:44-83), SURVEY A.4 ----
def test_unrolled_reduce_association():
    rng = np.random.default_rng(3)
    x = rng.standard_normal(43)
    p = [0.0] * 8
    for j in range(0, 40, 8):
        for k in range(8):
            p[k] = p[k] + x[j + k]
    acc = 0.0
    acc = acc + (p[0] + p[4])
    acc = acc + (p[1] + p[5])
    acc = acc + (p[2] + p[6])
    acc = acc + (p[3] + p[7])
    for v in x[40:]:
        acc = acc + v
    want = 0.0 + acc  # reduce_all_cpu_serial: acc = f_sum(init, unrolled_reduce(run))
    assert oracle.reduce_all("sum", x, L.c_contig_layout([43])) == want


# ---- cross-library check of rstsr-core/tests/tensor_sum.rs:13-107: (4,512,512) f64, vs an independent sum ----
def test_tensor_sum_cross_library():
    rng = np.random.default_rng(42)
    a = rng.random(4 * 512 * 512)
    l = L.c_contig_layout([4, 512, 512])
    for device in ("serial", "rayon"):
        out, lo = oracle.reduce_axes("sum", a, l, [0], device)
        assert np.abs(oracle.to_numpy(out, lo) - a.reshape(4, 512, 512).sum(0)).max() < 1e-6
        out, lo = oracle.reduce_axes("sum", a, l, [-1, -2], device)
        assert np.abs(oracle.to_numpy(out, lo) - a.reshape(4, 512, 512).sum((-1, -2))).max() < 1e-6


def test_vecdot_and_isclose_kats():
    """rstsr-core/src/tensor/linalg/vecdot.rs:41-80 (doc tests) and rstsr-dtype-traits/src/isclose.rs:152-169."""
    c, lc = oracle.tensor_vecdot(np.array([1, 2, 3]), L.c_contig_layout([3]), np.array([4, 5, 6]), L.c_contig_layout([3]), [-1], [-1])
    assert lc.ndim == 0 and int(oracle.to_numpy(c, lc)) == 32
    c, lc = oracle.tensor_vecdot(np.array([1, 2, 3, 4]), L.c_contig_layout([2, 2]), np.array([5, 6, 7, 8]),
                                 L.c_contig_layout([2, 2]), [1], [1])
    assert oracle.to_numpy(c, lc).tolist() == [17, 53]
    a = np.array([0., 5., 0., 0., 0., 10., 0., 6., 8.])
    c, lc = oracle.tensor_vecdot(a, L.c_contig_layout([3, 3]), np.array([0., 0.6, 0.8]), L.c_contig_layout([3]), [1], [0])
    assert np.allclose(oracle.to_numpy(c, lc), [3., 8., 10.], rtol=1e-15)
    f = np.float64
    assert oracle.isclose_scalar(f(1.00001), f(1.00002), 1e-5, 1e-8, False) is True
    assert oracle.isclose_scalar(f(1.00001), f(1.00002), 1e-6, 1e-9, False) is False
    assert oracle.isclose_scalar(np.uint64(100), np.uint64(102), 1e-5, 1e-8, False) is False
    l1 = L.c_contig_layout([1])
    assert oracle.tensor_allclose(np.array([1.00001]), l1, np.array([1.00002]), l1) is True
    assert oracle.tensor_allclose(np.array([100], dtype=np.uint64), l1, np.array([102], dtype=np.uint64), l1) is False


def test_pack_unpack_tri_kats():
    """rstsr-core/src/tensor/operators/op_tri.rs:187-228 (test_pack_tri), both default orders."""
    a = np.arange(48.0)
    lb, la = L.f_contig_layout([3, 4, 4]), L.f_contig_layout([3, 10])
    packed = np.zeros(30)
    oracle.pack_tri(packed, la, a, lb, "L", "row")
    assert oracle.to_numpy(packed, la)[1].tolist() == [1., 4., 16., 7., 19., 31., 10., 22., 34., 46.]
    full = np.zeros(48)
    oracle.unpack_tri(full, lb, packed, la, "L", "Sy", "row")
    assert oracle.to_numpy(full, lb)[0, 1].tolist() == [3., 15., 18., 21.]
    lb, la = L.c_contig_layout([4, 4, 3]), L.c_contig_layout([10, 3])
    packed = np.zeros(30)
    oracle.pack_tri(packed, la, a, lb, "U", "col")
    assert oracle.to_numpy(packed, la)[:, 1].tolist() == [1., 4., 16., 7., 19., 31., 10., 22., 34., 46.]
    full = np.zeros(48)
    oracle.unpack_tri(full, lb, packed, la, "U", "Sy", "col")
    assert oracle.to_numpy(full, lb)[:, 1, 0].tolist() == [3., 15., 18., 21.]
    # antisymmetric: zero diagonal, sign flip across it; N: only the stored triangle
    p = np.arange(1.0, 7.0)
    l6, l33 = L.c_contig_layout([6]), L.c_contig_layout([3, 3])
    out = np.full(9, -1.0)
    oracle.unpack_tri(out, l33, p, l6, "L", "Ay", "row")
    assert out.reshape(3, 3).tolist() == [[0., -2., -4.], [2., 0., -5.], [4., 5., 0.]]
    out = np.full(9, -1.0)
    oracle.unpack_tri(out, l33, p, l6, "U", "N", "row")
    assert out.reshape(3, 3).tolist() == [[1., 2., 3.], [-1., 4., 5.], [-1., -1., 6.]]
    sel, ls = oracle.tensor_index_select(np.arange(12), L.c_contig_layout([3, 4]), 1, [3, -4, 1])
    assert oracle.to_numpy(sel, ls).tolist() == [[3, 0, 1], [7, 4, 5], [11, 8, 9]]
