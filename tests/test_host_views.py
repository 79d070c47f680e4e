"""View manipulations and shape broadcasting of the Python mirror that never reach the GPU (layouts only, no device
needed): the bodies of the reference's NumPy-derived conformance tests rstsr-core/tests/core_func/manipulation/test_{expand_dims,
squeeze,moveaxis,flip,broadcast_shapes}.rs, plus a seeded cross-check of every view against NumPy on an index array."""
import itertools

import numpy as np
import pytest
from numpy.lib.stride_tricks import as_strided

import rstsr_b200 as rt


def view(shape, order="C", offset=0):
    """layout-only tensor over a contiguous buffer"""
    shape = tuple(shape)
    item = 1
    stride = [0] * len(shape)
    for i in (range(len(shape) - 1, -1, -1) if order == "C" else range(len(shape))):
        stride[i] = item
        item *= max(shape[i], 1)
    return rt.Tensor(None, rt.Layout(shape, tuple(stride), offset))


def realise(t, n):
    """the buffer indices a view addresses, as an array of its shape"""
    base = np.arange(n, dtype=np.int64)
    return as_strided(base[t.layout.offset:], shape=t.shape, strides=[s * 8 for s in t.stride])


def err_kind(fn):
    with pytest.raises(rt.RstsrCudaError) as e:
        fn()
    return e.value.kind


# ---- test_expand_dims.rs ----
def test_expand_dims_functionality():
    s = (2, 3, 4, 5)
    a = view(s)
    for axis in range(-5, 4):
        b = a.expand_dims(axis)
        assert b.shape[axis] == 1
        assert b.squeeze(axis % b.ndim).shape == s


def test_expand_dims_axis_tuple():
    a = view((3, 3, 3))
    assert a.expand_dims([0, 1, 2]).shape == (1, 1, 1, 3, 3, 3)
    assert a.expand_dims([0, -1, -2]).shape == (1, 3, 3, 3, 1, 1)
    assert a.expand_dims([0, 3, 5]).shape == (1, 3, 3, 1, 3, 1)
    assert a.expand_dims([0, -3, -5]).shape == (1, 1, 3, 1, 3, 3)
    assert a.unsqueeze([0, -3, -5]).shape == (1, 1, 3, 1, 3, 3)


def test_expand_dims_errors():
    a = view((2, 3, 4, 5))
    assert err_kind(lambda: a.expand_dims(-6)) == "InvalidValue"
    assert err_kind(lambda: a.expand_dims(5)) == "InvalidValue"
    b = view((3, 3, 3))
    assert err_kind(lambda: b.expand_dims([0, -6])) == "InvalidValue"
    assert err_kind(lambda: b.expand_dims([0, 5])) == "InvalidValue"
    assert err_kind(lambda: b.expand_dims([1, 1])) == "InvalidValue"  # test_repeated_axis


# ---- test_squeeze.rs ----
def test_squeeze_basic():
    assert view((1, 3, 3)).squeeze().shape == (3, 3)
    z = view((1, 3, 1))
    assert z.squeeze().shape == (3,)
    assert z.squeeze(0).shape == (3, 1)
    assert z.squeeze(-1).shape == (1, 3)
    assert z.squeeze(2).shape == (1, 3)
    a = view((3, 1)).expand_dims(0)
    assert a.squeeze().shape == (3,)
    assert a.squeeze(0).shape == (3, 1)
    assert a.squeeze(2).shape == (1, 3)
    assert a.squeeze(-1).shape == (1, 3)


def test_squeeze_axis_and_errors():
    a = view((1, 3, 1))
    assert np.array_equal(realise(a.squeeze(), 3), np.arange(3))
    assert a.squeeze([0]).shape == (3, 1) and a.squeeze([2]).shape == (1, 3)
    assert err_kind(lambda: a.squeeze(1)) == "InvalidValue"          # the axis is not of extent 1
    assert err_kind(lambda: a.squeeze([0, 0])) == "InvalidValue"     # "Same axes is not allowed here."
    assert err_kind(lambda: a.squeeze(-4)) == "InvalidValue"         # "Some negative index is too small."
    assert err_kind(lambda: a.squeeze(3)) == "InvalidValue"


def test_squeeze_keeps_the_addressed_elements():
    for shape, want in (((20, 10, 10, 1, 1), (20, 10, 10)), ((20, 1, 10, 20, 1), (20, 10, 20)), ((1, 1, 20, 10), (20, 10))):
        n = int(np.prod(shape))
        got = realise(view(shape).squeeze(), n)
        assert got.shape == want and np.array_equal(got, np.arange(n).reshape(want))
    assert view((1, 1, 1)).squeeze().shape == () and view((1,)).squeeze().shape == ()


# ---- test_moveaxis.rs ----
def test_moveaxis_move_to_end_and_new_position():
    x = view((5, 6, 7))
    for source, expected in ((0, (6, 7, 5)), (1, (5, 7, 6)), (2, (5, 6, 7)), (-1, (5, 6, 7))):
        assert x.moveaxis(source, -1).shape == expected
    y = view((1, 2, 3, 4))
    for source, destination, expected in ((0, 1, (2, 1, 3, 4)), (1, 2, (1, 3, 2, 4)), (1, -1, (1, 3, 4, 2))):
        assert y.moveaxis(source, destination).shape == expected


def test_moveaxis_preserve_order_and_many_axes():
    x = view((1, 2, 3, 4))
    for source, destination in ((0, 0), (3, -1), (-1, 3), ([0, -1], [0, -1]), ([2, 0], [2, 0]), (range(4), range(4))):
        assert x.moveaxis(source, destination).shape == (1, 2, 3, 4)
    y = view((0, 1, 2, 3))
    for source, destination, expected in (([0, 1], [2, 3], (2, 3, 0, 1)), ([2, 3], [0, 1], (2, 3, 0, 1)),
                                          ([0, 1, 2], [2, 3, 0], (2, 3, 0, 1)), ([3, 0], [1, 0], (0, 3, 1, 2)),
                                          ([0, 3], [0, 1], (0, 3, 1, 2))):
        assert y.moveaxis(source, destination).shape == expected
    z = view((3, 4, 5))
    assert z.transpose().shape == z.swapaxes(0, -1).shape == (5, 4, 3)
    assert z.moveaxis([0, 1], [-1, -2]).shape == (5, 4, 3)
    assert z.moveaxis([0, 1, 2], [-1, -2, -3]).shape == (5, 4, 3)


def test_moveaxis_errors():
    x = view((1, 2, 3, 4, 5))
    for bad in (lambda: x.moveaxis([0, 1], [0]), lambda: x.moveaxis([0, 0], [1, 2]), lambda: x.moveaxis([0, 1], [2, 2]),
                lambda: x.moveaxis(5, 0), lambda: x.moveaxis(-6, 0), lambda: x.moveaxis(0, 5), lambda: x.moveaxis(0, -6)):
        assert err_kind(bad) == "InvalidValue"


# ---- test_flip.rs ----
def test_flip_axes_and_errors():
    assert err_kind(lambda: view((4,)).flip(1)) == "InvalidValue"
    a = view((4, 4))
    for bad in (lambda: a.flip(2), lambda: a.flip(-3), lambda: a.flip([0, 3]), lambda: a.flip([1, 1])):
        assert err_kind(bad) == "InvalidValue"
    n = 2 * 3 * 4
    idx = np.arange(n).reshape(2, 3, 4)
    t = view((2, 3, 4))
    for i in range(3):
        assert np.array_equal(realise(t.flip(i), n), np.flip(idx, i))
    assert np.array_equal(realise(t.flip(), n), np.flip(idx))               # every axis
    assert np.array_equal(realise(t.flip([]), n), idx)                       # axis = (): nothing
    assert np.array_equal(realise(t.flip([0, 2]), n), np.flip(idx, (0, 2)))
    assert np.array_equal(realise(t.flip([1, -1]), n), np.flip(idx, (1, 2)))


# ---- every view against NumPy on an index array ----
@pytest.mark.parametrize("seed", range(6))
def test_views_address_what_numpy_addresses(seed):
    rng = np.random.default_rng(seed)
    for _ in range(40):
        nd = int(rng.integers(1, 5))
        shape = tuple(int(rng.integers(1, 4)) for _ in range(nd))
        n = int(np.prod(shape))
        order = "C" if rng.random() < 0.5 else "F"
        t, ref = view(shape, order), np.arange(n).reshape(shape, order=order)
        for _ in range(4):
            op = rng.choice(["flip", "moveaxis", "expand_dims", "squeeze", "transpose"])
            if op == "flip":
                ax = [int(a) for a in rng.permutation(t.ndim)[: int(rng.integers(0, t.ndim + 1))]]
                t, ref = t.flip(ax), np.flip(ref, tuple(ax))
            elif op == "moveaxis" and t.ndim:
                k = int(rng.integers(1, t.ndim + 1))
                src = [int(a) - (t.ndim if rng.random() < 0.5 else 0) for a in rng.permutation(t.ndim)[:k]]
                dst = [int(a) for a in rng.permutation(t.ndim)[:k]]
                t, ref = t.moveaxis(src, dst), np.moveaxis(ref, src, dst)
            elif op == "expand_dims" and t.ndim < 6:
                k = int(rng.integers(1, 3))
                ax = [int(a) for a in rng.permutation(t.ndim + k)[:k]]
                t, ref = t.expand_dims(ax), np.expand_dims(ref, tuple(ax))
            elif op == "squeeze":
                units = [i for i, d in enumerate(t.shape) if d == 1]
                if units and rng.random() < 0.5:
                    ax = [units[int(rng.integers(0, len(units)))]]
                    t, ref = t.squeeze(ax), np.squeeze(ref, tuple(ax))
                else:
                    t, ref = t.squeeze(), np.squeeze(ref)
            elif op == "transpose" and t.ndim:
                perm = [int(a) for a in rng.permutation(t.ndim)]
                t, ref = t.transpose(perm), np.transpose(ref, perm)
            assert t.shape == ref.shape
            assert np.array_equal(realise(t, n), ref), (shape, order, op)


def test_moveaxis_all_pairs_match_numpy():
    shape = (2, 3, 4, 5)
    n = int(np.prod(shape))
    t, ref = view(shape), np.arange(n).reshape(shape)
    for s, d in itertools.product(range(-4, 4), repeat=2):
        assert np.array_equal(realise(t.moveaxis(s, d), n), np.moveaxis(ref, s, d))


# ---- test_broadcast_shapes.rs (through the library's host algebra, rc_layout_broadcast, and the oracle) ----
BROADCAST_OK = [
    ([], ()), ([()], ()), ([(7,)], (7,)), ([(1, 2), (2,)], (1, 2)), ([(1, 1)], (1, 1)), ([(1, 1), (3, 4)], (3, 4)),
    ([(6, 7), (5, 6, 1), (7,), (5, 1, 7)], (5, 6, 7)), ([(5, 6, 1)], (5, 6, 1)), ([(1, 3), (3, 1)], (3, 3)),
    ([(1, 0), (0, 0)], (0, 0)), ([(0, 1), (0, 0)], (0, 0)), ([(1, 0), (0, 1)], (0, 0)), ([(1, 1), (0, 0)], (0, 0)),
    ([(1, 1), (1, 0)], (1, 0)), ([(1, 1), (0, 1)], (0, 1)), ([(), (0,)], (0,)), ([(0,), (0, 0)], (0, 0)),
    ([(0,), (0, 1)], (0, 0)), ([(1,), (0, 0)], (0, 0)), ([(), (0, 0)], (0, 0)), ([(1, 1), (0,)], (1, 0)),
    ([(1,), (0, 1)], (0, 1)), ([(1,), (1, 0)], (1, 0)), ([(), (1, 0)], (1, 0)), ([(), (0, 1)], (0, 1)),
    ([(1,), (3,)], (3,)), ([(2,), (3, 2)], (3, 2)),
    ([(1, 2)] * 32, (1, 2)), ([(1, 2)] * 100, (1, 2)), ([(2,)] * 32 + [()], (2,)),
]
BROADCAST_BAD = [[(3,), (4,)], [(2, 3), (2,)], [(3,), (3,), (4,)], [(1, 3, 4), (2, 3, 3)],
                 [(1, 2), (3, 1), (3, 2), (10, 5)], [(2,), (2, 3)], [(2,)] * 32 + [(3,)]]


def _oracle_broadcast_shapes(shapes, order):
    from oracle import layout as OL
    out = ()
    for s in shapes:
        out = tuple(OL.broadcast_shape(out, tuple(s), order)[0])
    return out


def test_broadcast_shapes_row_major():
    from oracle import layout as OL
    for shapes, want in BROADCAST_OK:
        assert rt.broadcast_shapes(shapes, rt.ROW_MAJOR) == want, shapes
        assert _oracle_broadcast_shapes(shapes, OL.ROW_MAJOR) == want, shapes
        if all(len(s) for s in shapes):  # NumPy agrees (row-major rule)
            assert np.broadcast_shapes(*shapes) == want
    for shapes in BROADCAST_BAD:
        assert err_kind(lambda: rt.broadcast_shapes(shapes, rt.ROW_MAJOR)) == "InvalidLayout"
        with pytest.raises(OL.LayoutError):
            _oracle_broadcast_shapes(shapes, OL.ROW_MAJOR)


def test_broadcast_shapes_col_major():
    """left-aligned rule of a column-major device (broadcast.rs:34-37)"""
    from oracle import layout as OL
    for shapes, want in (([(1, 6, 1, 8), (5, 1, 7)], (5, 6, 7, 8)), ([(4, 5), (4,)], (4, 5)),
                         ([(5, 3, 15), (5, 1, 15)], (5, 3, 15)), ([(5, 3, 15), (5, 3)], (5, 3, 15))):
        assert rt.broadcast_shapes(shapes, rt.COL_MAJOR) == want
        assert _oracle_broadcast_shapes(shapes, OL.COL_MAJOR) == want
    assert err_kind(lambda: rt.broadcast_shapes([(3,), (2, 1)], rt.COL_MAJOR)) == "InvalidLayout"
    with pytest.raises(OL.LayoutError):
        _oracle_broadcast_shapes([(3,), (2, 1)], OL.COL_MAJOR)


# ---- test_transpose.rs ----
def test_transpose_and_swapaxes_kats():
    a = view((2, 2))
    assert np.array_equal(realise(a.transpose(), 4), np.arange(4).reshape(2, 2).T)
    for bad in ([0], [0, 0], [0, 1, 2]):
        err_kind(lambda: a.transpose(bad))
    arr = view((2, 3))
    want = np.arange(6).reshape(2, 3).T
    assert np.array_equal(realise(arr.transpose([1, 0]), 6), want)
    assert np.array_equal(realise(arr.transpose([-1, -2]), 6), want)
    # swapaxes: every pair of axes (negative ones too) is a view of the same buffer with NumPy's shape and elements
    shape = (3, 4, 5, 6)
    n = int(np.prod(shape))
    src, t = np.arange(n).reshape(shape), view(shape)
    for bad in ((-5, 0), (4, 0), (0, -5), (0, 4)):
        assert err_kind(lambda: t.swapaxes(*bad)) == "InvalidValue"
    for i in range(-4, 4):
        for j in range(-4, 4):
            c = t.swapaxes(i, j)
            assert c.layout.offset == 0 and not c.owned
            assert np.array_equal(realise(c, n), np.swapaxes(src, i, j))
