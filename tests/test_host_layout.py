"""Product host-side layout algebra (C++ inside librstsr_cuda.so, reached through the C ABI) against the oracle
(independent Python restatement) and the reference KATs.  Runs without a GPU."""
import numpy as np
import pytest

import rstsr_b200 as rt
from oracle import layout as L
from rstsr_b200 import _ffi

from helpers import O, P, random_view, same

ORDERS = [(rt.ROW_MAJOR, L.ROW_MAJOR), (rt.COL_MAJOR, L.COL_MAJOR)]


def test_kats_through_the_c_abi():
    # rearrangement.rs:465-496 via layout_for_array_copy(K) = F-contig of the greedy permutation, permuted back
    l = rt.Layout((3, 2, 6), (3, -180, 15), 782)
    assert l.bounds_index() == (602, 864)  # layoutbase.rs test_bounds_index
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.Layout((3, 2, 6), (3, -180, 15), 15).bounds_index()
    assert e.value.kind == "ValueOutOfRange"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.Layout((3, 2, 6), (3, 4, 7), 1000).check()
    assert e.value.kind == "InvalidLayout"
    rt.Layout((3, 2, 6), (3, -300, 0), 1000).check()
    a, b = rt.broadcast_layout(rt.Layout.contig([8, 1, 6, 3, 1], rt.ROW_MAJOR), rt.Layout.contig([7, 1, 3, 5], rt.COL_MAJOR),
                               rt.ROW_MAJOR)
    assert a.stride == (18, 0, 3, 1, 0) and b.stride == (0, 1, 0, 7, 21)  # broadcast.rs test_broadcast_layout
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.broadcast_layout(rt.Layout.contig([3], 0), rt.Layout.contig([4], 0), rt.ROW_MAJOR)
    assert e.value.kind == "InvalidLayout" and "Broadcasting failed" in str(e.value)


@pytest.mark.parametrize("seed", range(4))
def test_array_copy_reduce_bounds_match_oracle(seed):
    rng = np.random.default_rng(seed)
    for _ in range(600):
        l, _ = random_view(rng, allow_broadcast=True)
        for it, name in ((_ffi.ITER_K, "K"), (_ffi.ITER_C, "C"), (_ffi.ITER_F, "F"), (_ffi.ITER_A, "A")):
            assert same(rt.layout_for_array_copy(P(l), it), L.layout_for_array_copy(l, name)), (l, name)
        assert P(l).bounds_index() == L.bounds_index(l)
        assert P(l).c_contig() == L.c_contig(l) and P(l).f_contig() == L.f_contig(l)
        if l.ndim:
            k = int(rng.integers(1, l.ndim + 1))
            axes = [int(a) for a in rng.permutation(l.ndim)[:k]]
            axes = [a if rng.random() < 0.5 else a - l.ndim for a in axes]
            try:
                want = L.layout_for_reduce(l, axes)
            except L.LayoutError:
                with pytest.raises(rt.RstsrCudaError):
                    rt.layout_for_reduce(P(l), axes)
                continue
            assert same(rt.layout_for_reduce(P(l), axes), want), (l, axes)


@pytest.mark.parametrize("seed", range(4))
def test_broadcast_and_binary_layout_match_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(600):
        l1, _ = random_view(rng, allow_broadcast=True)
        l2, _ = random_view(rng, allow_broadcast=True)
        for po, oo in ORDERS:
            try:
                e1, e2 = L.broadcast_layout(l1, l2, oo)
            except L.LayoutError:
                with pytest.raises(rt.RstsrCudaError) as e:
                    rt.broadcast_layout(P(l1), P(l2), po)
                assert e.value.kind == "InvalidLayout"
                continue
            a1, a2 = rt.broadcast_layout(P(l1), P(l2), po)
            assert same(a1, e1) and same(a2, e2)
            assert same(rt.layout_for_binary_op(a1, a2, po), L.get_layout_for_binary_op(e1, e2, oo)), (e1, e2)


@pytest.mark.parametrize("seed", range(3))
def test_reshapeable_matches_oracle(seed):
    rng = np.random.default_rng(200 + seed)
    for _ in range(800):
        l, _ = random_view(rng, allow_broadcast=False)
        n = l.size
        # random factorisation of n (or a wrong size now and then)
        shape, rem = [], n
        while rem > 1 and len(shape) < 4:
            divs = [d for d in range(1, rem + 1) if rem % d == 0]
            d = int(rng.choice(divs))
            shape.append(d)
            rem //= d
        shape.append(rem)
        rng.shuffle(shape)
        if rng.random() < 0.1:
            shape[0] += 1
        for po, oo in ORDERS:
            try:
                want = L.layout_reshapeable(l, shape, oo)
            except L.LayoutError as err:
                with pytest.raises(rt.RstsrCudaError) as e:
                    rt.layout_reshapeable(P(l), shape, po)
                assert e.value.kind == err.kind
                continue
            got = rt.layout_reshapeable(P(l), shape, po)
            if want is None:
                assert got is None, (l, shape)
            else:
                assert got is not None and same(got, want), (l, shape, got, want)


def test_layout_equal_ignores_strides_of_unit_axes():
    a = rt.Layout((2, 1, 3), (3, 99, 1), 0)
    b = rt.Layout((2, 1, 3), (3, 1, 1), 0)
    assert a.same_as(b)
    assert not a.same_as(rt.Layout((2, 1, 3), (3, 1, 1), 1))  # offsets matter


def test_reduce_axes_errors():
    l = rt.Layout.contig([2, 3, 4], rt.ROW_MAJOR)
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.layout_for_reduce(l, [0, 0])
    assert e.value.kind == "InvalidValue" and "Duplicate" in str(e.value)
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.layout_for_reduce(l, [3])
    assert e.value.kind == "InvalidValue"
    assert rt.layout_for_reduce(l, [-1, 0]).shape == (3,)
