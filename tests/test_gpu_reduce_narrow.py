"""GPU parity for reductions of the narrow integer types (i8 / u8 / i16 / u16): bit-exact against the oracle's C loops
(sum / prod wrap in the element type, as release-mode Rust does) and against NumPy for the index / count ops."""
import numpy as np
import pytest

import oracle
import rstsr_b200 as rt
from oracle import layout as L

from helpers import O, P, random_view, same, seed_of, upload, view_np

pytestmark = pytest.mark.gpu

NARROW = [np.int8, np.uint8, np.int16, np.uint16]


def full_range(rng, n, dtype):
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, n, endpoint=True).astype(dtype)


@pytest.mark.parametrize("dtype", NARROW)
@pytest.mark.parametrize("op", ["sum", "prod", "max", "min"])
def test_reduce_all_and_axes_random_views(dev, dev_col, op, dtype):
    rng = np.random.default_rng(seed_of("narrow", op, np.dtype(dtype).name))
    for it in range(16):
        la, na = random_view(rng, max_ndim=4, max_extent=9)
        a = full_range(rng, na, dtype)
        d = dev if it % 2 == 0 else dev_col
        if la.size or op in ("sum", "prod"):
            got = d.reduce_all(op, upload(d, a), P(la))
            want = oracle.reduce_all(op, a, la)
            assert got == want and np.asarray(got).dtype == np.dtype(dtype), (op, dtype, la)
        if la.ndim == 0 or (la.size == 0 and op in ("max", "min")):
            continue
        k = int(rng.integers(1, la.ndim + 1))
        axes = [int(x) for x in rng.permutation(la.ndim)[:k]]
        raw, lo = d.reduce_axes(op, upload(d, a), P(la), axes)
        want, lw = oracle.reduce_axes(op, a, la, axes)
        assert same(lo, lw)
        assert np.array_equal(view_np(d.to_cpu_vec(raw), O(lo)), view_np(want, lw)), (op, dtype, la, axes)


@pytest.mark.parametrize("dtype", NARROW)
@pytest.mark.parametrize("shape,axis", [((1 << 20,), 0), ((513, 1031), 1), ((1031, 513), 0), ((64, 33, 65), 1)])
def test_reduce_large_shapes(dev, dtype, shape, axis):
    """vectorised (32-byte pack) row and column kernels, split + second pass"""
    rng = np.random.default_rng(seed_of("narrowbig", shape, axis, np.dtype(dtype).name))
    a = full_range(rng, int(np.prod(shape)), dtype)
    t = rt.asarray(a, dev).reshape(list(shape))
    v = a.reshape(shape)
    with np.errstate(over="ignore"):
        assert t.sum_all() == v.sum(dtype=dtype)
        assert t.max_all() == v.max() and t.min_all() == v.min()
        assert np.array_equal(t.sum_axes(axis).to_numpy(), v.sum(axis=axis, dtype=dtype))
        assert np.array_equal(t.max_axes(axis).to_numpy(), v.max(axis=axis))
        assert np.array_equal(t.min_axes(axis).to_numpy(), v.min(axis=axis))
        small = (v % 3).astype(dtype)  # products that do not vanish at once
        ts = rt.asarray(small.reshape(-1), dev).reshape(list(shape))
        assert np.array_equal(ts.prod_axes(axis).to_numpy(), small.prod(axis=axis, dtype=dtype))
    assert t.argmax_all() == int(np.argmax(v)) and t.argmin_all() == int(np.argmin(v))
    assert np.array_equal(t.argmax_axes(axis).to_numpy(), np.argmax(v, axis=axis).astype(np.uint64))
    assert np.array_equal(t.argmin_axes(axis).to_numpy(), np.argmin(v, axis=axis).astype(np.uint64))
    assert t.count_nonzero_all() == int(np.count_nonzero(v))
    assert np.array_equal(t.count_nonzero_axes(axis).to_numpy(), np.count_nonzero(v, axis=axis).astype(np.uint64))


def test_narrow_edge_cases(dev):
    a = rt.asarray(np.array([127, 1], dtype=np.int8), dev)
    assert a.sum_all() == np.int8(-128)  # wraps
    assert rt.asarray(np.array([16, 16], dtype=np.uint8), dev).prod_all() == np.uint8(0)
    assert rt.asarray(np.array([-128, 127, 0], dtype=np.int8), dev).max_all() == 127
    assert rt.asarray(np.array([-128, 127, 0], dtype=np.int8), dev).min_all() == -128
    assert rt.asarray(np.array([65535, 0, 7], dtype=np.uint16), dev).max_all() == 65535
    z = rt.zeros([0], dev, dtype=np.int16)
    assert z.sum_all() == 0 and z.prod_all() == 1
    with pytest.raises(rt.RstsrCudaError) as e:
        z.max_all()
    assert e.value.kind == "InvalidValue"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.asarray(np.array([1, 2], dtype=np.int16), dev).mean_all()
    assert e.value.kind == "UnImplemented"
