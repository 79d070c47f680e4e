"""GPU parity: device-side creation -- arange / linspace / tril / triu / zeros / ones / full.
Reference: rstsr-native-impl/src/cpu_rayon/creation.rs:8-131, cpu_serial/op_tri.rs:524-590 and the KATs of
rstsr-core/tests/core_func/creation/test_{arange,linspace,tril_triu,full,zeros,ones}.rs."""
import math

import numpy as np
import pytest

import rstsr_b200 as rt

pytestmark = pytest.mark.gpu


def ref_arange(start, end, step, dtype):
    """arange_by_primitive_f64 / _isize (cpu_rayon/creation.rs:8-55), restated with numpy scalars."""
    dt = np.dtype(dtype)
    if dt.kind == "f":
        s, e, st = float(dt.type(start)), float(dt.type(end)), float(dt.type(step))
        n = max(int(math.ceil((e - s) / st)), 0)
        out = np.array([dt.type(s + i * st) for i in range(n)], dtype=dt)
        if n and ((st > 0 and out[-1] >= dt.type(end)) or (st < 0 and out[-1] <= dt.type(end))):
            out = out[:-1]
        return out
    s, e, st = int(start), int(end), int(step)
    n = max(int(math.ceil((e - s) / st)), 0)
    out = [s + i * st for i in range(n)]
    if out and ((st > 0 and out[-1] >= e) or (st < 0 and out[-1] <= e)):
        out.pop()
    return np.array(out, dtype=dt)


@pytest.mark.parametrize("dtype,args", [
    (np.float64, (0.0, 1.0, 0.1)), (np.float64, (1.0, -2.0, -0.3)), (np.float64, (0.0, 0.0, 1.0)), (np.float64, (5.0, 1.0, 1.0)),
    (np.float32, (0.0, 1.0, 0.1)), (np.float32, (0.5, 100.25, 0.75)), (np.float64, (0.0, 1e6, 1.0)),
    (np.int32, (0, 10, 1)), (np.int32, (3, 40, 7)), (np.int64, (10, -10, -3)), (np.uint64, (0, 3240, 1)), (np.uint32, (5, 5, 2)),
])
def test_arange(dev, dtype, args):
    raw = dev.arange_impl(*args, dtype)
    want = ref_arange(*args, dtype)
    assert raw.len == want.size
    assert np.array_equal(dev.to_cpu_vec(raw), want)


def test_arange_zero_step_is_invalid(dev):
    with pytest.raises(rt.RstsrCudaError) as e:
        dev.arange_impl(0.0, 1.0, 0.0, np.float64)
    assert e.value.kind == "InvalidValue"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,endpoint", [(0, True), (1, True), (5, True), (5, False), (1001, True), (64, False)])
def test_linspace(dev, dtype, n, endpoint):
    dt = np.dtype(dtype)
    start, end = dt.type(1.0), dt.type(5.0)
    raw = dev.linspace_impl(start, end, n, endpoint, dtype)
    if n == 0:
        assert raw.len == 0
        return
    if n == 1:
        want = np.array([start], dtype=dt)
    else:
        step = (end - start) / dt.type(n - 1 if endpoint else n)
        want = np.array([start + dt.type(i) * step for i in range(n)], dtype=dt)
    assert np.array_equal(dev.to_cpu_vec(raw), want)
    if n == 5 and endpoint:  # auto_impl/creation.rs test_linspace: linspace(1, 5, 5) == [1, 2, 3, 4, 5]
        assert dev.to_cpu_vec(raw).tolist() == [1.0, 2.0, 3.0, 4.0, 5.0]


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.uint8])
@pytest.mark.parametrize("shape", [(5, 5), (4, 7), (7, 4), (3, 6, 5), (2, 3, 4, 4)])
@pytest.mark.parametrize("k", [0, 1, -1, 3, -5])
def test_tril_triu(dev, dtype, shape, k):
    rng = np.random.default_rng(abs(k) + len(shape))
    a = (rng.integers(1, 50, int(np.prod(shape)))).astype(dtype)
    for fn, ref in (("tril_impl", np.tril), ("triu_impl", np.triu)):
        raw = dev.outof_cpu_vec(a)
        getattr(dev, fn)(raw, rt.Layout.contig(shape, rt.ROW_MAJOR), k)
        assert np.array_equal(dev.to_cpu_vec(raw).reshape(shape), ref(a.reshape(shape), k)), (fn, shape, k)
    # on a transposed (strided) view of the last two axes: tril of the view == triu of the base, transposed
    raw = dev.outof_cpu_vec(a)
    l = rt.Layout.contig(shape, rt.ROW_MAJOR)
    lt = rt.Layout(l.shape[:-2] + (l.shape[-1], l.shape[-2]), l.stride[:-2] + (l.stride[-1], l.stride[-2]), 0)
    dev.tril_impl(raw, lt, k)
    want = np.swapaxes(np.tril(np.swapaxes(a.reshape(shape), -1, -2), k), -1, -2)
    assert np.array_equal(dev.to_cpu_vec(raw).reshape(shape), want)


def test_zeros_ones_full(dev):
    assert (dev.to_cpu_vec(dev.zeros_impl(np.float64, 1000)) == 0).all()
    assert (dev.to_cpu_vec(dev.ones_impl(np.int32, 77)) == 1).all()
    assert (dev.to_cpu_vec(dev.full_impl(np.float32, 129, 2.5)) == np.float32(2.5)).all()
    assert dev.to_cpu_vec(dev.full_impl(np.bool_, 5, True)).tolist() == [True] * 5
    t = rt.full([3, 4], 7, dev, dtype=np.int64)
    assert t.to_numpy().tolist() == [[7] * 4] * 3
