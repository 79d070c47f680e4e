"""GPU parity: sum / prod / max / min / mean over all or selected axes vs the oracle.
Integers and max/min: bit-exact.  Float sum/prod/mean: |got - want| <= tol * sum|x| with tol = 1e-12 (f64) /
1e-5 (f32) -- the stated tolerance of the north star (summation order differs from the reference's)."""
import numpy as np
import pytest

import oracle
import rstsr_b200 as rt
from oracle import layout as L

from helpers import O, P, rand_data, random_view, same, seed_of, upload, view_np

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def check_close(got, want, a_abs_sum, dtype, op):
    got, want = np.asarray(got), np.asarray(want)
    if np.dtype(dtype).kind in "iub" or op in ("max", "min"):
        assert np.array_equal(got, want), (op, dtype)
    else:
        tol = TOL[np.dtype(dtype)]
        scale = np.maximum(np.asarray(a_abs_sum, dtype=np.float64), 1e-300)
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        assert (err <= tol * scale).all(), (op, dtype, float(err.max()), float(np.min(scale)))


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int64, np.uint32, np.uint64])
@pytest.mark.parametrize("op", ["sum", "prod", "max", "min", "mean"])
def test_reduce_all_random_views(dev, op, dtype):
    if op == "mean" and np.dtype(dtype).kind != "f":
        pytest.skip("mean of integers is not supported by the reference (numpy_differences.md:234-244)")
    rng = np.random.default_rng(seed_of(op, np.dtype(dtype).name))
    for _ in range(12):
        la, na = random_view(rng, max_extent=9)
        a = rand_data(rng, na, dtype)
        if op == "prod" and np.dtype(dtype).kind == "f":
            a = (1 + a * np.dtype(dtype).type(0.01)).astype(dtype)
        got = dev.reduce_all(op, upload(dev, a), P(la))
        want = oracle.reduce_all(op, a, la)
        av = np.abs(view_np(a, la).astype(np.float64))
        scale = av.sum() if op in ("sum", "mean") else np.abs(float(want)) * 8
        if op == "mean":
            scale = scale / max(la.size, 1)
        check_close(got, want, scale, dtype, op)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.uint64])
@pytest.mark.parametrize("op", ["sum", "prod", "max", "min", "mean"])
def test_reduce_axes_random_views(dev, op, dtype):
    if op == "mean" and np.dtype(dtype).kind != "f":
        pytest.skip("mean of integers is not supported by the reference")
    rng = np.random.default_rng(seed_of("axes", op, np.dtype(dtype).name))
    done = 0
    while done < 25:
        la, na = random_view(rng, max_ndim=5, max_extent=7)
        if la.ndim == 0:
            continue
        k = int(rng.integers(1, la.ndim + 1))
        axes = [int(x) for x in rng.permutation(la.ndim)[:k]]
        axes = [x if rng.random() < 0.5 else x - la.ndim for x in axes]
        a = rand_data(rng, na, dtype)
        if op == "prod" and np.dtype(dtype).kind == "f":
            a = (1 + a * np.dtype(dtype).type(0.01)).astype(dtype)
        raw, lo = dev.reduce_axes(op, upload(dev, a), P(la), axes)
        ref, lo_ref = oracle.reduce_axes(op, a, la, axes)
        assert same(lo, lo_ref), (la, axes, lo, lo_ref)  # the CALLEE chooses the layout: must match the reference
        got = view_np(dev.to_cpu_vec(raw), O(lo))
        want = view_np(ref, lo_ref)
        nax = [x % la.ndim for x in axes]
        av = np.abs(view_np(a, la).astype(np.float64))
        scale = av.sum(axis=tuple(nax)) if op in ("sum", "mean") else np.abs(want.astype(np.float64)) * 8
        if op == "mean":
            scale = scale / max(int(np.prod([la.shape[x] for x in nax])), 1)
        check_close(got, want, scale, dtype, op)
        done += 1


def test_reference_reduction_kats_on_device(dev, dev_col):
    """rstsr-core/src/tensor/reduction.rs:417-613 through the Tensor mirror, row- and col-major devices."""
    a = rt.arange(3240, dev, dtype=np.uint64).reshape([12, 15, 18]).swapaxes(-1, -2)[2:-3, 1:-4:2, -1:3:-2]
    assert rt.arange(24, dev, dtype=np.uint64).sum_all() == 276
    assert a.sum_all() == 446586
    s = rt.arange(3240, dev, dtype=np.uint64).reshape([4, 6, 15, 9]).transpose([2, 0, 3, 1]).sum_axes([0, -2])
    sn = s.to_numpy()
    assert (sn[0, 1], sn[1, 2], sn[3, 5]) == (27270, 154845, 428220)
    # col-major device: into_shape is F-ordered -> different numbers (reduction.rs:453-530)
    a = rt.arange(3240, dev_col, dtype=np.uint64).reshape([12, 15, 18]).swapaxes(-1, -2)[2:-3, 1:-4:2, -1:3:-2]
    assert a.sum_all() == 403662
    s = rt.arange(3240, dev_col, dtype=np.uint64).reshape([4, 6, 15, 9]).transpose([2, 0, 3, 1]).sum_axes([0, -2])
    sn = s.to_numpy()
    assert (sn[0, 1], sn[1, 2], sn[3, 5]) == (217620, 218295, 220185)
    v = rt.asarray(np.array([8, 4, 2, 9, 3, 7, 2, 8, 1, 6, 10, 5]), dev).reshape([4, 3])
    assert v.min_axes(0).to_vec().tolist() == [2, 3, 1]
    assert v.min_axes(1).to_vec().tolist() == [2, 3, 1, 5]
    assert v.min_all() == 1
    m = rt.arange(24, dev, dtype=np.float64).reshape([2, 3, 4])
    assert m.mean_all() == 11.5
    assert m.mean_axes([0, 2]).to_vec().tolist() == [7.5, 11.5, 15.5]
    assert m[::-1, :, ::-2].mean_axes([-1, 1]).to_vec().tolist() == [18.0, 6.0]
    m = rt.arange(24, dev_col, dtype=np.float64).reshape([2, 3, 4])
    assert m.mean_axes([0, 2]).to_vec().tolist() == [9.5, 11.5, 13.5]
    assert m[::-1, :, ::-2].mean_axes([-1, 1]).to_vec().tolist() == [15.0, 14.0]


def test_tensor_sum_cross_library_on_device(dev):
    """rstsr-core/tests/tensor_sum.rs:13-107: (4,512,512) f64, seed 42, max abs diff < 1e-6."""
    rng = np.random.default_rng(42)
    a = rng.random(4 * 512 * 512)
    t = rt.asarray(a, dev).reshape([4, 512, 512])
    assert np.abs(t.sum_axes(0).to_numpy() - a.reshape(4, 512, 512).sum(0)).max() < 1e-6
    assert np.abs(t.sum_axes([-1, -2]).to_numpy() - a.reshape(4, 512, 512).sum((-1, -2))).max() < 1e-6


def test_edge_semantics(dev):
    """Empty input, NaN handling and the finite start value (auto_impl/reduction.rs:51-53, ext_real.rs:70-87)."""
    z = rt.asarray(np.zeros(0), dev).reshape([0, 3])
    for op in ("max", "min"):
        with pytest.raises(rt.RstsrCudaError) as e:
            z._reduce(op)
        assert e.value.kind == "InvalidValue" and "zero-size" in str(e.value)
        with pytest.raises(rt.RstsrCudaError):
            z._reduce(op, [0])
    assert z.sum_all() == 0.0 and z.prod_all() == 1.0
    assert z.sum_axes(0).to_numpy().tolist() == [0.0, 0.0, 0.0]
    assert z.sum_axes(1).shape == (0,)
    a = rt.asarray(np.array([np.nan, 2.0, np.nan, -1.0]), dev)
    assert a.max_all() == 2.0 and a.min_all() == -1.0
    allnan = rt.asarray(np.array([np.nan, -np.inf, np.nan]), dev)
    assert allnan.max_all() == np.finfo(np.float64).min  # init is f64::MIN (finite), NaN skipped
    assert np.isnan(rt.asarray(np.array([1.0, np.nan]), dev).sum_all())
    # integer sums wrap (release-mode Rust)
    big = rt.asarray(np.full(4, 2**62, dtype=np.int64), dev)
    assert big.sum_all() == np.int64(0)
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.arange(4, dev, dtype=np.int32).mean_all()
    assert e.value.kind == "UnImplemented"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.arange(6, dev).reshape([2, 3]).sum_axes([0, 0])
    assert e.value.kind == "InvalidValue"


def test_broadcast_axes_use_the_mathematical_result(dev):
    """Reductions over / along stride-0 axes: the reference's repeat count is suspect; the product
    and the oracle both give the mathematically intended result."""
    v = np.array([1.0, 2.0, 3.0])
    t = rt.asarray(v, dev).broadcast_to([4, 3])
    assert t.sum_axes(0).to_numpy().tolist() == [4.0, 8.0, 12.0]
    assert t.sum_axes(1).to_numpy().tolist() == [6.0] * 4
    assert t.sum_all() == 24.0 and t.max_all() == 3.0


RED_SHAPES = [
    ("row kernel vector", (513, 4096), [1]),
    ("row kernel scalar tail", (37, 1001), [1]),
    ("row kernel split (few outputs)", (3, 1 << 18), [1]),
    ("col kernel vector + split", (5000, 512), [0]),
    ("col kernel narrow", (4097, 24), [0]),
    ("col kernel with kept outer", (7, 300, 256), [1]),
    ("row kernel multi-dim reduced", (40, 6, 64), [0, 2]),
    ("generic (no unit stride)", (30, 40, 10), "stepped"),
    ("reduce everything via axes", (64, 64, 16), [0, 1, 2]),
    ("tiny", (3, 2), [0]),
    # few reduced rows over a contiguous kept axis: reduce_cols_small_kernel (one thread per column pack, no row-lanes)
    ("small rows 3 x packs", (3, 4096), [0]),
    ("small rows 3 x odd columns (scalar)", (3, 4099), [0]),
    ("small rows 2, kept outer", (50, 2, 1028), [1]),
    ("small rows 4, kept outer, odd columns", (9, 4, 333), [1]),
    ("small rows multi-dim reduced (2 x 3)", (2, 7, 3, 260), [0, 2]),
    ("small rows 16 x short kept rows", (300, 16, 64), [1]),
    ("small rows 1", (1, 5000), [0]),
    ("17 rows: general column kernel", (17, 5000), [0]),
]


@pytest.mark.parametrize("name,shape,axes", RED_SHAPES, ids=[r[0] for r in RED_SHAPES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64])
@pytest.mark.parametrize("op", ["sum", "max", "mean"])
def test_reduce_kernel_variants(dev, name, shape, axes, dtype, op):
    if op == "mean" and np.dtype(dtype).kind != "f":
        pytest.skip("mean of integers is not supported by the reference (numpy_differences.md:234-244)")
    rng = np.random.default_rng(seed_of(name, op))
    n = int(np.prod(shape))
    a = rand_data(rng, n, dtype)
    la = L.c_contig_layout(shape)
    if axes == "stepped":
        la = la.narrow(2, slice(None, None, 2)).narrow(0, slice(None, None, 3))
        axes = [1]
    raw, lo = dev.reduce_axes(op, upload(dev, a), P(la), axes)
    ref, lo_ref = oracle.reduce_axes(op, a, la, axes)
    assert same(lo, lo_ref)
    got, want = view_np(dev.to_cpu_vec(raw), O(lo)), view_np(ref, lo_ref)
    av = np.abs(view_np(a, la).astype(np.float64)).sum(axis=tuple(axes))
    check_close(got, want, av, dtype, op)
    # F-contiguous input: the output follows the input's memory order (K order, rearrangement.rs:144-148)
    lf = L.f_contig_layout(shape)
    raw, lo = dev.reduce_axes(op, upload(dev, a), P(lf), axes if axes != "stepped" else [1])
    ref, lo_ref = oracle.reduce_axes(op, a, lf, axes)
    assert same(lo, lo_ref)
    check_close(view_np(dev.to_cpu_vec(raw), O(lo)), view_np(ref, lo_ref),
                np.abs(view_np(a, lf).astype(np.float64)).sum(axis=tuple(axes)), dtype, op)


def test_results_are_run_to_run_deterministic(dev):
    rng = np.random.default_rng(1)
    a = rng.standard_normal(1 << 22)
    raw = upload(dev, a)
    l = rt.Layout((1 << 22,), (1,))
    first = dev.reduce_all("sum", raw, l)
    for _ in range(5):
        assert dev.reduce_all("sum", raw, l) == first  # fixed two-pass order, no float atomics


def test_reduce_axes_into_strided_output(dev):
    rng = np.random.default_rng(2)
    a = rng.random(200 * 300)
    out0 = np.full(600, -1.0)
    raw_o = upload(dev, out0)
    lo = rt.Layout((300,), (2,), 1)
    dev.reduce_axes_into("sum", upload(dev, a), rt.Layout((200, 300), (300, 1)), [0], raw_o, lo)
    got = dev.to_cpu_vec(raw_o)
    assert np.allclose(got[1::2], a.reshape(200, 300).sum(0), rtol=1e-12) and (got[0::2] == -1.0).all()
