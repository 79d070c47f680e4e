"""GPU parity for the "next" reductions: var / std / l2_norm / argmin / argmax / all / any /
count_nonzero -- same kernels as sum/max with other monoids.  KATs: rstsr-core/src/tensor/reduction.rs:615-909."""
import numpy as np
import pytest

import oracle
import rstsr_b200 as rt
from oracle import layout as L

from helpers import O, P, rand_data, random_view, same, seed_of, upload, view_np

pytestmark = pytest.mark.gpu


def test_reference_kats(dev):
    v = np.array([8, 4, 2, 9, 3, 7, 2, 8, 1, 6, 10, 5], dtype=np.float64)
    a = rt.asarray(v, dev).reshape([4, 3])
    # tensor/reduction.rs test_var / test_std: np.var / np.std of the same data
    assert abs(a.var_all() - np.var(v)) < 1e-12
    assert np.allclose(a.var_axes(0).to_numpy(), np.var(v.reshape(4, 3), axis=0), rtol=1e-12)
    assert np.allclose(a.var_axes(1).to_numpy(), np.var(v.reshape(4, 3), axis=1), rtol=1e-12)
    assert abs(a.std_all() - np.std(v)) < 1e-12
    assert np.allclose(a.std_axes(0).to_numpy(), np.std(v.reshape(4, 3), axis=0), rtol=1e-12)
    assert abs(a.l2_norm_all() - np.linalg.norm(v)) < 1e-12
    assert np.allclose(a.l2_norm_axes(0).to_numpy(), np.linalg.norm(v.reshape(4, 3), axis=0), rtol=1e-13)
    # argmin / argmax (tensor/reduction.rs:816-866): first occurrence, row-major flattened index
    i = rt.asarray(v.astype(np.int64), dev).reshape([4, 3])
    assert i.argmin_all() == 8 and i.argmax_all() == 10
    assert i.argmin_axes(0).to_numpy().tolist() == [2, 1, 2]   # columns: [8,9,2,6] [4,3,8,10] [2,7,1,5]
    assert i.argmin_axes(1).to_numpy().tolist() == [2, 1, 2, 2]
    assert i.argmax_axes(0).to_numpy().tolist() == [1, 3, 1]
    ties = rt.asarray(np.array([3, 1, 1, 3, 1, 3]), dev)
    assert ties.argmin_all() == 1 and ties.argmax_all() == 0
    b = rt.asarray(np.array([True, True, False, True]), dev).reshape([2, 2])
    assert b.all_all() is False and b.any_all() is True
    assert b.all_axes(0).to_numpy().tolist() == [False, True]
    assert b.any_axes(1).to_numpy().tolist() == [True, True]
    assert b.count_nonzero_all() == 3                      # OpSumBoolAPI: sum of a bool tensor
    z = rt.asarray(np.array([0.0, 1.5, 0.0, -2.0, 0.0, 3.0]), dev).reshape([2, 3])
    assert z.count_nonzero_all() == 3
    assert z.count_nonzero_axes(0).to_numpy().tolist() == [1, 1, 1]
    assert z.count_nonzero_axes(1).to_numpy().tolist() == [1, 2]


def test_arg_nan_rule(dev):
    """y < NaN is false and f_comp(None, y) is true (cpu_serial/reduction.rs:436-470): NaN is skipped unless first."""
    a = rt.asarray(np.array([2.0, np.nan, 1.0, np.nan, 1.0]), dev)
    assert a.argmin_all() == 2 and a.argmax_all() == 0
    a = rt.asarray(np.array([np.nan, 5.0, 1.0]), dev)
    assert a.argmin_all() == 0 and a.argmax_all() == 0
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.asarray(np.zeros(0), dev).argmin_all()
    assert e.value.kind == "InvalidLayout"


def test_arg_worst_value_rows(dev):
    """Rows whose extreme IS the type's worst value (the empty accumulator of the kernels carries it): all -inf / +inf,
    all INT_MIN / INT_MAX, NaN in front of them, long rows split over many threads and both kernel families (rows, columns)
    -- the answer is the reference's fold (element 0 unconditionally, then strictly better values only)."""
    import oracle
    cases = []
    for n in (5, 300, 5000):
        cases += [np.full(n, -np.inf), np.full(n, np.inf), np.r_[np.nan, np.full(n - 1, -np.inf)],
                  np.r_[-np.inf, np.nan, np.full(n - 2, -np.inf)], np.r_[np.full(n - 1, -np.inf), 1.0],
                  np.r_[np.nan, np.arange(n - 1.0)], np.r_[np.full(n // 2, np.inf), -np.inf, np.full(n - n // 2 - 1, np.inf)]]
    for row in cases:
        for dt in (np.float64, np.float32):
            v = row.astype(dt)
            t = rt.asarray(v, dev)
            for op, is_max in (("argmax", True), ("argmin", False)):
                want = oracle._arg_fold(list(v), is_max)
                assert getattr(t, op + "_all")() == want, (op, dt, len(v), v[:3])
                # the same row as one of many rows (rows kernel) and as one of many columns (column kernels)
                m = np.tile(v, (7, 1))
                tm = rt.asarray(m.reshape(-1), dev).reshape([7, len(v)])
                assert getattr(tm, op + "_axes")(1).to_numpy().tolist() == [want] * 7
                tc = rt.asarray(np.ascontiguousarray(m.T).reshape(-1), dev).reshape([len(v), 7])
                assert getattr(tc, op + "_axes")(0).to_numpy().tolist() == [want] * 7
    for dt in (np.int32, np.int64, np.uint32):
        info = np.iinfo(dt)
        for n in (4, 3000):
            for fillv in (info.min, info.max):
                v = np.full(n, fillv, dtype=dt)
                t = rt.asarray(v, dev)
                assert t.argmax_all() == 0 and t.argmin_all() == 0
                v[n // 2] = 7 if fillv == info.min else 3
                t = rt.asarray(v, dev)
                assert t.argmax_all() == (n // 2 if fillv == info.min else 0)
                assert t.argmin_all() == (n // 2 if fillv == info.max else 0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("op", ["var", "std", "l2_norm"])
def test_float_moments_random_views(dev, op, dtype):
    rng = np.random.default_rng(seed_of(op, np.dtype(dtype).name))
    tol = 1e-12 if dtype == np.float64 else 2e-5
    done = 0
    while done < 20:
        la, na = random_view(rng, max_ndim=4, max_extent=9)
        if la.ndim == 0 or la.size == 0:
            continue
        a = rand_data(rng, na, dtype)
        k = int(rng.integers(1, la.ndim + 1))
        axes = [int(x) for x in rng.permutation(la.ndim)[:k]]
        raw, lo = dev.reduce_axes(op, upload(dev, a), P(la), axes)
        ref, lo_ref = oracle.reduce_ext(op, a, la, axes)
        assert same(lo, lo_ref)
        got, want = view_np(dev.to_cpu_vec(raw), O(lo)).astype(np.float64), view_np(ref, lo_ref).astype(np.float64)
        v = view_np(a, la).astype(np.float64)
        msq = (v * v).mean(axis=tuple(axes))
        scale = msq if op == "var" else np.sqrt(np.maximum(msq, 1e-300)) * (1 if op == "std" else np.sqrt(v.size))
        if op == "std":  # d sqrt(x) blows up near 0: compare variances instead
            assert (np.abs(got * got - want * want) <= 4 * tol * msq + 1e-300).all()
        else:
            assert (np.abs(got - want) <= 4 * tol * np.maximum(scale, 1e-300)).all(), (op, la, axes)
        s_got = dev.reduce_all(op, upload(dev, a), P(la))
        s_want = oracle.reduce_ext(op, a, la, None)
        assert abs(float(s_got) - float(s_want)) <= 8 * tol * max(float((v * v).mean()), 1e-300) * (v.size if op == "l2_norm" else 1) + 1e-300 \
            or abs(float(s_got) ** 2 - float(s_want) ** 2) <= 8 * tol * float((v * v).sum())
        done += 1


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int64, np.uint32, np.uint64])
@pytest.mark.parametrize("op", ["argmin", "argmax", "count_nonzero"])
def test_index_and_count_random_views(dev, op, dtype):
    rng = np.random.default_rng(seed_of(op, np.dtype(dtype).name))
    done = 0
    while done < 25:
        la, na = random_view(rng, max_ndim=4, max_extent=8)
        if la.ndim == 0 or la.size == 0:
            continue
        a = rand_data(rng, na, dtype)
        a = (a.astype(np.float64) // 3).astype(dtype) if np.dtype(dtype).kind == "f" else (a // 200).astype(dtype)  # many ties / zeros
        k = int(rng.integers(1, la.ndim + 1))
        axes = [int(x) for x in rng.permutation(la.ndim)[:k]]  # order matters for arg*: flattened in the order given
        raw, lo = dev.reduce_axes(op, upload(dev, a), P(la), axes)
        ref, lo_ref = oracle.reduce_ext(op, a, la, axes)
        assert same(lo, lo_ref)
        assert raw.dtype == np.uint64
        assert np.array_equal(view_np(dev.to_cpu_vec(raw), O(lo)), view_np(ref, lo_ref)), (op, la, axes)
        assert dev.reduce_all(op, upload(dev, a), P(la)) == oracle.reduce_ext(op, a, la, None)
        done += 1


@pytest.mark.parametrize("op", ["all", "any", "count_nonzero"])
def test_bool_reductions_random_views(dev, op):
    rng = np.random.default_rng(seed_of("bool", op))
    done = 0
    while done < 25:
        la, na = random_view(rng, max_ndim=4, max_extent=8)
        if la.ndim == 0:
            continue
        a = (rng.random(na) < (0.9 if op == "all" else 0.1)).astype(np.bool_)
        k = int(rng.integers(1, la.ndim + 1))
        axes = [int(x) for x in rng.permutation(la.ndim)[:k]]
        raw, lo = dev.reduce_axes(op, upload(dev, a), P(la), axes)
        ref, lo_ref = oracle.reduce_ext(op, a, la, axes)
        assert same(lo, lo_ref)
        assert np.array_equal(view_np(dev.to_cpu_vec(raw), O(lo)), view_np(ref, lo_ref)), (op, la, axes)
        assert dev.reduce_all(op, upload(dev, a), P(la)) == oracle.reduce_ext(op, a, la, None)
        done += 1


LARGE = [((2048, 4096), [1]), ((4096, 2048), [0]), ((64, 96, 128), [0, 2]), ((1 << 21,), [0])]


@pytest.mark.parametrize("shape,axes", LARGE)
def test_large_shapes_hit_vector_and_split_paths(dev, shape, axes):
    rng = np.random.default_rng(seed_of("large", shape))
    a = rng.standard_normal(int(np.prod(shape)))
    t = rt.asarray(a, dev).reshape(list(shape))
    an = a.reshape(shape)
    ax = tuple(axes)
    assert np.allclose(t.var_axes(axes).to_numpy(), an.var(axis=ax), rtol=1e-9, atol=1e-12)
    assert np.allclose(t.l2_norm_axes(axes).to_numpy(), np.sqrt((an * an).sum(axis=ax)), rtol=1e-12)
    moved = np.moveaxis(an, axes, list(range(an.ndim - len(axes), an.ndim)))
    flat = moved.reshape(moved.shape[:an.ndim - len(axes)] + (-1,))
    assert np.array_equal(t.argmax_axes(axes).to_numpy(), flat.argmax(-1).astype(np.uint64))
    assert np.array_equal(t.argmin_axes(axes).to_numpy(), flat.argmin(-1).astype(np.uint64))
    assert np.array_equal((t.binary("gt", 0.5)).count_nonzero_axes(axes).to_numpy(), (an > 0.5).sum(axis=ax).astype(np.uint64))
    assert t.argmax_all() == int(an.argmax()) and t.argmin_all() == int(an.argmin())


def test_unraveled_arg_all(dev, dev_col):
    """unraveled_argmin_all / unraveled_argmax_all: multi-index of the first extreme element in row-major order"""
    rng = np.random.default_rng(seed_of("unravel"))
    v = rng.integers(-50, 50, (5, 6, 7)).astype(np.int64)
    for d in (dev, dev_col):
        t = rt.asarray(v, d)
        assert t.unraveled_argmin_all() == tuple(int(i) for i in np.unravel_index(np.argmin(v), v.shape))
        assert t.unraveled_argmax_all() == tuple(int(i) for i in np.unravel_index(np.argmax(v), v.shape))
        tt = t.transpose([2, 0, 1])[::-1]
        vv = v.transpose(2, 0, 1)[::-1]
        assert tt.unraveled_argmax_all() == tuple(int(i) for i in np.unravel_index(np.argmax(vv), vv.shape))


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int16])
@pytest.mark.parametrize("inner", [1, 2, 3, 4, 5, 8, 16])
def test_many_outputs_short_runs(dev, dev_col, dtype, inner):
    """(N, 3)-style reductions: a thread owns several outputs (reduce_tiny_kernel) -- all ops, both orders, views"""
    rng = np.random.default_rng(seed_of("tiny", inner, np.dtype(dtype).name))
    n = 5003
    v = (rng.standard_normal((n, inner)) * 50).astype(dtype)
    for d in (dev, dev_col):
        t = rt.asarray(v, d)
        with np.errstate(over="ignore"):
            got = t.sum_axes(-1).to_numpy()
            want = v.sum(-1, dtype=dtype)
        if np.dtype(dtype).kind == "f":
            assert np.all(np.abs(got - want) <= (1e-12 if dtype == np.float64 else 1e-5) * np.abs(v).sum(-1) + 1e-300)
        else:
            assert np.array_equal(got, want)
        assert np.array_equal(t.max_axes(-1).to_numpy(), v.max(-1))
        assert np.array_equal(t.min_axes(1).to_numpy(), v.min(1))
        assert np.array_equal(t.argmax_axes(-1).to_numpy(), np.argmax(v, -1).astype(np.uint64))
        assert np.array_equal(t.argmin_axes(-1).to_numpy(), np.argmin(v, -1).astype(np.uint64))
        assert np.array_equal(t.count_nonzero_axes(-1).to_numpy(), np.count_nonzero(v, axis=-1).astype(np.uint64))
        if np.dtype(dtype).kind == "f":
            tol = 1e-12 if dtype == np.float64 else 1e-5
            assert np.all(np.abs(t.mean_axes(-1).to_numpy() - v.mean(-1, dtype=np.float64)) <= tol * np.abs(v).mean(-1) + 1e-300)
            # the reference's one-pass formula q/n - (s/n)^2 in the element type (auto_impl/reduction.rs:207-317)
            q = (v * v).sum(-1, dtype=dtype) / dtype(inner)
            m1 = v.sum(-1, dtype=dtype) / dtype(inner)
            tol = 1e-12 if dtype == np.float64 else 1e-5
            assert np.all(np.abs(t.var_axes(-1).to_numpy() - (q - m1 * m1)) <= 4 * tol * q + 1e-300)
            vv = (v.astype(np.float64) ** 2).sum(-1)
            assert np.all(np.abs(rt.vecdot(t, t).to_numpy() - vv) <= 4 * tol * vv + 1e-300)
        # the same through a transposed and a sliced view (short run not contiguous / offset rows)
        tt = rt.asarray(np.ascontiguousarray(v.T), d).reverse_axes()
        assert np.array_equal(tt.max_axes(-1).to_numpy(), v.max(-1))
        ts = t[3:-2]
        assert np.array_equal(ts.min_axes(-1).to_numpy(), v[3:-2].min(-1))


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int16, np.uint8])
def test_misaligned_contiguous_rows(dev, dtype):
    """rows that are contiguous but not pack-aligned from their first element (a[:, 1:-1], odd pitch / length):
    reduce_rows_peel_kernel -- head / aligned body / tail per row, every op"""
    rng = np.random.default_rng(seed_of("peel", np.dtype(dtype).name))
    for rows, cols, sl in ((700, 1003, slice(0, None)), (640, 1024, slice(1, -1)), (900, 777, slice(3, None)),
                           (1200, 4096, slice(5, 4001))):
        v = (rng.standard_normal((rows, cols)) * 40).astype(dtype)
        t = rt.asarray(v, dev)[:, sl]
        w = v[:, sl]
        with np.errstate(over="ignore"):
            got, want = t.sum_axes(-1).to_numpy(), w.sum(-1, dtype=dtype)
        if np.dtype(dtype).kind == "f":
            tol = 1e-12 if dtype == np.float64 else 1e-5
            assert np.all(np.abs(got.astype(np.float64) - w.astype(np.float64).sum(-1)) <= tol * np.abs(w).astype(np.float64).sum(-1) + 1e-300)
            assert np.all(np.abs(t.mean_axes(-1).to_numpy() - w.mean(-1, dtype=np.float64)) <= tol * np.abs(w).mean(-1) + 1e-300)
        else:
            assert np.array_equal(got, want)
        assert np.array_equal(t.max_axes(-1).to_numpy(), w.max(-1))
        assert np.array_equal(t.min_axes(-1).to_numpy(), w.min(-1))
        assert np.array_equal(t.argmax_axes(-1).to_numpy(), np.argmax(w, -1).astype(np.uint64))
        assert np.array_equal(t.argmin_axes(-1).to_numpy(), np.argmin(w, -1).astype(np.uint64))
        assert np.array_equal(t.count_nonzero_axes(-1).to_numpy(), np.count_nonzero(w, axis=-1).astype(np.uint64))
    # NaN rules survive the peel: NaN in the head / tail is skipped by max, accepted by argmax only at index 0
    f = rng.standard_normal((600, 515))
    f[:, 1] = np.nan
    f[5, 514] = np.nan
    tf = rt.asarray(f, dev)[:, 1:]
    assert np.array_equal(tf.max_axes(-1).to_numpy(), np.nanmax(f[:, 1:], axis=-1))
    assert np.all(tf.argmax_axes(-1).to_numpy() == 0)   # element 0 of every row is NaN: it sticks (reference rule)
    tg = rt.asarray(f, dev)[:, 2:]
    assert np.array_equal(tg.argmax_axes(-1).to_numpy(), np.nanargmax(f[:, 2:], axis=-1).astype(np.uint64))
