"""The reference's device-agnostic conformance bodies (`rstsr-core/tests/core_func/**`, re-run per device through
`crates-device/*/tests/tests_core_row.rs`) with `DeviceType = DeviceCuda`: same test names, same literals.
`assert_equal` = shape equality + `rt::allclose` (rstsr-core/tests/test_utils/equality.rs:12-27) -- evaluated here on the
device as well, so these tests exercise allclose_all exactly the way the reference's own harness does."""
import numpy as np
import pytest

import rstsr_b200 as rt

pytestmark = pytest.mark.gpu


@pytest.fixture()
def device(dev):
    dev.set_default_order(rt.ROW_MAJOR)
    return dev


def T(nested, device, dtype=None):
    """rt::tensor_from_nested!"""
    return rt.asarray(np.array(nested, dtype=dtype), device)


def assert_equal(a, b):
    """test_utils/equality.rs: shapes equal, then allclose on the device (exact compare for bool / empty)."""
    assert tuple(a.shape) == tuple(b.shape), (a.shape, b.shape)
    if a.size == 0:
        return
    if a.dtype == np.bool_ or b.dtype == np.bool_:
        assert np.array_equal(a.to_numpy(), b.to_numpy())
        return
    if a.dtype != b.dtype:
        b = b.astype(a.dtype)
    assert rt.allclose(a, b)


def is_err(fn, *args):
    try:
        fn(*args)
    except rt.RstsrCudaError:
        return True
    return False


# ---- linalg/test_vecdot.rs ----
def test_vecdot_basic(device):
    arr1 = rt.arange(6, device).reshape([2, 3])
    arr2 = rt.arange(3, device).reshape([1, 3])
    expected = T([5, 14], device)
    assert_equal(rt.vecdot(arr1, arr2), expected)
    assert_equal(rt.vecdot(arr1.reverse_axes(), arr2.reverse_axes(), -2), expected)


def test_vecdot_broadcast(device):
    a = rt.arange(4, device).reshape([2, 1, 2])
    b = rt.arange(4, device).reshape([1, 2, 2])
    assert_equal(rt.vecdot(a, b), (a * b).sum_axes(-1))
    b2 = rt.arange(4, device).reshape([2, 2])
    assert_equal(rt.vecdot(a, b2), (a * b2).sum_axes(-1))


def test_vecdot_broadcast_fails(device):
    a = rt.arange(8, device).reshape([4, 2])
    b = rt.arange(4, device).reshape([4, 1])
    assert is_err(rt.vecdot, a, b)
    assert is_err(rt.vecdot, a, rt.asarray(np.array(7), device))
    a2 = rt.arange(2, device).reshape([2, 1, 1])
    b3 = rt.arange(3, device).reshape([3, 1, 1])
    assert is_err(rt.vecdot, a2, b3)


def test_vecdot_empty(device):
    a = rt.zeros([0], device, dtype=np.int32)
    assert rt.vecdot(a, a).to_scalar() == 0


def test_vecdot_non_contiguous(device):
    a = rt.arange(12, device).reshape([3, 4])
    b = rt.asarray(np.arange(8, 20), device).reshape([3, 4])
    result = rt.vecdot(a[:, ::2], b[:, ::2])
    assert rt.allclose(result, T([20, 132, 308], device))


# ---- creation_from_tensor/test_concat.rs ----
def test_returns_copy(device):
    a = rt.full([3, 3], 1.0, device)
    b = rt.concatenate([a])
    assert b.raw.ptr != a.raw.ptr


def test_exceptions(device):
    for ndim in (1, 2, 3):
        a = rt.full([1] * ndim, 1.0, device)
        rt.concatenate([a, a], 0)
        assert is_err(rt.concatenate, [a, a], ndim)
        assert is_err(rt.concatenate, [a, a], -(ndim + 1))
    assert is_err(rt.concatenate, [rt.asarray(np.array(0), device)], 0)
    assert is_err(rt.concatenate, [rt.zeros([1], device), rt.zeros([1, 1], device)], 0)
    a, b = rt.full([1, 2, 3], 1.0, device), rt.full([2, 2, 3], 1.0, device)
    rt.concatenate([a, b], 0)
    assert is_err(rt.concatenate, [a, b], 1)
    assert is_err(rt.concatenate, [a, b], 2)
    assert is_err(rt.concatenate, [], 0)


def test_concatenate(device):
    r4, r3 = T([0, 1, 2, 3], device), T([0, 1, 2], device)
    assert_equal(rt.concatenate([r4]), r4)
    expected = T([0, 1, 2, 3, 0, 1, 2], device)
    assert_equal(rt.concatenate([r4, r3]), expected)
    assert_equal(rt.concatenate([r4, r3], 0), expected)
    assert_equal(rt.concatenate([r4, r3], -1), expected)
    a23, a13 = T([[10, 11, 12], [13, 14, 15]], device), T([[0, 1, 2]], device)
    res = T([[10, 11, 12], [13, 14, 15], [0, 1, 2]], device)
    assert_equal(rt.concatenate([a23, a13]), res)
    assert_equal(rt.concatenate([a23, a13], 0), res)
    assert_equal(rt.concatenate([a23.reverse_axes(), a13.reverse_axes()], 1), res.reverse_axes())
    assert_equal(rt.concatenate([a23.reverse_axes(), a13.reverse_axes()], -1), res.reverse_axes())
    assert is_err(rt.concatenate, [a23.reverse_axes(), a13.reverse_axes()], 0)


# ---- creation_from_tensor/test_diag.rs ----
def test_vector(device):
    vals = T([0, 100, 200, 300, 400], device)
    v = np.array([0, 100, 200, 300, 400])
    assert_equal(rt.diag(vals), T(np.diag(v), device))
    assert_equal(rt.diag(vals, 2), T(np.diag(v, 2), device))
    assert_equal(rt.diag(vals, -2), T(np.diag(v, -2), device))


def test_matrix(device):
    m = np.array([[100 * (i + j) + 1 for j in range(5)] for i in range(5)])
    vals = T(m, device)
    assert_equal(rt.diag(vals), T([1, 201, 401, 601, 801], device))
    assert_equal(rt.diag(vals, 2), T([201, 401, 601], device))
    assert_equal(rt.diag(vals, -2), T([201, 401, 601], device))


def test_fortran_order(device):
    m = np.array([[100 * (i + j) + 1 for j in range(5)] for i in range(5)])
    vals_f = T(m, device).to_contig(rt.COL_MAJOR)
    assert vals_f.layout.f_contig()
    assert_equal(rt.diag(vals_f), T([1, 201, 401, 601, 801], device))


def test_diag_bounds(device):
    a = T([[1, 2], [3, 4], [5, 6]], device)
    assert rt.diag(a, 2).shape == (0,)
    assert_equal(rt.diag(a, 1), T([2], device))
    assert_equal(rt.diag(a), T([1, 4], device))
    assert_equal(rt.diag(a, -1), T([3, 6], device))
    assert_equal(rt.diag(a, -2), T([5], device))
    assert rt.diag(a, -3).shape == (0,)


def test_failure(device):
    assert is_err(rt.diag, T([[[1]]], device))


# ---- creation_from_tensor/test_stack.rs, test_hstack.rs, test_vstack.rs, test_unstack.rs ----
def test_1d_input(device):
    a, b = T([1, 2, 3], device), T([4, 5, 6], device)
    r1 = T([[1, 2, 3], [4, 5, 6]], device)
    assert_equal(rt.stack([a, b]), r1)
    assert_equal(rt.stack([a, b], 1), r1.reverse_axes())


def test_shapes(device):
    arrays1 = [rt.asarray(np.random.default_rng(i).standard_normal(3), device) for i in range(10)]
    for axis, shape in ((0, (10, 3)), (1, (3, 10)), (-1, (3, 10)), (-2, (10, 3))):
        assert rt.stack(arrays1, axis).shape == shape
    assert is_err(rt.stack, arrays1, 2)
    assert is_err(rt.stack, arrays1, -3)
    arrays2 = [rt.asarray(np.random.default_rng(i).standard_normal(12), device).reshape([3, 4]) for i in range(10)]
    for axis, shape in ((0, (10, 3, 4)), (1, (3, 10, 4)), (2, (3, 4, 10)), (-1, (3, 4, 10)), (-2, (3, 10, 4)), (-3, (10, 3, 4))):
        assert rt.stack(arrays2, axis).shape == shape


def test_empty_arrays(device):
    e = rt.zeros([0], device, dtype=np.int64)
    assert rt.stack([e, e, e]).shape == (3, 0)
    assert rt.stack([e, e, e], 1).shape == (0, 3)


def test_edge_cases(device):
    assert is_err(rt.stack, [], 0)
    a, b = rt.arange(3, device), rt.arange(2, device)
    assert is_err(rt.stack, [a, b], 0)
    assert is_err(rt.stack, [a, b], 1)
    assert is_err(rt.stack, [rt.zeros([3, 3], device, dtype=np.int64), a], 1)


def test_0d_input(device):
    a, b, c = (rt.asarray(np.array(v, dtype=np.int32), device) for v in (1, 2, 3))
    assert_equal(rt.stack([a, b, c]), T([1, 2, 3], device, np.int32))
    assert_equal(rt.hstack([a, b]), T([1, 2], device, np.int32))
    assert_equal(rt.vstack([a, b]), T([[1], [2]], device, np.int32))


def test_hstack_vstack_arrays(device):
    assert is_err(rt.hstack, [])
    assert is_err(rt.vstack, [])
    a, b = T([1], device), T([2], device)
    assert_equal(rt.hstack([a, b]), T([1, 2], device))
    assert_equal(rt.vstack([a, b]), T([[1], [2]], device))
    a2, b2 = T([[1], [2]], device), T([[1], [2]], device)
    assert_equal(rt.hstack([a2, b2]), T([[1, 1], [2, 2]], device))
    assert_equal(rt.vstack([a2, b2]), T([[1], [2], [1], [2]], device))
    a3, b3 = T([1, 2], device), T([1, 2], device)
    assert_equal(rt.vstack([a3, b3]), T([[1, 2], [1, 2]], device))


def test_unstack(device):
    a = rt.arange(24, device).reshape([2, 3, 4])
    for axis in (0, -3):
        stacks = rt.unstack(a, axis)
        assert len(stacks) == 2
        assert_equal(stacks[0], a[0, :, :])
        assert_equal(stacks[1], a[1, :, :])
    for axis in (1, -2):
        stacks = rt.unstack(a, axis)
        assert len(stacks) == 3
        for k in range(3):
            assert_equal(stacks[k], a[:, k, :])
    for axis in (2, -1):
        stacks = rt.unstack(a, axis)
        assert len(stacks) == 4
        for k in range(4):
            assert_equal(stacks[k], a[:, :, k])
    assert is_err(rt.unstack, a, 3)
    assert is_err(rt.unstack, a, -4)
    assert is_err(rt.unstack, rt.asarray(np.array(0), device), 0)


# ---- reduction/test_*.rs ----
def test_argmax_combinations(device):
    assert T([1] * 8 + [0] * 7, device).argmax_all() == 0
    assert T([3, 3, 3, 3, 2, 2, 2, 2], device).argmax_all() == 0
    assert T([0, 1, 2, 3, 4, 5, 6, 7], device).argmax_all() == 7
    assert T([7, 6, 5, 4, 3, 2, 1, 0], device).argmax_all() == 0


def test_argmax_regression(device):
    a = rt.arange(4 * 5 * 6 * 7 * 8, device).reshape([4, 5, 6, 7, 8])
    v = np.arange(4 * 5 * 6 * 7 * 8).reshape(4, 5, 6, 7, 8)
    for i in range(5):
        assert np.array_equal(a.argmax_axes(i).to_numpy(), np.argmax(v, axis=i).astype(np.uint64))
        assert np.array_equal(a.argmin_axes(i).to_numpy(), np.argmin(v, axis=i).astype(np.uint64))


def test_argmax_axes_2d(device):
    b = T([[3, 6, 9], [4, 10, 5], [8, 3, 2]], device)
    assert b.argmax_axes(0).to_vec().tolist() == [2, 1, 0]
    assert b.argmax_axes(1).to_vec().tolist() == [2, 1, 0]


def test_count_nonzero_numeric(device):
    arr = T([[0, 1, 7, 0, 0], [3, 0, 0, 2, 19]], device)
    assert arr.count_nonzero_axes(1).to_vec().tolist() == [2, 3]
    assert arr.count_nonzero_all() == 5


def test_all_basic_and_nd(device):
    y1, y2, y3 = T([False, True, True, False], device), T([False] * 4, device), T([True] * 4, device)
    assert not y1.all_all() and y3.all_all() and not y2.all_all()
    assert (~y2).all_all()
    n = T([[False, False, True], [False, True, True], [True, True, True]], device)
    assert not n.all_all()
    assert n.all_axes(0).to_vec().tolist() == [False, False, True]
    assert n.all_axes(1).to_vec().tolist() == [False, False, True]
    assert n.any_all() and n.any_axes(0).to_vec().tolist() == [True, True, True]


def test_prod_numeric_and_basic(device):
    arr = T([[1, 2, 3, 4], [5, 6, 7, 9], [10, 3, 4, 5]], device)
    assert_equal(arr.prod_axes(-1), T([24, 1890, 600], device))
    assert_equal(arr.prod_axes(0), T([50, 36, 84, 180], device))
    assert T([1, 2, 10, 11, 6, 5, 4], device).prod_all() == 26400
    assert T([[1, 2, 3], [4, 5, 6]], device).prod_all() == 720


def test_std_var_numeric(device):
    a = T([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], device)
    assert abs(a.std_all() - 1.707825127659933) < 1e-14
    assert_equal(a.std_axes(0), T([1.5, 1.5, 1.5], device))
    assert_equal(a.std_axes(1), T([0.816496580927726, 0.816496580927726], device))
    assert abs(a.var_all() - 2.9166666666666665) < 1e-14
    assert_equal(a.var_axes(0), T([2.25, 2.25, 2.25], device))
    assert_equal(a.var_axes(1), T([0.6666666666666666, 0.6666666666666666], device))


def test_sum_numeric(device):
    m = T([[1, 2, 3], [4, 5, 6], [7, 8, 9]], device)
    assert_equal(m.sum_axes(1), T([6, 15, 24], device))
    assert m.sum_all() == 45
    a = rt.arange(24, device).reshape([2, 3, 4])
    assert_equal(a.sum_axes([-2, -1]), T([66, 210], device))
    assert T([[1, 2, 3], [4, 5, 6]], device).sum_all() == 21


# ---- indexing/test_indexing.rs ----
def test_indexing(device):
    a = rt.arange(10, device)
    assert a[-1].shape == () and a[-1].to_scalar() == 9
    b = rt.arange(24, device).reshape([2, 3, 4])
    assert b[..., 0].shape == (2, 3)
    assert_equal(b[0, ..., 1], T([1, 5, 9], device))
    m = rt.arange(12, device).reshape([3, 4])
    assert_equal(m[1], T([4, 5, 6, 7], device))
    assert_equal(m[-1], T([8, 9, 10, 11], device))
    assert m[1, 2].to_scalar() == 6
    assert_equal(m[:, 1], T([1, 5, 9], device))
    assert_equal(m[:, 1:3], T([[1, 2], [5, 6], [9, 10]], device))
    v = rt.arange(3, device)
    assert v[None].shape == (1, 3) and v[:, None].shape == (3, 1)
    assert_equal(a.index_select(0, [2, 4, 8]), T([2, 4, 8], device))
