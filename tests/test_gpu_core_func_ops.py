"""More of the reference's conformance bodies with `DeviceType = DeviceCuda` (see test_gpu_core_func.py):
manipulation/test_to_contig.rs, manipulation/test_to_layout.rs, operators/test_arithmetic.rs,
operators/test_comparison.rs, math/test_unary_math.rs -- same test names (prefixed by their file), same literals."""
import math

import numpy as np
import pytest

import rstsr_b200 as rt

from test_gpu_core_func import T, assert_equal, is_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def device(dev):
    dev.set_default_order(rt.ROW_MAJOR)
    return dev


def is_view_of(result, a):
    return (not result.owned) and result.raw.ptr == a.raw.ptr


# ---- manipulation/test_to_contig.rs ----
def test_to_contig_already_c_contig(device):
    a = rt.arange(24, device).reshape([2, 3, 4])
    assert a.layout.c_contig()
    assert is_view_of(a.to_contig(rt.ROW_MAJOR), a)
    r = a.to_contig(rt.COL_MAJOR)
    assert r.owned and r.layout.f_contig()


def test_to_contig_already_f_contig(dev_col):
    a = rt.arange(24, dev_col).reshape([2, 3, 4])
    assert a.layout.f_contig()
    assert is_view_of(a.to_contig(rt.COL_MAJOR), a)
    r = a.to_contig(rt.ROW_MAJOR)
    assert r.owned and r.layout.c_contig()


def test_to_contig_transposed_tensor(device):
    t = rt.arange(12, device).reshape([3, 4]).reverse_axes()
    assert not t.layout.c_contig() and t.layout.f_contig()
    rc = t.to_contig(rt.ROW_MAJOR)
    assert rc.owned and rc.layout.c_contig() and rc.shape == (4, 3)
    rf = t.to_contig(rt.COL_MAJOR)
    assert not rf.owned and rf.layout.f_contig() and rf.shape == (4, 3)
    assert rt.allclose(rc, T([[0, 4, 8], [1, 5, 9], [2, 6, 10], [3, 7, 11]], device))


def test_to_contig_sliced_tensor(device):
    s = rt.arange(24, device).reshape([4, 6])[::2, ::2]
    assert s.shape == (2, 3) and s.stride == (12, 2) and not s.layout.c_contig()
    c = s.to_contig(rt.ROW_MAJOR)
    assert c.owned and c.layout.c_contig() and c.stride == (3, 1)
    assert rt.allclose(c, T([[0, 2, 4], [12, 14, 16]], device))


def test_to_contig_f_order_sliced_and_preserves_values(device):
    s = rt.arange(24, device).reshape([4, 6])[::2, ::2]
    cc, cf = s.to_contig(rt.ROW_MAJOR), s.to_contig(rt.COL_MAJOR)
    assert cc.layout.c_contig() and cf.layout.f_contig()
    assert np.array_equal(cc.to_numpy(), cf.to_numpy())
    data = [1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5, 8.5, 9.5, 10.5, 11.5, 12.5]
    sl = rt.asarray(np.array(data), device).reshape([3, 4])[1:, ::2]
    assert np.array_equal(sl.to_contig(rt.ROW_MAJOR).to_numpy(), sl.to_numpy())
    assert np.array_equal(sl.to_contig(rt.COL_MAJOR).to_numpy(), sl.to_numpy())


def test_to_contig_scalar_and_1d_tensor(device):
    a = rt.asarray(np.array(42, dtype=np.int32), device)
    for order in (rt.ROW_MAJOR, rt.COL_MAJOR):
        assert not a.to_contig(order).owned
    v = rt.arange(10, device)
    assert v.layout.c_contig() and v.layout.f_contig() and not v.to_contig(rt.ROW_MAJOR).owned
    st = v[::2]
    assert not st.layout.c_contig() and not st.layout.f_contig()
    c = st.to_contig(rt.ROW_MAJOR)
    assert c.owned and c.stride == (1,)


def test_to_contig_prefer(device, dev_col):
    a = rt.arange(24, device).reshape([2, 3, 4])
    assert is_view_of(a.to_prefer(rt.ROW_MAJOR), a)                       # test_prefer_c_on_c_contig
    r = a.to_prefer(rt.COL_MAJOR)                                          # test_prefer_f_on_c_contig
    assert r.owned and r.layout.f_contig()
    f = rt.arange(24, dev_col).reshape([2, 3, 4])
    assert f.layout.f_contig() and not f.to_prefer(rt.COL_MAJOR).owned     # test_prefer_f_on_f_contig
    r = f.to_prefer(rt.ROW_MAJOR)                                          # test_prefer_c_on_f_contig
    assert r.owned and r.layout.c_contig()
    s = rt.arange(24, device).reshape([4, 6])[::2, ::2]                    # test_prefer_on_non_contig
    assert s.to_prefer(rt.ROW_MAJOR).owned and s.to_prefer(rt.COL_MAJOR).owned
    assert rt.allclose(a.to_prefer(rt.ROW_MAJOR), a.to_contig(rt.ROW_MAJOR))  # test_prefer_vs_contig


def test_to_contig_edge_shapes(device):
    e = rt.zeros([0, 5], device).to_contig(rt.ROW_MAJOR)                   # test_empty_tensor
    assert e.shape == (0, 5) and e.layout.c_contig()
    one = rt.asarray(np.array([42.0]), device).reshape([1, 1]).to_contig(rt.ROW_MAJOR)  # test_single_element
    assert one.shape == (1, 1) and one.layout.c_contig()
    s = rt.arange(120, device).reshape([2, 3, 4, 5])[:, ::2, :, ::2]       # test_high_dim
    cc, cf = s.to_contig(rt.ROW_MAJOR), s.to_contig(rt.COL_MAJOR)
    assert cc.layout.c_contig() and cf.layout.f_contig() and cc.shape == s.shape == cf.shape
    assert np.array_equal(cc.to_numpy(), np.arange(120).reshape(2, 3, 4, 5)[:, ::2, :, ::2])
    fl = rt.arange(12, device).reshape([3, 4]).flip(0).to_contig(rt.ROW_MAJOR)  # test_reverse_stride
    assert fl.layout.c_contig() and fl.stride[0] > 0
    assert rt.allclose(fl, T([[8, 9, 10, 11], [4, 5, 6, 7], [0, 1, 2, 3]], device))


# ---- manipulation/test_to_layout.rs ----
def test_to_layout_bodies(device):
    a = rt.arange(12, device).reshape([3, 4])
    assert is_view_of(a.to_layout(a.layout), a)                            # test_same_layout_no_copy
    lf = rt.Layout.contig([3, 4], rt.COL_MAJOR)
    r = a.to_layout(lf)                                                    # test_different_layout_copies
    assert r.owned and r.layout.f_contig()
    flat = a.to_layout(rt.Layout.contig([12], rt.ROW_MAJOR))              # test_dimensionality_change
    assert flat.shape == (12,) and flat.layout.c_contig()
    assert flat.to_numpy().tolist() == a.to_numpy().reshape(-1).tolist()
    s = a[:, ::2]                                                          # test_strided_tensor
    rs = s.to_layout(rt.Layout.contig([3, 2], rt.ROW_MAJOR))
    assert rs.owned and rs.layout.c_contig() and rs.shape == (3, 2)
    assert rt.allclose(rs, T([[0, 2], [4, 6], [8, 10]], device))
    assert is_err(a.to_layout, rt.Layout.contig([3, 3], rt.ROW_MAJOR))     # test_size_mismatch_error


# ---- operators/test_arithmetic.rs, test_comparison.rs ----
def test_arithmetic_bodies(device):
    a, b = T([[1, 2], [3, 4]], device), T([[10, 20], [30, 40]], device)
    assert_equal(a + b, T([[11, 22], [33, 44]], device))
    assert_equal(T([[1, 2, 3], [4, 5, 6]], device) + T([10, 20, 30], device), T([[11, 22, 33], [14, 25, 36]], device))
    assert_equal(b - a, T([[9, 18], [27, 36]], device))
    assert_equal(a * T([10, 100], device), T([[10, 200], [30, 400]], device))
    assert_equal(T([[10.0, 20.0], [30.0, 40.0]], device) / T([[2.0, 4.0], [5.0, 8.0]], device), T([[5.0, 5.0], [6.0, 5.0]], device))
    assert_equal(T([[10, 21], [33, 44]], device).binary("rem", T([[3, 4], [5, 7]], device)), T([[1, 1], [3, 2]], device))


def test_comparison_bodies(device):
    a, b = T([1, 2, 3, 4], device), T([2, 2, 2, 2], device)
    assert a.binary("gt", b).to_vec().tolist() == [False, False, True, True]
    assert a.binary("ge", b).to_vec().tolist() == [False, True, True, True]
    assert a.binary("lt", b).to_vec().tolist() == [True, False, False, False]
    assert a.binary("le", b).to_vec().tolist() == [True, True, False, False]
    assert a.binary("eq", b).to_vec().tolist() == [False, True, False, False]
    assert a.binary("ne", b).to_vec().tolist() == [True, False, True, True]
    x, y = T([1, 5, 3], device), T([4, 2, 6], device)
    assert_equal(x.binary("maximum", y), T([4, 5, 6], device))
    assert_equal(x.binary("minimum", y), T([1, 2, 3], device))


# ---- math/test_unary_math.rs ----
def test_abs_sqrt_sign_rounding(device):
    assert_equal(abs(T([-1, 2, -3], device)), T([1, 2, 3], device))
    assert_equal(T([0.0, 1.0, 4.0, 9.0], device).unary("sqrt"), T([0.0, 1.0, 2.0, 3.0], device))
    assert_equal(T([-2.0, 0.0, 3.0], device).unary("sign"), T([-1.0, 0.0, 1.0], device))
    assert T([-2, 0, 3], device).unary("sign").to_vec().tolist() == [-1, 0, 1]
    assert T([0, 5], device, np.uint8).unary("sign").to_vec().tolist() == [0, 1]
    f = T([-1.5, 0.5, 2.4], device)
    assert_equal(f.unary("floor"), T([-2.0, 0.0, 2.0], device))
    assert_equal(f.unary("ceil"), T([-1.0, 1.0, 3.0], device))
    assert_equal(f.unary("trunc"), T([-1.0, 0.0, 2.0], device))


def test_sign_special_values(device):
    s = T([np.nan, np.inf, -np.inf, -0.0, 0.0], device).unary("sign").to_vec()
    assert np.isnan(s[0]) and s[1] == 1.0 and s[2] == -1.0
    assert s[3] == 0.0 and not np.signbit(s[3])  # both zeros map to +0.0 (NumPy convention, ext_num.rs:166-178)
    assert s[4] == 0.0


def test_exp_log_trig(device):
    x = T([0.0, 1.0, 2.0], device)
    ex = x.unary("exp")
    assert_equal(ex, T([1.0, math.e, math.e * math.e], device))
    assert_equal(ex.unary("log"), x)
    ang = T([0.0, math.pi / 2.0, math.pi], device)
    assert_equal(ang.unary("sin"), T([0.0, 1.0, 0.0], device))
    assert_equal(ang.unary("cos"), T([1.0, 0.0, -1.0], device))
    th1 = math.tanh(1.0)
    assert_equal(T([-1.0, 0.0, 1.0], device).unary("tanh"), T([-th1, 0.0, th1], device))


def test_is_nan_is_finite_is_inf(device):
    a = T([0.0, np.nan, np.inf, -np.inf], device)
    assert a.unary("isnan").to_vec().tolist() == [False, True, False, False]
    assert a.unary("isfinite").to_vec().tolist() == [True, False, False, False]
    assert a.unary("isinf").to_vec().tolist() == [False, False, True, True]


# ---- creation/test_eye.rs, test_ones.rs, test_zeros.rs, test_full.rs ----
def test_eye_basic_and_rect_offset(device, dev_col):
    expected = T([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], device, np.int32)
    assert_equal(rt.eye(4, device, dtype=np.int32), expected)
    assert_equal(rt.eye(4, device, dtype=np.float32), expected.astype(np.float32))
    assert_equal(rt.eye(3, device, 4, 1, dtype=np.int32), T([[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], device, np.int32))
    for n, m, k in ((3, 5, 0), (5, 3, -1), (4, 4, 2), (4, 4, -3), (2, 6, 4)):
        # Layout::diagonal only accepts offsets in (-rows, rows) (layoutbase.rs:352-363): beyond that the reference's
        # eye stays all-zero even where NumPy's would not (wide matrices, k >= rows)
        want = np.eye(n, m, k) if -n < k < n else np.zeros((n, m))
        assert np.array_equal(rt.eye(n, device, m, k).to_numpy(), want)
    # ColMajor builds [n_cols, n_rows].f() (tensor/creation.rs:418-421)
    e = rt.eye(3, dev_col, 4, 0)
    assert e.shape == (4, 3) and e.layout.f_contig() and np.array_equal(e.to_numpy(), np.eye(4, 3))


def test_ones_zeros_full(device):
    assert np.array_equal(rt.ones([2, 3], device, dtype=np.int32).to_numpy(), np.ones((2, 3), np.int32))
    assert np.array_equal(rt.zeros([2, 3], device).to_numpy(), np.zeros((2, 3)))
    assert np.array_equal(rt.full([2, 2], 7.5, device).to_numpy(), np.full((2, 2), 7.5))
    assert rt.ones([0, 3], device).shape == (0, 3)
