"""Host-side decisions of the Python mirror that never reach the GPU (no device needed): Layout::c_prefer / f_prefer
(known answers of rstsr-common/src/layout/layoutbase.rs:754-806) and Layout::diagonal (layoutbase.rs:322-384)."""
import numpy as np

import rstsr_b200 as rt
from rstsr_b200.tensor import _diagonal_layout, _kept_layout


def view(shape, stride, offset=0):
    return rt.Tensor(None, rt.Layout(tuple(shape), tuple(stride), offset))


def test_is_f_prefer_kats():
    shape = [3, 5, 7]
    assert view(shape, [1, 10, 100])._prefer(True)
    assert view(shape, [1, 3, 15])._prefer(True)
    assert not view(shape, [1, 3, -15], 1000)._prefer(True)
    assert not view(shape, [1, 21, 3])._prefer(True)
    assert not view(shape, [35, 7, 1])._prefer(True)
    assert not view(shape, [2, 6, 30])._prefer(True)
    assert view([], [])._prefer(True)
    assert view([2, 0, 4], [1, 10, 100])._prefer(True)
    assert view([2, 1, 4], [1, 1, 2])._prefer(True)


def test_is_c_prefer_kats():
    shape = [3, 5, 7]
    assert view(shape, [100, 10, 1])._prefer(False)
    assert view(shape, [35, 7, 1])._prefer(False)
    assert not view(shape, [-35, 7, 1], 1000)._prefer(False)
    assert not view(shape, [7, 21, 1])._prefer(False)
    assert not view(shape, [1, 3, 15])._prefer(False)
    assert not view(shape, [70, 14, 2])._prefer(False)
    assert view([], [])._prefer(False)
    assert view([2, 0, 4], [1, 10, 100])._prefer(False)
    assert view([2, 1, 4], [4, 1, 1])._prefer(False)


def test_diagonal_layout_matches_numpy_inside_the_reference_range():
    rng = np.random.default_rng(5)
    for _ in range(200):
        d1, d2 = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        base = np.arange(4 * d1 * d2 + 10)
        t1, t2 = (d2, 1) if rng.random() < 0.5 else (1, d1)
        off0 = int(rng.integers(0, 5))
        a = np.lib.stride_tricks.as_strided(base[off0:], shape=(d1, d2), strides=(t1 * 8, t2 * 8))
        for k in range(-d1 - 1, d2 + 2):
            l = _diagonal_layout(rt.Layout((d1, d2), (t1, t2), off0), k)
            got = [int(base[l.offset + i * l.stride[0]]) for i in range(l.shape[0])]
            if -d1 < k < d1:  # Layout::diagonal accepts offsets in (-rows, rows) only
                want = np.diagonal(a, k).tolist()
                if k >= d2:
                    want = []
                assert got == want, (d1, d2, k)
            else:
                assert l.shape == (0,)


def test_kept_layout():
    l = rt.Layout((2, 3, 4, 5), (60, 20, 5, 1), 7)
    k = _kept_layout(l, [1, 3])
    assert (k.shape, k.stride, k.offset) == ((2, 4), (60, 5), 7)
