"""GPU parity for the tensor-level compositions of `assign`: concat / stack / hstack / vstack / unstack /
diag (rstsr-core/src/tensor/creation_from_tensor.rs) against NumPy's functions of the same names -- bit-exact."""
import numpy as np
import pytest

import rstsr_b200 as rt
from oracle import layout as L

from helpers import P, rand_data, random_view, seed_of, upload, view_np

pytestmark = pytest.mark.gpu


def views_like(rng, dev, shape, axis, sizes, dtype):
    """tensors whose shapes equal `shape` except along `axis`, each seen through a random permuted / strided view"""
    out_t, out_np = [], []
    for n in sizes:
        shp = list(shape)
        shp[axis] = n
        nd = len(shp)
        perm = [int(p) for p in rng.permutation(nd)]
        inv = [perm.index(i) for i in range(nd)]
        big = [shp[p] * 2 + 1 for p in perm]
        l = L.c_contig_layout(big)
        for ax in range(nd):
            l = l.narrow(ax, slice(1, 1 + 2 * shp[perm[ax]], 2))
        l = l.transpose(inv)
        buf = rand_data(rng, int(np.prod(big)), dtype)
        out_t.append(rt.Tensor(upload(dev, buf), P(l)))
        out_np.append(view_np(buf, l))
    return out_t, out_np


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int16, np.uint8, np.bool_])
def test_concat_and_stack_match_numpy(dev, dev_col, dtype):
    rng = np.random.default_rng(seed_of("concat", np.dtype(dtype).name))
    for d in (dev, dev_col):
        for shape in ([5], [3, 4], [2, 3, 4], [2, 1, 3, 2]):
            for axis in range(len(shape)):
                ts, ns = views_like(rng, d, shape, axis, [shape[axis], 1, 0, 3], dtype)
                got = rt.concat(ts, axis)
                assert np.array_equal(got.to_numpy(), np.concatenate(ns, axis))
                assert np.array_equal(rt.concat(ts, axis - len(shape)).to_numpy(), np.concatenate(ns, axis))
                contig = got.layout.c_contig() if d is dev else got.layout.f_contig()
                assert contig  # empty_f((shape, &device)): contiguous in the device's default order
            for axis in range(-len(shape) - 1, len(shape) + 1):
                ts, ns = views_like(rng, d, shape, 0, [shape[0]] * 3, dtype)
                assert np.array_equal(rt.stack(ts, axis).to_numpy(), np.stack(ns, axis))


def test_hstack_vstack_unstack_diag(dev):
    rng = np.random.default_rng(seed_of("hv"))
    a, b = rng.standard_normal(5), rng.standard_normal(7)
    ta, tb = rt.asarray(a, dev), rt.asarray(b, dev)
    assert np.array_equal(rt.hstack([ta, tb]).to_numpy(), np.hstack([a, b]))
    assert np.array_equal(rt.vstack([ta, ta[::-1]]).to_numpy(), np.vstack([a, a[::-1]]))
    m, k = rng.standard_normal((3, 4)), rng.standard_normal((3, 2))
    tm, tk = rt.asarray(m, dev), rt.asarray(k, dev)
    assert np.array_equal(rt.hstack([tm, tk]).to_numpy(), np.hstack([m, k]))
    assert np.array_equal(rt.vstack([tm, tm.flip(0)]).to_numpy(), np.vstack([m, m[::-1]]))
    s = rt.asarray(np.array(2.5), dev)
    assert rt.hstack([s, s]).to_numpy().tolist() == [2.5, 2.5]
    assert rt.vstack([s, s]).shape == (2, 1)
    parts = rt.unstack(tm, 1)
    assert len(parts) == 4 and all(p.shape == (3,) for p in parts)
    assert all(np.array_equal(p.to_numpy(), m[:, i]) for i, p in enumerate(parts))
    assert parts[0].raw is tm.raw  # views, no copy
    def ref_diag(x, off):
        # Layout::diagonal (rstsr-common/src/layout/layoutbase.rs:352-363) accepts offsets in (-d1, d1) only -- d1 = rows
        # for BOTH signs -- so a wide matrix loses its outermost super-diagonals (NumPy keeps them): reference behaviour
        return np.diag(x, off) if -x.shape[0] < off < x.shape[0] else np.zeros(0)

    for off in (-3, -2, -1, 0, 1, 2, 3, 5):
        assert np.array_equal(rt.diag(tm, off).to_numpy(), ref_diag(m, off)), off
        assert np.array_equal(rt.diag(tm.reverse_axes(), off).to_numpy(), ref_diag(m.T, off)), off
    for off in (-2, 0, 3):
        assert np.array_equal(rt.diag(ta, off).to_numpy(), np.diag(a, off))
        assert np.array_equal(rt.diag(ta[::-2], off).to_numpy(), np.diag(a[::-2], off))


def test_concat_errors(dev, dev_col):
    a, b = rt.zeros([2, 3], dev), rt.zeros([2, 4], dev)
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.concat([], 0)
    assert e.value.kind == "InvalidValue"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.concat([a, b], 0)
    assert e.value.kind == "InvalidLayout" and "same shape except for the concatenation axis" in str(e.value)
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.concat([a, rt.zeros([2], dev)], 0)
    assert e.value.kind == "InvalidLayout"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.concat([a, a], 2)
    assert e.value.kind == "InvalidValue"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.concat([a, rt.zeros([2, 3], dev_col)], 0)
    assert e.value.kind == "DeviceMismatch"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.stack([a, b], 0)
    assert e.value.kind == "InvalidLayout"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.concat([rt.asarray(np.array(1.0), dev)], 0)
    assert e.value.kind == "InvalidLayout"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.diag(rt.zeros([2, 2, 2], dev))
    assert e.value.kind == "InvalidLayout"


def test_concat_large(dev):
    """three (2048, 4096) f64 slabs along each axis: each assign is a full-bandwidth strided copy"""
    rng = np.random.default_rng(seed_of("concatbig"))
    arrs = [rng.standard_normal((2048, 4096)) for _ in range(3)]
    ts = [rt.asarray(x, dev) for x in arrs]
    assert np.array_equal(rt.concat(ts, 0).to_numpy(), np.concatenate(arrs, 0))
    assert np.array_equal(rt.concat(ts, 1).to_numpy(), np.concatenate(arrs, 1))
    assert np.array_equal(rt.stack(ts, 2).to_numpy(), np.stack(arrs, 2))


def test_meshgrid_matches_numpy(dev, dev_col):
    """creation_from_tensor/test_meshgrid.rs: 'xy' / 'ij' indexing, copies and broadcast views, 1..4 inputs"""
    rng = np.random.default_rng(seed_of("mesh"))
    xs = [rng.standard_normal(n) for n in (5, 3, 4, 2)]
    for d in (dev, dev_col):
        ts = [rt.asarray(x, d) for x in xs]
        for k in (1, 2, 3, 4):
            for indexing in ("xy", "ij"):
                want = np.meshgrid(*xs[:k], indexing=indexing)
                got = rt.meshgrid(ts[:k], indexing)
                assert len(got) == k
                for g, w in zip(got, want):
                    assert g.shape == w.shape and np.array_equal(g.to_numpy(), w)
                    assert g.layout.c_contig() if d is dev else g.layout.f_contig()
                views = rt.meshgrid(ts[:k], indexing, copy=False)
                for g, w in zip(views, want):
                    assert np.array_equal(g.to_numpy(), w) and g.raw.ptr in [t.raw.ptr for t in ts]
        # strided / reversed inputs
        got = rt.meshgrid([ts[0][::-2], ts[1][1:]], "ij")
        want = np.meshgrid(xs[0][::-2], xs[1][1:], indexing="ij")
        assert all(np.array_equal(g.to_numpy(), w) for g, w in zip(got, want))
    assert rt.meshgrid([]) == []
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.meshgrid([rt.zeros([2, 2], dev)])
    assert e.value.kind == "InvalidLayout"
    with pytest.raises(rt.RstsrCudaError) as e:
        rt.meshgrid([rt.zeros([2], dev)], "yx")
    assert e.value.kind == "InvalidValue"
