"""GPU parity for index-driven data movement: index_select / take, pack_tri, unpack_tri through the
C ABI vs the oracle -- bit-exact (pure movement; the antisymmetric unpack only flips a sign).
KATs: rstsr-core/src/tensor/operators/op_tri.rs:187-228 (test_pack_tri, both default orders)."""
import numpy as np
import pytest

import oracle
import rstsr_b200 as rt
from oracle import layout as L

from helpers import O, P, rand_data, random_view, same, seed_of, upload, view_np

pytestmark = pytest.mark.gpu

ALL_DTYPES = [np.float64, np.float32, np.int64, np.int32, np.int16, np.int8, np.uint8, np.bool_]


# ---------------- index_select ----------------
def check_select(dev, order, a, la, axis, indices):
    want, lw = oracle.tensor_index_select(a, la, axis, list(indices), order)
    t = rt.Tensor(upload(dev, a), P(la))
    got = t.index_select(axis, list(indices))
    assert same(got.layout, lw)
    assert np.array_equal(got.to_numpy(), view_np(want, lw))
    return got


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_index_select_each_axis(dev, dev_col, dtype):
    rng = np.random.default_rng(seed_of("sel", np.dtype(dtype).name))
    for shape in ([17], [6, 40], [33, 5], [4, 6, 8], [3, 1, 5, 7], [256, 512]):
        a = rand_data(rng, int(np.prod(shape)), dtype)
        for d, order, mk in ((dev, "row", L.c_contig_layout), (dev_col, "col", L.f_contig_layout)):
            la = mk(shape)
            for axis in range(len(shape)):
                n = shape[axis]
                for idx in ([0], list(range(n)), list(range(n))[::-1], list(rng.integers(-n, n, 2 * n + 3)), []):
                    check_select(d, order, a, la, axis, [int(i) for i in idx])
                check_select(d, order, a, la, axis - len(shape), [n - 1, 0, 0])


@pytest.mark.parametrize("seed", range(30))
def test_index_select_random_views(dev, dev_col, seed):
    rng = np.random.default_rng(seed_of("selv", seed))
    dtype = ALL_DTYPES[seed % len(ALL_DTYPES)]
    la, na = random_view(rng, max_ndim=4, max_extent=9, allow_broadcast=True)
    if la.ndim == 0:
        return
    a = rand_data(rng, na, dtype)
    axis = int(rng.integers(0, la.ndim))
    n = la.shape[axis]
    idx = [int(i) for i in rng.integers(-n, n, int(rng.integers(0, 3 * n + 1)))]
    d, order = (dev, "row") if seed % 2 == 0 else (dev_col, "col")
    check_select(d, order, a, la, axis, idx)


def test_index_select_into_strided_output_and_errors(dev):
    rng = np.random.default_rng(seed_of("selo"))
    a = rng.standard_normal((12, 20))
    ta = rt.asarray(a, dev)
    out = rt.full([5, 40], -1.0, dev)
    oc = out[:, ::2]  # (5, 20) with stride 2 along the kept axis
    dev.index_select(oc.raw, oc.layout, ta.raw, ta.layout, 0, [3, 3, 11, 0, 7])
    o = out.to_numpy()
    assert np.array_equal(o[:, ::2], a[[3, 3, 11, 0, 7]]) and np.all(o[:, 1::2] == -1.0)
    with pytest.raises(rt.RstsrCudaError) as ei:
        ta.index_select(0, [12])
    assert ei.value.kind == "IndexError"
    with pytest.raises(rt.RstsrCudaError) as ei:
        ta.index_select(1, [-21])
    assert ei.value.kind == "IndexError"
    with pytest.raises(rt.RstsrCudaError) as ei:
        ta.index_select(2, [0])
    assert ei.value.kind == "InvalidValue"
    with pytest.raises(rt.RstsrCudaError) as ei:  # device level: index list length must equal lc.shape[axis]
        dev.index_select(oc.raw, oc.layout, ta.raw, ta.layout, 0, [1, 2])
    assert ei.value.kind == "InvalidLayout" and "Invalid index length." in str(ei.value)
    with pytest.raises(rt.RstsrCudaError) as ei:
        dev.index_select(oc.raw, oc.layout, ta.raw, ta.layout, 0, [1, 2, 3, 4, 12])
    assert ei.value.kind == "IndexError" and "Index out of range." in str(ei.value)
    # take(indices, axis) is index_select(axis, indices)
    assert np.array_equal(ta.take([1, -1], 1).to_numpy(), a[:, [1, -1]])


def test_index_select_large_rows(dev):
    """row gather of a (4096, 4096) f64 matrix: the 16-byte word path"""
    rng = np.random.default_rng(seed_of("selbig"))
    a = rng.standard_normal((4096, 4096))
    idx = rng.integers(0, 4096, 4096)
    t = rt.asarray(a, dev)
    assert np.array_equal(t.index_select(0, idx).to_numpy(), a[idx])
    assert np.array_equal(t.index_select(1, idx[:100]).to_numpy(), a[:, idx[:100]])
    # gather along the contiguous axis with most of each row wanted: rows staged in shared memory
    assert np.array_equal(t.index_select(1, idx).to_numpy(), a[:, idx])
    assert np.array_equal(t[::2, 7:4000].index_select(1, idx[:3000] % 3993).to_numpy(), a[::2, 7:4000][:, idx[:3000] % 3993])
    for dt in (np.float32, np.int16, np.uint8):
        b = (a * 100).astype(dt)
        assert np.array_equal(rt.asarray(b, dev).index_select(-1, idx).to_numpy(), b[:, idx])
    out = rt.full([4096, 8192], -3.0, dev)
    oc = out[:, ::2]
    dev.index_select(oc.raw, oc.layout, t.raw, t.layout, 1, idx)
    o = out.to_numpy()
    assert np.array_equal(o[:, ::2], a[:, idx]) and np.all(o[:, 1::2] == -3.0)
    f = rt.asarray(a.astype(np.float32), dev)
    assert np.array_equal(f[1:, 1:].index_select(0, idx[:500] % 4095).to_numpy(), a.astype(np.float32)[1:, 1:][idx[:500] % 4095])


def test_index_select_long_rows_clustered_indices(dev):
    """gather along the contiguous axis of rows too long for shared memory: monotone / clustered selections take the
    windowed kernel (source span of every 2048-index window staged in smem), scattered ones the direct kernel"""
    rng = np.random.default_rng(seed_of("selwin"))
    n_src = 40000
    a = rng.standard_normal((37, n_src))
    t = rt.asarray(a, dev)
    every_other = np.arange(0, n_src, 2)
    dropped = np.delete(np.arange(n_src), rng.integers(0, n_src, 500))
    mask = np.nonzero(rng.random(n_src) < 0.7)[0]
    locally_shuffled = np.arange(20000).reshape(-1, 50)[:, ::-1].reshape(-1)      # clustered but not monotone
    repeated = np.repeat(np.arange(0, 12000, 3), 3)
    scattered = rng.integers(0, n_src, 30000)                                      # falls back to the direct kernel
    for idx in (every_other, dropped, mask, locally_shuffled, repeated, scattered):
        assert np.array_equal(t.index_select(1, idx).to_numpy(), a[:, idx])
    for dt in (np.float32, np.int16, np.uint8):
        b = (a * 50).astype(dt)
        assert np.array_equal(rt.asarray(b, dev).index_select(-1, mask).to_numpy(), b[:, mask])
    # sliced source rows (offset, pitch) and a strided output
    assert np.array_equal(t[3:, 11:].index_select(1, every_other[:15000]).to_numpy(), a[3:, 11:][:, every_other[:15000]])
    out = rt.full([37, 2 * mask.size], -1.0, dev)
    oc = out[:, ::2]
    dev.index_select(oc.raw, oc.layout, t.raw, t.layout, 1, mask)
    o = out.to_numpy()
    assert np.array_equal(o[:, ::2], a[:, mask]) and np.all(o[:, 1::2] == -1.0)


# ---------------- pack_tri / unpack_tri ----------------
def test_reference_kats(dev, dev_col):
    a = np.arange(48.0)
    t = rt.Tensor(upload(dev, a), P(L.f_contig_layout([3, 4, 4])))
    p = t.pack_tril()
    assert p.shape == (3, 10) and p.stride == (1, 3)  # f-preferred input -> f-contiguous output
    assert p.to_numpy()[1].tolist() == [1., 4., 16., 7., 19., 31., 10., 22., 34., 46.]
    b = p.unpack_tril("Sy")
    assert b.shape == (3, 4, 4) and b.to_numpy()[0, 1].tolist() == [3., 15., 18., 21.]
    t = rt.Tensor(upload(dev_col, a), P(L.c_contig_layout([4, 4, 3])))
    p = t.pack_triu()
    assert p.shape == (10, 3) and p.stride == (3, 1)
    assert p.to_numpy()[:, 1].tolist() == [1., 4., 16., 7., 19., 31., 10., 22., 34., 46.]
    b = p.unpack_triu("Sy")
    assert b.shape == (4, 4, 3) and b.to_numpy()[:, 1, 0].tolist() == [3., 15., 18., 21.]
    # test_correctness: packing a matrix and its f-contiguous copy give the same values
    m = rt.arange(16, dev, dtype=np.float64).reshape([4, 4])
    assert np.array_equal(m.pack_tril().to_numpy(), m.to_contig(rt.COL_MAJOR).pack_tril().to_numpy())


def full_layouts(rng, rest, n, order):
    """a few layouts of a (rest.., n, n) [row] / (n, n, rest..) [col] tensor and the buffer size they need"""
    shape = list(rest) + [n, n] if order == "row" else [n, n] + list(rest)
    outs = [L.c_contig_layout(shape), L.f_contig_layout(shape)]
    nd = len(shape)
    perm = [int(p) for p in rng.permutation(nd)]
    inv = [perm.index(i) for i in range(nd)]
    outs.append(L.c_contig_layout([shape[p] for p in perm]).transpose(inv))
    big = L.c_contig_layout([s + 2 for s in shape])
    for ax, s in enumerate(shape):
        big = big.narrow(ax, slice(1, 1 + s))
    outs.append(big)
    if n > 1:
        outs.append(L.c_contig_layout(shape).narrow(nd - 1 if order == "row" else 0, slice(None, None, -1)))
    return shape, outs


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int16, np.uint8])
@pytest.mark.parametrize("n", [1, 2, 5, 32, 33, 100])
def test_pack_tri_matches_oracle(dev, dev_col, dtype, n):
    rng = np.random.default_rng(seed_of("pack", n, np.dtype(dtype).name))
    for rest in ([], [3], [2, 5]):
        for d, order in ((dev, "row"), (dev_col, "col")):
            shape, layouts = full_layouts(rng, rest, n, order)
            for lb in layouts:
                nbuf = max(L.bounds_index(lb)[1], 1)
                b = rand_data(rng, nbuf, dtype)
                t = rt.Tensor(upload(d, b), P(lb))
                for uplo in ("L", "U"):
                    got = t.pack_tri(uplo)
                    want = np.zeros(max(got.layout.size, 1), dtype=b.dtype)
                    oracle.pack_tri(want, O(got.layout), b, lb, uplo, order)
                    assert np.array_equal(got.to_numpy(), view_np(want, O(got.layout))), (order, rest, uplo, lb)
                    # against numpy's own triangle extraction
                    vb = view_np(b, lb)
                    if order == "col":
                        vb = vb.transpose(list(range(vb.ndim))[::-1])
                        ref = np.tril(np.ones((n, n), bool)) if uplo == "U" else np.triu(np.ones((n, n), bool))
                    else:
                        ref = np.tril(np.ones((n, n), bool)) if uplo == "L" else np.triu(np.ones((n, n), bool))
                    g = got.to_numpy()
                    if order == "col":
                        g = g.transpose(list(range(g.ndim))[::-1])
                    assert np.array_equal(g, vb[..., ref])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 7, 32, 33, 70])
def test_unpack_tri_matches_oracle(dev, dev_col, dtype, n):
    rng = np.random.default_rng(seed_of("unpack", n, np.dtype(dtype).name))
    n_tp = n * (n + 1) // 2
    for rest in ([], [3], [2, 4]):
        for d, order in ((dev, "row"), (dev_col, "col")):
            pshape = list(rest) + [n_tp] if order == "row" else [n_tp] + list(rest)
            nd = len(pshape)
            lps = [L.c_contig_layout(pshape), L.f_contig_layout(pshape)]
            if n_tp > 1:
                lps.append(L.c_contig_layout(pshape).narrow(nd - 1 if order == "row" else 0, slice(None, None, -1)))
            for lb in lps:
                b = rand_data(rng, max(L.bounds_index(lb)[1], 1), dtype)
                t = rt.Tensor(upload(d, b), P(lb))
                for uplo in ("L", "U"):
                    for symm in ("Sy", "He", "Ay", "Ah"):
                        got = t.unpack_tri(uplo, symm)
                        want = np.zeros(max(got.layout.size, 1), dtype=b.dtype)
                        oracle.unpack_tri(want, O(got.layout), b, lb, uplo, symm, order)
                        assert np.array_equal(got.to_numpy(), view_np(want, O(got.layout))), (order, rest, uplo, symm)
                        # pack(unpack(x)) == x for the symmetric flavours; the antisymmetric ones zero the diagonal
                        back = got.pack_tri(uplo).to_numpy()
                        src = view_np(b, lb)
                        if symm in ("Sy", "He"):
                            assert np.array_equal(back, src)
                        else:
                            g = got.to_numpy()
                            gt = g.swapaxes(-1, -2) if order == "row" else g.swapaxes(0, 1)
                            assert np.array_equal(g, -gt)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("rest", [[8], [5], [4, 6], [6, 3], [16, 2]])
def test_tri_batch_axis_fastest(dev, dev_col, dtype, rest):
    """the batch axes are the memory-fastest ones ([.., n, n].f() on a row-major handle and its col-major twin): the
    run-moving kernels, with every pack width the alignment allows"""
    rng = np.random.default_rng(seed_of("trirest", rest, np.dtype(dtype).name))
    for n in (1, 2, 7, 33):
        n_tp = n * (n + 1) // 2
        for d, order in ((dev, "row"), (dev_col, "col")):
            fshape = rest + [n, n] if order == "row" else [n, n] + rest[::-1]
            pshape = rest + [n_tp] if order == "row" else [n_tp] + rest[::-1]
            mk = L.f_contig_layout if order == "row" else L.c_contig_layout   # the NON-default order
            lfull, lpacked = mk(fshape), mk(pshape)
            full = rng.standard_normal(int(np.prod(fshape))).astype(dtype)
            tf = rt.Tensor(upload(d, full), P(lfull))
            for uplo in ("L", "U"):
                got = tf.pack_tri(uplo)
                if n > 1:
                    assert same(got.layout, lpacked)  # the output follows the input's f- / c-preference
                lpacked = O(got.layout)
                want = np.zeros(max(lpacked.size, 1), dtype=dtype)
                oracle.pack_tri(want, lpacked, full, lfull, uplo, order)
                assert np.array_equal(got.to_numpy(), view_np(want, lpacked))
                for symm in ("Sy", "Ah", "N"):
                    out = rt.Tensor(upload(d, np.full(full.size, -5, dtype=dtype)), P(lfull))
                    d.unpack_tri(out.raw, out.layout, got.raw, got.layout, uplo, symm)
                    wfull = np.full(full.size, -5, dtype=dtype)
                    oracle.unpack_tri(wfull, lfull, want, lpacked, uplo, symm, order)
                    assert np.array_equal(out.to_numpy(), view_np(wfull, lfull)), (order, rest, n, uplo, symm)
                # sliced batch (offset, odd pitch): narrower packs
                if rest[0] > 4:
                    ax = 0 if order == "row" else len(fshape) - 1
                    sub_f = lfull.narrow(ax, slice(1, None))
                    tsub = rt.Tensor(tf.raw, P(sub_f))
                    gp = tsub.pack_tri(uplo)
                    wp = np.zeros(max(gp.layout.size, 1), dtype=dtype)
                    oracle.pack_tri(wp, O(gp.layout), full, sub_f, uplo, order)
                    assert np.array_equal(gp.to_numpy(), view_np(wp, O(gp.layout)))
                    un = gp.unpack_tri(uplo, "Ay")
                    wu = np.zeros(max(un.layout.size, 1), dtype=dtype)
                    oracle.unpack_tri(wu, O(un.layout), wp, O(gp.layout), uplo, "Ay", order)
                    assert np.array_equal(un.to_numpy(), view_np(wu, O(un.layout)))


def test_unpack_tri_n_leaves_other_triangle(dev, dev_col):
    rng = np.random.default_rng(seed_of("unpackN"))
    n, n_tp = 37, 37 * 38 // 2
    b = rng.standard_normal(2 * n_tp)
    for d, order, pshape, fshape in ((dev, "row", [2, n_tp], [2, n, n]), (dev_col, "col", [n_tp, 2], [n, n, 2])):
        lb = L.c_contig_layout(pshape) if order == "row" else L.f_contig_layout(pshape)
        tb = rt.Tensor(upload(d, b), P(lb))
        for uplo in ("L", "U"):
            out = rt.full(fshape, -9.0, d)
            d.unpack_tri(out.raw, out.layout, tb.raw, tb.layout, uplo, "N")
            want = np.full(2 * n * n, -9.0)
            oracle.unpack_tri(want, O(out.layout), b, lb, uplo, "N", order)
            assert np.array_equal(out.to_numpy(), view_np(want, O(out.layout)))
            assert (out.to_numpy() == -9.0).sum() == 2 * (n * n - n_tp)


def test_tri_errors(dev):
    x = rt.zeros([3, 4, 5], dev)
    with pytest.raises(rt.RstsrCudaError) as ei:
        x.pack_tril()
    assert ei.value.kind == "InvalidLayout" and "Last two dimensions should be the same" in str(ei.value)
    with pytest.raises(rt.RstsrCudaError) as ei:
        rt.zeros([3, 7], dev).unpack_tril("Sy")
    assert ei.value.kind == "InvalidLayout" and "triangular number" in str(ei.value)
    with pytest.raises(rt.RstsrCudaError) as ei:  # ComplexFloat only in the reference
        rt.zeros([6], dev, dtype=np.int32).unpack_tril("Sy")
    assert ei.value.kind == "UnImplemented"
    p, f = rt.zeros([2, 6], dev), rt.zeros([2, 4, 4], dev)
    with pytest.raises(rt.RstsrCudaError):
        dev.pack_tri(p.raw, p.layout, f.raw, f.layout, "L")
    with pytest.raises(rt.RstsrCudaError):
        dev.unpack_tri(f.raw, f.layout, p.raw, p.layout, "L", "Sy")
    e = rt.zeros([0, 3, 3], dev).pack_tril()
    assert e.shape == (0, 6)
    z = rt.zeros([2, 0, 0], dev).pack_triu()
    assert z.shape == (2, 0)


def test_tri_large_roundtrip(dev):
    """(6, 1500, 1500) f64: unpack(pack(x)) restores the stored triangle and mirrors it"""
    rng = np.random.default_rng(seed_of("tribig"))
    n = 1500
    a = rng.standard_normal((6, n, n))
    t = rt.asarray(a, dev)
    for uplo, tri in (("L", np.tril), ("U", np.triu)):
        p = t.pack_tri(uplo)
        assert p.shape == (6, n * (n + 1) // 2)
        s = p.unpack_tri(uplo, "Sy").to_numpy()
        want = tri(a) + np.swapaxes(tri(a, -1 if uplo == "L" else 1), -1, -2)
        assert np.array_equal(s, want)
        y = p.unpack_tri(uplo, "Ay").to_numpy()
        strict = tri(a, -1 if uplo == "L" else 1)
        assert np.array_equal(y, strict - np.swapaxes(strict, -1, -2))
