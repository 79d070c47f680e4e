"""GPU parity: assign / assign_arbitary / fill / to_contig / reshape through the C ABI vs the oracle.  Bit-exact."""
import numpy as np
import pytest

import oracle
import rstsr_b200 as rt
from oracle import layout as L

from helpers import O, P, rand_data, random_view, same, seed_of, upload, view_np

pytestmark = pytest.mark.gpu

ALL_DTYPES = [np.bool_, np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32,
              np.float64]


def _pair_same_shape(rng):
    """Two random views of one shape (dst must not self-overlap: no broadcast on dst)."""
    lc, nc = random_view(rng)
    shape = lc.shape
    stor_perm = [int(p) for p in rng.permutation(len(shape))]
    stor = L.c_contig_layout([shape[p] for p in stor_perm])
    la = stor.transpose([stor_perm.index(i) for i in range(len(shape))]) if shape else stor
    for ax in range(la.ndim):
        if rng.random() < 0.25:
            la = la.narrow(ax, slice(None, None, -1))
    if la.ndim and rng.random() < 0.3:  # broadcast source axis
        ax = int(rng.integers(0, la.ndim))
        st = list(la.stride)
        st[ax] = 0
        la = L.Layout(la.shape, tuple(st), la.offset)
    na = 1
    for d in shape:
        na *= d
    return lc, nc, la, max(na, 1)


@pytest.mark.parametrize("tc", ALL_DTYPES)
@pytest.mark.parametrize("ta", ALL_DTYPES)
def test_assign_with_cast_all_dtype_pairs(dev, tc, ta):
    """c[idx] = cast(a[idx]) with Rust `as` semantics (promotion.rs): saturating float->int, NaN -> 0, != 0 to bool."""
    rng = np.random.default_rng(seed_of(np.dtype(tc).name, np.dtype(ta).name))
    for it in range(3):
        lc, nc, la, na = _pair_same_shape(rng)
        a = rand_data(rng, na, ta)
        if np.dtype(ta).kind == "f":
            a = (a * np.dtype(ta).type(1e3)).astype(ta)
            a[::4] = np.nan
            a[1::5] = np.inf
            a[2::6] = -np.inf
            a[3::7] = np.dtype(ta).type(3e38 if ta == np.float32 else 1e300)
        c0 = rand_data(rng, nc, tc)
        raw_c = upload(dev, c0)
        dev.assign(raw_c, P(lc), upload(dev, a), P(la))
        ref = c0.copy()
        oracle.assign(ref, lc, a, la)
        got = dev.to_cpu_vec(raw_c)
        assert np.array_equal(got.view(np.uint8), ref.view(np.uint8)), (tc, ta, lc, la)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int16, np.uint8])
@pytest.mark.parametrize("order", [(rt.ROW_MAJOR, L.ROW_MAJOR), (rt.COL_MAJOR, L.COL_MAJOR)])
def test_assign_arbitary_random_views(dev, dev_col, dtype, order):
    """Flattened-order pairing incl. shapes with and without a common refinement (cpu_serial/assignment.rs:28-67)."""
    d = dev if order[0] == rt.ROW_MAJOR else dev_col
    rng = np.random.default_rng(seed_of(np.dtype(dtype).name, order[1]))
    done = 0
    while done < 30:
        lc, nc = random_view(rng, max_extent=6)
        la, na = random_view(rng, max_extent=6)
        if lc.size != la.size:
            # make la's size match by viewing a flat buffer
            la = L.c_contig_layout([lc.size]) if rng.random() < 0.5 else la
            if lc.size != la.size:
                continue
            na = max(lc.size, 1)
        a = rand_data(rng, na, dtype)
        c0 = rand_data(rng, nc, dtype)
        raw_c = upload(d, c0)
        d.assign_arbitary(raw_c, P(lc), upload(d, a), P(la))
        ref = c0.copy()
        oracle.assign_arbitary(ref, lc, a, la, order[1])
        assert np.array_equal(d.to_cpu_vec(raw_c).view(np.uint8), ref.view(np.uint8)), (lc, la)
        done += 1


def test_assign_arbitary_size_mismatch_is_an_error(dev):
    a = upload(dev, np.zeros(6))
    with pytest.raises(rt.RstsrCudaError) as e:
        dev.assign_arbitary(a, rt.Layout((2, 3), (3, 1)), a, rt.Layout((5,), (1,)))
    assert e.value.kind == "InvalidLayout"


@pytest.mark.parametrize("dtype", [np.float64, np.int32, np.uint8, np.bool_])
def test_fill(dev, dtype):
    rng = np.random.default_rng(seed_of("fill", np.dtype(dtype).name))
    for _ in range(8):
        lc, nc = random_view(rng, allow_broadcast=True)
        c0 = rand_data(rng, nc, dtype)
        raw = upload(dev, c0)
        dev.fill(raw, P(lc), 1)
        ref = c0.copy()
        oracle.fill(ref, lc, 1)
        assert np.array_equal(dev.to_cpu_vec(raw), ref), lc
    # fill with a float into an int tensor casts like Rust `as` (fill_promote, cpu_rayon/assignment.rs:181-225)
    raw = upload(dev, np.zeros(4, dtype=np.int32))
    dev.fill(raw, rt.Layout((4,), (1,)), 2.9)
    assert dev.to_cpu_vec(raw).tolist() == [2, 2, 2, 2]
    dev.fill(raw, rt.Layout((4,), (1,)), float("nan"))
    assert dev.to_cpu_vec(raw).tolist() == [0, 0, 0, 0]
    dev.fill(raw, rt.Layout((4,), (1,)), 1e20)
    assert dev.to_cpu_vec(raw).tolist() == [2**31 - 1] * 4


def test_reference_to_contig_kats_on_device(dev, dev_col):
    """rstsr-core/tests/core_func/manipulation/test_to_contig.rs:12-116."""
    a = rt.arange(24, dev).reshape([2, 3, 4])
    assert a.layout.c_contig()
    v = a.to_contig(rt.ROW_MAJOR)
    assert not v.owned and v.raw.ptr == a.raw.ptr  # already contiguous: a view, no copy
    f = a.to_contig(rt.COL_MAJOR)
    assert f.owned and f.layout.f_contig() and np.array_equal(f.to_numpy(), np.arange(24).reshape(2, 3, 4))
    a = rt.arange(24, dev_col).reshape([2, 3, 4])  # col-major device: into_shape is F-ordered
    assert a.layout.f_contig()
    assert not a.to_contig(rt.COL_MAJOR).owned
    c = a.to_contig(rt.ROW_MAJOR)
    assert c.owned and c.layout.c_contig() and np.array_equal(c.to_numpy(), np.arange(24).reshape((2, 3, 4), order="F"))
    t = rt.arange(12, dev).reshape([3, 4]).reverse_axes()
    assert not t.layout.c_contig() and t.layout.f_contig()
    rc = t.to_contig(rt.ROW_MAJOR)
    assert rc.owned and rc.shape == (4, 3)
    assert rc.to_numpy().tolist() == [[0, 4, 8], [1, 5, 9], [2, 6, 10], [3, 7, 11]]
    assert not t.to_contig(rt.COL_MAJOR).owned
    s = rt.arange(24, dev).reshape([4, 6])[::2, ::2]
    assert (s.shape, s.stride) == ((2, 3), (12, 2))
    sc = s.to_contig(rt.ROW_MAJOR)
    assert sc.owned and sc.stride == (3, 1) and sc.to_numpy().tolist() == [[0, 2, 4], [12, 14, 16]]
    # issue 77: a contiguous slice with an offset still copies
    sl = rt.arange(24, dev).reshape([4, 6])[1:]
    out = sl.to_contig(rt.ROW_MAJOR)
    assert out.owned and out.layout.offset == 0 and np.array_equal(out.to_numpy(), np.arange(24).reshape(4, 6)[1:])


def test_reference_reshape_and_assign_kats_on_device(dev, dev_col):
    a = rt.arange(24, dev)
    assert not a.reshape([2, 3, 4]).owned
    t = a.reshape([4, 6]).reverse_axes()
    r = t.reshape([24])
    assert r.owned and np.array_equal(r.to_numpy(), np.arange(24).reshape(4, 6).T.reshape(-1))
    r = a.reshape([4, 6]).reshape([3, 8], order=rt.COL_MAJOR)
    assert np.array_equal(r.to_numpy(), np.arange(24).reshape(4, 6).reshape((3, 8), order="F"))
    with pytest.raises(rt.RstsrCudaError):
        a.reshape([5, 5])
    # tensor/assignment.rs:150-178: assign i32 -> f32 and fill
    x = rt.zeros([3, 5], dev, dtype=np.float32)
    x.assign(rt.arange(15, dev, dtype=np.int32).reshape([3, 5]))
    assert x.to_numpy().reshape(-1).tolist() == [float(i) for i in range(15)]
    x.assign(rt.arange(5, dev, dtype=np.int32))
    assert x.to_numpy().tolist() == [[0, 1, 2, 3, 4]] * 3
    x.fill(1.5)
    assert (x.to_numpy() == 1.5).all()


COPY_SHAPES = [
    ("2-D transpose", (300, 500), (1, 0)),
    ("3-D (2,0,1)", (40, 50, 60), (2, 0, 1)),
    ("3-D (1,0,2) inner kept", (33, 65, 128), (1, 0, 2)),
    ("4-D (3,1,0,2)", (6, 17, 9, 70), (3, 1, 0, 2)),
    ("thin", (2, 100000), (1, 0)),
    ("ragged tiles", (65, 129), (1, 0)),
    # one short and one long extent: ew_tile_rect_kernel (short extent whole, linear walk of the tile)
    ("rect narrow X=17", (17, 1000), (1, 0)),
    ("rect narrow X=40", (40, 700), (1, 0)),
    ("rect narrow Y=17", (1000, 17), (1, 0)),
    ("rect narrow Y=48", (700, 48), (1, 0)),
    ("rect narrow X batched ragged", (3, 20, 300), (0, 2, 1)),
    ("rect narrow Y batched ragged", (3, 300, 20), (0, 2, 1)),
    ("rect narrow Y=7 long X", (5000, 7), (1, 0)),
    ("rect narrow X=3", (3, 1000), (1, 0)),
    ("rect narrow X=31", (31, 333), (1, 0)),
    ("rect narrow Y=32", (300, 32), (1, 0)),
    ("rect narrow Y=4 batched", (2, 1500, 4), (0, 2, 1)),
    ("rect narrow Y=2", (100000, 2), (1, 0)),
    ("rect narrow Y=3", (4097, 3), (1, 0)),
]


@pytest.mark.parametrize("name,shape,perm", COPY_SHAPES, ids=[c[0] for c in COPY_SHAPES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int16, np.uint8])
@pytest.mark.parametrize("target", [rt.ROW_MAJOR, rt.COL_MAJOR])
def test_permuted_to_contig(dev, name, shape, perm, dtype, target):
    """Tile-transposing copy: to_contig of a permuted view, row- and col-major targets (config 2 in small)."""
    rng = np.random.default_rng(seed_of(name, np.dtype(dtype).name))
    src = rand_data(rng, int(np.prod(shape)), dtype)
    t = rt.asarray(src, dev).reshape(list(shape)).transpose(list(perm)).to_contig(target)
    want = src.reshape(shape).transpose(perm)
    assert t.layout.c_contig() if target == rt.ROW_MAJOR else t.layout.f_contig()
    assert np.array_equal(t.to_numpy(), want)
    lsrc = L.c_contig_layout(shape).transpose(list(perm))
    ref, lref, _ = oracle.tensor_to_contig(src, lsrc, L.ROW_MAJOR if target == rt.ROW_MAJOR else L.COL_MAJOR)
    assert same(t.layout, lref)
    assert np.array_equal(dev.to_cpu_vec(t.raw)[:ref.size], ref)


def test_copy_into_strided_destination(dev):
    """assign into a transposed / stepped destination view: only the addressed elements change."""
    rng = np.random.default_rng(9)
    dst0 = rand_data(rng, 70 * 90, np.float64)
    src = rand_data(rng, 35 * 90, np.float64)
    ld = L.c_contig_layout([90, 70]).reverse_axes().narrow(0, slice(None, None, 2))  # (35, 90) view
    ls = L.c_contig_layout([35, 90])
    raw = upload(dev, dst0)
    dev.assign(raw, P(ld), upload(dev, src), P(ls))
    ref = dst0.copy()
    oracle.assign(ref, ld, src, ls)
    assert np.array_equal(dev.to_cpu_vec(raw), ref)


def test_to_owned_uses_k_order(dev):
    a = rt.arange(24, dev, dtype=np.float64).reshape([2, 3, 4]).transpose([2, 0, 1])
    o = a.to_owned()
    want = L.layout_for_array_copy(O(a.layout), "K")
    assert same(o.layout, want) and np.array_equal(o.to_numpy(), a.to_numpy())


def test_change_device_between_handles(dev):
    """DeviceChangeAPI between two DeviceCuda handles: another stream of the same GPU, and a second GPU if present."""
    import rstsr_b200 as rt
    rng = np.random.default_rng(seed_of("chdev"))
    a = rng.standard_normal(1 << 20)
    t = rt.asarray(a, dev).reshape([1024, 1024])[::2, ::-1]
    other = rt.DeviceCuda(0, rt.COL_MAJOR)
    targets = [other]
    if rt.DeviceCuda.device_count() > 1:
        targets.append(rt.DeviceCuda(1, rt.ROW_MAJOR))
    try:
        for tgt in targets:
            moved = (t + 1.0).to_device(tgt)           # produced on dev's stream, consumed on the target's
            assert moved.device is tgt and moved.shape == (512, 1024)
            back = (moved * 2.0).to_device(dev)        # temporaries freed right after the copy is enqueued
            want = (a.reshape(1024, 1024)[::2, ::-1] + 1.0) * 2.0
            assert np.array_equal(back.to_numpy(), want)
            v = t.to_device(tgt)
            assert same(O(v.layout), O(t.layout)) and np.array_equal(v.to_numpy(), t.to_numpy())
    finally:
        for tgt in targets:
            tgt.close()


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.bool_, np.int16, np.uint16])
def test_narrow_permuted_copies(dev, dtype):
    """1- / 2-byte permuted copies take the word-granular tile (rc_tile_narrow.cuh) when extents and strides allow,
    the one-element-per-lane kernels otherwise: both must be bit-exact, partial tiles and batches included."""
    rng = np.random.default_rng(seed_of(("narrow", np.dtype(dtype).name)))
    cases = [((256, 384), (1, 0)), ((132, 260), (1, 0)), ((130, 258), (1, 0)), ((131, 257), (1, 0)), ((5, 200, 136), (0, 2, 1)),
             ((72, 3, 520), (2, 1, 0)), ((3, 140, 2, 148), (2, 3, 0, 1)), ((64, 64), (1, 0)), ((1000, 12), (1, 0)), ((12, 1000), (1, 0))]
    for shape, perm in cases:
        n = int(np.prod(shape))
        a = rng.integers(0, 2 if dtype == np.bool_ else 120, n).astype(dtype)
        la = L.c_contig_layout(list(shape)).transpose(list(perm))
        t = rt.Tensor(upload(dev, a), P(la)).to_contig(rt.ROW_MAJOR)
        want = np.ascontiguousarray(a.reshape(shape).transpose(perm))
        assert np.array_equal(t.to_numpy(), want), (dtype, shape, perm)
        # a sliced source (offset, pitch not a multiple of 4 bytes in general)
        sl = la.narrow(0, slice(1, None)).narrow(1, slice(2, None))
        t2 = rt.Tensor(upload(dev, a), P(sl)).to_contig(rt.ROW_MAJOR)
        idx = [slice(None)] * len(shape)
        v = a.reshape(shape).transpose(perm)[1:, 2:]
        assert np.array_equal(t2.to_numpy(), np.ascontiguousarray(v)), (dtype, shape, perm, "sliced")


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int16, np.uint8])
def test_short_axis_packed_transposes(dev, dtype):
    """(n, k) <-> (k, n) word copies with a packed short axis take ew_tile_short_kernel (rc_tile_short.cuh) in both
    directions -- 16-byte vectors when extents, strides and pointers allow, element by element otherwise; unpacked
    (sliced) short rows fall back to the rectangular / square tile.  All bit-exact."""
    rng = np.random.default_rng(seed_of(("strip", np.dtype(dtype).name)))
    for k in (2, 3, 4, 5, 8, 12, 17, 31, 32, 33, 48, 64, 65):
        for n in (128, 130, 1024, 4112, 5003):
            a = rng.integers(0, 120, n * k + 3).astype(dtype)
            for off in (0, 1):  # off = 1: base pointer not 16-byte aligned -> scalar variant
                # short Y: source rows of k elements packed -> k long output rows
                t = rt.Tensor(upload(dev, a), rt.Layout((k, n), (1, k), off)).to_contig(rt.ROW_MAJOR)
                assert np.array_equal(t.to_numpy(), a[off:off + n * k].reshape(n, k).T), (dtype, k, n, off, "short y")
                # short X: k long source rows -> output rows of k elements packed
                t = rt.Tensor(upload(dev, a), rt.Layout((n, k), (1, n), off)).to_contig(rt.ROW_MAJOR)
                assert np.array_equal(t.to_numpy(), a[off:off + n * k].reshape(k, n).T), (dtype, k, n, off, "short x")
    # batched (both directions), a pitched rows side, and a sliced source whose short rows are NOT packed
    a = rng.integers(0, 120, 3 * 4000 * 6).astype(dtype)
    t = rt.Tensor(upload(dev, a), rt.Layout((3, 6, 4000), (24000, 1, 6))).to_contig(rt.ROW_MAJOR)
    assert np.array_equal(t.to_numpy(), a.reshape(3, 4000, 6).transpose(0, 2, 1))
    t = rt.Tensor(upload(dev, a), rt.Layout((3, 4000, 6), (24000, 1, 4000))).to_contig(rt.ROW_MAJOR)
    assert np.array_equal(t.to_numpy(), a.reshape(3, 6, 4000).transpose(0, 2, 1))
    t = rt.Tensor(upload(dev, a), rt.Layout((2048, 6), (1, 4000), 16)).to_contig(rt.ROW_MAJOR)  # k rows at pitch 4000
    assert np.array_equal(t.to_numpy(), a[:24000].reshape(6, 4000)[:, 16:2064].T)
    t = rt.Tensor(upload(dev, a), rt.Layout((5, 4000), (1, 6), 1)).to_contig(rt.ROW_MAJOR)
    assert np.array_equal(t.to_numpy(), a[:24000].reshape(4000, 6)[:, 1:6].T)
    # de-interleave INTO a pitched destination: k output rows of a wider array
    src = rng.integers(0, 120, 2048 * 8).astype(dtype)
    raw = upload(dev, np.zeros(8 * 2304, dtype=dtype))
    dev.assign(raw, rt.Layout((8, 2048), (2304, 1), 128), upload(dev, src), rt.Layout((8, 2048), (1, 8)))
    want = np.zeros((8, 2304), dtype=dtype)
    want[:, 128:2176] = src.reshape(2048, 8).T
    assert np.array_equal(dev.to_cpu_vec(raw).reshape(8, 2304), want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int16, np.uint8])
def test_random_permutations_through_the_tile_kernels(dev, dtype):
    """Seeded random 3- / 4-D permutations whose extents straddle the selection thresholds of the square, wide, word,
    rectangular and short-axis tiles (short extents 2..70, long ones 100..700, sliced / flipped sources, both targets):
    whichever kernel the launcher picks, the copy is bit-exact."""
    rng = np.random.default_rng(seed_of(("randperm", np.dtype(dtype).name)))
    for case in range(48):
        nd = int(rng.integers(2, 5))
        shape = [int(rng.choice([rng.integers(2, 71), rng.integers(100, 700)])) for _ in range(nd)]
        while int(np.prod(shape)) > (1 << 22):
            shape[int(np.argmax(shape))] //= 2
        perm = [int(p) for p in rng.permutation(nd)]
        a = rng.integers(0, 120, int(np.prod(shape))).astype(dtype)
        view = a.reshape(shape).transpose(perm)
        la = L.c_contig_layout(shape).transpose(perm)
        if case % 3 == 1:      # a slice: offsets and pitches that break the vector / flat preconditions
            ax = int(rng.integers(0, nd))
            lo = int(rng.integers(0, 3))
            la = la.narrow(ax, slice(lo, None))
            view = view[(slice(None),) * ax + (slice(lo, None),)]
        elif case % 3 == 2:    # a flipped axis
            ax = int(rng.integers(0, nd))
            la = la.narrow(ax, slice(None, None, -1))
            view = view[(slice(None),) * ax + (slice(None, None, -1),)]
        for target, order in ((rt.ROW_MAJOR, "C"), (rt.COL_MAJOR, "F")):
            t = rt.Tensor(upload(dev, a), P(la)).to_contig(target)
            got = t.to_numpy()
            assert got.shape == view.shape and np.array_equal(got, view), (dtype, case, shape, perm, order)
