"""GPU parity at BASELINE.json's FULL sizes, through size-independent properties (the oracle would take minutes):
round trips, comparison against independently computed slabs, known-answer reductions, linearity."""
import numpy as np
import pytest

import rstsr_b200 as rt
from rstsr_b200 import Layout

pytestmark = pytest.mark.gpu


def _device_random(dev, n, dtype, seed):
    """Deterministic pseudo-random device data without a 2^33-byte host buffer: upload one 2^24 block and tile it
    with different affine twists per block (values stay in [0, 1))."""
    rng = np.random.default_rng(seed)
    block = rng.random(1 << 24).astype(dtype)
    raw = dev.uninit_impl(dtype, n)
    blk = dev.outof_cpu_vec(block)
    nb = n // block.size
    lb = Layout((block.size,), (1,))
    for i in range(nb):
        dst = Layout((block.size,), (1,), i * block.size)
        # block_i = frac(block * (1 + i/nb))  -- computed on device: mul then rem 1.0
        dev.op_mutc_refa_numb("mul", raw, dst, blk, lb, 1.0 + i / nb)
        dev.op_muta_numb("rem", raw, dst, 1.0)
    host = lambda i: np.fmod(block * np.dtype(dtype).type(1.0 + i / nb), np.dtype(dtype).type(1.0))
    return raw, host, block.size


def test_cfg1_broadcast_add_full_size_bit_exact(dev):
    n = 8192
    rng = np.random.default_rng(43)
    a = rng.standard_normal(n * n)
    b = rng.standard_normal(n)
    c = rt.asarray(a, dev).reshape([n, n]) + rt.asarray(b, dev)
    assert c.layout.shape == (n, n) and c.layout.stride == (n, 1)
    assert np.array_equal(c.to_numpy(), a.reshape(n, n) + b)


@pytest.mark.parametrize("target", [rt.ROW_MAJOR, rt.COL_MAJOR])
def test_cfg2_permuted_copy_full_size(dev, target):
    """(1024,1024,512) f64 viewed transpose(2,0,1) -> to_contig: sampled slabs equal numpy's, and permuting back
    with a second to_contig reproduces the source bit for bit (round trip over all 2^29 elements)."""
    shp = (1024, 1024, 512)
    n = shp[0] * shp[1] * shp[2]
    raw, host_block, bs = _device_random(dev, n, np.float64, 44)
    src = rt.Tensor(raw, Layout.contig(shp, rt.ROW_MAJOR))
    out = src.transpose([2, 0, 1]).to_contig(target)
    assert out.shape == (512, 1024, 1024)
    assert out.layout.c_contig() if target == rt.ROW_MAJOR else out.layout.f_contig()
    # slab check: out[:, i, :] == src[i, :, :].T for a few i (one i spans 2^19 elements = 1/32 of a block)
    for i in (0, 517, 1023):
        blk = host_block((i * shp[1] * shp[2]) // bs)
        off = (i * shp[1] * shp[2]) % bs
        want = blk[off:off + shp[1] * shp[2]].reshape(shp[1], shp[2]).T
        got = out[:, i, :].to_contig(rt.ROW_MAJOR).to_numpy()
        assert np.array_equal(got, want)
    back = out.transpose([1, 2, 0]).to_contig(rt.ROW_MAJOR)  # (1024,1024,512) again
    diff = back.binary("ne", src)
    assert diff.astype(np.int32).sum_all() == 0
    del out, back, diff


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-12)])
def test_cfg3_axis_reductions_full_size(dev, dtype, tol):
    n = 16384
    rng = np.random.default_rng(45)
    a = rng.random(n * n).astype(dtype)
    a[12345 * n + 678] = 7.5  # planted maximum
    t = rt.asarray(a, dev).reshape([n, n])
    an = a.reshape(n, n)
    l1 = {0: np.abs(an.astype(np.float64)).sum(0), 1: np.abs(an.astype(np.float64)).sum(1)}
    for axis in (0, -1):
        got = t.sum_axes(axis).to_numpy().astype(np.float64)
        want = an.sum(axis=axis, dtype=np.float64)
        assert (np.abs(got - want) <= tol * l1[axis % 2]).all(), (dtype, axis)
        mx = t.max_axes(axis).to_numpy()
        assert np.array_equal(mx, an.max(axis=axis))
        assert mx.max() == dtype(7.5)
    assert t.max_all() == dtype(7.5)
    s = float(t.sum_all())
    assert abs(s - an.sum(dtype=np.float64)) <= tol * l1[0].sum()
    # F-contiguous input (col-major data): reducing axis 0 is now the contiguous case; output layout follows
    tf = rt.Tensor(t.raw, Layout((n, n), (1, n)))
    got = tf.sum_axes(0).to_numpy().astype(np.float64)
    assert (np.abs(got - an.T.sum(axis=0, dtype=np.float64)) <= tol * l1[1]).all()


def test_cfg4_mixed_stride_ternary_full_size(dev):
    """c = a + b.transpose(1,0,3,2) on (64,64,512,512) f64 (24 GiB live): slabs vs numpy, then (c - a) equals the
    permuted b everywhere (checked on device)."""
    shp = (64, 64, 512, 512)
    n = 64 * 64 * 512 * 512
    ra, host_a, bs = _device_random(dev, n, np.float64, 46)
    rb, host_b, _ = _device_random(dev, n, np.float64, 47)
    a = rt.Tensor(ra, Layout.contig(shp, rt.ROW_MAJOR))
    b = rt.Tensor(rb, Layout.contig(shp, rt.ROW_MAJOR))
    bt = b.transpose([1, 0, 3, 2])
    c = a + bt
    assert c.layout.stride == Layout.contig(shp, rt.ROW_MAJOR).stride  # get_layout_for_binary_op -> C-contiguous
    slab = 512 * 512
    for (i, j) in ((0, 0), (3, 5), (63, 62)):
        ia, ib = (i * 64 + j) * slab, (j * 64 + i) * slab
        wa = host_a(ia // bs)[ia % bs: ia % bs + slab].reshape(512, 512)
        wb = host_b(ib // bs)[ib % bs: ib % bs + slab].reshape(512, 512)
        got = c[i, j].to_contig(rt.ROW_MAJOR).to_numpy()
        assert np.array_equal(got, wa + wb.T)
    # secondary: broadcast operand with strides (0,0,512,1), in place into c:  c = a * v
    v = rt.Tensor(rb, Layout(shp, (0, 0, 512, 1)))
    dev.op_mutc_refa_refb("mul", c.raw, c.layout, a.raw, a.layout, v.raw, v.layout)
    wv = host_b(0)[:slab].reshape(512, 512)
    ia = (17 * 64 + 9) * slab
    wa = host_a(ia // bs)[ia % bs: ia % bs + slab].reshape(512, 512)
    assert np.array_equal(c[17, 9].to_contig(rt.ROW_MAJOR).to_numpy(), wa * wv)
    del c


def test_cfg5_full_reduction_known_answer(dev):
    """2^31 f64 (16 GiB: one GPU's share of the 64 GiB config at 4 GPUs): sum and max with a closed-form answer."""
    n = 1 << 31
    raw = dev.uninit_impl(np.float64, n)
    l = Layout((n,), (1,))
    dev.fill(raw, l, 0.5)
    assert dev.reduce_all("sum", raw, l) == 0.5 * n       # exact: every partial is a multiple of 0.5 below 2^53
    dev.set_index(raw, 1234567891, 3.25)
    assert dev.reduce_all("max", raw, l) == 3.25
    assert dev.reduce_all("min", raw, l) == 0.5
    assert dev.reduce_all("sum", raw, l) == 0.5 * n + 2.75
    # as a (2^15, 2^16) C-contiguous matrix: the canonicaliser merges it to the same flat reduction
    assert dev.reduce_all("sum", raw, Layout((1 << 15, 1 << 16), (1 << 16, 1))) == 0.5 * n + 2.75
    assert dev.reduce_all("mean", raw, l) == (0.5 * n + 2.75) / n
