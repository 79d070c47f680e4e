"""Shared test helpers: seeded strided views over flat buffers, oracle <-> product conversions."""
import numpy as np

import oracle
from oracle import layout as OL


def seed_of(*args):
    """Stable (process-independent) seed from any printable arguments."""
    import zlib
    return zlib.crc32(repr(args).encode())


def P(l):
    """oracle Layout -> product Layout"""
    import rstsr_b200 as rt
    return rt.Layout(l.shape, l.stride, l.offset)


def O(l):
    """product Layout -> oracle Layout"""
    return OL.Layout(l.shape, l.stride, l.offset)


def same(l1, l2):
    return (tuple(l1.shape), tuple(l1.stride), l1.offset) == (tuple(l2.shape), tuple(l2.stride), l2.offset)


def random_view(rng, max_ndim=4, max_extent=7, allow_neg=True, allow_step=True, allow_broadcast=False):
    """A random valid view (oracle Layout) and the size of the buffer it lives in."""
    nd = int(rng.integers(0, max_ndim + 1))
    shape = [int(rng.integers(1, max_extent + 1)) for _ in range(nd)]
    base = OL.c_contig_layout(shape) if rng.random() < 0.5 else OL.f_contig_layout(shape)
    perm = list(rng.permutation(nd)) if nd else []
    l = base.transpose([int(p) for p in perm]) if nd else base
    for ax in range(nd):
        r = rng.random()
        if allow_neg and r < 0.25:
            l = l.narrow(ax, slice(None, None, -1))
        elif allow_step and r < 0.45 and l.shape[ax] > 1:
            l = l.narrow(ax, slice(int(rng.integers(0, 2)), None, 2))
        elif r < 0.55 and l.shape[ax] > 2:
            l = l.narrow(ax, slice(1, -1))
    if allow_broadcast and nd and rng.random() < 0.3:
        ax = int(rng.integers(0, nd))
        st = list(l.stride)
        st[ax] = 0
        l = OL.Layout(l.shape, tuple(st), l.offset)
    size = 1
    for d in shape:
        size *= d
    return l, max(size, 1)


def rand_data(rng, n, dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        return rng.standard_normal(n).astype(dtype)
    if dtype.kind == "b":
        return rng.integers(0, 2, n).astype(np.bool_)
    info = np.iinfo(dtype)
    lo, hi = max(info.min, -1000), min(info.max, 1000)
    return rng.integers(lo, hi, n, endpoint=True).astype(dtype)


def upload(dev, arr):
    return dev.outof_cpu_vec(np.ascontiguousarray(arr))


def view_np(raw, l):
    return oracle.to_numpy(raw, l)
