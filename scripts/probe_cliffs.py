"""Scratch probe: hunt for layout regimes where a kernel falls off its roofline -- awkward permutations, tiny inner
extents, sliced / misaligned views, small reduced extents -- next to torch on the same views (not product)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)
NP = {torch.float64: np.float64, torch.float32: np.float32, torch.uint8: np.uint8, torch.int16: np.int16, torch.int64: np.int64}


def timeit(fn, iters=10, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def wrap(t):
    """torch tensor (any strides) -> rt.Tensor view of the same memory"""
    base = t.untyped_storage()
    item = t.element_size()
    n = base.nbytes() // item
    raw = dev.wrap(base.data_ptr(), n, NP[t.dtype])
    return rt.Tensor(raw, rt.Layout(tuple(t.shape), tuple(t.stride()), t.storage_offset()), owned=False)


def rep(name, nbytes, ours, theirs):
    so, st = timeit(ours), timeit(theirs)
    go, gt = nbytes / so / 1e9, nbytes / st / 1e9
    flag = "  <-- CLIFF" if go < 0.8 * gt and go < 5000 else ""
    print(f"{name:64s} ours {go:7.0f} GB/s ({so * 1e6:7.1f} us)  torch {gt:7.0f} GB/s{flag}", flush=True)


def rand(shape, dt):
    if dt.is_floating_point:
        return torch.rand(*shape, dtype=dt, device="cuda")
    return torch.randint(0, 100, shape, dtype=dt, device="cuda")


# ---- permuted copies ----
for dt in (torch.float64, torch.float32, torch.uint8):
    a = rand((256, 512, 256), dt)
    nb = 2 * a.numel() * a.element_size()
    for perm in ((0, 2, 1), (1, 0, 2), (2, 1, 0), (1, 2, 0), (2, 0, 1)):
        v = a.permute(*perm)
        tv = wrap(v)
        rep(f"to_contig {str(dt)[6:]} (256,512,256).permute{perm}", nb, lambda: tv.to_contig(rt.ROW_MAJOR), lambda: v.contiguous())
for dt, n in ((torch.float32, 8192), (torch.int16, 16384), (torch.uint8, 16384)):
    a = rand((n, n), dt)
    v = a.t()
    tv = wrap(v)
    rep(f"transpose copy {str(dt)[6:]} ({n},{n})", 2 * a.numel() * a.element_size(), lambda: tv.to_contig(rt.ROW_MAJOR), lambda: v.contiguous())

# ---- tiny inner extents ----
for shape in ((1 << 24, 3), (1 << 22, 5), (1 << 20, 17)):
    a = rand(shape, torch.float64)
    b = rand(shape, torch.float64)
    ta, tb = wrap(a), wrap(b)
    nb = a.numel() * 8
    rep(f"add f64 {shape}", 3 * nb, lambda: ta + tb, lambda: a + b)
    at = rand((shape[1], shape[0]), torch.float64).t()
    tat = wrap(at)
    rep(f"to_contig f64 of ({shape[1]},{shape[0]}).T", 2 * nb, lambda: tat.to_contig(rt.ROW_MAJOR), lambda: at.contiguous())
    rep(f"add f64 {shape} + transposed operand", 3 * nb, lambda: ta + tat, lambda: a + at)

# ---- broadcasts ----
n = 8192
col, row = rand((n, 1), torch.float64), rand((1, n), torch.float64)
tcol, trow = wrap(col), wrap(row)
rep("outer sum (n,1)+(1,n) f64", n * n * 8, lambda: tcol + trow, lambda: col + row)
m = rand((n, n), torch.float64)
tm = wrap(m)
scratch = rand((n, n), torch.float64)
tscr = wrap(scratch)
rep("fill (n,n) f64 (write-only)", n * n * 8, lambda: tscr.fill(1.5), lambda: scratch.fill_(1.5))
rep("(n,n) * (n,1) f64", 2 * n * n * 8, lambda: tm * tcol, lambda: m * col)
rep("(n,n).T + (n,n) f64", 3 * n * n * 8, lambda: wrap(m.t()) + tm, lambda: m.t() + m)
rep("(n,n).T + (n,n).T f64", 3 * n * n * 8, lambda: wrap(m.t()) + wrap(m.t()), lambda: m.t() + m.t())

# ---- sliced / misaligned views ----
s1 = m[:, 1:-1]
rep("copy of a[:, 1:-1] (misaligned rows) f64", 2 * s1.numel() * 8, lambda: wrap(s1).to_contig(rt.ROW_MAJOR), lambda: s1.contiguous())
s2 = m[1::2]
rep("copy of a[1::2] f64", 2 * s2.numel() * 8, lambda: wrap(s2).to_contig(rt.ROW_MAJOR), lambda: s2.contiguous())
s3 = m[:, ::2]
rep("copy of a[:, ::2] f64 (useful bytes)", 2 * s3.numel() * 8, lambda: wrap(s3).to_contig(rt.ROW_MAJOR), lambda: s3.contiguous())
s4 = m.flip(1)
rep("copy of a.flip(1) f64", 2 * s4.numel() * 8, lambda: wrap(m)[:, ::-1].to_contig(rt.ROW_MAJOR), lambda: m.flip(1))
f32 = rand((n, n), torch.float32)
rep("copy of f32 a[:, 1:-1]", 2 * (n * (n - 2)) * 4, lambda: wrap(f32[:, 1:-1]).to_contig(rt.ROW_MAJOR), lambda: f32[:, 1:-1].contiguous())

# ---- reductions ----
c3 = rand((256, 512, 512), torch.float64)
t3 = wrap(c3)
nb3 = c3.numel() * 8
rep("sum axes (0,2) of (256,512,512) f64", nb3, lambda: t3.sum_axes([0, 2]), lambda: c3.sum((0, 2)))
rep("sum axes (0,1) of (256,512,512) f64", nb3, lambda: t3.sum_axes([0, 1]), lambda: c3.sum((0, 1)))
rep("sum axes (1,2) of (256,512,512) f64", nb3, lambda: t3.sum_axes([1, 2]), lambda: c3.sum((1, 2)))
rep("sum axis 1 of (256,512,512).permute(2,0,1) f64", nb3, lambda: wrap(c3.permute(2, 0, 1)).sum_axes(1), lambda: c3.permute(2, 0, 1).sum(1))
rep("max axis -1 of a[:, 1:-1] f64", s1.numel() * 8, lambda: wrap(s1).max_axes(-1), lambda: s1.amax(-1))
rep("argmax axis 0 of (n,n) f64", n * n * 8, lambda: tm.argmax_axes(0), lambda: m.argmax(0))
rep("argmax axis 1 of (n,n) f64", n * n * 8, lambda: tm.argmax_axes(1), lambda: m.argmax(1))
for shape in ((1 << 24, 4), (1 << 22, 16), (1 << 20, 100)):
    a = rand(shape, torch.float64)
    ta = wrap(a)
    rep(f"sum axis -1 of {shape} f64", a.numel() * 8, lambda: ta.sum_axes(-1), lambda: a.sum(-1))
    rep(f"var axis -1 of {shape} f64", a.numel() * 8, lambda: ta.var_axes(-1), lambda: a.var(-1, unbiased=False))

# ---- misaligned contiguous rows, more dtypes ----
for dt in (torch.float64, torch.float32, torch.int16):
    a = rand((8192, 8192), dt)
    sl = a[:, 1:-1]
    ts = wrap(sl)
    nb = sl.numel() * a.element_size()
    rep(f"sum axis -1 of a[:, 1:-1] {str(dt)[6:]}", nb, lambda: ts.sum_axes(-1), lambda: sl.sum(-1))
    rep(f"max axis -1 of a[:, 1:-1] {str(dt)[6:]}", nb, lambda: ts.max_axes(-1), lambda: sl.amax(-1))
