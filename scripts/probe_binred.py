"""Scratch probe: achieved HBM bandwidth of the binary reductions (vecdot / allclose) next to torch (not product)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import rstsr_b200 as rt
from rstsr_b200 import Layout

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def rep(name, nbytes, fn):
    s = timeit(fn)
    print(f"{name:58s} {nbytes / s / 1e9:8.1f} GB/s  {s * 1e6:8.1f} us", flush=True)


for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32)):
    item = np.dtype(ndt).itemsize
    n = 16384 if item == 4 else 8192
    m = 8192
    N = n * m
    a = torch.rand(N, dtype=tdt, device="cuda")
    b = a.clone()
    out = torch.empty(max(n, m), dtype=tdt, device="cuda")
    ra, rb = dev.wrap(a.data_ptr(), N, ndt), dev.wrap(b.data_ptr(), N, ndt)
    ro = dev.wrap(out.data_ptr(), max(n, m), ndt)
    flat = Layout((N,), (1,))
    full = Layout((m, n), (n, 1))
    scalar = Layout((), ())
    tag = np.dtype(ndt).name
    rep(f"{tag} vecdot 1-D ({N},)", 2 * N * item, lambda: dev.vecdot(ro, scalar, ra, flat, rb, flat, [0], [0]))
    rep(f"{tag} vecdot rows ({m},{n}) axis -1", 2 * N * item, lambda: dev.vecdot(ro, Layout((m,), (1,)), ra, full, rb, full, [1], [1]))
    rep(f"{tag} vecdot cols ({m},{n}) axis 0", 2 * N * item, lambda: dev.vecdot(ro, Layout((n,), (1,)), ra, full, rb, full, [0], [0]))
    rep(f"{tag} vecdot rows x broadcast vector", N * item,
        lambda: dev.vecdot(ro, Layout((m,), (1,)), ra, full, rb, Layout((n,), (1,)), [1], [0]))
    rep(f"{tag} allclose_all (true: reads everything, incl. sync)", 2 * N * item, lambda: dev.allclose_all(ra, flat, rb, flat))
    rep(f"{tag} sum_all (one stream, for reference)", N * item, lambda: dev.reduce_all("sum", ra, flat))
    A, B = a.view(m, n), b.view(m, n)
    rep(f"{tag} torch.dot", 2 * N * item, lambda: torch.dot(a, b))
    rep(f"{tag} torch (A*B).sum(-1) [2 kernels]", 2 * N * item, lambda: (A * B).sum(-1))
    rep(f"{tag} torch.linalg.vecdot(A, B)", 2 * N * item, lambda: torch.linalg.vecdot(A, B))
    rep(f"{tag} torch.allclose", 2 * N * item, lambda: torch.allclose(a, b))
    del a, b, A, B
