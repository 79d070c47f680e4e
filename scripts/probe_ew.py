"""Scratch probe: isolate which elementwise shapes are slow (not product)."""
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


def wrap(t):
    return dev.wrap(t.data_ptr(), t.numel(), np.float64)


n = 8192
pitch = n + 64
a = torch.rand(n * pitch, dtype=torch.float64, device="cuda")
b = torch.rand(n * pitch, dtype=torch.float64, device="cuda")
c = torch.empty(n * pitch, dtype=torch.float64, device="cuda")
ra, rb, rc = wrap(a), wrap(b), wrap(c)
N = n * n
flat = Layout((N,), (1,))
full = Layout((n, n), (n, 1))
brow = Layout((n, n), (0, 1))
bcol = Layout((n, n), (1, 0))
pitched = Layout((n, n), (pitch, 1))


def rep(name, nbytes, fn):
    s = timeit(fn)
    print(f"{name:58s} {nbytes / s / 1e9:8.1f} GB/s  {s * 1e6:8.1f} us", flush=True)


rep("assign 1-D contiguous (copy)", 2 * N * 8, lambda: dev.assign(rc, flat, ra, flat))
rep("add 1-D contiguous a+b", 3 * N * 8, lambda: dev.op_mutc_refa_refb("add", rc, flat, ra, flat, rb, flat))
rep("add scalar a+2.0 1-D", 2 * N * 8, lambda: dev.op_mutc_refa_numb("add", rc, flat, ra, flat, 2.0))
rep("add in-place a+=b 1-D", 3 * N * 8, lambda: dev.op_muta_refb("add", ra, flat, rb, flat))
rep("cfg1 add (n,n)+(n,) row broadcast", 2 * N * 8, lambda: dev.op_mutc_refa_refb("add", rc, full, ra, full, rb, brow))
rep("add (n,n)+(n,1) column broadcast", 2 * N * 8, lambda: dev.op_mutc_refa_refb("add", rc, full, ra, full, rb, bcol))
rep("add pitched rows (no broadcast, 2-D)", 3 * N * 8,
    lambda: dev.op_mutc_refa_refb("add", rc, pitched, ra, pitched, rb, pitched))
rep("copy pitched rows (2-D)", 2 * N * 8, lambda: dev.assign(rc, pitched, ra, pitched))
rep("torch copy", 2 * N * 8, lambda: c[:N].copy_(a[:N]))
rep("torch add", 3 * N * 8, lambda: torch.add(a[:N], b[:N], out=c[:N]))
rep("torch add bcast", 2 * N * 8, lambda: torch.add(a[:N].view(n, n), b[:n], out=c[:N].view(n, n)))
