"""Scratch probe: reductions whose contiguous kept axis is short (middle-axis sums of (a, b, c) with small c)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32)):
    item = np.dtype(ndt).itemsize
    for shape in ((4096, 4096, 2), (4096, 4096, 4), (2048, 4096, 8), (1024, 4096, 16), (512, 4096, 32), (1 << 24, 4), (1 << 22, 16),
                  (4, 1 << 24), (16, 1 << 22)):
        n = int(np.prod(shape))
        a = torch.rand(n, dtype=tdt, device="cuda")
        ra = dev.wrap(a.data_ptr(), n, ndt)
        la = Layout.contig(shape, rt.ROW_MAJOR)
        axis = len(shape) - 2
        oshape = tuple(s for i, s in enumerate(shape) if i != axis)
        out = torch.empty(int(np.prod(oshape)), dtype=tdt, device="cuda")
        ro = dev.wrap(out.data_ptr(), out.numel(), ndt)
        lo = Layout.contig(oshape, rt.ROW_MAJOR)
        s = timeit(lambda: dev.reduce_axes_into("sum", ra, la, [axis], ro, lo))
        A = a.view(*shape)
        st = timeit(lambda: A.sum(axis))
        print(f"{np.dtype(ndt).name} sum axis {axis} of {shape}: {n * item / s / 1e9:8.1f} GB/s ({s * 1e6:7.1f} us)   torch {n * item / st / 1e9:8.1f} GB/s",
              flush=True)
        del a, out
