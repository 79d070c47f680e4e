"""Scratch probe: transposed operands with one short axis (tile kernel vs flat kernel; RC_TILE_MIN selects)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)


def timeit(fn, iters=8, warmup=2):
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


for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32), (torch.int16, np.int16), (torch.uint8, np.uint8)):
    item = np.dtype(ndt).itemsize
    for k in (3, 4, 6, 8, 12, 16, 17, 24, 32, 48, 64, 100):
        n = (1 << 25) // k
        src = (torch.rand(k * n, device="cuda") * 100).to(tdt)
        dst = torch.empty(k * n, dtype=tdt, device="cuda")
        rs, rd = dev.wrap(src.data_ptr(), k * n, ndt), dev.wrap(dst.data_ptr(), k * n, ndt)
        # (k, n) row-major viewed transposed -> (n, k) contiguous output, and the other way round
        a = timeit(lambda: dev.assign(rd, rt.Layout((n, k), (k, 1)), rs, rt.Layout((n, k), (1, n))))
        b = timeit(lambda: dev.assign(rd, rt.Layout((k, n), (n, 1)), rs, rt.Layout((k, n), (1, k))))
        nb = 2 * k * n * item
        print(f"{np.dtype(ndt).name} k={k:4d}: (k,n).T -> (n,k) {nb / a / 1e9:7.0f} GB/s   (n,k).T -> (k,n) {nb / b / 1e9:7.0f} GB/s", flush=True)
